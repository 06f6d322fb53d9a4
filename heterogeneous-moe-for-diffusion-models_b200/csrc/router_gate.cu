// R-GATE: fused router tail (adaLN modulate -> linear -> +zeta*noise -> mask -> softmax -> top-k ->
// softmax(top-k) -> sparse scatter) with warp-shuffle reductions for the load-balance / z-loss
// statistics.  Replaces ~14 ATen launches of Router.forward (models/model_components.py:148-168) and
// ~8 of EDM_LOSS.load_balance / z_loss (Utils/utils.py:158-172).
//
// Mapping: one warp per token.  Experts are padded to EP = 2^n.  EP <= 32: the 32 lanes form EP groups
// of G = 32/EP lanes, lane (e, g) accumulates channels c = g, g+G, ... of expert e and a log2(G)-step
// shuffle finishes the dot product.  EP == 64: each lane owns experts lane and lane+32.  The weight
// matrix sits in shared memory with row stride WS == G (mod 32) so that the 32 lanes of one read hit 32
// distinct banks; the modulated feature row is staged per warp and read as a broadcast.
// HBM-bound: per token (C + 2C) fp32 in, 3E fp32 + k(4+4) B out.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace hdmoe {

constexpr int kGateWarps = 8;
constexpr int kGateThreads = kGateWarps * 32;

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return (v > bv) || (v == bv && i < bi); }

template <int EP>
__global__ void __launch_bounds__(kGateThreads)
router_gate_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ cond,
                       const float* __restrict__ w_hat, const float* __restrict__ noise, float zeta,
                       const float* __restrict__ mask, const float* __restrict__ logits_in, int T, int C, int E,
                       int top_k, float* __restrict__ logits, float* __restrict__ gate_probs,
                       float* __restrict__ sparse_w, int32_t* __restrict__ topk_idx, float* __restrict__ topk_w,
                       float* __restrict__ stats, float* __restrict__ partial, unsigned* __restrict__ ticket) {
    constexpr int G = EP <= 32 ? 32 / EP : 1;
    constexpr int EL = EP <= 32 ? 1 : EP / 32;
    extern __shared__ float smem[];
    const int WS = ((C + 31) / 32) * 32 + G;
    float* ws = smem;                                   // [EP][WS]
    float* xs_all = ws + (logits_in ? 0 : EP * WS);     // [warps][C]
    float* red = xs_all + (logits_in ? 0 : kGateWarps * C);  // [warps][2*EP+1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstat = 2 * E + 1;

    if (!logits_in) {
        for (int i = threadIdx.x; i < EP * C; i += kGateThreads) {
            int e = i / C, c = i - e * C;
            ws[e * WS + c] = e < E ? w_hat[(size_t)e * C + c] : 0.f;
        }
    }
    __syncthreads();

    int e[EL];
    bool own[EL];
    const int g = lane % G;
#pragma unroll
    for (int q = 0; q < EL; ++q) {
        e[q] = (EP <= 32 ? lane / G : lane) + 32 * q;
        own[q] = (g == 0) && e[q] < E;
    }
    float colsum[EL], cnt[EL], zacc = 0.f;
#pragma unroll
    for (int q = 0; q < EL; ++q) colsum[q] = cnt[q] = 0.f;

    float* xs = xs_all + warp * C;
    for (int t = blockIdx.x * kGateWarps + warp; t < T; t += gridDim.x * kGateWarps) {
        float l[EL];
        if (logits_in) {
#pragma unroll
            for (int q = 0; q < EL; ++q) l[q] = e[q] < E ? logits_in[(size_t)t * E + e[q]] : -INFINITY;
        } else {
            // modulated features -> per-warp smem (coalesced loads)
            for (int c = lane; c < C; c += 32) {
                float x = pooled[(size_t)t * C + c];
                if (cond) x = x * (1.f + cond[(size_t)t * 2 * C + c]) + cond[(size_t)t * 2 * C + C + c];
                xs[c] = x;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < EL; ++q) {
                float acc = 0.f;
                const float* wr = ws + e[q] * WS;
                for (int c = g; c < C; c += G) acc = fmaf(xs[c], wr[c], acc);
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                l[q] = acc;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < EL; ++q) {
                if (e[q] < E) {
                    if (noise) l[q] += noise[(size_t)t * E + e[q]] * zeta;   // :156
                } else {
                    l[q] = -INFINITY;
                }
            }
        }
        if (mask) {
#pragma unroll
            for (int q = 0; q < EL; ++q)
                if (e[q] < E && mask[(size_t)t * E + e[q]] == 0.f) l[q] = -INFINITY;   // :160
        }

        // ---- softmax over experts
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < EL; ++q)
            if (own[q]) m = fmaxf(m, l[q]);
        m = warp_max(m);
        float p[EL], s = 0.f;
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            p[q] = own[q] ? expf(l[q] - m) : 0.f;  // all-masked row: -inf - -inf = NaN like torch
            s += p[q];
        }
        s = warp_sum(s);
#pragma unroll
        for (int q = 0; q < EL; ++q) p[q] = p[q] / s;

        // ---- z-loss term (Utils/utils.py:169-171)
        float m2 = -INFINITY, cl[EL], s2 = 0.f;
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            cl[q] = fminf(fmaxf(l[q], -50.f), 50.f);
            if (own[q]) m2 = fmaxf(m2, cl[q]);
        }
        m2 = warp_max(m2);
#pragma unroll
        for (int q = 0; q < EL; ++q) s2 += own[q] ? expf(cl[q] - m2) : 0.f;
        s2 = warp_sum(s2);
        const float lse = m2 + logf(s2);
        zacc += fminf(lse * lse, 100.f);

        // ---- top-k by iterative arg-max (lowest index wins ties), then softmax over the k values
        bool taken[EL];
        float sw[EL];
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            taken[q] = false;
            sw[q] = 0.f;
        }
        float v0 = 0.f, wsum = 0.f, my_w = 0.f;
        int my_i = -1;
        int slot_of[EL];
#pragma unroll
        for (int q = 0; q < EL; ++q) slot_of[q] = -1;
        for (int j = 0; j < top_k; ++j) {
            float bv = -INFINITY;
            int bi = INT_MAX;
#pragma unroll
            for (int q = 0; q < EL; ++q)
                if (own[q] && !taken[q] && better(l[q], e[q], bv, bi)) {
                    bv = l[q];
                    bi = e[q];
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better(ov, oi, bv, bi)) {
                    bv = ov;
                    bi = oi;
                }
            }
            if (j == 0) v0 = bv;
            const float ej = expf(bv - v0);
            wsum += ej;
#pragma unroll
            for (int q = 0; q < EL; ++q)
                if (e[q] == bi) {
                    taken[q] = true;
                    sw[q] = ej;
                    slot_of[q] = j;
                }
            if (lane == j) {
                my_w = ej;
                my_i = bi;
            }
        }
#pragma unroll
        for (int q = 0; q < EL; ++q) sw[q] = slot_of[q] >= 0 ? sw[q] / wsum : 0.f;
        if (lane < top_k) {
            topk_idx[(size_t)t * top_k + lane] = my_i;
            topk_w[(size_t)t * top_k + lane] = my_w / wsum;
        }
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            if (own[q]) {
                const size_t o = (size_t)t * E + e[q];
                logits[o] = l[q];
                gate_probs[o] = p[q];
                sparse_w[o] = sw[q];
                colsum[q] += p[q];
                cnt[q] += sw[q] > 0.f ? 1.f : 0.f;
            }
        }
    }

    // ---- deterministic statistics: warp -> block (fixed order) -> grid (last block, fixed order)
    float* r = red + warp * (2 * EP + 1);
#pragma unroll
    for (int q = 0; q < EL; ++q)
        if (own[q]) {
            r[e[q]] = colsum[q];
            r[EP + e[q]] = cnt[q];
        }
    if (lane == 0) r[2 * EP] = zacc;
    __syncthreads();
    if (threadIdx.x < nstat) {
        const int i = threadIdx.x;
        const int src = i < E ? i : (i < 2 * E ? EP + (i - E) : 2 * EP);
        float a = 0.f;
        for (int w = 0; w < kGateWarps; ++w) a += red[w * (2 * EP + 1) + src];
        partial[(size_t)blockIdx.x * nstat + i] = a;
    }
    __threadfence();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        __threadfence();
        if (threadIdx.x < nstat) {
            float a = 0.f;
            for (unsigned b = 0; b < gridDim.x; ++b) a += partial[(size_t)b * nstat + threadIdx.x];
            stats[threadIdx.x] = a;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
template <int EP>
__global__ void __launch_bounds__(kGateThreads)
router_gate_bwd_kernel(const float* __restrict__ pooled, const float* __restrict__ cond,
                       const float* __restrict__ w_hat, const float* __restrict__ logits,
                       const int32_t* __restrict__ topk_idx, const float* __restrict__ g_sparse,
                       const float* __restrict__ g_probs, const float* __restrict__ g_logits,
                       const float* __restrict__ g_stats, int T, int C, int E, int top_k,
                       float* __restrict__ d_pooled, float* __restrict__ d_cond, float* __restrict__ d_w_hat,
                       float* __restrict__ d_logits_out) {
    constexpr int G = EP <= 32 ? 32 / EP : 1;
    constexpr int EL = EP <= 32 ? 1 : EP / 32;
    extern __shared__ float smem[];
    const bool lin = pooled != nullptr;
    const int WS = ((C + 31) / 32) * 32 + 1;
    float* ws = smem;                                         // [EP][WS]
    float* dws = ws + (lin ? EP * WS : 0);                    // [EP][WS]
    float* xs_all = dws + (lin ? EP * WS : 0);                // [warps][C]
    float* dls_all = xs_all + (lin ? kGateWarps * C : 0);     // [warps][EP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lin) {
        for (int i = threadIdx.x; i < EP * C; i += kGateThreads) {
            int e = i / C, c = i - e * C;
            ws[e * WS + c] = e < E ? w_hat[(size_t)e * C + c] : 0.f;
            dws[e * WS + c] = 0.f;
        }
    }
    __syncthreads();
    int e[EL];
    bool own[EL];
    const int g = lane % G;
#pragma unroll
    for (int q = 0; q < EL; ++q) {
        e[q] = (EP <= 32 ? lane / G : lane) + 32 * q;
        own[q] = (g == 0) && e[q] < E;
    }
    float* xs = xs_all + warp * C;
    float* dls = dls_all + warp * EP;
    const float g_z = g_stats ? g_stats[2 * E] : 0.f;

    for (int t = blockIdx.x * kGateWarps + warp; t < T; t += gridDim.x * kGateWarps) {
        float l[EL], p[EL], dl[EL];
#pragma unroll
        for (int q = 0; q < EL; ++q) l[q] = e[q] < E ? logits[(size_t)t * E + e[q]] : -INFINITY;
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < EL; ++q)
            if (own[q]) m = fmaxf(m, l[q]);
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            p[q] = own[q] ? expf(l[q] - m) : 0.f;
            s += p[q];
        }
        s = warp_sum(s);
        // softmax backward: dl = p * (gp - sum(p*gp)), gp = g_probs + g_colsum
        float gp[EL], dot = 0.f;
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            p[q] /= s;
            gp[q] = 0.f;
            if (own[q]) {
                if (g_probs) gp[q] += g_probs[(size_t)t * E + e[q]];
                if (g_stats) gp[q] += g_stats[e[q]];
            }
            dot += own[q] ? p[q] * gp[q] : 0.f;
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int q = 0; q < EL; ++q) dl[q] = own[q] ? p[q] * (gp[q] - dot) : 0.f;

        // z-loss backward: d/dl min(lse^2, 100) = 2*lse*softmax(clamp(l)) where |l| <= 50 and lse^2 <= 100
        if (g_z != 0.f) {
            float m2 = -INFINITY, cl[EL], s2 = 0.f, ex2[EL];
#pragma unroll
            for (int q = 0; q < EL; ++q) {
                cl[q] = fminf(fmaxf(l[q], -50.f), 50.f);
                if (own[q]) m2 = fmaxf(m2, cl[q]);
            }
            m2 = warp_max(m2);
#pragma unroll
            for (int q = 0; q < EL; ++q) {
                ex2[q] = own[q] ? expf(cl[q] - m2) : 0.f;
                s2 += ex2[q];
            }
            s2 = warp_sum(s2);
            const float lse = m2 + logf(s2);
            if (lse * lse <= 100.f) {
#pragma unroll
                for (int q = 0; q < EL; ++q)
                    if (own[q] && l[q] >= -50.f && l[q] <= 50.f) dl[q] += g_z * 2.f * lse * (ex2[q] / s2);
            }
        }
        // top-k softmax backward
        if (g_sparse) {
            float v0 = -INFINITY, wsum = 0.f, wj[HDMOE_MAX_TOPK], gj[HDMOE_MAX_TOPK];
            int ij[HDMOE_MAX_TOPK];
            for (int j = 0; j < top_k; ++j) {
                ij[j] = topk_idx[(size_t)t * top_k + j];
                const float v = logits[(size_t)t * E + ij[j]];
                if (j == 0) v0 = v;
                wj[j] = expf(v - v0);
                wsum += wj[j];
                gj[j] = g_sparse[(size_t)t * E + ij[j]];
            }
            float d2 = 0.f;
            for (int j = 0; j < top_k; ++j) {
                wj[j] /= wsum;
                d2 += wj[j] * gj[j];
            }
            for (int j = 0; j < top_k; ++j) {
#pragma unroll
                for (int q = 0; q < EL; ++q)
                    if (own[q] && e[q] == ij[j]) dl[q] += wj[j] * (gj[j] - d2);
            }
        }
#pragma unroll
        for (int q = 0; q < EL; ++q) {
            if (own[q]) {
                if (g_logits) dl[q] += g_logits[(size_t)t * E + e[q]];
                // masked_fill backward: no gradient through a masked (-inf) logit; an all-masked row has NaN
                // probabilities whose products are dropped here exactly as autograd's masked_fill does.
                if (l[q] == -INFINITY) dl[q] = 0.f;
                if (d_logits_out) d_logits_out[(size_t)t * E + e[q]] = dl[q];
            }
        }
        if (!lin) continue;
        // the masked logit had no gradient w.r.t. the linear output; NaN rows contribute nothing.
#pragma unroll
        for (int q = 0; q < EL; ++q)
            if (e[q] < EP && g == 0) dls[e[q]] = (e[q] < E && dl[q] == dl[q]) ? dl[q] : 0.f;
        for (int c = lane; c < C; c += 32) {
            float x = pooled[(size_t)t * C + c];
            if (cond) x = x * (1.f + cond[(size_t)t * 2 * C + c]) + cond[(size_t)t * 2 * C + C + c];
            xs[c] = x;
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
            float dx = 0.f;
            const float xm = xs[c];
            for (int ee = 0; ee < E; ++ee) {
                const float d = dls[ee];
                dx = fmaf(d, ws[ee * WS + c], dx);
                if (d != 0.f) atomicAdd(&dws[ee * WS + c], d * xm);
            }
            if (cond) {
                const float gm = cond[(size_t)t * 2 * C + c];
                d_pooled[(size_t)t * C + c] = dx * (1.f + gm);
                d_cond[(size_t)t * 2 * C + c] = dx * pooled[(size_t)t * C + c];
                d_cond[(size_t)t * 2 * C + C + c] = dx;
            } else {
                d_pooled[(size_t)t * C + c] = dx;
            }
        }
        __syncwarp();
    }
    if (!lin) return;
    __syncthreads();
    for (int i = threadIdx.x; i < E * C; i += kGateThreads) {
        int ee = i / C, c = i - ee * C;
        const float v = dws[ee * WS + c];
        if (v != 0.f) atomicAdd(&d_w_hat[i], v);
    }
}

static int pad_experts(int E) {
    int ep = 4;
    while (ep < E) ep <<= 1;
    return ep;
}
static int gate_grid(int T) {
    int b = (T + kGateWarps - 1) / kGateWarps;
    const int cap = kNumSMs * 4;
    return b < 1 ? 1 : (b > cap ? cap : b);
}

}  // namespace hdmoe

using namespace hdmoe;

extern "C" size_t hdmoe_router_gate_workspace_bytes(int T, int E) {
    return ((size_t)gate_grid(T) * (2 * E + 1) + 4) * sizeof(float);
}

#define GATE_DISPATCH(EPV, KERNEL, SMEM, ...)                                                              \
    case EPV: {                                                                                            \
        auto kfn = KERNEL<EPV>;                                                                            \
        if ((SMEM) > 48 * 1024)                                                                            \
            HDMOE_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
        kfn<<<grid, kGateThreads, (SMEM), st>>>(__VA_ARGS__);                                              \
    } break;

extern "C" int hdmoe_router_gate_fwd(const float* pooled, const float* cond, const float* w_hat, const float* noise,
                                     float zeta, const float* mask, const float* logits_in, int T, int C, int E,
                                     int top_k, float* logits, float* gate_probs, float* sparse_w,
                                     int32_t* topk_idx, float* topk_w, float* stats, void* workspace,
                                     hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(T >= 0 && E >= 1 && E <= HDMOE_MAX_EXPERTS, "router_gate: need 1 <= E <= %d (got %d)",
                    HDMOE_MAX_EXPERTS, E);
    HDMOE_CHECK_ARG(top_k >= 1 && top_k <= HDMOE_MAX_TOPK && top_k <= E, "router_gate: bad top_k %d (E=%d)", top_k, E);
    HDMOE_CHECK_ARG(logits_in || (pooled && w_hat && C >= 1 && C <= 1024), "router_gate: need pooled/w_hat, C<=1024");
    HDMOE_CHECK_ARG(logits && gate_probs && sparse_w && topk_idx && topk_w && stats && workspace,
                    "router_gate: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int EP = pad_experts(E);
    const int grid = gate_grid(T);
    // the ticket sits at a FIXED offset (word 0) so that calls with different grids share one zeroed counter
    unsigned* ticket = (unsigned*)workspace;
    float* partial = (float*)workspace + 4;
    if (T == 0) {
        HDMOE_CHECK_CUDA(cudaMemsetAsync(stats, 0, (2 * E + 1) * sizeof(float), st));
        return HDMOE_OK;
    }
    const int G = EP <= 32 ? 32 / EP : 1;
    const int WS = ((C + 31) / 32) * 32 + G;
    size_t smem = (size_t)kGateWarps * (2 * EP + 1) * sizeof(float);
    if (!logits_in) smem += ((size_t)EP * WS + (size_t)kGateWarps * C) * sizeof(float);
    HDMOE_CHECK_ARG(smem <= 200 * 1024, "router_gate: E*C too large for shared memory");
    switch (EP) {
        GATE_DISPATCH(4, router_gate_fwd_kernel, smem, pooled, cond, w_hat, noise, zeta, mask, logits_in, T, C, E,
                      top_k, logits, gate_probs, sparse_w, topk_idx, topk_w, stats, partial, ticket)
        GATE_DISPATCH(8, router_gate_fwd_kernel, smem, pooled, cond, w_hat, noise, zeta, mask, logits_in, T, C, E,
                      top_k, logits, gate_probs, sparse_w, topk_idx, topk_w, stats, partial, ticket)
        GATE_DISPATCH(16, router_gate_fwd_kernel, smem, pooled, cond, w_hat, noise, zeta, mask, logits_in, T, C, E,
                      top_k, logits, gate_probs, sparse_w, topk_idx, topk_w, stats, partial, ticket)
        GATE_DISPATCH(32, router_gate_fwd_kernel, smem, pooled, cond, w_hat, noise, zeta, mask, logits_in, T, C, E,
                      top_k, logits, gate_probs, sparse_w, topk_idx, topk_w, stats, partial, ticket)
        GATE_DISPATCH(64, router_gate_fwd_kernel, smem, pooled, cond, w_hat, noise, zeta, mask, logits_in, T, C, E,
                      top_k, logits, gate_probs, sparse_w, topk_idx, topk_w, stats, partial, ticket)
        default:
            HDMOE_CHECK_ARG(false, "router_gate: unsupported padded expert count %d", EP);
    }
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_router_gate_bwd(const float* pooled, const float* cond, const float* w_hat, const float* logits,
                                     const int32_t* topk_idx, const float* g_sparse, const float* g_probs,
                                     const float* g_logits, const float* g_stats, int T, int C, int E, int top_k,
                                     float* d_pooled, float* d_cond, float* d_w_hat, float* d_logits_out,
                                     hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(T >= 0 && E >= 1 && E <= HDMOE_MAX_EXPERTS && top_k >= 1 && top_k <= HDMOE_MAX_TOPK && top_k <= E,
                    "router_gate_bwd: bad E/top_k");
    HDMOE_CHECK_ARG(logits && topk_idx, "router_gate_bwd: logits/topk_idx required");
    const bool lin = pooled != nullptr;
    HDMOE_CHECK_ARG(!lin || (w_hat && d_pooled && d_w_hat && (!cond || d_cond) && C >= 1 && C <= 1024),
                    "router_gate_bwd: linear part needs w_hat, d_pooled, d_w_hat (and d_cond with cond)");
    cudaStream_t st = (cudaStream_t)stream;
    if (lin) HDMOE_CHECK_CUDA(cudaMemsetAsync(d_w_hat, 0, (size_t)E * C * sizeof(float), st));
    if (T == 0) return HDMOE_OK;
    const int EP = pad_experts(E);
    const int grid = gate_grid(T);
    const int WS = ((C + 31) / 32) * 32 + 1;
    size_t smem = 0;
    if (lin) smem = ((size_t)2 * EP * WS + (size_t)kGateWarps * C + (size_t)kGateWarps * EP) * sizeof(float);
    HDMOE_CHECK_ARG(smem <= 200 * 1024, "router_gate_bwd: E*C too large for shared memory");
    switch (EP) {
        GATE_DISPATCH(4, router_gate_bwd_kernel, smem, pooled, cond, w_hat, logits, topk_idx, g_sparse, g_probs,
                      g_logits, g_stats, T, C, E, top_k, d_pooled, d_cond, d_w_hat, d_logits_out)
        GATE_DISPATCH(8, router_gate_bwd_kernel, smem, pooled, cond, w_hat, logits, topk_idx, g_sparse, g_probs,
                      g_logits, g_stats, T, C, E, top_k, d_pooled, d_cond, d_w_hat, d_logits_out)
        GATE_DISPATCH(16, router_gate_bwd_kernel, smem, pooled, cond, w_hat, logits, topk_idx, g_sparse, g_probs,
                      g_logits, g_stats, T, C, E, top_k, d_pooled, d_cond, d_w_hat, d_logits_out)
        GATE_DISPATCH(32, router_gate_bwd_kernel, smem, pooled, cond, w_hat, logits, topk_idx, g_sparse, g_probs,
                      g_logits, g_stats, T, C, E, top_k, d_pooled, d_cond, d_w_hat, d_logits_out)
        GATE_DISPATCH(64, router_gate_bwd_kernel, smem, pooled, cond, w_hat, logits, topk_idx, g_sparse, g_probs,
                      g_logits, g_stats, T, C, E, top_k, d_pooled, d_cond, d_w_hat, d_logits_out)
        default:
            HDMOE_CHECK_ARG(false, "router_gate_bwd: unsupported padded expert count %d", EP);
    }
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
