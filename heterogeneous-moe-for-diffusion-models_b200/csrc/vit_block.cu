// Fused DiffiT block of the ViT experts (Vit_block.forward, models/model_components.py:525-562, with the
// MP_Attention self-attention of models/model_internals.py:354-409 inside): GroupNorm(4) -> mp_silu -> linear1 ->
// LayerNorm -> time-conditioned multi-head self-attention with rel_pos_bias -> mp_sum -> mp_sum -> LayerNorm ->
// linear2 -> mp_silu -> linear3 -> mp_sum -> mp_sum, for emb = 32, 8 heads of dimension 4, time_dim = 64, hidden
// 128 and up to 64 tokens -- the shipped ViT experts at 32x32 (S = 64 / 16 / 16 / 4).
//
// The reference runs ~85 launches per block and expert (5-13 MFLOP per sample: pure launch latency, SURVEY §8a
// a10).  Here ONE CTA owns one dispatched row: its tokens (<= 8 KB) and every intermediate live in shared memory,
// the weights of the row's expert are staged from L2 per matrix, attention probabilities stay in registers
// (one warp per (head, query) row).  All experts of the layer run in the same launch -- the row's expert index is
// read on the device -- so rows are processed once instead of once per expert, and the path needs no host
// synchronisation.  fp32 throughout.
//
// Backward recomputes the block from its saved input.  Parameter gradients are accumulated without atomics in a
// CTA-private scratch slice (persistent CTAs: CTA c serves expert c % E and walks that expert's rows with stride
// G / E); the caller sums the G / E slices of each expert.
#include "common.cuh"

namespace hdmoe {

constexpr int kVD = 32;        // embedding
constexpr int kVH = 8;         // heads (dimension 4)
constexpr int kVT = 64;        // time embedding
constexpr int kVHid = 128;     // MLP hidden
constexpr int kVS = 64;        // max tokens
constexpr int kVThreads = 256;
constexpr int kVKT = kVS + 1;  // padded row stride of the transposed K / V buffers
constexpr float kVc = 0.70710678118654752f;     // mp_sum(a, b, 0.5) = (a + b) / sqrt(2)
constexpr float kVSiluInv = 1.f / 0.596f;
constexpr float kVEps = 1e-5f;

// prepared-weight block of one (expert, block): the PreparedGroup order (prepared.py: vit_expert_group)
constexpr int W_L1 = 0, W_Q = 1024, W_K = 2048, W_V = 3072, W_O = 4096, W_QT = 5120, W_KT = 7168, W_VT = 9216,
              W_L2 = 11264, W_L3 = 15360, W_TOTAL = 19456;
// aux block of one (expert, block): norms (+ final LayerNorm of the expert) and rel_pos_bias [8, S, S]
constexpr int A_GN = 0, A_LN1 = 64, A_LN2 = 128, A_LNF = 192, A_BIAS = 256;

struct VitTables {
    long long w_off[8];      // float offset of the expert's weight block inside w_hat_flat
    long long a_off[8];      // float offset of the expert's aux block inside aux
    int S[8];                // tokens of the expert
    int n_experts;
};

__device__ __forceinline__ float vsilu(float u) { return u / (1.f + __expf(-u)) * kVSiluInv; }
__device__ __forceinline__ float vdsilu(float u) {
    const float s = 1.f / (1.f + __expf(-u));
    return s * (1.f + u * (1.f - s)) * kVSiluInv;
}
// out[s][n] = epi(s, n, sum_k in[s][k] * W[n][k]) for s < S.  in: smem, row stride K.  W: global [N][K], staged
// transposed into wbuf[k][N+1].  Thread: lane -> columns n = lane + 32 j, warp -> rows s = warp + 8 i.
template <int N, int K, typename Epi>
__device__ __forceinline__ void lin_fwd(const float* __restrict__ in, const float* __restrict__ Wg, float* wbuf, int S,
                                        Epi epi) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * K; idx += kVThreads) {
        const int n = idx / K, k = idx - n * K;
        wbuf[k * (N + 1) + n] = Wg[idx];
    }
    __syncthreads();
    constexpr int NJ = N / 32, NI = kVS / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[NI][NJ];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
    for (int k4 = 0; k4 < K; k4 += 4) {
        float w[4][NJ];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int j = 0; j < NJ; ++j) w[kk][j] = wbuf[(k4 + kk) * (N + 1) + lane + 32 * j];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int s = warp + 8 * i;
            if (s < S) {
                const float4 a = *reinterpret_cast<const float4*>(in + s * K + k4);
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    acc[i][j] += a.x * w[0][j] + a.y * w[1][j] + a.z * w[2][j] + a.w * w[3][j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int s = warp + 8 * i;
        if (s < S) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) epi(s, lane + 32 * j, acc[i][j]);
        }
    }
}

// LayerNorm over the 32 channels of every token: warp per token, lane = channel.  xhat optional.
__device__ __forceinline__ void layer_norm(const float* in, float* out, float* xhat, float* rstd_out, const float* __restrict__ gb,
                                           int S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float g = gb[lane], b = gb[32 + lane];
    for (int s = warp; s < S; s += 8) {
        const float v = in[s * kVD + lane];
        const float mean = warp_sum(v) * (1.f / kVD);
        const float d = v - mean;
        const float rstd = rsqrtf(warp_sum(d * d) * (1.f / kVD) + kVEps);
        const float xh = d * rstd;
        if (xhat) xhat[s * kVD + lane] = xh;
        if (rstd_out && lane == 0) rstd_out[s] = rstd;
        out[s * kVD + lane] = xh * g + b;
    }
}

// GroupNorm(4 groups of 8 channels over all S tokens) statistics: warp g < 4 -> (mean, rstd) of group g
__device__ __forceinline__ void group_stats(const float* X, float* stats /*[8]*/, int S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 4) {
        float s = 0.f;
        for (int i = lane; i < S * 8; i += 32) s += X[(i >> 3) * kVD + warp * 8 + (i & 7)];
        const float mean = warp_sum(s) / (float)(S * 8);
        float ss = 0.f;
        for (int i = lane; i < S * 8; i += 32) {
            const float d = X[(i >> 3) * kVD + warp * 8 + (i & 7)] - mean;
            ss += d * d;
        }
        const float var = warp_sum(ss) / (float)(S * 8);
        if (lane == 0) {
            stats[warp] = mean;
            stats[4 + warp] = rsqrtf(var + kVEps);
        }
    }
}

// time projections: tp[0..95] = [Wqt; Wkt; Wvt] t  (each [32][64])
__device__ __forceinline__ void time_proj(const float* __restrict__ Wqt, const float* t, float* tp) {
    if (threadIdx.x < 96) {
        const float4* w = reinterpret_cast<const float4*>(Wqt + (size_t)threadIdx.x * kVT);     // W_QT, W_KT, W_VT are contiguous
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kVT / 4; ++k) {
            const float4 ww = w[k];
            a += ww.x * t[4 * k] + ww.y * t[4 * k + 1] + ww.z * t[4 * k + 2] + ww.w * t[4 * k + 3];
        }
        tp[threadIdx.x] = a;
    }
}

// self-attention of one row: warp per (head, query); lanes over keys (2 per lane for S = 64).
// Q [S][32]; KT, VT [32][kVKT] (transposed); bias [8][S][S]; O [S][32]
__device__ __forceinline__ void attention_fwd(const float* Q, const float* KT, const float* VT, const float* __restrict__ bias,
                                              float* O, int S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int pair = warp; pair < kVH * S; pair += 8) {
        const int h = pair / S, i = pair - h * S;
        const float4 q = *reinterpret_cast<const float4*>(Q + i * kVD + 4 * h);
        float sc[2], mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = lane + 32 * u;
            sc[u] = -INFINITY;
            if (j < S) {
                const float* kt = KT + (4 * h) * kVKT + j;
                sc[u] = 0.5f * (q.x * kt[0] + q.y * kt[kVKT] + q.z * kt[2 * kVKT] + q.w * kt[3 * kVKT]) +
                        bias[((size_t)h * S + i) * S + j];
            }
            mx = fmaxf(mx, sc[u]);
        }
        mx = warp_max(mx);
        float p[2], sum = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = lane + 32 * u;
            p[u] = j < S ? __expf(sc[u] - mx) : 0.f;
            sum += p[u];
            if (j < S) {
                const float* vt = VT + (4 * h) * kVKT + j;
                o0 += p[u] * vt[0]; o1 += p[u] * vt[kVKT]; o2 += p[u] * vt[2 * kVKT]; o3 += p[u] * vt[3 * kVKT];
            }
        }
        sum = warp_sum(sum); o0 = warp_sum(o0); o1 = warp_sum(o1); o2 = warp_sum(o2); o3 = warp_sum(o3);
        if (lane == 0) {
            const float inv = 1.f / sum;
            *reinterpret_cast<float4*>(O + i * kVD + 4 * h) = make_float4(o0 * inv, o1 * inv, o2 * inv, o3 * inv);
        }
    }
}

struct VitSmemFwd {
    float X[kVS * kVD], A[kVS * kVD], H1[kVS * kVD], Y[kVS * kVD], Q[kVS * kVD], O[kVS * kVD], Y2[kVS * kVD];
    float KT[kVD * kVKT], VT[kVD * kVKT];
    float M[kVS * kVHid];
    float wbuf[kVHid * 33];        // >= 32 * 129
    float t[kVT], tp[96], stats[8];
};

// forward of rows [0, R): tok_in / tok_out [R][kVS][32]; final_ln: apply the expert's final LayerNorm (A_LNF)
__device__ void vit_block_row_fwd(VitSmemFwd& sm, const float* __restrict__ W, const float* __restrict__ aux, int S, int final_ln,
                                  float* __restrict__ out_row) {
    const int tid = threadIdx.x;
    group_stats(sm.X, sm.stats, S);
    __syncthreads();
    for (int i = tid; i < S * kVD; i += kVThreads) {
        const int c = i & 31, g = c >> 3;
        sm.A[i] = vsilu((sm.X[i] - sm.stats[g]) * sm.stats[4 + g] * aux[A_GN + c] + aux[A_GN + 32 + c]);
    }
    lin_fwd<kVD, kVD>(sm.A, W + W_L1, sm.wbuf, S, [&](int s, int n, float v) { sm.H1[s * kVD + n] = v; });
    __syncthreads();
    layer_norm(sm.H1, sm.Y, nullptr, nullptr, aux + A_LN1, S);
    time_proj(W + W_QT, sm.t, sm.tp);
    lin_fwd<kVD, kVD>(sm.Y, W + W_Q, sm.wbuf, S, [&](int s, int n, float v) { sm.Q[s * kVD + n] = v + sm.tp[n]; });
    lin_fwd<kVD, kVD>(sm.Y, W + W_K, sm.wbuf, S, [&](int s, int n, float v) { sm.KT[n * kVKT + s] = v + sm.tp[32 + n]; });
    lin_fwd<kVD, kVD>(sm.Y, W + W_V, sm.wbuf, S, [&](int s, int n, float v) { sm.VT[n * kVKT + s] = v + sm.tp[64 + n]; });
    __syncthreads();
    attention_fwd(sm.Q, sm.KT, sm.VT, aux + A_BIAS, sm.O, S);
    // Y2 = mp_sum(mp_sum(Y, out_proj(O)), H1)
    lin_fwd<kVD, kVD>(sm.O, W + W_O, sm.wbuf, S, [&](int s, int n, float v) {
        const int i = s * kVD + n;
        sm.Y2[i] = kVc * (kVc * (sm.Y[i] + v) + sm.H1[i]);
    });
    __syncthreads();
    layer_norm(sm.Y2, sm.A, nullptr, nullptr, aux + A_LN2, S);               // A = LN2(Y2)
    lin_fwd<kVHid, kVD>(sm.A, W + W_L2, sm.wbuf, S, [&](int s, int n, float v) { sm.M[s * kVHid + n] = vsilu(v); });
    // out = mp_sum(X, mp_sum(linear3(M), Y2))
    lin_fwd<kVD, kVHid>(sm.M, W + W_L3, sm.wbuf, S, [&](int s, int n, float v) {
        const int i = s * kVD + n;
        sm.O[i] = kVc * (sm.X[i] + kVc * (v + sm.Y2[i]));
    });
    __syncthreads();
    if (final_ln) {
        layer_norm(sm.O, sm.A, nullptr, nullptr, aux + A_LNF, S);
        __syncthreads();
    }
    const float* res = final_ln ? sm.A : sm.O;
    for (int i = tid; i < kVS * kVD; i += kVThreads) out_row[i] = i < S * kVD ? res[i] : 0.f;
}

__global__ void __launch_bounds__(kVThreads)
vit_block_fwd_kernel(const float* __restrict__ tok_in, const float* __restrict__ time, const int* __restrict__ row_expert,
                     const float* __restrict__ w_hat, const float* __restrict__ aux, VitTables tb, int final_ln,
                     float* __restrict__ tok_out) {
    extern __shared__ __align__(16) unsigned char smraw[];
    VitSmemFwd& sm = *reinterpret_cast<VitSmemFwd*>(smraw);
    const int r = blockIdx.x, tid = threadIdx.x;
    const int e = row_expert[r];
    float* out_row = tok_out + (size_t)r * kVS * kVD;
    if (e < 0 || e >= tb.n_experts) {
        for (int i = tid; i < kVS * kVD; i += kVThreads) out_row[i] = 0.f;
        return;
    }
    const int S = tb.S[e];
    const float4* src = reinterpret_cast<const float4*>(tok_in + (size_t)r * kVS * kVD);
    for (int i = tid; i < S * kVD / 4; i += kVThreads) reinterpret_cast<float4*>(sm.X)[i] = src[i];
    if (tid < kVT) sm.t[tid] = time[(size_t)r * kVT + tid];
    __syncthreads();
    vit_block_row_fwd(sm, w_hat + tb.w_off[e], aux + tb.a_off[e], S, final_ln, out_row);
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_vit_block_fwd(const float* tok_in, const float* time, const int32_t* row_expert, const float* w_hat,
                                   const float* aux, const int64_t* w_off, const int64_t* a_off, const int32_t* tokens,
                                   int n_experts, int64_t rows, int final_ln, float* tok_out, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(tok_in && time && row_expert && w_hat && aux && w_off && a_off && tokens && tok_out && rows >= 1,
                    "vit_block_fwd: null pointer / no rows");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= 8, "vit_block_fwd: 1..8 experts (got %d)", n_experts);
    VitTables tb;
    tb.n_experts = n_experts;
    for (int e = 0; e < n_experts; ++e) {
        HDMOE_CHECK_ARG(tokens[e] >= 1 && tokens[e] <= kVS, "vit_block_fwd: 1..64 tokens per row (got %d)", tokens[e]);
        tb.w_off[e] = w_off[e];
        tb.a_off[e] = a_off[e];
        tb.S[e] = tokens[e];
    }
    static bool attr = false;
    if (!attr) {
        HDMOE_CHECK_ARG(cudaFuncSetAttribute(vit_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(VitSmemFwd)) == cudaSuccess, "vit_block_fwd: shared memory attribute");
        attr = true;
    }
    vit_block_fwd_kernel<<<(unsigned)rows, kVThreads, sizeof(VitSmemFwd), (cudaStream_t)stream>>>(
        tok_in, time, row_expert, w_hat, aux, tb, final_ln, tok_out);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
