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


// ================================================================================================ backward
// General per-row product out[s][n] = epi(s, n, sum_k f(in[s][k]) * Wt[k][n]), s < S, with the weight staged in
// wbuf[k][N + 1].  TRANS: Wg is [N][K] (forward use, staged transposed); !TRANS: Wg is [K][N] (data gradient:
// out = dOut * W with W stored [N_out = K][K_in = N], staged as it lies).  SILU_IN applies mp_silu to the input.
template <int N, int K, bool TRANS, bool SILU_IN, typename Epi>
__device__ __forceinline__ void lin_gen(const float* __restrict__ in, const float* __restrict__ Wg, float* wbuf, int S, Epi epi) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * K; idx += kVThreads) {
        if (TRANS) {
            const int n = idx / K, k = idx - n * K;
            wbuf[k * (N + 1) + n] = Wg[idx];
        } else {
            const int k = idx / N, n = idx - k * N;
            wbuf[k * (N + 1) + n] = Wg[idx];
        }
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
                float4 a = *reinterpret_cast<const float4*>(in + s * K + k4);
                if (SILU_IN) {
                    a.x = vsilu(a.x); a.y = vsilu(a.y); a.z = vsilu(a.z); a.w = vsilu(a.w);
                }
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

__device__ __forceinline__ void vred4(float* addr, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// weight gradient of one linear layer: dW[n][k] += sum_s d[s][n] * f(in[s][k]);  d: [S][N], in: [S][K] (shared memory),
// dW: global [N][K].  A thread owns float4 pieces of dW and adds them with one vector atomic each.
template <int N, int K, bool SILU_IN>
__device__ __forceinline__ void outer_acc(const float* d, const float* in, float* __restrict__ dW, int S) {
    constexpr int K4 = K / 4;
    for (int idx = threadIdx.x; idx < N * K4; idx += kVThreads) {
        const int n = idx / K4, k4 = (idx - n * K4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
            const float g = d[s * N + n];
            float4 a = *reinterpret_cast<const float4*>(in + s * K + k4);
            if (SILU_IN) {
                a.x = vsilu(a.x); a.y = vsilu(a.y); a.z = vsilu(a.z); a.w = vsilu(a.w);
            }
            acc.x += g * a.x; acc.y += g * a.y; acc.z += g * a.z; acc.w += g * a.w;
        }
        vred4(dW + n * K + k4, acc);
    }
}

// LayerNorm backward: dx[s][c] (+)= rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)), dxh = dy * gamma;
// d_gamma += sum_s dy * xhat, d_beta += sum_s dy (atomics).  scale_old: dx = scale_old * dx_old + ln_bwd.
// dy and dx may alias.
__device__ __forceinline__ void layer_norm_bwd(const float* dy, const float* xhat, const float* rstd, const float* __restrict__ gb,
                                               float* dx, float scale_old, bool accumulate, float* __restrict__ d_gb, int S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float g = gb[lane];
    float dg = 0.f, db = 0.f;
    for (int s = warp; s < S; s += 8) {
        const float y = dy[s * kVD + lane], xh = xhat[s * kVD + lane];
        dg += y * xh;
        db += y;
        const float dxh = y * g;
        const float m1 = warp_sum(dxh) * (1.f / kVD), m2 = warp_sum(dxh * xh) * (1.f / kVD);
        const float v = rstd[s] * (dxh - m1 - xh * m2);
        dx[s * kVD + lane] = accumulate ? scale_old * dx[s * kVD + lane] + v : v;
    }
    atomicAdd(d_gb + lane, dg);
    atomicAdd(d_gb + 32 + lane, db);
}

struct VitSmemBwd {
    float X[kVS * kVD], A[kVS * kVD], XH1[kVS * kVD], Y[kVS * kVD], Q[kVS * kVD], O[kVS * kVD], XH2[kVS * kVD], A2[kVS * kVD];
    float KT[kVD * kVKT], VT[kVD * kVKT];
    float P2[kVS * kVHid];         // pre-activation of linear2, later its gradient
    float wbuf[kVHid * 33];
    float G0[kVS * kVD], G1[kVS * kVD], G2[kVS * kVD], G3[kVS * kVD], G4[kVS * kVD];
    float DK[kVD * kVKT], DV[kVD * kVKT];      // used row-major [S][32]
    float t[kVT], tp[96], dtp[96], stats[8], rstd1[kVS], rstd2[kVS], rstdF[kVS], gsum[8];
};

// attention backward of one row: warp h owns head h and walks the queries; its lanes own the keys j = lane, lane + 32
// and keep dK / dV of those keys in registers.  dO: [S][32]; outputs dQ, dK, dV row-major [S][32]; d_bias: global.
__device__ __forceinline__ void attention_bwd(const float* Q, const float* KT, const float* VT, const float* __restrict__ bias,
                                              const float* dO, float* dQ, float* dK, float* dV, float* __restrict__ d_bias,
                                              int S) {
    const int lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    float k_[2][4], v_[2][4], dk[2][4], dv[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            k_[u][d] = j < S ? KT[(4 * h + d) * kVKT + j] : 0.f;
            v_[u][d] = j < S ? VT[(4 * h + d) * kVKT + j] : 0.f;
            dk[u][d] = dv[u][d] = 0.f;
        }
    }
    for (int i = 0; i < S; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(Q + i * kVD + 4 * h);
        const float4 go = *reinterpret_cast<const float4*>(dO + i * kVD + 4 * h);
        float sc[2], mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = lane + 32 * u;
            sc[u] = j < S ? 0.5f * (q.x * k_[u][0] + q.y * k_[u][1] + q.z * k_[u][2] + q.w * k_[u][3]) +
                                bias[((size_t)h * S + i) * S + j]
                          : -INFINITY;
            mx = fmaxf(mx, sc[u]);
        }
        mx = warp_max(mx);
        float pr[2], sum = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            pr[u] = (lane + 32 * u) < S ? __expf(sc[u] - mx) : 0.f;
            sum += pr[u];
        }
        const float inv = 1.f / warp_sum(sum);
        float dp[2], dot = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            pr[u] *= inv;
            dp[u] = go.x * v_[u][0] + go.y * v_[u][1] + go.z * v_[u][2] + go.w * v_[u][3];
            dot += pr[u] * dp[u];
        }
        dot = warp_sum(dot);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = lane + 32 * u;
            const float ds = pr[u] * (dp[u] - dot);
            if (j < S) atomicAdd(d_bias + ((size_t)h * S + i) * S + j, ds);
            dv[u][0] += pr[u] * go.x; dv[u][1] += pr[u] * go.y; dv[u][2] += pr[u] * go.z; dv[u][3] += pr[u] * go.w;
            const float hs = 0.5f * ds;
            dk[u][0] += hs * q.x; dk[u][1] += hs * q.y; dk[u][2] += hs * q.z; dk[u][3] += hs * q.w;
            q0 += hs * k_[u][0]; q1 += hs * k_[u][1]; q2 += hs * k_[u][2]; q3 += hs * k_[u][3];
        }
        q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2); q3 = warp_sum(q3);
        if (lane == 0) *reinterpret_cast<float4*>(dQ + i * kVD + 4 * h) = make_float4(q0, q1, q2, q3);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        if (j < S) {
            *reinterpret_cast<float4*>(dK + j * kVD + 4 * h) = make_float4(dk[u][0], dk[u][1], dk[u][2], dk[u][3]);
            *reinterpret_cast<float4*>(dV + j * kVD + 4 * h) = make_float4(dv[u][0], dv[u][1], dv[u][2], dv[u][3]);
        }
    }
}

// One CTA per dispatched row: recompute the block from its input, then back-propagate.  d_w / d_aux: global fp32,
// laid out like w_hat / aux, zeroed by the caller, accumulated with atomics (rows of one expert meet there).
__global__ void __launch_bounds__(kVThreads)
vit_block_bwd_kernel(const float* __restrict__ tok_in, const float* __restrict__ time, const int* __restrict__ row_expert,
                     const float* __restrict__ w_hat, const float* __restrict__ aux, VitTables tb, int final_ln,
                     const float* __restrict__ d_out, float* __restrict__ d_tok, float* __restrict__ d_time,
                     float* __restrict__ d_w, float* __restrict__ d_aux) {
    extern __shared__ __align__(16) unsigned char smraw[];
    VitSmemBwd& sm = *reinterpret_cast<VitSmemBwd*>(smraw);
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = row_expert[r];
    float* dtok_row = d_tok + (size_t)r * kVS * kVD;
    if (e < 0 || e >= tb.n_experts) {
        for (int i = tid; i < kVS * kVD; i += kVThreads) dtok_row[i] = 0.f;
        if (tid < kVT) d_time[(size_t)r * kVT + tid] = 0.f;
        return;
    }
    const int S = tb.S[e];
    const float* W = w_hat + tb.w_off[e];
    const float* ax = aux + tb.a_off[e];
    float* dW = d_w + tb.w_off[e];
    float* dax = d_aux + tb.a_off[e];
    {
        const float4* src = reinterpret_cast<const float4*>(tok_in + (size_t)r * kVS * kVD);
        const float4* gsrc = reinterpret_cast<const float4*>(d_out + (size_t)r * kVS * kVD);
        for (int i = tid; i < S * kVD / 4; i += kVThreads) {
            reinterpret_cast<float4*>(sm.X)[i] = src[i];
            reinterpret_cast<float4*>(sm.G3)[i] = gsrc[i];
        }
        if (tid < kVT) sm.t[tid] = time[(size_t)r * kVT + tid];
    }
    __syncthreads();
    // ---------------------------------------------------------------- forward recompute
    group_stats(sm.X, sm.stats, S);
    __syncthreads();
    for (int i = tid; i < S * kVD; i += kVThreads) {
        const int c = i & 31, g = c >> 3;
        sm.A[i] = vsilu((sm.X[i] - sm.stats[g]) * sm.stats[4 + g] * ax[A_GN + c] + ax[A_GN + 32 + c]);
    }
    lin_gen<kVD, kVD, true, false>(sm.A, W + W_L1, sm.wbuf, S, [&](int s, int n, float v) { sm.G0[s * kVD + n] = v; });   // H1
    __syncthreads();
    layer_norm(sm.G0, sm.Y, sm.XH1, sm.rstd1, ax + A_LN1, S);
    time_proj(W + W_QT, sm.t, sm.tp);
    lin_gen<kVD, kVD, true, false>(sm.Y, W + W_Q, sm.wbuf, S, [&](int s, int n, float v) { sm.Q[s * kVD + n] = v + sm.tp[n]; });
    lin_gen<kVD, kVD, true, false>(sm.Y, W + W_K, sm.wbuf, S, [&](int s, int n, float v) { sm.KT[n * kVKT + s] = v + sm.tp[32 + n]; });
    lin_gen<kVD, kVD, true, false>(sm.Y, W + W_V, sm.wbuf, S, [&](int s, int n, float v) { sm.VT[n * kVKT + s] = v + sm.tp[64 + n]; });
    __syncthreads();
    attention_fwd(sm.Q, sm.KT, sm.VT, ax + A_BIAS, sm.O, S);
    lin_gen<kVD, kVD, true, false>(sm.O, W + W_O, sm.wbuf, S, [&](int s, int n, float v) {
        const int i = s * kVD + n;
        sm.G1[i] = kVc * (kVc * (sm.Y[i] + v) + sm.G0[i]);                                          // Y2
    });
    __syncthreads();
    layer_norm(sm.G1, sm.A2, sm.XH2, sm.rstd2, ax + A_LN2, S);
    lin_gen<kVHid, kVD, true, false>(sm.A2, W + W_L2, sm.wbuf, S, [&](int s, int n, float v) { sm.P2[s * kVHid + n] = v; });
    if (final_ln) {
        // Out = mp_sum(X, mp_sum(linear3(mp_silu(P2)), Y2)); only its normalised form is needed (final LayerNorm backward)
        lin_gen<kVD, kVHid, true, true>(sm.P2, W + W_L3, sm.wbuf, S, [&](int s, int n, float v) {
            const int i = s * kVD + n;
            sm.G2[i] = kVc * (sm.X[i] + kVc * (v + sm.G1[i]));
        });
        __syncthreads();
        layer_norm(sm.G2, sm.G4, sm.G2, sm.rstdF, ax + A_LNF, S);      // G2 <- xhat (in place), G4 scratch
        __syncthreads();
        // G3 = d(Out) from d(LNF(Out))
        layer_norm_bwd(sm.G3, sm.G2, sm.rstdF, ax + A_LNF, sm.G3, 0.f, false, dax + A_LNF, S);
    }
    __syncthreads();
    // ---------------------------------------------------------------- backward
    // Out = c (X + c (U + Y2)):  dX += c dOut (G4);  dU = dY2 = c^2 dOut (G3 in place, G1 copy)
    for (int i = tid; i < S * kVD; i += kVThreads) {
        const float g = sm.G3[i];
        sm.G4[i] = kVc * g;
        sm.G3[i] = 0.5f * g;
        sm.G1[i] = 0.5f * g;
    }
    __syncthreads();
    outer_acc<kVD, kVHid, true>(sm.G3, sm.P2, dW + W_L3, S);                               // dW3 += dU^T mp_silu(P2)
    // dP2 = (dU W3) * mp_silu'(P2), in place of P2
    lin_gen<kVHid, kVD, false, false>(sm.G3, W + W_L3, sm.wbuf, S, [&](int s, int n, float v) {
        const int i = s * kVHid + n;
        sm.P2[i] = v * vdsilu(sm.P2[i]);
    });
    __syncthreads();
    outer_acc<kVHid, kVD, false>(sm.P2, sm.A2, dW + W_L2, S);                              // dW2 += dP2^T A2
    lin_gen<kVD, kVHid, false, false>(sm.P2, W + W_L2, sm.wbuf, S, [&](int s, int n, float v) { sm.G2[s * kVD + n] = v; });   // dA2
    __syncthreads();
    layer_norm_bwd(sm.G2, sm.XH2, sm.rstd2, ax + A_LN2, sm.G1, 1.f, true, dax + A_LN2, S);  // G1 = dY2 (complete)
    __syncthreads();
    // Y2 = c (c (Y + Z) + H1):  dZ = dY(part) = c^2 dY2 -> G3;  dH1(part) = c dY2 stays in G1 (scaled in the LN1 step)
    for (int i = tid; i < S * kVD; i += kVThreads) sm.G3[i] = 0.5f * sm.G1[i];
    __syncthreads();
    outer_acc<kVD, kVD, false>(sm.G3, sm.O, dW + W_O, S);                                  // dWo += dZ^T O
    lin_gen<kVD, kVD, false, false>(sm.G3, W + W_O, sm.wbuf, S, [&](int s, int n, float v) { sm.G2[s * kVD + n] = v; });      // dO
    __syncthreads();
    attention_bwd(sm.Q, sm.KT, sm.VT, ax + A_BIAS, sm.G2, sm.G0, sm.DK, sm.DV, dax + A_BIAS, S);   // dQ -> G0
    __syncthreads();
    outer_acc<kVD, kVD, false>(sm.G0, sm.Y, dW + W_Q, S);
    outer_acc<kVD, kVD, false>(sm.DK, sm.Y, dW + W_K, S);
    outer_acc<kVD, kVD, false>(sm.DV, sm.Y, dW + W_V, S);
    // time projections: d tp = column sums of dQ / dK / dV; d Wqt += dtp t^T; d t = Wqt^T dtp
    if (tid < 96) {
        const float* src = tid < 32 ? sm.G0 : (tid < 64 ? sm.DK : sm.DV);
        const int n = tid & 31;
        float a = 0.f;
        for (int s = 0; s < S; ++s) a += src[s * kVD + n];
        sm.dtp[tid] = a;
    }
    __syncthreads();
    for (int idx = tid; idx < 96 * kVT / 4; idx += kVThreads) {
        const int n = idx / (kVT / 4), j4 = (idx - n * (kVT / 4)) * 4;
        const float g = sm.dtp[n];
        vred4(dW + W_QT + n * kVT + j4, make_float4(g * sm.t[j4], g * sm.t[j4 + 1], g * sm.t[j4 + 2], g * sm.t[j4 + 3]));
    }
    if (tid < kVT) {
        float a = 0.f;
        for (int n = 0; n < 96; ++n) a += W[W_QT + n * kVT + tid] * sm.dtp[n];
        d_time[(size_t)r * kVT + tid] = a;
    }
    // dY (G3) += dQ Wq + dK Wk + dV Wv
    lin_gen<kVD, kVD, false, false>(sm.G0, W + W_Q, sm.wbuf, S, [&](int s, int n, float v) { sm.G3[s * kVD + n] += v; });
    lin_gen<kVD, kVD, false, false>(sm.DK, W + W_K, sm.wbuf, S, [&](int s, int n, float v) { sm.G3[s * kVD + n] += v; });
    lin_gen<kVD, kVD, false, false>(sm.DV, W + W_V, sm.wbuf, S, [&](int s, int n, float v) { sm.G3[s * kVD + n] += v; });
    __syncthreads();
    // dH1 = c dY2 + LN1'(dY)   (G1 in place)
    layer_norm_bwd(sm.G3, sm.XH1, sm.rstd1, ax + A_LN1, sm.G1, kVc, true, dax + A_LN1, S);
    __syncthreads();
    outer_acc<kVD, kVD, false>(sm.G1, sm.A, dW + W_L1, S);                                 // dW1 += dH1^T A
    lin_gen<kVD, kVD, false, false>(sm.G1, W + W_L1, sm.wbuf, S, [&](int s, int n, float v) { sm.G2[s * kVD + n] = v; });     // dA
    __syncthreads();
    // GroupNorm + mp_silu backward: G = xh gamma + beta, dG = dA mp_silu'(G); per channel d gamma / d beta;
    // per group m1 = mean(dxh), m2 = mean(dxh xh), dX = rstd (dxh - m1 - xh m2).  G2 <- dxh, G0 <- xh.
    {
        float dg = 0.f, db = 0.f;
        const int c = lane, g = c >> 3;
        const float gam = ax[A_GN + c], bet = ax[A_GN + 32 + c], mean = sm.stats[g], rs = sm.stats[4 + g];
        for (int s = warp; s < S; s += 8) {
            const int i = s * kVD + c;
            const float xh = (sm.X[i] - mean) * rs;
            const float dG = sm.G2[i] * vdsilu(xh * gam + bet);
            dg += dG * xh;
            db += dG;
            sm.G2[i] = dG * gam;
            sm.G0[i] = xh;
        }
        atomicAdd(dax + A_GN + c, dg);
        atomicAdd(dax + A_GN + 32 + c, db);
    }
    __syncthreads();
    if (warp < 4) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = lane; i < S * 8; i += 32) {
            const int o = (i >> 3) * kVD + warp * 8 + (i & 7);
            s1 += sm.G2[o];
            s2 += sm.G2[o] * sm.G0[o];
        }
        s1 = warp_sum(s1) / (float)(S * 8);
        s2 = warp_sum(s2) / (float)(S * 8);
        if (lane == 0) {
            sm.gsum[warp] = s1;
            sm.gsum[4 + warp] = s2;
        }
    }
    __syncthreads();
    for (int i = tid; i < kVS * kVD; i += kVThreads) {
        float v = 0.f;
        if (i < S * kVD) {
            const int g = (i & 31) >> 3;
            v = sm.G4[i] + sm.stats[4 + g] * (sm.G2[i] - sm.gsum[g] - sm.G0[i] * sm.gsum[4 + g]);
        }
        dtok_row[i] = v;
    }
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
    // per (device, function) attribute: set on every call (cheap), a process may drive several GPUs
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(vit_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VitSmemFwd)));
    vit_block_fwd_kernel<<<(unsigned)rows, kVThreads, sizeof(VitSmemFwd), (cudaStream_t)stream>>>(
        tok_in, time, row_expert, w_hat, aux, tb, final_ln, tok_out);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_vit_block_bwd(const float* tok_in, const float* time, const int32_t* row_expert, const float* w_hat,
                                   const float* aux, const int64_t* w_off, const int64_t* a_off, const int32_t* tokens,
                                   int n_experts, int64_t rows, int final_ln, const float* d_out, float* d_tok, float* d_time,
                                   float* d_w, float* d_aux, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(tok_in && time && row_expert && w_hat && aux && w_off && a_off && tokens && d_out && d_tok && d_time &&
                        d_w && d_aux && rows >= 1,
                    "vit_block_bwd: null pointer / no rows");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= 8, "vit_block_bwd: 1..8 experts (got %d)", n_experts);
    VitTables tb;
    tb.n_experts = n_experts;
    for (int e = 0; e < n_experts; ++e) {
        HDMOE_CHECK_ARG(tokens[e] >= 1 && tokens[e] <= kVS, "vit_block_bwd: 1..64 tokens per row (got %d)", tokens[e]);
        HDMOE_CHECK_ARG(w_off[e] % 4 == 0 && a_off[e] % 4 == 0, "vit_block_bwd: block offsets must be multiples of 4 floats");
        tb.w_off[e] = w_off[e];
        tb.a_off[e] = a_off[e];
        tb.S[e] = tokens[e];
    }
    // per (device, function) attribute: set on every call (cheap), a process may drive several GPUs
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(vit_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VitSmemBwd)));
    vit_block_bwd_kernel<<<(unsigned)rows, kVThreads, sizeof(VitSmemBwd), (cudaStream_t)stream>>>(
        tok_in, time, row_expert, w_hat, aux, tb, final_ln, d_out, d_tok, d_time, d_w, d_aux);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
