// Trunk attention with head_dim = 4: flash-style fused softmax(Q K^T / sqrt(d)) V for
// MP_Attention.forward (models/model_internals.py:380-404) when there is no rel_pos_bias (the cross-attention of
// the fusion trunk, models/model_config2.py:279-289: S_q = H*W = 1024 / 4096, S_k = 1024 / 4096 / 77, 8 heads).
//
// The reference materialises (B, heads, S_q, S_k) fp32 scores (32 MiB per sample at 32^2, 512 MiB at 64^2) and the
// library memory-efficient kernels spend 43 ms per train step at B = 256.  With d = 4 the contraction is far too
// thin for tensor cores (K = 4); the kernel is bound by exp throughput and CUDA-core FMAs, so it is written for
// those: one thread owns a query (q, o, running max / sum in registers), a warp = 32 queries of ONE head, K / V
// tiles sit in shared memory and every lane of a warp reads the same key (a broadcast, no bank conflicts), the
// softmax is online with one rescale per 8 keys, exp2 with the scale folded into the logits.
// Layout: q [B, S_q, heads*4], k / v [B, S_k, heads*4], o [B, S_q, heads*4] fp32 -- exactly what the 1x1
// projections produce, so no head transposes are needed.  Backward: dQ kernel (thread per query) and dK/dV
// kernel (thread per key) recompute p from the saved log-sum-exp: no atomics, deterministic.
#include "common.cuh"

namespace hdmoe {

constexpr int kAtQPT = 2;        // queries per thread
constexpr int kAtTile = 64;      // keys (fwd, dQ) / queries (dKV) per shared-memory tile
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// grid (ceil(S_q / (32*QPT)), B), block 32 * heads
__global__ void __launch_bounds__(256)
attn_d4_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                   float* __restrict__ o, float* __restrict__ lse, int Sq, int Sk, int H, float scale) {
    extern __shared__ float4 sm4[];
    float4* Ks = sm4;                       // [kAtTile][H]
    float4* Vs = sm4 + kAtTile * H;
    const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int C = H * 4;
    const float c2 = scale * kLog2e;
    float4 qv[kAtQPT], acc[kAtQPT];
    float m[kAtQPT], l[kAtQPT];
    int qi[kAtQPT];
#pragma unroll
    for (int u = 0; u < kAtQPT; ++u) {
        qi[u] = blockIdx.x * 32 * kAtQPT + u * 32 + lane;
        const int qq = min(qi[u], Sq - 1);
        qv[u] = *reinterpret_cast<const float4*>(q + ((size_t)b * Sq + qq) * C + head * 4);
        qv[u].x *= c2; qv[u].y *= c2; qv[u].z *= c2; qv[u].w *= c2;
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        m[u] = -INFINITY;
        l[u] = 0.f;
    }
    for (int j0 = 0; j0 < Sk; j0 += kAtTile) {
        const int nk = min(kAtTile, Sk - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < kAtTile * H; i += blockDim.x) {
            const int j = i / H, h = i - j * H;
            if (j < nk) {
                Ks[i] = *reinterpret_cast<const float4*>(k + ((size_t)b * Sk + j0 + j) * C + h * 4);
                Vs[i] = *reinterpret_cast<const float4*>(v + ((size_t)b * Sk + j0 + j) * C + h * 4);
            }
        }
        __syncthreads();
        for (int jj = 0; jj < nk; jj += 8) {
            const int n8 = min(8, nk - jj);
#pragma unroll
            for (int u = 0; u < kAtQPT; ++u) {
                float s[8];
                float cm = -INFINITY;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    s[t] = t < n8 ? dot4(qv[u], Ks[(jj + t) * H + head]) : -INFINITY;
                    cm = fmaxf(cm, s[t]);
                }
                const float mn = fmaxf(m[u], cm);
                const float alpha = exp2f(m[u] - mn);
                float ps = 0.f;
                float4 a = make_float4(acc[u].x * alpha, acc[u].y * alpha, acc[u].z * alpha, acc[u].w * alpha);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float p = exp2f(s[t] - mn);
                    if (t < n8) {
                        const float4 vv = Vs[(jj + t) * H + head];
                        ps += p;
                        a.x += p * vv.x; a.y += p * vv.y; a.z += p * vv.z; a.w += p * vv.w;
                    }
                }
                acc[u] = a;
                l[u] = l[u] * alpha + ps;
                m[u] = mn;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kAtQPT; ++u) {
        if (qi[u] < Sq) {
            const float inv = 1.f / l[u];
            *reinterpret_cast<float4*>(o + ((size_t)b * Sq + qi[u]) * C + head * 4) =
                make_float4(acc[u].x * inv, acc[u].y * inv, acc[u].z * inv, acc[u].w * inv);
            lse[((size_t)b * H + head) * Sq + qi[u]] = m[u] + log2f(l[u]);      // log2 units of the scaled logits
        }
    }
}

// dQ (and D = <dO, O>): thread per query
__global__ void __launch_bounds__(256)
attn_d4_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ o, const float* __restrict__ dO, const float* __restrict__ lse,
                      float* __restrict__ dq, float* __restrict__ Dbuf, int Sq, int Sk, int H, float scale) {
    extern __shared__ float4 sm4[];
    float4* Ks = sm4;
    float4* Vs = sm4 + kAtTile * H;
    const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int C = H * 4;
    const float c2 = scale * kLog2e;
    float4 qv[kAtQPT], g[kAtQPT], acc[kAtQPT];
    float ls[kAtQPT], D[kAtQPT];
    int qi[kAtQPT];
#pragma unroll
    for (int u = 0; u < kAtQPT; ++u) {
        qi[u] = blockIdx.x * 32 * kAtQPT + u * 32 + lane;
        const int qq = min(qi[u], Sq - 1);
        const size_t off = ((size_t)b * Sq + qq) * C + head * 4;
        qv[u] = *reinterpret_cast<const float4*>(q + off);
        qv[u].x *= c2; qv[u].y *= c2; qv[u].z *= c2; qv[u].w *= c2;
        g[u] = *reinterpret_cast<const float4*>(dO + off);
        const float4 ov = *reinterpret_cast<const float4*>(o + off);
        D[u] = dot4(g[u], ov);
        ls[u] = lse[((size_t)b * H + head) * Sq + qq];
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qi[u] < Sq) Dbuf[((size_t)b * H + head) * Sq + qi[u]] = D[u];
    }
    for (int j0 = 0; j0 < Sk; j0 += kAtTile) {
        const int nk = min(kAtTile, Sk - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < kAtTile * H; i += blockDim.x) {
            const int j = i / H, h = i - j * H;
            if (j < nk) {
                Ks[i] = *reinterpret_cast<const float4*>(k + ((size_t)b * Sk + j0 + j) * C + h * 4);
                Vs[i] = *reinterpret_cast<const float4*>(v + ((size_t)b * Sk + j0 + j) * C + h * 4);
            }
        }
        __syncthreads();
        for (int j = 0; j < nk; ++j) {
            const float4 kk = Ks[j * H + head], vv = Vs[j * H + head];
#pragma unroll
            for (int u = 0; u < kAtQPT; ++u) {
                const float p = exp2f(dot4(qv[u], kk) - ls[u]);
                const float ds = p * (dot4(g[u], vv) - D[u]);
                acc[u].x += ds * kk.x; acc[u].y += ds * kk.y; acc[u].z += ds * kk.z; acc[u].w += ds * kk.w;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kAtQPT; ++u)
        if (qi[u] < Sq)
            *reinterpret_cast<float4*>(dq + ((size_t)b * Sq + qi[u]) * C + head * 4) =
                make_float4(acc[u].x * scale, acc[u].y * scale, acc[u].z * scale, acc[u].w * scale);
}

// dK, dV: thread per key; queries stream through shared memory
__global__ void __launch_bounds__(256)
attn_d4_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       const float* __restrict__ dO, const float* __restrict__ lse, const float* __restrict__ Dbuf,
                       float* __restrict__ dk, float* __restrict__ dv, int Sq, int Sk, int H, float scale) {
    extern __shared__ float4 sm4[];
    float4* Qs = sm4;                                   // [kAtTile][H]  (pre-scaled by scale*log2e)
    float4* Gs = sm4 + kAtTile * H;                     // dO
    float2* Ls = reinterpret_cast<float2*>(sm4 + 2 * kAtTile * H);   // (lse, D) [kAtTile][H]
    const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int C = H * 4;
    const float c2 = scale * kLog2e;
    float4 kv[kAtQPT], vv[kAtQPT], ak[kAtQPT], av[kAtQPT];
    int kj[kAtQPT];
#pragma unroll
    for (int u = 0; u < kAtQPT; ++u) {
        kj[u] = blockIdx.x * 32 * kAtQPT + u * 32 + lane;
        const int jj = min(kj[u], Sk - 1);
        const size_t off = ((size_t)b * Sk + jj) * C + head * 4;
        kv[u] = *reinterpret_cast<const float4*>(k + off);
        vv[u] = *reinterpret_cast<const float4*>(v + off);
        ak[u] = av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i0 = 0; i0 < Sq; i0 += kAtTile) {
        const int nq = min(kAtTile, Sq - i0);
        __syncthreads();
        for (int i = threadIdx.x; i < kAtTile * H; i += blockDim.x) {
            const int r = i / H, h = i - r * H;
            if (r < nq) {
                float4 qq = *reinterpret_cast<const float4*>(q + ((size_t)b * Sq + i0 + r) * C + h * 4);
                qq.x *= c2; qq.y *= c2; qq.z *= c2; qq.w *= c2;
                Qs[i] = qq;
                Gs[i] = *reinterpret_cast<const float4*>(dO + ((size_t)b * Sq + i0 + r) * C + h * 4);
                Ls[i] = make_float2(lse[((size_t)b * H + h) * Sq + i0 + r], Dbuf[((size_t)b * H + h) * Sq + i0 + r]);
            }
        }
        __syncthreads();
        for (int r = 0; r < nq; ++r) {
            const float4 qq = Qs[r * H + head], gg = Gs[r * H + head];
            const float2 ld = Ls[r * H + head];
#pragma unroll
            for (int u = 0; u < kAtQPT; ++u) {
                const float p = exp2f(dot4(qq, kv[u]) - ld.x);
                av[u].x += p * gg.x; av[u].y += p * gg.y; av[u].z += p * gg.z; av[u].w += p * gg.w;
                const float ds = p * (dot4(gg, vv[u]) - ld.y);
                ak[u].x += ds * qq.x; ak[u].y += ds * qq.y; ak[u].z += ds * qq.z; ak[u].w += ds * qq.w;
            }
        }
    }
    const float inv = 1.f / kLog2e;     // Qs carried scale*log2e; dK needs scale
#pragma unroll
    for (int u = 0; u < kAtQPT; ++u)
        if (kj[u] < Sk) {
            const size_t off = ((size_t)b * Sk + kj[u]) * C + head * 4;
            *reinterpret_cast<float4*>(dk + off) = make_float4(ak[u].x * inv, ak[u].y * inv, ak[u].z * inv, ak[u].w * inv);
            *reinterpret_cast<float4*>(dv + off) = av[u];
        }
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_attn_d4_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int B, int Sq,
                                 int Sk, int heads, float scale, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(q && k && v && o && lse && B >= 1 && Sq >= 1 && Sk >= 1, "attn_d4_fwd: bad args");
    HDMOE_CHECK_ARG(heads >= 1 && heads <= 8, "attn_d4_fwd: 1..8 heads of dimension 4 (got %d)", heads);
    HDMOE_CHECK_ARG(B <= 65535, "attn_d4_fwd: batch > 65535");
    dim3 grid((Sq + 32 * kAtQPT - 1) / (32 * kAtQPT), B);
    const size_t smem = 2 * kAtTile * heads * sizeof(float4);
    attn_d4_fwd_kernel<<<grid, 32 * heads, smem, (cudaStream_t)stream>>>(q, k, v, o, lse, Sq, Sk, heads, scale);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_attn_d4_bwd(const float* q, const float* k, const float* v, const float* o, const float* dO,
                                 const float* lse, float* dq, float* dk, float* dv, float* Dbuf, int B, int Sq, int Sk,
                                 int heads, float scale, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(q && k && v && o && dO && lse && dq && dk && dv && Dbuf, "attn_d4_bwd: null pointer");
    HDMOE_CHECK_ARG(heads >= 1 && heads <= 8 && B >= 1 && B <= 65535, "attn_d4_bwd: bad heads / batch");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 gq((Sq + 32 * kAtQPT - 1) / (32 * kAtQPT), B), gk((Sk + 32 * kAtQPT - 1) / (32 * kAtQPT), B);
    attn_d4_bwd_dq_kernel<<<gq, 32 * heads, 2 * kAtTile * heads * sizeof(float4), st>>>(q, k, v, o, dO, lse, dq, Dbuf, Sq,
                                                                                        Sk, heads, scale);
    HDMOE_CHECK_LAUNCH();
    attn_d4_bwd_dkv_kernel<<<gk, 32 * heads, kAtTile * heads * (2 * sizeof(float4) + sizeof(float2)), st>>>(
        q, k, v, dO, lse, Dbuf, dk, dv, Sq, Sk, heads, scale);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
