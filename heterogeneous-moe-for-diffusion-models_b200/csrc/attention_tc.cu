// Tensor-core head_dim = 4 trunk attention (the superseded CUDA-core kernels are archived in tools/legacy/attention.cu).
//
// With d = 4 the contraction is too thin for tcgen05 tiles, but the warp-level m16n8k8 TF32 MMA fits it exactly:
// its K = 8 holds the 4 features TWICE, which is used for split-operand ("3xTF32"-style) products instead of
// padding -- A = [x_hi | x_lo] against B = [y_hi | y_hi] gives x_hi*y_hi + x_lo*y_hi in one instruction, and a
// second one with B = [y_lo | 0] adds x_hi*y_lo: fp32-grade logits from two MMAs per 16x8 tile.  For the P*V type
// products the 8 output columns hold [V_hi | V_lo] (summed at the end).  The accumulator fragment of S (row g,
// keys 2t, 2t+1) is reused directly as the A fragment of the next product by permuting the key order of the B
// operand (keys 2t -> k = t, keys 2t+1 -> k = t + 4), so probabilities never leave registers.
// Per (query, key) pair the CUDA cores are left with max, subtract, exp2, row-sum and the tf32 rounding of p
// (5 instructions instead of ~14); `split_p` additionally splits p / dS into hi + lo (two more MMAs per tile) for
// strict fp32 parity.  All three kernels stage the "long" operand of one (batch, head) in shared memory in one
// layout X = even/odd item arrays of [hi(4) | lo(4)] floats that serves both fragment types without bank
// conflicts (the odd array is shifted by 4 banks).
// Same layouts, saved tensors (lse in log2 units of the scaled logits, D = <dO, O>) as the archived CUDA-core version.
#include "common.cuh"

namespace hdmoe {

constexpr int kTcWarps = 8;                 // 16 "short" rows per warp, 128 per CTA
constexpr int kTcThreads = kTcWarps * 32;
constexpr int kTcChunk = 1024;              // max items of the long operand staged at once
constexpr float kTcLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
// tf32 round-to-nearest (ties away) by integer add; the MMA ignores the low 13 mantissa bits
__device__ __forceinline__ uint32_t tf32_bits(float x) { return f2u(x) + 0x1000u; }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((f2u(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// X layout: item i -> (i & 1 ? Xo : Xe)[(i >> 1) * 8 + {0..3 hi, 4..7 lo}], Xo = Xe + n_items * 4 + 4
struct XView {
    const float* e;
    const float* o;
};
__device__ __forceinline__ int x_floats(int n_items) { return n_items * 8 + 4; }
__device__ __forceinline__ XView x_view(const float* base, int n_items) { return {base, base + n_items * 4 + 4}; }

// stage n_valid rows (4 floats of head h each, row stride C) * mul into the X layout; rows >= n_valid are zero
__device__ __forceinline__ void stage_x(float* base, int n_items, const float* __restrict__ src, int C, int n_valid, float mul) {
    float* xe = base;
    float* xo = base + n_items * 4 + 4;
    for (int i = threadIdx.x; i < n_items; i += kTcThreads) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_valid) v = *reinterpret_cast<const float4*>(src + (size_t)i * C);
        v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        float* dst = ((i & 1) ? xo : xe) + (i >> 1) * 8;
        *reinterpret_cast<float4*>(dst) = hi;
        *reinterpret_cast<float4*>(dst + 4) = lo;
    }
}
// B fragments for  S(16 x 8 items) = A(16 x [hi|lo]) * X^T : item = tile*8 + g, feature t
__device__ __forceinline__ void x_row_frag(const XView& x, int tile, int g, int t, uint32_t& hi, uint32_t& lo) {
    const float* p = ((g & 1) ? x.o : x.e) + (tile * 4 + (g >> 1)) * 8 + t;
    hi = f2u(p[0]);
    lo = tf32_bits(p[4]);
}
// B fragments for  acc(16 x [hi|lo]) += P(16 x 8 items) * X : b0 = X[item 2t][g], b1 = X[item 2t+1][g]
__device__ __forceinline__ void x_col_frag(const XView& x, int tile, int g, int t, uint32_t& b0, uint32_t& b1) {
    const int idx = (tile * 4 + t) * 8 + g;
    b0 = tf32_bits(x.e[idx]);
    b1 = tf32_bits(x.o[idx]);
}
// S += [a_hi | a_lo] * X^T: SPLIT -> fp32-grade accuracy (two MMAs); otherwise the long operand is rounded to
// TF32 (the short one stays split, it is free) and one MMA is enough
template <bool SPLIT>
__device__ __forceinline__ void mma_split(float (&s)[4], const uint32_t (&a)[4], uint32_t bhi, uint32_t blo) {
    mma_tf32(s, a[0], a[1], a[2], a[3], bhi, bhi);
    if (SPLIT) mma_tf32(s, a[0], a[1], a[2], a[3], blo, 0u);
}
// acc += P * X with P taken from an accumulator fragment (c0,c1: row g keys 2t,2t+1; c2,c3: row g+8)
template <bool SPLIT>
__device__ __forceinline__ void mma_from_acc(float (&acc)[4], const float (&p)[4], uint32_t b0, uint32_t b1) {
    if (SPLIT) {
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = tf32_hi(p[i]);
        mma_tf32(acc, f2u(h[0]), f2u(h[2]), f2u(h[1]), f2u(h[3]), b0, b1);
        mma_tf32(acc, tf32_bits(p[0] - h[0]), tf32_bits(p[2] - h[2]), tf32_bits(p[1] - h[1]), tf32_bits(p[3] - h[3]), b0, b1);
    } else {
        mma_tf32(acc, tf32_bits(p[0]), tf32_bits(p[2]), tf32_bits(p[1]), tf32_bits(p[3]), b0, b1);
    }
}
// A fragment [x_hi | x_lo] of two rows (g, g+8): this lane holds feature t of each
__device__ __forceinline__ void a_split(uint32_t (&a)[4], float x0, float x1) {
    const float h0 = tf32_hi(x0), h1 = tf32_hi(x1);
    a[0] = f2u(h0);
    a[1] = f2u(h1);
    a[2] = tf32_bits(x0 - h0);
    a[3] = tf32_bits(x1 - h1);
}
__device__ __forceinline__ float quad_sum(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    return x + __shfl_xor_sync(0xffffffffu, x, 2);
}
__device__ __forceinline__ float quad_max(float x) {
    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
    return fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
}
// acc columns are [hi(0..3) | lo(4..7)]: lanes t < 2 end up with the final columns 2t, 2t+1
__device__ __forceinline__ void fold_hi_lo(float (&acc)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], 2);
}

// ---------------------------------------------------------------------------------------------------------
// forward: grid (ceil(Sq/128), heads, B)
// ---------------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(kTcThreads)
attn_d4_tc_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      float* __restrict__ o, float* __restrict__ lse, int Sq, int Sk, int H, float scale, int chunk) {
    extern __shared__ __align__(16) float smf[];
    float* Kb = smf;
    float* Vb = smf + x_floats(chunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, b = blockIdx.z, C = H * 4;
    const int r0 = blockIdx.x * (kTcWarps * 16) + warp * 16 + g, r1 = r0 + 8;
    const float c2 = scale * kTcLog2e;
    const float* qb = q + (size_t)b * Sq * C + h * 4;
    uint32_t aq[4];
    a_split(aq, qb[(size_t)min(r0, Sq - 1) * C + t] * c2, qb[(size_t)min(r1, Sq - 1) * C + t] * c2);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int c0 = 0; c0 < Sk; c0 += chunk) {
        const int nk = min(chunk, Sk - c0);
        __syncthreads();
        stage_x(Kb, chunk, k + ((size_t)b * Sk + c0) * C + h * 4, C, nk, 1.f);
        stage_x(Vb, chunk, v + ((size_t)b * Sk + c0) * C + h * 4, C, nk, 1.f);
        __syncthreads();
        const XView K = x_view(Kb, chunk), V = x_view(Vb, chunk);
        for (int kb = 0; kb < nk; kb += 64) {
            float s[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bh, bl;
                x_row_frag(K, (kb >> 3) + j, g, t, bh, bl);
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
                mma_split<SPLIT>(s[j], aq, bh, bl);
            }
            if (kb + 64 > nk) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int key = kb + j * 8 + 2 * t;
                    if (key >= nk) s[j][0] = s[j][2] = -INFINITY;
                    if (key + 1 >= nk) s[j][1] = s[j][3] = -INFINITY;
                }
            }
            float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x0 = fmaxf(x0, fmaxf(s[j][0], s[j][1]));
                x1 = fmaxf(x1, fmaxf(s[j][2], s[j][3]));
            }
            const float n0 = fmaxf(m0, quad_max(x0)), n1 = fmaxf(m1, quad_max(x1));
            const float al0 = ex2(m0 - n0), al1 = ex2(m1 - n1);
            m0 = n0; m1 = n1;
            acc[0] *= al0; acc[1] *= al0; acc[2] *= al1; acc[3] *= al1;
            l0 *= al0; l1 *= al1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float p[4];
                p[0] = ex2(s[j][0] - n0); p[1] = ex2(s[j][1] - n0);
                p[2] = ex2(s[j][2] - n1); p[3] = ex2(s[j][3] - n1);
                l0 += p[0] + p[1];
                l1 += p[2] + p[3];
                uint32_t b0, b1;
                x_col_frag(V, (kb >> 3) + j, g, t, b0, b1);
                mma_from_acc<SPLIT>(acc, p, b0, b1);
            }
        }
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    fold_hi_lo(acc);
    if (t < 2) {
        float* ob = o + (size_t)b * Sq * C + h * 4 + 2 * t;
        if (r0 < Sq) *reinterpret_cast<float2*>(ob + (size_t)r0 * C) = make_float2(acc[0] / l0, acc[1] / l0);
        if (r1 < Sq) *reinterpret_cast<float2*>(ob + (size_t)r1 * C) = make_float2(acc[2] / l1, acc[3] / l1);
    }
    if (t == 0) {
        float* lb = lse + ((size_t)b * H + h) * Sq;
        if (r0 < Sq) lb[r0] = m0 + log2f(l0);
        if (r1 < Sq) lb[r1] = m1 + log2f(l1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// dQ (and D = <dO, O>): grid (ceil(Sq/128), heads, B); K and V of the (batch, head) staged in the X layout
// ---------------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(kTcThreads)
attn_d4_tc_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     const float* __restrict__ o, const float* __restrict__ dO, const float* __restrict__ lse,
                     float* __restrict__ dq, float* __restrict__ Dbuf, int Sq, int Sk, int H, float scale, int chunk) {
    extern __shared__ __align__(16) float smf[];
    float* Kb = smf;
    float* Vb = smf + x_floats(chunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, b = blockIdx.z, C = H * 4;
    const int r0 = blockIdx.x * (kTcWarps * 16) + warp * 16 + g, r1 = r0 + 8;
    const int rr0 = min(r0, Sq - 1), rr1 = min(r1, Sq - 1);
    const float c2 = scale * kTcLog2e;
    const size_t base = (size_t)b * Sq * C + h * 4 + t;
    uint32_t aq[4], ag[4];
    a_split(aq, q[base + (size_t)rr0 * C] * c2, q[base + (size_t)rr1 * C] * c2);
    const float g0 = dO[base + (size_t)rr0 * C], g1 = dO[base + (size_t)rr1 * C];
    a_split(ag, g0, g1);
    const float D0 = quad_sum(g0 * o[base + (size_t)rr0 * C]), D1 = quad_sum(g1 * o[base + (size_t)rr1 * C]);
    const float* lb = lse + ((size_t)b * H + h) * Sq;
    const float ls0 = lb[rr0], ls1 = lb[rr1];
    if (t == 0) {
        float* db = Dbuf + ((size_t)b * H + h) * Sq;
        if (r0 < Sq) db[r0] = D0;
        if (r1 < Sq) db[r1] = D1;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < Sk; c0 += chunk) {
        const int nk = min(chunk, Sk - c0);
        __syncthreads();
        stage_x(Kb, chunk, k + ((size_t)b * Sk + c0) * C + h * 4, C, nk, 1.f);
        stage_x(Vb, chunk, v + ((size_t)b * Sk + c0) * C + h * 4, C, nk, 1.f);
        __syncthreads();
        const XView K = x_view(Kb, chunk), V = x_view(Vb, chunk);
        const int ntile = (nk + 7) >> 3;
#pragma unroll 4
        for (int j = 0; j < ntile; ++j) {
            float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t bh, bl;
            x_row_frag(K, j, g, t, bh, bl);
            mma_split<SPLIT>(s, aq, bh, bl);
            x_row_frag(V, j, g, t, bh, bl);
            mma_split<SPLIT>(dp, ag, bh, bl);
            float ds[4];
            ds[0] = ex2(s[0] - ls0) * (dp[0] - D0);
            ds[1] = ex2(s[1] - ls0) * (dp[1] - D0);
            ds[2] = ex2(s[2] - ls1) * (dp[2] - D1);
            ds[3] = ex2(s[3] - ls1) * (dp[3] - D1);
            // keys >= nk are staged as zeros: K rows are 0 there, so their dS contributes nothing to dQ
            uint32_t b0, b1;
            x_col_frag(K, j, g, t, b0, b1);
            mma_from_acc<SPLIT>(acc, ds, b0, b1);
        }
    }
    fold_hi_lo(acc);
    if (t < 2) {
        float* ob = dq + (size_t)b * Sq * C + h * 4 + 2 * t;
        if (r0 < Sq) *reinterpret_cast<float2*>(ob + (size_t)r0 * C) = make_float2(acc[0] * scale, acc[1] * scale);
        if (r1 < Sq) *reinterpret_cast<float2*>(ob + (size_t)r1 * C) = make_float2(acc[2] * scale, acc[3] * scale);
    }
}

// ---------------------------------------------------------------------------------------------------------
// dK, dV: grid (ceil(Sk/128), heads, B); warp owns 16 keys, queries (scaled q, dO, lse, D) staged in chunks
// ---------------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(kTcThreads)
attn_d4_tc_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ dO, const float* __restrict__ lse, const float* __restrict__ Dbuf,
                      float* __restrict__ dk, float* __restrict__ dv, int Sq, int Sk, int H, float scale, int chunk) {
    extern __shared__ __align__(16) float smf[];
    float* Qb = smf;
    float* Gb = smf + x_floats(chunk);
    float4* Ls = reinterpret_cast<float4*>(smf + 2 * x_floats(chunk));     // [chunk/2] (lse0, D0, lse1, D1)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, b = blockIdx.z, C = H * 4;
    const int r0 = blockIdx.x * (kTcWarps * 16) + warp * 16 + g, r1 = r0 + 8;
    const int rr0 = min(r0, Sk - 1), rr1 = min(r1, Sk - 1);
    const float c2 = scale * kTcLog2e;
    const size_t base = (size_t)b * Sk * C + h * 4 + t;
    uint32_t ak[4], av[4];
    a_split(ak, k[base + (size_t)rr0 * C], k[base + (size_t)rr1 * C]);
    a_split(av, v[base + (size_t)rr0 * C], v[base + (size_t)rr1 * C]);
    float acck[4] = {0.f, 0.f, 0.f, 0.f}, accv[4] = {0.f, 0.f, 0.f, 0.f};
    const float* lb = lse + ((size_t)b * H + h) * Sq;
    const float* db = Dbuf + ((size_t)b * H + h) * Sq;
    for (int c0 = 0; c0 < Sq; c0 += chunk) {
        const int nq = min(chunk, Sq - c0);
        __syncthreads();
        stage_x(Qb, chunk, q + ((size_t)b * Sq + c0) * C + h * 4, C, nq, c2);
        stage_x(Gb, chunk, dO + ((size_t)b * Sq + c0) * C + h * 4, C, nq, 1.f);
        for (int i = threadIdx.x; i < chunk / 2; i += kTcThreads) {
            const int i0 = 2 * i, i1 = 2 * i + 1;
            // queries beyond Sq: lse = +inf makes p = exp2(-inf) = 0
            Ls[i] = make_float4(i0 < nq ? lb[c0 + i0] : INFINITY, i0 < nq ? db[c0 + i0] : 0.f,
                                i1 < nq ? lb[c0 + i1] : INFINITY, i1 < nq ? db[c0 + i1] : 0.f);
        }
        __syncthreads();
        const XView Q = x_view(Qb, chunk), G = x_view(Gb, chunk);
        const int ntile = (nq + 7) >> 3;
#pragma unroll 4
        for (int j = 0; j < ntile; ++j) {
            float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t bh, bl;
            x_row_frag(Q, j, g, t, bh, bl);
            mma_split<SPLIT>(s, ak, bh, bl);                      // S^T tile: rows = keys, columns = queries 2t, 2t+1
            x_row_frag(G, j, g, t, bh, bl);
            mma_split<SPLIT>(dp, av, bh, bl);                     // dP^T = V dO^T
            const float4 ld = Ls[j * 4 + t];
            float p[4], ds[4];
            p[0] = ex2(s[0] - ld.x); p[1] = ex2(s[1] - ld.z);
            p[2] = ex2(s[2] - ld.x); p[3] = ex2(s[3] - ld.z);
            ds[0] = p[0] * (dp[0] - ld.y); ds[1] = p[1] * (dp[1] - ld.w);
            ds[2] = p[2] * (dp[2] - ld.y); ds[3] = p[3] * (dp[3] - ld.w);
            uint32_t b0, b1;
            x_col_frag(G, j, g, t, b0, b1);
            mma_from_acc<SPLIT>(accv, p, b0, b1);          // dV += P^T dO
            x_col_frag(Q, j, g, t, b0, b1);
            mma_from_acc<SPLIT>(acck, ds, b0, b1);         // dK += dS^T Q'
        }
    }
    fold_hi_lo(acck);
    fold_hi_lo(accv);
    if (t < 2) {
        const float inv = 1.f / kTcLog2e;                  // Q' carried scale*log2e; dK needs scale
        const size_t ob = (size_t)b * Sk * C + h * 4 + 2 * t;
        if (r0 < Sk) {
            *reinterpret_cast<float2*>(dk + ob + (size_t)r0 * C) = make_float2(acck[0] * inv, acck[1] * inv);
            *reinterpret_cast<float2*>(dv + ob + (size_t)r0 * C) = make_float2(accv[0], accv[1]);
        }
        if (r1 < Sk) {
            *reinterpret_cast<float2*>(dk + ob + (size_t)r1 * C) = make_float2(acck[2] * inv, acck[3] * inv);
            *reinterpret_cast<float2*>(dv + ob + (size_t)r1 * C) = make_float2(accv[2], accv[3]);
        }
    }
}

static int tc_chunk(int n) { return std::min(kTcChunk, ((n + 63) / 64) * 64); }

template <typename Kern>
static int tc_smem(Kern kern, size_t bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_attn_d4_tc_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int B, int Sq,
                                    int Sk, int heads, float scale, int split_p, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(q && k && v && o && lse && B >= 1 && Sq >= 1 && Sk >= 1, "attn_d4_tc_fwd: bad args");
    HDMOE_CHECK_ARG(heads >= 1 && heads <= 65535 && B <= 65535, "attn_d4_tc_fwd: heads / batch out of range");
    const int chunk = tc_chunk(Sk);
    const size_t smem = 2 * (size_t)(chunk * 8 + 4) * sizeof(float);
    dim3 grid((Sq + kTcWarps * 16 - 1) / (kTcWarps * 16), heads, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (split_p) {
        HDMOE_CHECK_ARG(tc_smem(attn_d4_tc_fwd_kernel<true>, smem), "attn_d4_tc_fwd: shared memory attribute");
        attn_d4_tc_fwd_kernel<true><<<grid, kTcThreads, smem, st>>>(q, k, v, o, lse, Sq, Sk, heads, scale, chunk);
    } else {
        HDMOE_CHECK_ARG(tc_smem(attn_d4_tc_fwd_kernel<false>, smem), "attn_d4_tc_fwd: shared memory attribute");
        attn_d4_tc_fwd_kernel<false><<<grid, kTcThreads, smem, st>>>(q, k, v, o, lse, Sq, Sk, heads, scale, chunk);
    }
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_attn_d4_tc_bwd(const float* q, const float* k, const float* v, const float* o, const float* dO,
                                    const float* lse, float* dq, float* dk, float* dv, float* Dbuf, int B, int Sq, int Sk,
                                    int heads, float scale, int split_p, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(q && k && v && o && dO && lse && dq && dk && dv && Dbuf, "attn_d4_tc_bwd: null pointer");
    HDMOE_CHECK_ARG(heads >= 1 && heads <= 65535 && B >= 1 && B <= 65535 && Sq >= 1 && Sk >= 1, "attn_d4_tc_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int chunk = tc_chunk(Sk);
        const size_t smem = 2 * (size_t)(chunk * 8 + 4) * sizeof(float);
        dim3 grid((Sq + kTcWarps * 16 - 1) / (kTcWarps * 16), heads, B);
        if (split_p) {
            HDMOE_CHECK_ARG(tc_smem(attn_d4_tc_dq_kernel<true>, smem), "attn_d4_tc_bwd: shared memory attribute");
            attn_d4_tc_dq_kernel<true><<<grid, kTcThreads, smem, st>>>(q, k, v, o, dO, lse, dq, Dbuf, Sq, Sk, heads, scale, chunk);
        } else {
            HDMOE_CHECK_ARG(tc_smem(attn_d4_tc_dq_kernel<false>, smem), "attn_d4_tc_bwd: shared memory attribute");
            attn_d4_tc_dq_kernel<false><<<grid, kTcThreads, smem, st>>>(q, k, v, o, dO, lse, dq, Dbuf, Sq, Sk, heads, scale, chunk);
        }
        HDMOE_CHECK_LAUNCH();
    }
    {
        const int chunk = tc_chunk(Sq);
        const size_t smem = (2 * (size_t)(chunk * 8 + 4) + 4 + (size_t)chunk * 2) * sizeof(float);
        dim3 grid((Sk + kTcWarps * 16 - 1) / (kTcWarps * 16), heads, B);
        if (split_p) {
            HDMOE_CHECK_ARG(tc_smem(attn_d4_tc_dkv_kernel<true>, smem), "attn_d4_tc_bwd: shared memory attribute");
            attn_d4_tc_dkv_kernel<true><<<grid, kTcThreads, smem, st>>>(q, k, v, dO, lse, Dbuf, dk, dv, Sq, Sk, heads, scale, chunk);
        } else {
            HDMOE_CHECK_ARG(tc_smem(attn_d4_tc_dkv_kernel<false>, smem), "attn_d4_tc_bwd: shared memory attribute");
            attn_d4_tc_dkv_kernel<false><<<grid, kTcThreads, smem, st>>>(q, k, v, dO, lse, Dbuf, dk, dv, Sq, Sk, heads, scale, chunk);
        }
        HDMOE_CHECK_LAUNCH();
    }
    return HDMOE_OK;
}
