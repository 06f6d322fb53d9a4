// Router trunk normalisation (Router.hard_route, models/model_components.py:92-103): GroupNorm(1, C) + ReLU, and
// for the last layer + AdaptiveAvgPool2d((1,1)), on channels-last activations [B, HW, C], fp32 (library-convolution trunk)
// or bf16 (tcgen05 trunk: the activations between the grouped convolutions; statistics and the pooled output stay fp32).
//
// The library path costs ~5 tensor passes forward and ~11 backward per layer (moments, unvectorised broadcast
// affine, ReLU, their backward reductions) plus NCHW<->NHWC conversions around every convolution.  GroupNorm with
// ONE group reduces over the whole sample, so one CTA owns one sample: pass 1 reads it for the statistics, pass 2
// re-reads it (512 KB, L2-resident) and writes the result -- HBM sees one read and one write.  The pooled variant
// never writes the activation at all.  Backward has the same two-pass shape.  HBM-bound, float4 everywhere,
// deterministic (no atomics; per-sample dgamma / dbeta partials are summed by the caller).
#include "common.cuh"

namespace hdmoe {

constexpr int kGnThreads = 1024;

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = l < (kGnThreads >> 5) ? red[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;       // every thread holds the total
}

// y may be NULL (pooled-only); pooled may be NULL.  stats[b] = (mean, rstd).
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
gn1_relu_fwd_kernel(const T* __restrict__ x, const float4* __restrict__ gamma, const float4* __restrict__ beta,
                    T* __restrict__ y, float* __restrict__ pooled, float2* __restrict__ stats, int HW, int C, float eps,
                    int rpg) {
    __shared__ double red[32];
    __shared__ float4 pr[kGnThreads];
    const int b = blockIdx.x, cq = C >> 2;
    const long long nvec = (long long)HW * cq;
    const T* xs = x + (size_t)b * nvec * 4;
    float s = 0.f, ss = 0.f;
#pragma unroll 4
    for (long long i = threadIdx.x; i < nvec; i += kGnThreads) {
        const float4 v = Vec4<T>::load(xs + 4 * i);
        s += (v.x + v.y) + (v.z + v.w);
        ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    const double N = (double)nvec * 4.0;
    const double mean = block_sum_d((double)s, red) / N;
    const double var = fmax(block_sum_d((double)ss, red) / N - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps)), mu = (float)mean;
    if (threadIdx.x == 0) stats[b] = make_float2(mu, rstd);
    const int c4 = threadIdx.x % cq;                       // kGnThreads % cq == 0 (host check): fixed channel quad
    const int grp = b / rpg;                               // rows [grp * rpg, +rpg) share one (gamma, beta): one router each
    const float4 g = gamma[grp * cq + c4], be = beta[grp * cq + c4];
    const float4 a = make_float4(rstd * g.x, rstd * g.y, rstd * g.z, rstd * g.w);
    const float4 sh = make_float4(be.x - mu * a.x, be.y - mu * a.y, be.z - mu * a.z, be.w - mu * a.w);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    T* ys = y ? y + (size_t)b * nvec * 4 : nullptr;
#pragma unroll 4
    for (long long i = threadIdx.x; i < nvec; i += kGnThreads) {
        const float4 v = Vec4<T>::load(xs + 4 * i);
        float4 r;
        r.x = fmaxf(fmaf(v.x, a.x, sh.x), 0.f);
        r.y = fmaxf(fmaf(v.y, a.y, sh.y), 0.f);
        r.z = fmaxf(fmaf(v.z, a.z, sh.z), 0.f);
        r.w = fmaxf(fmaf(v.w, a.w, sh.w), 0.f);
        if (ys) Vec4<T>::store(ys + 4 * i, r);
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
    }
    if (pooled) {
        pr[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < cq) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = threadIdx.x; j < kGnThreads; j += cq) {
                const float4 u = pr[j];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            const float inv = 1.f / (float)HW;
            reinterpret_cast<float4*>(pooled + (size_t)b * C)[threadIdx.x] = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
        }
    }
}

// dy (activation gradient) or dpooled (gradient of the pooled output, broadcast / HW) -- exactly one is non-NULL
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
gn1_relu_bwd_kernel(const T* __restrict__ x, const float4* __restrict__ gamma, const float4* __restrict__ beta,
                    const float2* __restrict__ stats, const T* __restrict__ dy, const float* __restrict__ dpooled,
                    T* __restrict__ dx, float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, int HW, int C,
                    int rpg) {
    __shared__ double red[32];
    __shared__ float4 pr[kGnThreads];
    const int b = blockIdx.x, cq = C >> 2;
    const long long nvec = (long long)HW * cq;
    const T* xs = x + (size_t)b * nvec * 4;
    const T* gs = dy ? dy + (size_t)b * nvec * 4 : nullptr;
    const float2 st = stats[b];
    const float mu = st.x, rstd = st.y;
    const int c4 = threadIdx.x % cq;
    const int grp = b / rpg;
    const float4 g = gamma[grp * cq + c4], be = beta[grp * cq + c4];
    float4 gp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dpooled) {
        gp = reinterpret_cast<const float4*>(dpooled + (size_t)b * C)[c4];
        const float inv = 1.f / (float)HW;
        gp.x *= inv; gp.y *= inv; gp.z *= inv; gp.w *= inv;
    }
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s1 = 0.f, s2 = 0.f;
#define GN_ELEM(X, G, GA, BE, GO, DGA, DBE)                      \
    {                                                            \
        const float xh = ((X) - mu) * rstd;                      \
        const float gg = fmaf(xh, GA, BE) > 0.f ? (G) : 0.f;     \
        DGA += gg * xh;                                          \
        DBE += gg;                                               \
        GO = gg * (GA);                                          \
        s1 += GO;                                                \
        s2 += GO * xh;                                           \
    }
#pragma unroll 4
    for (long long i = threadIdx.x; i < nvec; i += kGnThreads) {
        const float4 v = Vec4<T>::load(xs + 4 * i);
        const float4 gi = gs ? Vec4<T>::load(gs + 4 * i) : gp;
        float t;
        GN_ELEM(v.x, gi.x, g.x, be.x, t, dg.x, db.x)
        GN_ELEM(v.y, gi.y, g.y, be.y, t, dg.y, db.y)
        GN_ELEM(v.z, gi.z, g.z, be.z, t, dg.z, db.z)
        GN_ELEM(v.w, gi.w, g.w, be.w, t, dg.w, db.w)
    }
#undef GN_ELEM
    const double N = (double)nvec * 4.0;
    const float m1 = (float)(block_sum_d((double)s1, red) / N);
    const float m2 = (float)(block_sum_d((double)s2, red) / N);
    // per-channel partials of this sample
    pr[threadIdx.x] = dg;
    __syncthreads();
    if (threadIdx.x < cq) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = threadIdx.x; j < kGnThreads; j += cq) {
            const float4 u = pr[j];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        reinterpret_cast<float4*>(dgamma_part + (size_t)b * C)[threadIdx.x] = t;
    }
    __syncthreads();
    pr[threadIdx.x] = db;
    __syncthreads();
    if (threadIdx.x < cq) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = threadIdx.x; j < kGnThreads; j += cq) {
            const float4 u = pr[j];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        reinterpret_cast<float4*>(dbeta_part + (size_t)b * C)[threadIdx.x] = t;
    }
    T* ds = dx + (size_t)b * nvec * 4;
#pragma unroll 4
    for (long long i = threadIdx.x; i < nvec; i += kGnThreads) {
        const float4 v = Vec4<T>::load(xs + 4 * i);
        const float4 gi = gs ? Vec4<T>::load(gs + 4 * i) : gp;
        float4 r;
#define GN_DX(X, G, GA, BE, R)                                   \
    {                                                            \
        const float xh = ((X) - mu) * rstd;                      \
        const float gg = fmaf(xh, GA, BE) > 0.f ? (G) * (GA) : 0.f; \
        R = rstd * (gg - m1 - xh * m2);                          \
    }
        GN_DX(v.x, gi.x, g.x, be.x, r.x)
        GN_DX(v.y, gi.y, g.y, be.y, r.y)
        GN_DX(v.z, gi.z, g.z, be.z, r.z)
        GN_DX(v.w, gi.w, g.w, be.w, r.w)
#undef GN_DX
        Vec4<T>::store(ds + 4 * i, r);
    }
}

}  // namespace hdmoe
using namespace hdmoe;

static bool gn_shape_ok(int C) { return C >= 4 && C % 4 == 0 && kGnThreads % (C / 4) == 0; }

extern "C" int hdmoe_gn1_relu_fwd_t(const void* x, int dtype, const float* gamma, const float* beta, void* y, float* pooled,
                                    float* stats, int B, int HW, int C, float eps, int rows_per_group,
                                    hdmoe_stream_t stream) {
    const int rpg = rows_per_group > 0 ? rows_per_group : B;
    HDMOE_CHECK_ARG(x && gamma && beta && stats && (y || pooled) && B >= 1 && HW >= 1, "gn1_relu_fwd: bad args");
    HDMOE_CHECK_ARG(gn_shape_ok(C), "gn1_relu_fwd: C / 4 must divide 1024 (got C = %d)", C);
    HDMOE_CHECK_ARG(dtype == HDMOE_F32 || dtype == HDMOE_BF16, "gn1_relu_fwd: dtype must be f32 or bf16");
    if (dtype == HDMOE_F32)
        gn1_relu_fwd_kernel<float><<<B, kGnThreads, 0, (cudaStream_t)stream>>>(
            (const float*)x, (const float4*)gamma, (const float4*)beta, (float*)y, pooled, (float2*)stats, HW, C, eps, rpg);
    else
        gn1_relu_fwd_kernel<__nv_bfloat16><<<B, kGnThreads, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, (const float4*)gamma, (const float4*)beta, (__nv_bfloat16*)y, pooled, (float2*)stats, HW,
            C, eps, rpg);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_gn1_relu_bwd_t(const void* x, int dtype, const float* gamma, const float* beta, const float* stats,
                                    const void* dy, const float* dpooled, void* dx, float* dgamma_part, float* dbeta_part,
                                    int B, int HW, int C, int rows_per_group, hdmoe_stream_t stream) {
    const int rpg = rows_per_group > 0 ? rows_per_group : B;
    HDMOE_CHECK_ARG(x && gamma && beta && stats && dx && dgamma_part && dbeta_part && B >= 1 && HW >= 1, "gn1_relu_bwd: bad args");
    HDMOE_CHECK_ARG((dy != nullptr) != (dpooled != nullptr), "gn1_relu_bwd: exactly one of dy / dpooled");
    HDMOE_CHECK_ARG(gn_shape_ok(C), "gn1_relu_bwd: C / 4 must divide 1024 (got C = %d)", C);
    HDMOE_CHECK_ARG(dtype == HDMOE_F32 || dtype == HDMOE_BF16, "gn1_relu_bwd: dtype must be f32 or bf16");
    if (dtype == HDMOE_F32)
        gn1_relu_bwd_kernel<float><<<B, kGnThreads, 0, (cudaStream_t)stream>>>(
            (const float*)x, (const float4*)gamma, (const float4*)beta, (const float2*)stats, (const float*)dy, dpooled,
            (float*)dx, dgamma_part, dbeta_part, HW, C, rpg);
    else
        gn1_relu_bwd_kernel<__nv_bfloat16><<<B, kGnThreads, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, (const float4*)gamma, (const float4*)beta, (const float2*)stats,
            (const __nv_bfloat16*)dy, dpooled, (__nv_bfloat16*)dx, dgamma_part, dbeta_part, HW, C, rpg);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_gn1_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* pooled, float* stats,
                                  int B, int HW, int C, float eps, hdmoe_stream_t stream) {
    return hdmoe_gn1_relu_fwd_t(x, HDMOE_F32, gamma, beta, y, pooled, stats, B, HW, C, eps, 0, stream);
}

extern "C" int hdmoe_gn1_relu_bwd(const float* x, const float* gamma, const float* beta, const float* stats, const float* dy,
                                  const float* dpooled, float* dx, float* dgamma_part, float* dbeta_part, int B, int HW, int C,
                                  hdmoe_stream_t stream) {
    return hdmoe_gn1_relu_bwd_t(x, HDMOE_F32, gamma, beta, stats, dy, dpooled, dx, dgamma_part, dbeta_part, B, HW, C, 0, stream);
}
