// EDM-STEP: EDM preconditioning (models/model_config2.py:431-449) and the Heun 2nd-order step
// (Utils/EDM_sampler.py:98-107, CFG lerp :70) as fused elementwise kernels.  ~12 + ~10 eager launches
// per NFE become 2.  HBM-bound: ~36 B per latent element per Heun step at bf16 model I/O.
// fp32 arithmetic follows the reference's operation order with explicit roundings (no FMA contraction)
// so that, given the same denoiser output, the ODE state is bit-identical to the reference's.
#include "common.cuh"

namespace hdmoe {

struct Coef {
    float c_skip, c_out, c_in;
};
// models/model_config2.py:432-434 (fp32, same operation order)
__device__ __forceinline__ Coef edm_coef(float sigma, float sd) {
    const float s2 = __fmul_rn(sigma, sigma), d2 = __fmul_rn(sd, sd);
    const float sum = __fadd_rn(s2, d2);
    Coef c;
    c.c_skip = __fdiv_rn(d2, sum);
    c.c_out = __fdiv_rn(__fmul_rn(sigma, sd), __fsqrt_rn(sum));
    c.c_in = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(d2, s2)));
    return c;
}

template <typename TO>
__global__ void __launch_bounds__(256)
precond_in_kernel(const float* __restrict__ x, const float* __restrict__ sigma, int n_sigma, float sd,
                  TO* __restrict__ x_in, long long B, long long per) {
    const long long vper = per >> 2, total = B * vper;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / vper;
        const float ci = edm_coef(sigma[n_sigma == 1 ? 0 : b], sd).c_in;
        float4 v = *reinterpret_cast<const float4*>(x + (i << 2));
        v.x = __fmul_rn(v.x, ci); v.y = __fmul_rn(v.y, ci); v.z = __fmul_rn(v.z, ci); v.w = __fmul_rn(v.w, ci);
        Vec4<TO>::store(x_in + (i << 2), v);
    }
}

template <typename TI, typename TF>
__global__ void __launch_bounds__(256)
precond_out_kernel(const TI* __restrict__ x_in, const TF* __restrict__ F, const float* __restrict__ sigma,
                   int n_sigma, float sd, float* __restrict__ D, long long B, long long per) {
    const long long vper = per >> 2, total = B * vper;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / vper;
        const Coef c = edm_coef(sigma[n_sigma == 1 ? 0 : b], sd);
        const float4 xi = Vec4<TI>::load(x_in + (i << 2)), f = Vec4<TF>::load(F + (i << 2));
        float4 d;
        d.x = __fadd_rn(__fmul_rn(c.c_skip, xi.x), __fmul_rn(c.c_out, f.x));
        d.y = __fadd_rn(__fmul_rn(c.c_skip, xi.y), __fmul_rn(c.c_out, f.y));
        d.z = __fadd_rn(__fmul_rn(c.c_skip, xi.z), __fmul_rn(c.c_out, f.z));
        d.w = __fadd_rn(__fmul_rn(c.c_skip, xi.w), __fmul_rn(c.c_out, f.w));
        *reinterpret_cast<float4*>(D + (i << 2)) = d;
    }
}

template <typename TF, typename TG>
__global__ void __launch_bounds__(256)
precond_out_bwd_kernel(const float* __restrict__ dD, const float* __restrict__ sigma, int n_sigma, float sd,
                       TF* __restrict__ dF, TG* __restrict__ d_xin, long long B, long long per) {
    const long long vper = per >> 2, total = B * vper;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / vper;
        const Coef c = edm_coef(sigma[n_sigma == 1 ? 0 : b], sd);
        const float4 g = *reinterpret_cast<const float4*>(dD + (i << 2));
        if (dF) Vec4<TF>::store(dF + (i << 2), make_float4(c.c_out * g.x, c.c_out * g.y, c.c_out * g.z, c.c_out * g.w));
        if (d_xin)
            Vec4<TG>::store(d_xin + (i << 2), make_float4(c.c_skip * g.x, c.c_skip * g.y, c.c_skip * g.z, c.c_skip * g.w));
    }
}

template <typename TG>
__global__ void __launch_bounds__(256)
precond_in_bwd_kernel(const TG* __restrict__ d_xin, const float* __restrict__ sigma, int n_sigma, float sd,
                      float* __restrict__ dx, long long B, long long per) {
    const long long vper = per >> 2, total = B * vper;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / vper;
        const float ci = edm_coef(sigma[n_sigma == 1 ? 0 : b], sd).c_in;
        const float4 g = Vec4<TG>::load(d_xin + (i << 2));
        *reinterpret_cast<float4*>(dx + (i << 2)) = make_float4(ci * g.x, ci * g.y, ci * g.z, ci * g.w);
    }
}

template <typename TO>
__global__ void __launch_bounds__(256)
heun_pre_kernel(const float* __restrict__ x_cur, const float* __restrict__ eps, float noise_scale, float t_hat,
                float sd, float* __restrict__ x_hat, TO* __restrict__ x_in, long long n4) {
    const float ci = edm_coef(t_hat, sd).c_in;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 v = *reinterpret_cast<const float4*>(x_cur + (i << 2));
        if (eps) {   // x_hat = x_cur + (sqrt(t_hat^2 - t_cur^2) * S_noise) * randn   (EDM_sampler.py:99)
            const float4 e = *reinterpret_cast<const float4*>(eps + (i << 2));
            v.x = __fadd_rn(v.x, __fmul_rn(noise_scale, e.x));
            v.y = __fadd_rn(v.y, __fmul_rn(noise_scale, e.y));
            v.z = __fadd_rn(v.z, __fmul_rn(noise_scale, e.z));
            v.w = __fadd_rn(v.w, __fmul_rn(noise_scale, e.w));
        }
        *reinterpret_cast<float4*>(x_hat + (i << 2)) = v;
        Vec4<TO>::store(x_in + (i << 2), make_float4(__fmul_rn(v.x, ci), __fmul_rn(v.y, ci), __fmul_rn(v.z, ci),
                                                      __fmul_rn(v.w, ci)));
    }
}

__device__ __forceinline__ float denoised(const Coef& c, float xi, float f) {
    return __fadd_rn(__fmul_rn(c.c_skip, xi), __fmul_rn(c.c_out, f));
}
// ref.lerp(cond, w) = ref + w*(cond - ref)   (Utils/EDM_sampler.py:70; torch lerp for |w| >= 0.5 uses
// cond - (cond - ref)*(1 - w); both forms are evaluated exactly as ATen's lerp kernel does)
__device__ __forceinline__ float lerp_aten(float a, float b, float w) {
    const float diff = __fsub_rn(b, a);
    return fabsf(w) < 0.5f ? __fadd_rn(a, __fmul_rn(w, diff)) : __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.f, w)));
}

template <typename TI, typename TF>
__global__ void __launch_bounds__(256)
heun_euler_kernel(const float* __restrict__ x_hat, const TI* __restrict__ x_in, const TF* __restrict__ F,
                  const TF* __restrict__ Fg, float guidance, float t_hat, float t_next, float sd,
                  float* __restrict__ d_cur, float* __restrict__ x_next, TI* __restrict__ x_in_next, long long n) {
    const Coef c = edm_coef(t_hat, sd);
    const float ci_next = x_in_next ? edm_coef(t_next, sd).c_in : 0.f;
    const float dt = __fsub_rn(t_next, t_hat);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float xh = x_hat[i];
        float D;
        if (x_in) {   // F is the raw network output: apply the preconditioning here
            const float xi = to_f32<TI>(x_in[i]);
            D = denoised(c, xi, to_f32<TF>(F[i]));
            if (Fg) D = lerp_aten(denoised(c, xi, to_f32<TF>(Fg[i])), D, guidance);
        } else {      // F already is the denoised estimate D(x; sigma) of a foreign model
            D = to_f32<TF>(F[i]);
            if (Fg) D = lerp_aten(to_f32<TF>(Fg[i]), D, guidance);
        }
        const float d = __fdiv_rn(__fsub_rn(xh, D), t_hat);          // :101
        const float xn = __fadd_rn(xh, __fmul_rn(dt, d));            // :102
        d_cur[i] = d;
        x_next[i] = xn;
        if (x_in_next) x_in_next[i] = from_f32<TI>(__fmul_rn(xn, ci_next));
    }
}

template <typename TI, typename TF>
__global__ void __launch_bounds__(256)
heun_correct_kernel(const float* __restrict__ x_hat, const float* __restrict__ x_next, const TI* __restrict__ x_in_next,
                    const TF* __restrict__ F, const TF* __restrict__ Fg, float guidance, float t_hat, float t_next,
                    float sd, const float* __restrict__ d_cur, float* __restrict__ x_out, long long n) {
    const Coef c = edm_coef(t_next, sd);
    const float dt = __fsub_rn(t_next, t_hat);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float D;
        if (x_in_next) {
            const float xi = to_f32<TI>(x_in_next[i]);
            D = denoised(c, xi, to_f32<TF>(F[i]));
            if (Fg) D = lerp_aten(denoised(c, xi, to_f32<TF>(Fg[i])), D, guidance);
        } else {
            D = to_f32<TF>(F[i]);
            if (Fg) D = lerp_aten(to_f32<TF>(Fg[i]), D, guidance);
        }
        const float dp = __fdiv_rn(__fsub_rn(x_next[i], D), t_next);                                  // :106
        const float avg = __fadd_rn(__fmul_rn(0.5f, d_cur[i]), __fmul_rn(0.5f, dp));
        x_out[i] = __fadd_rn(x_hat[i], __fmul_rn(dt, avg));                                           // :107
    }
}

}  // namespace hdmoe
using namespace hdmoe;

#define EDM_ARGS_OK(B, per) HDMOE_CHECK_ARG((B) >= 1 && (per) >= 4 && (per) % 4 == 0, "edm: per-sample size must be a multiple of 4")

extern "C" int hdmoe_edm_precond_in(const float* x, const float* sigma, int n_sigma, float sigma_data, void* x_in,
                                    int x_in_dtype, int64_t B, int64_t per, hdmoe_stream_t stream) {
    EDM_ARGS_OK(B, per);
    HDMOE_CHECK_ARG(x && sigma && x_in && (n_sigma == 1 || n_sigma == B), "edm_precond_in: bad pointers / n_sigma");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * (per >> 2), 256, 8);
    if (x_in_dtype == HDMOE_F32)
        precond_in_kernel<float><<<grid, 256, 0, st>>>(x, sigma, n_sigma, sigma_data, (float*)x_in, B, per);
    else
        precond_in_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, sigma, n_sigma, sigma_data, (__nv_bfloat16*)x_in, B, per);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

#define DISPATCH2(K, a_dt, b_dt, ...)                                                        \
    do {                                                                                     \
        if (a_dt == HDMOE_F32 && b_dt == HDMOE_F32) K<float, float> __VA_ARGS__;             \
        else if (a_dt == HDMOE_F32 && b_dt == HDMOE_BF16) K<float, __nv_bfloat16> __VA_ARGS__; \
        else if (a_dt == HDMOE_BF16 && b_dt == HDMOE_F32) K<__nv_bfloat16, float> __VA_ARGS__; \
        else K<__nv_bfloat16, __nv_bfloat16> __VA_ARGS__;                                    \
    } while (0)

extern "C" int hdmoe_edm_precond_out(const void* x_in, int x_in_dtype, const void* F, int f_dtype, const float* sigma,
                                     int n_sigma, float sigma_data, float* D, int64_t B, int64_t per,
                                     hdmoe_stream_t stream) {
    EDM_ARGS_OK(B, per);
    HDMOE_CHECK_ARG(x_in && F && sigma && D && (n_sigma == 1 || n_sigma == B), "edm_precond_out: bad pointers / n_sigma");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * (per >> 2), 256, 8);
#define ARGS(TI, TF) <<<grid, 256, 0, st>>>((const TI*)x_in, (const TF*)F, sigma, n_sigma, sigma_data, D, B, per)
    if (x_in_dtype == HDMOE_F32 && f_dtype == HDMOE_F32) precond_out_kernel<float, float> ARGS(float, float);
    else if (x_in_dtype == HDMOE_F32) precond_out_kernel<float, __nv_bfloat16> ARGS(float, __nv_bfloat16);
    else if (f_dtype == HDMOE_F32) precond_out_kernel<__nv_bfloat16, float> ARGS(__nv_bfloat16, float);
    else precond_out_kernel<__nv_bfloat16, __nv_bfloat16> ARGS(__nv_bfloat16, __nv_bfloat16);
#undef ARGS
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_edm_precond_out_bwd(const float* dD, const float* sigma, int n_sigma, float sigma_data, void* dF,
                                         int df_dtype, void* d_xin, int dxin_dtype, int64_t B, int64_t per,
                                         hdmoe_stream_t stream) {
    EDM_ARGS_OK(B, per);
    HDMOE_CHECK_ARG(dD && sigma && (n_sigma == 1 || n_sigma == B), "edm_precond_out_bwd: bad pointers / n_sigma");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * (per >> 2), 256, 8);
#define ARGS(TF, TG) <<<grid, 256, 0, st>>>(dD, sigma, n_sigma, sigma_data, (TF*)dF, (TG*)d_xin, B, per)
    if (df_dtype == HDMOE_F32 && dxin_dtype == HDMOE_F32) precond_out_bwd_kernel<float, float> ARGS(float, float);
    else if (df_dtype == HDMOE_F32) precond_out_bwd_kernel<float, __nv_bfloat16> ARGS(float, __nv_bfloat16);
    else if (dxin_dtype == HDMOE_F32) precond_out_bwd_kernel<__nv_bfloat16, float> ARGS(__nv_bfloat16, float);
    else precond_out_bwd_kernel<__nv_bfloat16, __nv_bfloat16> ARGS(__nv_bfloat16, __nv_bfloat16);
#undef ARGS
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_edm_precond_in_bwd(const void* d_xin, int dxin_dtype, const float* sigma, int n_sigma,
                                        float sigma_data, float* dx, int64_t B, int64_t per, hdmoe_stream_t stream) {
    EDM_ARGS_OK(B, per);
    HDMOE_CHECK_ARG(d_xin && sigma && dx && (n_sigma == 1 || n_sigma == B), "edm_precond_in_bwd: bad pointers / n_sigma");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * (per >> 2), 256, 8);
    if (dxin_dtype == HDMOE_F32)
        precond_in_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)d_xin, sigma, n_sigma, sigma_data, dx, B, per);
    else
        precond_in_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)d_xin, sigma, n_sigma, sigma_data, dx, B, per);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_edm_heun_pre(const float* x_cur, const float* eps, float noise_scale, float t_hat, float sigma_data,
                                  float* x_hat, void* x_in, int x_in_dtype, int64_t n, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x_cur && x_hat && x_in && n >= 4 && n % 4 == 0, "edm_heun_pre: n must be a multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n >> 2, 256, 8);
    if (x_in_dtype == HDMOE_F32)
        heun_pre_kernel<float><<<grid, 256, 0, st>>>(x_cur, eps, noise_scale, t_hat, sigma_data, x_hat, (float*)x_in, n >> 2);
    else
        heun_pre_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x_cur, eps, noise_scale, t_hat, sigma_data, x_hat,
                                                             (__nv_bfloat16*)x_in, n >> 2);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_edm_heun_euler(const float* x_hat, const void* x_in, int x_in_dtype, const void* F,
                                    const void* F_guide, int f_dtype, float guidance, float t_hat, float t_next,
                                    float sigma_data, float* d_cur, float* x_next, void* x_in_next, int64_t n,
                                    hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x_hat && F && d_cur && x_next && n >= 1, "edm_heun_euler: null pointer");
    HDMOE_CHECK_ARG(t_hat > 0.f, "edm_heun_euler: t_hat must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n, 256, 8);
#define ARGS(TI, TF) <<<grid, 256, 0, st>>>(x_hat, (const TI*)x_in, (const TF*)F, (const TF*)F_guide, guidance, t_hat, t_next, sigma_data, d_cur, x_next, (TI*)x_in_next, n)
    if (x_in_dtype == HDMOE_F32 && f_dtype == HDMOE_F32) heun_euler_kernel<float, float> ARGS(float, float);
    else if (x_in_dtype == HDMOE_F32) heun_euler_kernel<float, __nv_bfloat16> ARGS(float, __nv_bfloat16);
    else if (f_dtype == HDMOE_F32) heun_euler_kernel<__nv_bfloat16, float> ARGS(__nv_bfloat16, float);
    else heun_euler_kernel<__nv_bfloat16, __nv_bfloat16> ARGS(__nv_bfloat16, __nv_bfloat16);
#undef ARGS
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_edm_heun_correct(const float* x_hat, const float* x_next, const void* x_in_next, int x_in_dtype,
                                      const void* F, const void* F_guide, int f_dtype, float guidance, float t_hat,
                                      float t_next, float sigma_data, const float* d_cur, float* x_out, int64_t n,
                                      hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x_hat && x_next && F && d_cur && x_out && n >= 1, "edm_heun_correct: null pointer");
    HDMOE_CHECK_ARG(t_next > 0.f, "edm_heun_correct: t_next must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n, 256, 8);
#define ARGS(TI, TF) <<<grid, 256, 0, st>>>(x_hat, x_next, (const TI*)x_in_next, (const TF*)F, (const TF*)F_guide, guidance, t_hat, t_next, sigma_data, d_cur, x_out, n)
    if (x_in_dtype == HDMOE_F32 && f_dtype == HDMOE_F32) heun_correct_kernel<float, float> ARGS(float, float);
    else if (x_in_dtype == HDMOE_F32) heun_correct_kernel<float, __nv_bfloat16> ARGS(float, __nv_bfloat16);
    else if (f_dtype == HDMOE_F32) heun_correct_kernel<__nv_bfloat16, float> ARGS(__nv_bfloat16, float);
    else heun_correct_kernel<__nv_bfloat16, __nv_bfloat16> ARGS(__nv_bfloat16, __nv_bfloat16);
#undef ARGS
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// EDM_LOSS data term (Utils/utils.py:135-146): every loss term that touches the images depends on them only through the
// per-sample squared error se[b] = sum_i (D[b,i] - x0[b,i])^2 (log_var is one scalar per sample), so the activation-
// sized part of the loss is ONE reduction kernel forward and ONE scaling kernel backward; the remaining arithmetic runs
// on [B]-sized vectors.  One CTA per sample, fixed summation order (deterministic).
// ---------------------------------------------------------------------------------------------------------------------
namespace hdmoe {
__global__ void __launch_bounds__(256)
sqerr_rows_kernel(const float4* __restrict__ d, const float4* __restrict__ x, float* __restrict__ se, long long per4) {
    __shared__ float red[8];
    const float4* dr = d + (size_t)blockIdx.x * per4;
    const float4* xr = x + (size_t)blockIdx.x * per4;
    float s = 0.f;
    for (long long i = threadIdx.x; i < per4; i += 256) {
        const float4 a = dr[i], b = xr[i];
        const float e0 = a.x - b.x, e1 = a.y - b.y, e2 = a.z - b.z, e3 = a.w - b.w;
        s += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        se[blockIdx.x] = t;
    }
}
// dD[b, i] = 2 * (D - x0) * g[b]
__global__ void __launch_bounds__(256)
sqerr_rows_bwd_kernel(const float4* __restrict__ d, const float4* __restrict__ x, const float* __restrict__ g,
                      float4* __restrict__ dd, long long per4, long long n4) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float c = 2.f * g[i / per4];
        const float4 a = d[i], b = x[i];
        dd[i] = make_float4(c * (a.x - b.x), c * (a.y - b.y), c * (a.z - b.z), c * (a.w - b.w));
    }
}
}  // namespace hdmoe

extern "C" int hdmoe_sqerr_rows(const float* d, const float* x, float* se, int B, int64_t per, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(d && x && se && B >= 1 && per >= 4 && per % 4 == 0, "sqerr_rows: bad args (row length must be a multiple of 4)");
    hdmoe::sqerr_rows_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const float4*)d, (const float4*)x, se, per / 4);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_sqerr_rows_bwd(const float* d, const float* x, const float* g_se, float* dd, int B, int64_t per,
                                    hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(d && x && g_se && dd && B >= 1 && per >= 4 && per % 4 == 0, "sqerr_rows_bwd: bad args");
    const long long n4 = (long long)B * per / 4;
    hdmoe::sqerr_rows_bwd_kernel<<<hdmoe::grid_for(n4, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)d, (const float4*)x, g_se, (float4*)dd, per / 4, n4);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// On-device producers of the train step's inputs: the noise add of Utils/training.py:133-134
// (noise = eps * sigma; x = latent + noise, fp32, two roundings as in the reference) and BOTH MaskGenerator band masks
// (Utils/utils.py:281-309) in one launch: ~25 small launches and no host round trip per step.
// ---------------------------------------------------------------------------------------------------------------
namespace hdmoe {
struct MaskGenParams {
    float centers[HDMOE_MAX_MASK_EXPERTS];
    float p_mean, div, bandwidth;            // div = p_std * sqrt(2); a true fp32 division, as the CPU reference does
    int n_experts, min_active;
};

__device__ __forceinline__ void band_mask_row(float sigma, const MaskGenParams& g, float* __restrict__ out) {
    // pct = clamp(0.5 * (1 + erf((log sigma - p_mean) / (p_std * sqrt 2))), 0, 1)
    const float ls = logf(sigma);
    float pct = __fmul_rn(0.5f, __fadd_rn(1.f, erff(__fdiv_rn(__fsub_rn(ls, g.p_mean), g.div))));
    pct = fminf(fmaxf(pct, 0.f), 1.f);
    float dist[HDMOE_MAX_MASK_EXPERTS];
    unsigned picked = 0;
#pragma unroll
    for (int e = 0; e < HDMOE_MAX_MASK_EXPERTS; ++e)
        if (e < g.n_experts) {
            dist[e] = fabsf(__fsub_rn(pct, g.centers[e]));
            out[e] = dist[e] <= g.bandwidth ? 1.f : 0.f;
        }
    // the min_active nearest experts are always live (topk(-dist); lowest index wins a tie)
    for (int m = 0; m < g.min_active && m < g.n_experts; ++m) {
        int best = -1;
        float bd = 0.f;
#pragma unroll
        for (int e = 0; e < HDMOE_MAX_MASK_EXPERTS; ++e)
            if (e < g.n_experts && !((picked >> e) & 1u) && (best < 0 || dist[e] < bd)) { best = e; bd = dist[e]; }
        picked |= 1u << best;
        out[best] = 1.f;
    }
}

__global__ void __launch_bounds__(256)
train_inputs_kernel(const float4* __restrict__ x0, const float4* __restrict__ eps, const float* __restrict__ sigma,
                    float4* __restrict__ x, long long per4, long long n4, int B, MaskGenParams ga, MaskGenParams gb,
                    float* __restrict__ mask_a, float* __restrict__ mask_b) {
    const long long tid = (long long)blockIdx.x * 256 + threadIdx.x;
    if (tid < B) {
        const float s = sigma[tid];
        if (mask_a) band_mask_row(s, ga, mask_a + tid * ga.n_experts);
        if (mask_b) band_mask_row(s, gb, mask_b + tid * gb.n_experts);
    }
    for (long long i = tid; i < n4; i += (long long)gridDim.x * 256) {
        const float s = sigma[i / per4];
        const float4 a = x0[i], e = eps[i];
        x[i] = make_float4(__fadd_rn(a.x, __fmul_rn(e.x, s)), __fadd_rn(a.y, __fmul_rn(e.y, s)),
                           __fadd_rn(a.z, __fmul_rn(e.z, s)), __fadd_rn(a.w, __fmul_rn(e.w, s)));
    }
}
}  // namespace hdmoe

static int fill_mask_params(hdmoe::MaskGenParams& g, const hdmoe_maskgen_t* m, const char* which) {
    if (!m) { g.n_experts = 0; g.min_active = 0; return HDMOE_OK; }
    HDMOE_CHECK_ARG(m->n_experts >= 1 && m->n_experts <= HDMOE_MAX_MASK_EXPERTS, "train_inputs: %s: 1 <= n_experts <= %d",
                    which, HDMOE_MAX_MASK_EXPERTS);
    HDMOE_CHECK_ARG(m->min_active >= 0 && m->min_active <= m->n_experts && m->p_std > 0.f, "train_inputs: %s: bad min_active / p_std", which);
    for (int e = 0; e < m->n_experts; ++e) g.centers[e] = m->centers[e];
    g.p_mean = m->p_mean;
    g.div = (float)((double)m->p_std * 1.4142135623730951);     // p_std * np.sqrt(2) in double, then the fp32 division
    g.bandwidth = m->bandwidth;
    g.n_experts = m->n_experts;
    g.min_active = m->min_active;
    return HDMOE_OK;
}

extern "C" int hdmoe_train_inputs(const float* x0, const float* eps, const float* sigma, float* x, int B, int64_t per,
                                  const hdmoe_maskgen_t* gen_a, float* mask_a, const hdmoe_maskgen_t* gen_b, float* mask_b,
                                  hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x0 && eps && sigma && x && B >= 1 && per >= 4 && per % 4 == 0, "train_inputs: bad args (row length must be a multiple of 4)");
    HDMOE_CHECK_ARG((gen_a != nullptr) == (mask_a != nullptr) && (gen_b != nullptr) == (mask_b != nullptr),
                    "train_inputs: a mask generator and its output go together");
    hdmoe::MaskGenParams ga{}, gb{};
    int rc = fill_mask_params(ga, gen_a, "gen_a");
    if (rc != HDMOE_OK) return rc;
    rc = fill_mask_params(gb, gen_b, "gen_b");
    if (rc != HDMOE_OK) return rc;
    const long long n4 = (long long)B * per / 4;
    int grid = hdmoe::grid_for(n4, 256, 8);
    if ((long long)grid * 256 < B) grid = (B + 255) / 256;
    hdmoe::train_inputs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)x0, (const float4*)eps, sigma, (float4*)x,
                                                                        per / 4, n4, B, ga, gb, mask_a, mask_b);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
