// Fused NHWC bf16 elementwise kernels of the U-Net expert block (Unet_block.forward,
// models/model_components.py:232-253, and the glue of Unet_expert.forward, :416,428): pixel-norm + mp_silu,
// emb-gain * mp_silu, mp_sum, mp_cat, and the NCHW <-> NHWC transposes at the dispatch / combine boundary.
// Each replaces 2-6 elementwise ATen passes over activation-sized tensors (forward and backward).  HBM-bound:
// every thread moves 16-byte vectors (8 bf16), consecutive lanes on consecutive vectors.
#include "common.cuh"

namespace hdmoe {

constexpr float kSiluGain = 1.f / 0.596f;
constexpr float kPixEps = 1e-4f;

struct V8 {
    float v[8];
};
__device__ __forceinline__ V8 ld8(const __nv_bfloat16* p) {
    const int4 u = *reinterpret_cast<const int4*>(p);
    const uint32_t w[4] = {(uint32_t)u.x, (uint32_t)u.y, (uint32_t)u.z, (uint32_t)u.w};
    V8 r;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        r.v[2 * q] = __uint_as_float(w[q] << 16);
        r.v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
    return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const V8& a) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __nv_bfloat162 o = __floats2bfloat162_rn(a.v[2 * q], a.v[2 * q + 1]);
        w[q] = *reinterpret_cast<uint32_t*>(&o);
    }
    *reinterpret_cast<int4*>(p) = make_int4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float silu_f(float u) { return u / (1.f + __expf(-u)); }
__device__ __forceinline__ float dsilu_f(float u) {
    const float s = 1.f / (1.f + __expf(-u));
    return s * (1.f + u * (1.f - s));
}

// ---- pixel-norm (+ mp_silu): LP = C/8 lanes share a pixel ------------------------------------------------
template <int LP>
__global__ void __launch_bounds__(256)
pixnorm_silu_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ xn, __nv_bfloat16* __restrict__ a,
                        long long nvec) {
    constexpr int C = LP * 8;
    const float alpha = rsqrtf((float)C);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        V8 v = ld8(x + i * 8);
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) ss += v.v[q] * v.v[q];
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float s = 1.f / (kPixEps + alpha * sqrtf(ss));
        V8 n, act;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            n.v[q] = v.v[q] * s;
            act.v[q] = silu_f(n.v[q]) * kSiluGain;
        }
        st8(xn + i * 8, n);
        st8(a + i * 8, act);
    }
}
// dx from (g_xn, g_a):  t = g_xn + g_a * silu'(xn)/0.596 ;  dx = s*t - s^2*alpha/n * x * <x, t>
template <int LP>
__global__ void __launch_bounds__(256)
pixnorm_silu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g_xn,
                        const __nv_bfloat16* __restrict__ g_a, __nv_bfloat16* __restrict__ dx, long long nvec) {
    constexpr int C = LP * 8;
    const float alpha = rsqrtf((float)C);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const V8 v = ld8(x + i * 8);
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) ss += v.v[q] * v.v[q];
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float nrm = sqrtf(ss);
        const float s = 1.f / (kPixEps + alpha * nrm);
        V8 t;
        const V8 ga = ld8(g_a + i * 8);
        if (g_xn) t = ld8(g_xn + i * 8);
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float u = v.v[q] * s;
            t.v[q] = (g_xn ? t.v[q] : 0.f) + ga.v[q] * dsilu_f(u) * kSiluGain;
            dot += v.v[q] * t.v[q];
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float k = nrm > 0.f ? s * s * alpha / nrm * dot : 0.f;
        V8 o8;
#pragma unroll
        for (int q = 0; q < 8; ++q) o8.v[q] = s * t.v[q] - k * v.v[q];
        st8(dx + i * 8, o8);
    }
}

// ---- y = mp_silu(z * gain[row, c])  (gain may be NULL) --------------------------------------------------
__global__ void __launch_bounds__(256)
gain_silu_fwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ gain, __nv_bfloat16* __restrict__ y,
                     long long nvec, int C, long long vec_per_row) {
    const int cvec = C / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        V8 v = ld8(z + i * 8);
        const float* g = gain ? gain + (i / vec_per_row) * C + (int)(i % cvec) * 8 : nullptr;
        V8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) o.v[q] = silu_f(g ? v.v[q] * g[q] : v.v[q]) * kSiluGain;
        st8(y + i * 8, o);
    }
}
// one CTA per row: dz = dy * silu'(u)/0.596 * gain ; dgain[row, c] = sum_pixels dy * silu'(u)/0.596 * z
__global__ void __launch_bounds__(256)
gain_silu_bwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ gain, const __nv_bfloat16* __restrict__ dy,
                     __nv_bfloat16* __restrict__ dz, float* __restrict__ dgain, int C, long long vec_per_row) {
    __shared__ float red[256][9];
    const int cvec = C / 8;
    const long long row = blockIdx.x;
    const int chunk = threadIdx.x % cvec;               // 256 % cvec == 0 for C in {32, 64, 128}; see host check
    float gl[8], acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        gl[q] = gain ? gain[row * C + chunk * 8 + q] : 1.f;
        acc[q] = 0.f;
    }
    for (long long j = threadIdx.x; j < vec_per_row; j += blockDim.x) {
        const long long i = row * vec_per_row + j;
        const V8 v = ld8(z + i * 8), g = ld8(dy + i * 8);
        V8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float du = g.v[q] * dsilu_f(v.v[q] * gl[q]) * kSiluGain;
            o.v[q] = du * gl[q];
            acc[q] += du * v.v[q];
        }
        st8(dz + i * 8, o);
    }
    if (!dgain) return;
#pragma unroll
    for (int q = 0; q < 8; ++q) red[threadIdx.x][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < C) {
        const int ch = threadIdx.x / 8, q = threadIdx.x % 8;
        float s = 0.f;
        for (int t = ch; t < 256; t += cvec) s += red[t][q];
        dgain[row * C + threadIdx.x] = s;
    }
}

// dz = dy * silu'(z)/0.596 for any channel count (no gain)
__global__ void __launch_bounds__(256)
silu_bwd_flat_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dz,
                     long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const V8 v = ld8(z + i * 8), g = ld8(dy + i * 8);
        V8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) o.v[q] = g.v[q] * dsilu_f(v.v[q]) * kSiluGain;
        st8(dz + i * 8, o);
    }
}

// ---- out = ca*x + cb*y ;  backward: (gx, gy) = (ca*g, cb*g) ---------------------------------------------
__global__ void __launch_bounds__(256)
axpby_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y, float ca, float cb,
             __nv_bfloat16* __restrict__ out, long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const V8 a = ld8(x + i * 8), b = ld8(y + i * 8);
        V8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) o.v[q] = ca * a.v[q] + cb * b.v[q];
        st8(out + i * 8, o);
    }
}
__global__ void __launch_bounds__(256)
scale2_kernel(const __nv_bfloat16* __restrict__ g, float ca, float cb, __nv_bfloat16* __restrict__ gx,
              __nv_bfloat16* __restrict__ gy, long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const V8 a = ld8(g + i * 8);
        V8 o1, o2;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            o1.v[q] = ca * a.v[q];
            o2.v[q] = cb * a.v[q];
        }
        st8(gx + i * 8, o1);
        st8(gy + i * 8, o2);
    }
}

// ---- mp_cat along channels: out[p, :Ca] = wa*a[p], out[p, Ca:] = wb*b[p]; split = backward -------------
__global__ void __launch_bounds__(256)
cat_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, float wa, float wb, int Ca, int Cb,
           __nv_bfloat16* __restrict__ out, long long npix) {
    const int va = Ca / 8, vo = (Ca + Cb) / 8;
    const long long nvec = npix * vo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / vo;
        const int c = (int)(i - p * vo);
        const bool fa = c < va;
        V8 v = fa ? ld8(a + (p * va + c) * 8) : ld8(b + (p * (vo - va) + (c - va)) * 8);
        const float w = fa ? wa : wb;
#pragma unroll
        for (int q = 0; q < 8; ++q) v.v[q] *= w;
        st8(out + i * 8, v);
    }
}
__global__ void __launch_bounds__(256)
split_kernel(const __nv_bfloat16* __restrict__ g, float wa, float wb, int Ca, int Cb, __nv_bfloat16* __restrict__ ga,
             __nv_bfloat16* __restrict__ gb, long long npix) {
    const int va = Ca / 8, vo = (Ca + Cb) / 8;
    const long long nvec = npix * vo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / vo;
        const int c = (int)(i - p * vo);
        const bool fa = c < va;
        V8 v = ld8(g + i * 8);
        const float w = fa ? wa : wb;
#pragma unroll
        for (int q = 0; q < 8; ++q) v.v[q] *= w;
        if (fa) st8(ga + (p * va + c) * 8, v);
        else st8(gb + (p * (vo - va) + (c - va)) * 8, v);
    }
}

// ---- NCHW [R, Cs, HW] <-> NHWC [R, HW, Cd] transposes through shared memory (32-pixel tiles) ------------
// to_nhwc: channels c < Cs copied, channel Cs set to `one_value` when one_channel (the appended ones channel,
// models/model_components.py:416), remaining channels zero.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cs, int Cd, int HW,
                    int one_channel) {
    __shared__ __nv_bfloat16 tile[32][136];     // [pixel][channel], padded
    const int r = blockIdx.y, p0 = blockIdx.x * 32;
    for (int i = threadIdx.x; i < Cs * 32; i += blockDim.x) {
        const int c = i / 32, p = i % 32;
        if (p0 + p < HW) tile[p][c] = src[((size_t)r * Cs + c) * HW + p0 + p];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * Cd; i += blockDim.x) {
        const int p = i / Cd, c = i % Cd;
        if (p0 + p < HW) {
            __nv_bfloat16 v = __float2bfloat16(0.f);
            if (c < Cs) v = tile[p][c];
            else if (one_channel && c == Cs) v = __float2bfloat16(1.f);
            dst[((size_t)r * HW + p0 + p) * Cd + c] = v;
        }
    }
}
// to_nchw: dst [R, Cd, HW] takes the first Cd channels of src [R, HW, Cs]
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cs, int Cd, int HW) {
    __shared__ __nv_bfloat16 tile[32][136];
    const int r = blockIdx.y, p0 = blockIdx.x * 32;
    for (int i = threadIdx.x; i < 32 * Cd; i += blockDim.x) {
        const int p = i / Cd, c = i % Cd;
        if (p0 + p < HW) tile[p][c] = src[((size_t)r * HW + p0 + p) * Cs + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Cd * 32; i += blockDim.x) {
        const int c = i / 32, p = i % 32;
        if (p0 + p < HW) dst[((size_t)r * Cd + c) * HW + p0 + p] = tile[p][c];
    }
}

}  // namespace hdmoe
using namespace hdmoe;
typedef __nv_bfloat16 bf16;

#define NHWC_GRID(nvec) grid_for((nvec), 256, 16)

extern "C" int hdmoe_nhwc_pixnorm_silu_fwd(const void* x, void* xn, void* a, int64_t npix, int C, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x && xn && a && npix >= 1 && (C == 32 || C == 64 || C == 128), "pixnorm_silu: C must be 32, 64 or 128");
    const long long nvec = npix * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 32) pixnorm_silu_fwd_kernel<4><<<NHWC_GRID(nvec), 256, 0, st>>>((const bf16*)x, (bf16*)xn, (bf16*)a, nvec);
    else if (C == 64) pixnorm_silu_fwd_kernel<8><<<NHWC_GRID(nvec), 256, 0, st>>>((const bf16*)x, (bf16*)xn, (bf16*)a, nvec);
    else pixnorm_silu_fwd_kernel<16><<<NHWC_GRID(nvec), 256, 0, st>>>((const bf16*)x, (bf16*)xn, (bf16*)a, nvec);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_pixnorm_silu_bwd(const void* x, const void* g_xn, const void* g_a, void* dx, int64_t npix, int C,
                                           hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x && g_a && dx && npix >= 1 && (C == 32 || C == 64 || C == 128), "pixnorm_silu_bwd: C must be 32, 64 or 128");
    const long long nvec = npix * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 32) pixnorm_silu_bwd_kernel<4><<<NHWC_GRID(nvec), 256, 0, st>>>((const bf16*)x, (const bf16*)g_xn, (const bf16*)g_a, (bf16*)dx, nvec);
    else if (C == 64) pixnorm_silu_bwd_kernel<8><<<NHWC_GRID(nvec), 256, 0, st>>>((const bf16*)x, (const bf16*)g_xn, (const bf16*)g_a, (bf16*)dx, nvec);
    else pixnorm_silu_bwd_kernel<16><<<NHWC_GRID(nvec), 256, 0, st>>>((const bf16*)x, (const bf16*)g_xn, (const bf16*)g_a, (bf16*)dx, nvec);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_gain_silu_fwd(const void* z, const float* gain, void* y, int64_t rows, int64_t pix_per_row, int C,
                                        hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(z && y && rows >= 1 && pix_per_row >= 1 && C % 8 == 0, "gain_silu: C %% 8 != 0");
    const long long vpr = pix_per_row * (C / 8), nvec = rows * vpr;
    gain_silu_fwd_kernel<<<NHWC_GRID(nvec), 256, 0, (cudaStream_t)stream>>>((const bf16*)z, gain, (bf16*)y, nvec, C, vpr);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_gain_silu_bwd(const void* z, const float* gain, const void* dy, void* dz, float* dgain, int64_t rows,
                                        int64_t pix_per_row, int C, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(z && dy && dz && rows >= 1 && pix_per_row >= 1 && C % 8 == 0, "gain_silu_bwd: C %% 8 != 0");
    HDMOE_CHECK_ARG(!dgain || gain, "gain_silu_bwd: dgain without gain");
    if (!gain) {
        const long long nvec = rows * pix_per_row * (C / 8);
        silu_bwd_flat_kernel<<<NHWC_GRID(nvec), 256, 0, (cudaStream_t)stream>>>((const bf16*)z, (const bf16*)dy, (bf16*)dz, nvec);
        HDMOE_CHECK_LAUNCH();
        return HDMOE_OK;
    }
    HDMOE_CHECK_ARG(C == 32 || C == 64 || C == 128, "gain_silu_bwd: gained C must be 32, 64 or 128");
    gain_silu_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const bf16*)z, gain, (const bf16*)dy, (bf16*)dz, dgain,
                                                                        C, pix_per_row * (C / 8));
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_axpby(const void* x, const void* y, float ca, float cb, void* out, int64_t n, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x && y && out && n >= 8 && n % 8 == 0, "axpby: n %% 8 != 0");
    axpby_kernel<<<NHWC_GRID(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)y, ca, cb, (bf16*)out, n / 8);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_scale2(const void* g, float ca, float cb, void* gx, void* gy, int64_t n, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(g && gx && gy && n >= 8 && n % 8 == 0, "scale2: n %% 8 != 0");
    scale2_kernel<<<NHWC_GRID(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)g, ca, cb, (bf16*)gx, (bf16*)gy, n / 8);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_cat(const void* a, const void* b, float wa, float wb, int Ca, int Cb, void* out, int64_t npix,
                              hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(a && b && out && Ca % 8 == 0 && Cb % 8 == 0 && npix >= 1, "cat: channels %% 8 != 0");
    cat_kernel<<<NHWC_GRID(npix * ((Ca + Cb) / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, wa, wb, Ca, Cb,
                                                                                 (bf16*)out, npix);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_split(const void* g, float wa, float wb, int Ca, int Cb, void* ga, void* gb, int64_t npix,
                                hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(g && ga && gb && Ca % 8 == 0 && Cb % 8 == 0 && npix >= 1, "split: channels %% 8 != 0");
    split_kernel<<<NHWC_GRID(npix * ((Ca + Cb) / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)g, wa, wb, Ca, Cb, (bf16*)ga,
                                                                                   (bf16*)gb, npix);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nchw_to_nhwc(const void* src, void* dst, int64_t rows, int Cs, int Cd, int64_t HW, int one_channel,
                                  hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(src && dst && rows >= 1 && rows <= 65535 && Cs >= 1 && Cs <= 128 && Cd >= Cs + (one_channel ? 1 : 0) && Cd <= 128,
                    "nchw_to_nhwc: channels out of range");
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)rows);
    nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, Cs, Cd, (int)HW, one_channel);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
extern "C" int hdmoe_nhwc_to_nchw(const void* src, void* dst, int64_t rows, int Cs, int Cd, int64_t HW, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(src && dst && rows >= 1 && rows <= 65535 && Cd >= 1 && Cd <= Cs && Cs <= 128, "nhwc_to_nchw: channels out of range");
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)rows);
    nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, Cs, Cd, (int)HW);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
