// W-PREP: multi-tensor magnitude-preserving weight preparation.  The reference re-normalises every
// MP_Conv weight on every call with ~8 tiny ATen kernels (models/model_internals.py:253-260, 26-30):
// 271 layers -> ~2000 launches per step.  Here ONE launch covers any number of layers: a CTA per
// (tensor, output row) computes the row norm with a warp-shuffle reduction, optionally rewrites the fp32
// master row in place (training-mode forced weight norm, quirk Q6) and emits the scaled row in the
// dtype / layout its consumer wants (plain [rows][fan_in], or the K-major tap layout of the implicit-GEMM
// convolution: [tap][rows][cin_pad], zero padded).
#include "common.cuh"

namespace hdmoe {

constexpr int kWprepThreads = 128;
constexpr float kEps = 1e-4f;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kWprepThreads / 32; ++q) s += red[q];
    return s;
}

template <typename T>
__device__ __forceinline__ void put(void* base, size_t i, float v);
template <>
__device__ __forceinline__ void put<float>(void* base, size_t i, float v) { ((float*)base)[i] = v; }
template <>
__device__ __forceinline__ void put<__nv_bfloat16>(void* base, size_t i, float v) {
    ((__nv_bfloat16*)base)[i] = __float2bfloat16_rn(v);
}

template <typename T>
__device__ __forceinline__ void emit_row_t(const hdmoe_wprep_desc& d, int layout, void* out, const float* w, int row,
                                           float scale) {
    if (layout == HDMOE_WLAYOUT_SAME) {
        for (int i = threadIdx.x; i < d.fan_in; i += kWprepThreads) put<T>(out, (size_t)row * d.fan_in + i, w[i] * scale);
    } else if (layout == HDMOE_WLAYOUT_TAPS) {
        // source row is [cin][taps]; destination is [tap][rows][cin_pad]
        for (int i = threadIdx.x; i < d.taps * d.cin_pad; i += kWprepThreads) {
            const int tap = i / d.cin_pad, c = i - tap * d.cin_pad;
            const float v = c < d.cin ? w[(size_t)c * d.taps + tap] * scale : 0.f;
            put<T>(out, ((size_t)tap * d.rows + row) * d.cin_pad + c, v);
        }
    } else {
        // data-gradient operand: destination [taps-1-tap][cin_rows][cout_pad], this CTA owns column `row`
        for (int i = threadIdx.x; i < d.taps * d.cin_rows; i += kWprepThreads) {
            const int tap = i / d.cin_rows, c = i - tap * d.cin_rows;
            put<T>(out, ((size_t)(d.taps - 1 - tap) * d.cin_rows + c) * d.cout_pad + row, w[(size_t)c * d.taps + tap] * scale);
        }
    }
}
__device__ __forceinline__ void emit_row(const hdmoe_wprep_desc& d, int layout, void* out, const float* w, int row,
                                         float scale) {
    if (d.out_dtype == HDMOE_F32) emit_row_t<float>(d, layout, out, w, row, scale);
    else emit_row_t<__nv_bfloat16>(d, layout, out, w, row, scale);
}

__global__ void __launch_bounds__(kWprepThreads)
wprep_fwd_kernel(const hdmoe_wprep_desc* __restrict__ descs, int n, int force) {
    __shared__ float red[kWprepThreads / 32];
    // binary search: last descriptor with block_start <= blockIdx.x
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const hdmoe_wprep_desc d = descs[lo];
    const int row = blockIdx.x - d.block_start;
    if (row >= d.rows) return;
    float* w = d.w + (size_t)row * d.fan_in;
    const float alpha = rsqrtf((float)d.fan_in);   // sqrt(norm.numel()/w.numel()) = 1/sqrt(fan_in)
    float ss = 0.f;
    for (int i = threadIdx.x; i < d.fan_in; i += kWprepThreads) { const float v = w[i]; ss += v * v; }
    float nrm = sqrtf(block_sum(ss, red));
    float inv = 1.f / (kEps + alpha * nrm);
    if (force && (!d.active || *d.active != 0)) {
        // weights <- normalize(weights); the forward then normalises the REWRITTEN values once more
        float ss2 = 0.f;
        for (int i = threadIdx.x; i < d.fan_in; i += kWprepThreads) {
            const float v = w[i] * inv;
            w[i] = v;
            ss2 += v * v;
        }
        nrm = sqrtf(block_sum(ss2, red));
        inv = 1.f / (kEps + alpha * nrm);
    }
    const float gain = d.gain_ptr ? *d.gain_ptr : d.gain;
    const float scale = inv * gain * alpha;
    emit_row(d, d.layout, d.w_hat, w, row, scale);
    if (d.w_hat2) emit_row(d, d.layout2, d.w_hat2, w, row, scale);
}

// gradient through w_hat = w * s / (eps + a*||w||),  s = gain*a,  a = 1/sqrt(fan_in):
//   d_w = s/(eps + a n) * (g - w * a <w, g> / (n (eps + a n)))          d_gain = <w_hat, g> / gain
__global__ void __launch_bounds__(kWprepThreads)
wprep_bwd_kernel(const float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ gain_ptr,
                 float gain, int rows, int fan_in, float* __restrict__ d_w, float* __restrict__ d_gain) {
    __shared__ float red[kWprepThreads / 32];
    const int row = blockIdx.x;
    const float* wr = w + (size_t)row * fan_in;
    const float* gr = g + (size_t)row * fan_in;
    const float a = rsqrtf((float)fan_in);
    float ss = 0.f, wg = 0.f;
    for (int i = threadIdx.x; i < fan_in; i += kWprepThreads) {
        const float v = wr[i];
        ss += v * v;
        wg += v * gr[i];
    }
    const float n = sqrtf(block_sum(ss, red));
    wg = block_sum(wg, red);
    const float gn = gain_ptr ? *gain_ptr : gain;
    const float den = kEps + a * n;
    const float s = gn * a / den;
    const float k = n > 0.f ? a * wg / (n * den) : 0.f;
    for (int i = threadIdx.x; i < fan_in; i += kWprepThreads) d_w[(size_t)row * fan_in + i] = s * (gr[i] - wr[i] * k);
    if (d_gain && threadIdx.x == 0) atomicAdd(d_gain, wg * a / den);
}

// multi-tensor variant; d_w_hat may be in the tap-major layout the weight-gradient kernel accumulates
__global__ void __launch_bounds__(kWprepThreads)
wprep_bwd_multi_kernel(const hdmoe_wprep_bwd_desc* __restrict__ descs, int n) {
    __shared__ float red[kWprepThreads / 32];
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const hdmoe_wprep_bwd_desc d = descs[lo];
    const int row = blockIdx.x - d.block_start;
    if (row >= d.rows) return;
    const float* wr = d.w + (size_t)row * d.fan_in;
    auto g_at = [&](int i) -> float {
        if (d.layout == HDMOE_WLAYOUT_SAME) return d.d_w_hat[(size_t)row * d.fan_in + i];
        const int c = i / d.taps, tap = i - c * d.taps;      // master index i = c*taps + tap
        return d.d_w_hat[((size_t)tap * d.rows + row) * d.cin_pad + c];
    };
    const float a = rsqrtf((float)d.fan_in);
    float ss = 0.f, wg = 0.f;
    for (int i = threadIdx.x; i < d.fan_in; i += kWprepThreads) {
        const float v = wr[i];
        ss += v * v;
        wg += v * g_at(i);
    }
    const float nrm = sqrtf(block_sum(ss, red));
    wg = block_sum(wg, red);
    const float gn = d.gain_ptr ? *d.gain_ptr : d.gain;
    const float den = kEps + a * nrm;
    const float s = gn * a / den;
    const float k = nrm > 0.f ? a * wg / (nrm * den) : 0.f;
    for (int i = threadIdx.x; i < d.fan_in; i += kWprepThreads) d.d_w[(size_t)row * d.fan_in + i] = s * (g_at(i) - wr[i] * k);
    if (d.d_gain && threadIdx.x == 0) atomicAdd(d.d_gain, wg * a / den);
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_wprep_bwd_multi(hdmoe_wprep_bwd_desc* descs_host, void* descs_dev, int n, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(descs_host && descs_dev && n >= 1, "wprep_bwd_multi: null descriptor table");
    cudaStream_t st = (cudaStream_t)stream;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        hdmoe_wprep_bwd_desc& d = descs_host[i];
        HDMOE_CHECK_ARG(d.w && d.d_w_hat && d.d_w && d.rows >= 1 && d.fan_in >= 1, "wprep_bwd_multi: descriptor %d is empty", i);
        d.block_start = total;
        total += d.rows;
    }
    HDMOE_CHECK_CUDA(cudaMemcpyAsync(descs_dev, descs_host, (size_t)n * sizeof(hdmoe_wprep_bwd_desc), cudaMemcpyHostToDevice, st));
    wprep_bwd_multi_kernel<<<total, kWprepThreads, 0, st>>>((const hdmoe_wprep_bwd_desc*)descs_dev, n);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_wprep_fwd(hdmoe_wprep_desc* descs_host, void* descs_dev, int n, int force, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(descs_host && descs_dev && n >= 1, "wprep_fwd: null descriptor table");
    cudaStream_t st = (cudaStream_t)stream;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        hdmoe_wprep_desc& d = descs_host[i];
        HDMOE_CHECK_ARG(d.w && d.w_hat && d.rows >= 1 && d.fan_in >= 1, "wprep_fwd: descriptor %d is empty", i);
        if (d.layout == HDMOE_WLAYOUT_TAPS || (d.w_hat2 && d.layout2 == HDMOE_WLAYOUT_TAPS))
            HDMOE_CHECK_ARG(d.cin * d.taps == d.fan_in && d.cin_pad >= d.cin, "wprep_fwd: descriptor %d: bad tap layout", i);
        if (d.layout == HDMOE_WLAYOUT_TAPS_T || (d.w_hat2 && d.layout2 == HDMOE_WLAYOUT_TAPS_T))
            HDMOE_CHECK_ARG(d.cin * d.taps == d.fan_in && d.cin_rows >= 1 && d.cin_rows <= d.cin && d.cout_pad >= d.rows,
                            "wprep_fwd: descriptor %d: bad transposed tap layout", i);
        d.block_start = total;
        total += d.rows;
    }
    HDMOE_CHECK_CUDA(cudaMemcpyAsync(descs_dev, descs_host, (size_t)n * sizeof(hdmoe_wprep_desc), cudaMemcpyHostToDevice, st));
    wprep_fwd_kernel<<<total, kWprepThreads, 0, st>>>((const hdmoe_wprep_desc*)descs_dev, n, force);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_wprep_fwd_resident(const void* descs_dev, int n, int total_rows, int force, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(descs_dev && n >= 1 && total_rows >= 1, "wprep_fwd_resident: bad table");
    wprep_fwd_kernel<<<total_rows, kWprepThreads, 0, (cudaStream_t)stream>>>((const hdmoe_wprep_desc*)descs_dev, n, force);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_wprep_bwd_multi_resident(const void* descs_dev, int n, int total_rows, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(descs_dev && n >= 1 && total_rows >= 1, "wprep_bwd_multi_resident: bad table");
    wprep_bwd_multi_kernel<<<total_rows, kWprepThreads, 0, (cudaStream_t)stream>>>((const hdmoe_wprep_bwd_desc*)descs_dev, n);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_wprep_bwd(const float* w, const float* d_w_hat, const float* gain_ptr, float gain, int rows,
                               int fan_in, float* d_w, float* d_gain, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(w && d_w_hat && d_w && rows >= 1 && fan_in >= 1, "wprep_bwd: bad args");
    wprep_bwd_kernel<<<rows, kWprepThreads, 0, (cudaStream_t)stream>>>(w, d_w_hat, gain_ptr, gain, rows, fan_in, d_w, d_gain);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
