// Optimizer side of the train step (Utils/training.py:195-197 = clip_grad_norm_(1.0) + AdamW.step
// over all ~560 parameter tensors): three launches for the whole model instead of the library's foreach norm /
// scale / multi-tensor chain (~1 ms per step at 9 M parameters, most of it launch count and extra passes).
//
//   1. mt_sumsq_kernel     one CTA per chunk of <= kOptChunk elements: sum of g^2 -> partial[chunk]      (reads g)
//   2. mt_finalize_kernel  one CTA: total = sqrt(sum of partials) in a fixed order (deterministic), clip coefficient
//                          min(1, max_norm / (total + 1e-6)) (torch.nn.utils.clip_grad_norm_), step counter += 1
//   3. mt_adamw_kernel     one CTA per chunk: g *= clip (optionally written back, as clip_grad_norm_ does in place),
//                          decoupled weight decay, moment updates, bias-corrected update  (reads p g m v, writes p m v)
//
// Tensors are described by a device-resident table (pointer-stable buffers replay inside a CUDA graph; the host refreshes
// the table only when a pointer moved).  HBM-bound: 28 B per parameter in pass 3, 4 B in pass 1.
#include "common.cuh"

namespace hdmoe {

constexpr int kOptChunk = 8192;      // elements per CTA
constexpr int kOptThreads = 256;

struct OptTensor {                   // mirrors _lib.OptTensorDesc
    float* p;
    float* g;                        // may be NULL: parameter without a gradient this step (skipped, like torch)
    float* m;
    float* v;
    int64_t numel;
    float lr, weight_decay;
    int32_t chunk_start, pad;        // first chunk id of this tensor
};

__device__ __forceinline__ int find_tensor(const OptTensor* __restrict__ t, int n, int chunk) {
    int lo = 0, hi = n - 1;          // last tensor with chunk_start <= chunk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid].chunk_start <= chunk) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ float block_sum_f(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = l < (kOptThreads >> 5) ? red[l] : 0.f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kOptThreads)
mt_sumsq_kernel(const OptTensor* __restrict__ tab, int n, float* __restrict__ partial) {
    __shared__ float red[32];
    const int chunk = blockIdx.x;
    const OptTensor t = tab[find_tensor(tab, n, chunk)];
    float s = 0.f;
    if (t.g) {
        const int64_t lo = (int64_t)(chunk - t.chunk_start) * kOptChunk;
        const int64_t hi = min(lo + (int64_t)kOptChunk, t.numel);
        const float* g = t.g;
        if ((((uintptr_t)g) & 15) == 0 && (lo & 3) == 0) {
            const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
            for (int64_t i = lo + 4 * threadIdx.x; i < hi4; i += 4 * kOptThreads) {
                const float4 q = *reinterpret_cast<const float4*>(g + i);
                s += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
            }
            for (int64_t i = hi4 + threadIdx.x; i < hi; i += kOptThreads) s += g[i] * g[i];
        } else {
            for (int64_t i = lo + threadIdx.x; i < hi; i += kOptThreads) s += g[i] * g[i];
        }
    }
    s = block_sum_f(s, red);
    if (threadIdx.x == 0) partial[chunk] = s;
}

// state[0] = number of step() calls, state[1] = total gradient norm, state[2] = clip coefficient; steps[i] = number of
// updates tensor i received (torch.optim.AdamW counts steps per tensor: a parameter without gradient does not advance)
__global__ void __launch_bounds__(1024)
mt_finalize_kernel(const OptTensor* __restrict__ tab, int n, const float* __restrict__ partial, int n_chunks, float max_norm,
                   float* __restrict__ state, float* __restrict__ steps) {
    __shared__ double red[32];
    for (int i = threadIdx.x; i < n; i += 1024)
        if (tab[i].g) steps[i] += 1.f;
    double s = 0.0;
    for (int i = threadIdx.x; i < n_chunks; i += 1024) s += (double)partial[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = red[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            const float total = (float)sqrt(t);
            state[0] += 1.f;
            state[1] = total;
            float c = 1.f;
            if (max_norm > 0.f) c = fminf(max_norm / (total + 1e-6f), 1.f);
            state[2] = c;
        }
    }
}

__global__ void __launch_bounds__(kOptThreads)
mt_adamw_kernel(const OptTensor* __restrict__ tab, int n, const float* __restrict__ state, const float* __restrict__ steps,
                float beta1, float beta2, float eps, int write_back_grad) {
    const int chunk = blockIdx.x;
    const int ti = find_tensor(tab, n, chunk);
    const OptTensor t = tab[ti];
    if (!t.g) return;
    const float step = steps[ti], clip = state[2];
    const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
    const float step_size = t.lr / bc1, inv_bc2_sqrt = rsqrtf(bc2);
    const float decay = 1.f - t.lr * t.weight_decay;
    const int64_t lo = (int64_t)(chunk - t.chunk_start) * kOptChunk;
    const int64_t hi = min(lo + (int64_t)kOptChunk, t.numel);
    auto upd = [&](float& p, float& g, float& m, float& v) {
        g *= clip;
        p *= decay;
        m = beta1 * m + (1.f - beta1) * g;            // torch: exp_avg.lerp_(grad, 1 - beta1)
        v = beta2 * v + (1.f - beta2) * g * g;
        const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
        p -= step_size * (m / denom);
    };
    const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0) && (lo & 3) == 0;
    int64_t i0 = lo;
    if (vec) {
        const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
        for (int64_t i = lo + 4 * threadIdx.x; i < hi4; i += 4 * kOptThreads) {
            float4 p = *reinterpret_cast<float4*>(t.p + i), g = *reinterpret_cast<const float4*>(t.g + i);
            float4 m = *reinterpret_cast<float4*>(t.m + i), v = *reinterpret_cast<float4*>(t.v + i);
            upd(p.x, g.x, m.x, v.x);
            upd(p.y, g.y, m.y, v.y);
            upd(p.z, g.z, m.z, v.z);
            upd(p.w, g.w, m.w, v.w);
            *reinterpret_cast<float4*>(t.p + i) = p;
            *reinterpret_cast<float4*>(t.m + i) = m;
            *reinterpret_cast<float4*>(t.v + i) = v;
            if (write_back_grad) *reinterpret_cast<float4*>(t.g + i) = g;
        }
        i0 = hi4;
    }
    for (int64_t i = i0 + threadIdx.x; i < hi; i += kOptThreads) {
        float p = t.p[i], g = t.g[i], m = t.m[i], v = t.v[i];
        upd(p, g, m, v);
        t.p[i] = p;
        t.m[i] = m;
        t.v[i] = v;
        if (write_back_grad) t.g[i] = g;
    }
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_optim_chunk_elems(void) { return kOptChunk; }

extern "C" int hdmoe_adamw_step(const void* table_dev, int n_tensors, int n_chunks, float* partial, float* state,
                                float* steps, float max_norm, float beta1, float beta2, float eps, int write_back_grad,
                                hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(table_dev && partial && state && steps && n_tensors >= 1 && n_chunks >= 1, "adamw_step: bad args");
    HDMOE_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adamw_step: bad hyper-parameters");
    cudaStream_t st = (cudaStream_t)stream;
    const OptTensor* tab = (const OptTensor*)table_dev;
    mt_sumsq_kernel<<<n_chunks, kOptThreads, 0, st>>>(tab, n_tensors, partial);
    HDMOE_CHECK_LAUNCH();
    mt_finalize_kernel<<<1, 1024, 0, st>>>(tab, n_tensors, partial, n_chunks, max_norm, state, steps);
    HDMOE_CHECK_LAUNCH();
    mt_adamw_kernel<<<n_chunks, kOptThreads, 0, st>>>(tab, n_tensors, state, steps, beta1, beta2, eps, write_back_grad);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
