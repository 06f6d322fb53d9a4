// Shared device/host helpers for libhdmoe_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hdmoe_b200.h"

namespace hdmoe {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
constexpr int kSchedSlots = 256;
int32_t* sched_slot(cudaStream_t st);   // per-(device, stream) {next tile, finished CTAs} counter pair, see core.cu

#define HDMOE_CHECK_ARG(cond, ...)                \
    do {                                          \
        if (!(cond)) {                            \
            ::hdmoe::set_error(__VA_ARGS__);      \
            return HDMOE_ERR_ARG;                 \
        }                                         \
    } while (0)

#define HDMOE_CHECK_CUDA(expr)                                                                   \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::hdmoe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                        \
            return HDMOE_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

#define HDMOE_CHECK_LAUNCH()                       \
    do {                                           \
        ::hdmoe::count_launch();                   \
        HDMOE_CHECK_CUDA(cudaPeekAtLastError());   \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 128-bit streaming accesses: read-once / write-once payloads should not pollute L1.
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w));
}

// dtype-generic 4-element vector load/store in fp32 registers
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
        uint2 u = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = u;
    }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
    int64_t b = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)kNumSMs * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace hdmoe
