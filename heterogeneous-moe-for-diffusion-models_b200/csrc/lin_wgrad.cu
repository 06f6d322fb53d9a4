// Weight gradient of the thin trunk projections: dW[32][32] = dY^T X with dY, X [rows][32] fp32 and rows = B*S
// (262 144 at batch 256, 32x32 latents).  These are the 1x1 q / k / v / out projections of the trunk attention
// (MP_Attention._proj, models/model_internals.py:364-372,407); autograd computes the gradient as
// mm([32, rows], [rows, 32]) and the library picks a 64x64-tile kernel without split-K: ~120 us per call, six calls
// on the single-stream attention backward.  The product reads 2 * rows * 128 B once and does 2 * rows * 1024 flops:
// HBM-bound (67 MB, ~10 us), so: persistent CTAs stream 64-row tiles through shared memory, every thread keeps a 4x4
// register tile of dW (64 threads cover 32x32, the four thread quarters split the rows of a tile), the quarters are
// summed in shared memory and each CTA adds its 32x32 partial with vector atomics.
#include "common.cuh"

namespace hdmoe {

constexpr int kLwThreads = 256;
constexpr int kLwTile = 64;          // rows per shared-memory tile

__global__ void __launch_bounds__(kLwThreads)
lin32_wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ dW, long long rows) {
    __shared__ __align__(16) float sy[2][kLwTile * 32], sx[2][kLwTile * 32];
    __shared__ __align__(16) float part[4][32 * 32];
    const int tid = threadIdx.x, quarter = tid >> 6, q = tid & 63, ni = q >> 3, ki = q & 7;
    const long long n_tiles = (rows + kLwTile - 1) / kLwTile;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    // a tile is 64 rows x 32 floats = 512 float4 per matrix: two per thread and matrix
    int buf = 0;
    long long t = blockIdx.x;
    float4 ry[2], rx[2];
    auto fetch = [&](long long tt) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int f = tid + u * kLwThreads;
            const long long row = tt * kLwTile + (f >> 3);
            if (row < rows) {
                ry[u] = __ldg(reinterpret_cast<const float4*>(dY + row * 32) + (f & 7));
                rx[u] = __ldg(reinterpret_cast<const float4*>(X + row * 32) + (f & 7));
            } else {
                ry[u] = rx[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (t < n_tiles) fetch(t);
    for (; t < n_tiles; t += gridDim.x) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int f = tid + u * kLwThreads;
            reinterpret_cast<float4*>(sy[buf])[f] = ry[u];
            reinterpret_cast<float4*>(sx[buf])[f] = rx[u];
        }
        __syncthreads();                                        // tile visible; the other buffer is free again
        const long long nt = t + gridDim.x;
        if (nt < n_tiles) fetch(nt);                            // next tile's loads fly under this tile's FMAs
        const float* y = sy[buf] + quarter * 16 * 32;
        const float* x = sx[buf] + quarter * 16 * 32;
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(y + r * 32 + 4 * ni);
            const float4 b = *reinterpret_cast<const float4*>(x + r * 32 + 4 * ki);
            acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
            acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
            acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
            acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
        }
        buf ^= 1;
    }
    // sum the four row quarters, then one vector atomic per float4 of the CTA's partial
#pragma unroll
    for (int a = 0; a < 4; ++a)
        *reinterpret_cast<float4*>(&part[quarter][(4 * ni + a) * 32 + 4 * ki]) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    __syncthreads();
    {
        const float4 p0 = reinterpret_cast<const float4*>(part[0])[tid], p1 = reinterpret_cast<const float4*>(part[1])[tid];
        const float4 p2 = reinterpret_cast<const float4*>(part[2])[tid], p3 = reinterpret_cast<const float4*>(part[3])[tid];
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dW + 4 * tid),
                     "f"(p0.x + p1.x + p2.x + p3.x), "f"(p0.y + p1.y + p2.y + p3.y), "f"(p0.z + p1.z + p2.z + p3.z),
                     "f"(p0.w + p1.w + p2.w + p3.w)
                     : "memory");
    }
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_lin32_wgrad(const float* dY, const float* X, float* dW, int64_t rows, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(dY && X && dW && rows >= 1, "lin32_wgrad: null pointer / no rows");
    HDMOE_CHECK_ARG((((uintptr_t)dY | (uintptr_t)X | (uintptr_t)dW) & 15) == 0, "lin32_wgrad: 16-byte alignment required");
    const long long n_tiles = (rows + kLwTile - 1) / kLwTile;
    const int grid = (int)(n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs);
    lin32_wgrad_kernel<<<grid, kLwThreads, 0, (cudaStream_t)stream>>>(dY, X, dW, (long long)rows);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
