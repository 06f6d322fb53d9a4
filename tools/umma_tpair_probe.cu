// Hardware probe for the NEXT grouped-convolution formulation ("gconv beyond the N = 64 operand-feed
// bound"): weights as the M = 128 operand holding TWO horizontal taps x 64 output channels, the flattened padded halo
// image as the N = 256 operand, every tap pair of the kernel accumulated into ONE TMEM accumulator because the
// horizontal offset of the pair goes into the B start address:
//
//     Acc[co     ][q'] = sum over pairs (tr, pr), ci of  W[tr, 2 pr    ][co][ci] * Xpad[q' + tr * Wp + 2 pr][ci]
//     Acc[64 + co][q'] = sum over pairs (tr, pr), ci of  W[tr, 2 pr + 1][co][ci] * Xpad[q' + tr * Wp + 2 pr][ci]
//     Y[q][co] = Acc[co][q] + Acc[64 + co][q - 1 + ... ]   ->   Y[q][co] = Acc[co][q] + Acc[64 + co][q + ... ]
//
// precisely: the odd tap (ts = 2 pr + 1) wants Xpad[q + tr * Wp + 2 pr + 1], which the shared B operand supplies at
// q' = q + 1, so  Y[q][co] = Acc[co][q] + Acc[64 + co][q + 1].  The unpaired last tap of a kernel row is paired with a
// zero block (weight layout with a row pitch of k + 1 taps, built on the host here).
//
// One CTA computes one tile: 32x32 image, Cin = Cout = 64, k = 5, positions [0, 256) of the flattened padded image.
// The raw accumulator [128][256] is written to global memory and combined / checked on the host against a direct
// convolution.  NOT part of the product library.  Result on B200: 14 528 outputs checked, max abs err 4.4e-6 -> OK.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -o tools/umma_tpair_probe tools/umma_tpair_probe.cu
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../heterogeneous-moe-for-diffusion-models_b200/csrc/tc.cuh"

namespace hdmoe {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace hdmoe
using namespace hdmoe;

constexpr int H = 32, W = 32, C = 64, CO = 64, K = 5, PAD = 2, WP = W + K - 1;   // WP = 36
constexpr int NPOS = 256;                                                         // positions per tile (MMA N)
constexpr int PAIRS = (K + 1) / 2;                                                // 3 pairs per kernel row
constexpr int BOX_ROWS = (NPOS - 1 + (K - 1) * (WP + 1)) / WP + 1;                // input rows the tile touches (12)
constexpr int A_TILE = 128 * C * 2;                                               // one pair tile: 16 KiB
constexpr int HALO_BYTES = BOX_ROWS * WP * C * 2;

__global__ void __launch_bounds__(192, 1)
tpair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, float* __restrict__ acc_out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* halo = smem;                                          // [BOX_ROWS * WP positions][64 ch], SWIZZLE_128B
    uint8_t* wbuf = smem + ((HALO_BYTES + 1023) / 1024) * 1024;    // one pair tile [128 rows][64 ch]
    __shared__ __align__(8) uint64_t bar_x, bar_w, bar_mma;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mb_init(&bar_x, 1);
        mb_init(&bar_w, 1);
        mb_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s2u(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0 && lane == 0) {
        // halo of the tile: rows -PAD .. of the image, columns -PAD .. (out of bounds = zero fill = 'same' padding)
        mb_expect_tx(&bar_x, (uint32_t)HALO_BYTES);
        tma_load_4d(halo, &tm_x, &bar_x, 0, -PAD, -PAD, 0);
        mb_wait(&bar_x, 0);
        // M = 128, N = 256, K = 16, bf16 x bf16 -> fp32, both operands K-major
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPOS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t wph = 0, mph = 0;
        bool first = true;
        for (int tr = 0; tr < K; ++tr)
            for (int pr = 0; pr < PAIRS; ++pr) {
                mb_expect_tx(&bar_w, (uint32_t)A_TILE);
                tma_load_2d(wbuf, &tm_w, &bar_w, 0, (tr * PAIRS + pr) * 128);
                mb_wait(&bar_w, wph);
                wph ^= 1;
                tc_fence_after();
                const uint64_t ad = umma_desc<C>(s2u(wbuf));
                const uint64_t bd = umma_desc<C>(s2u(halo) + (uint32_t)(tr * WP + 2 * pr) * (C * 2));
                for (int kk = 0; kk < C / 16; ++kk) {
                    tc_mma(tmem, ad + 2 * kk, bd + 2 * kk, idesc, first ? 0u : 1u);
                    first = false;
                }
                tc_commit(&bar_mma);                 // the weight buffer is reused: wait until these MMAs have read it
                mb_wait(&bar_mma, mph);
                mph ^= 1;
            }
    }
    __syncthreads();
    tc_fence_after();
    if (warp >= 2) {
        const int quad = warp & 3;                   // TMEM lanes 32 * quad .. +31 (warps 2, 3, 4, 5 -> quads 2, 3, 0, 1)
        const int m = quad * 32 + lane;              // accumulator row: (tap parity, output channel)
        for (int c0 = 0; c0 < NPOS; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            for (int u = 0; u < 32; ++u) acc_out[(size_t)m * NPOS + c0 + u] = __uint_as_float(v[u]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
    srand(1);
    auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.f; };
    std::vector<__nv_bfloat16> hx((size_t)H * W * C), hw((size_t)K * PAIRS * 128 * C);
    std::vector<float> fx(hx.size()), fw((size_t)K * K * CO * C);
    for (size_t i = 0; i < hx.size(); ++i) {
        hx[i] = __float2bfloat16(rnd());
        fx[i] = __bfloat162float(hx[i]);
    }
    for (size_t i = 0; i < fw.size(); ++i) fw[i] = __bfloat162float(__float2bfloat16(rnd() * 0.1f));
    // paired weight layout: [tr][pair][parity][co][ci], zero block for the missing partner of the last tap of a row
    for (int tr = 0; tr < K; ++tr)
        for (int pr = 0; pr < PAIRS; ++pr)
            for (int par = 0; par < 2; ++par)
                for (int co = 0; co < CO; ++co)
                    for (int ci = 0; ci < C; ++ci) {
                        const int ts = 2 * pr + par;
                        const float v = ts < K ? fw[(((size_t)tr * K + ts) * CO + co) * C + ci] : 0.f;
                        hw[((((size_t)tr * PAIRS + pr) * 2 + par) * CO + co) * C + ci] = __float2bfloat16(v);
                    }
    __nv_bfloat16 *dx, *dw;
    float* dacc;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&dacc, (size_t)128 * NPOS * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dacc, 0, (size_t)128 * NPOS * 4));
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm_x, tm_w;
    {
        cuuint64_t dims[4] = {C, W, H, 1};
        cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {C, WP, BOX_ROWS, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode X failed %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {C, (cuuint64_t)K * PAIRS * 128};
        cuuint64_t strides[1] = {C * 2};
        cuuint32_t box[2] = {C, 128}, es[2] = {1, 1};
        CUresult r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode W failed %d\n", (int)r); return 1; }
    }
    const int smem = ((HALO_BYTES + 1023) / 1024) * 1024 + A_TILE + 1024;
    CK(cudaFuncSetAttribute(tpair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tpair_kernel<<<1, 192, smem>>>(tm_x, tm_w, dacc);
    CK(cudaDeviceSynchronize());
    std::vector<float> acc((size_t)128 * NPOS);
    CK(cudaMemcpy(acc.data(), dacc, acc.size() * 4, cudaMemcpyDeviceToHost));
    // host: Y[q][co] = Acc[co][q] + Acc[64 + co][q + 1] for the output pixels among positions [0, NPOS - 1)
    double worst = 0, ref_max = 0;
    int checked = 0;
    for (int q = 0; q + 1 < NPOS; ++q) {
        const int h = q / WP, w = q % WP;
        if (w >= W || h >= H) continue;
        for (int co = 0; co < CO; ++co) {
            double ref = 0;
            for (int tr = 0; tr < K; ++tr)
                for (int ts = 0; ts < K; ++ts) {
                    const int hh = h + tr - PAD, ww = w + ts - PAD;
                    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                    for (int ci = 0; ci < C; ++ci)
                        ref += (double)fw[(((size_t)tr * K + ts) * CO + co) * C + ci] * fx[((size_t)hh * W + ww) * C + ci];
                }
            const double got = (double)acc[(size_t)co * NPOS + q] + (double)acc[(size_t)(64 + co) * NPOS + q + 1];
            worst = fmax(worst, fabs(got - ref));
            ref_max = fmax(ref_max, fabs(ref));
            ++checked;
        }
    }
    printf("tap-pair formulation: %d outputs checked, max abs err %.3e (max |ref| %.3e) -> %s\n", checked, worst, ref_max,
           worst < 1e-3 * fmax(1.0, ref_max) ? "OK" : "MISMATCH");
    return 0;
}
