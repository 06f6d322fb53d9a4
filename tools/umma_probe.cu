// Hardware probe (bring-up tool, not part of the library): UMMA shared-memory descriptor semantics needed by
// the halo-reuse convolution and the weight-gradient kernel.
//   test 1: K-major SW128 A operand whose start address is shifted by `shift` rows (128 B each) inside a larger
//           TMA-written buffer, with base_offset = 0 and base_offset = (addr >> 7) & 7.
//   test 2: MN-major SW128 operands (A^T and B^T stored [K][64]), B shifted by `shift` K-rows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s2u(b)), "r"(parity) : "memory");
        if (spin > (1u << 24)) { printf("probe: wait timeout\n"); __trap(); }
    }
}

// mode 0: K-major A (rows = M, 128 B rows), K-major B.  mode 1: MN-major A and B ([K][64] tiles).
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                    int mode, int shift, int use_base_offset, int a_rows, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // a_rows x 128 B
    uint8_t* sb = smem + 64 * 1024;     // up to 256 x 128 B
    __shared__ __align__(8) uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s2u(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const int b_rows = mode == 0 ? 64 : a_rows;
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(a_rows + b_rows) * 128;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(&bar_full)), "r"(bytes) : "memory");
        for (int r0 = 0; r0 < a_rows; r0 += 64)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s2u(sa + r0 * 128)), "l"(&ta), "r"(s2u(&bar_full)), "r"(0), "r"(r0) : "memory");
        for (int r0 = 0; r0 < b_rows; r0 += 64)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s2u(sb + r0 * 128)), "l"(&tb), "r"(s2u(&bar_full)), "r"(0), "r"(r0) : "memory");
        mb_wait(&bar_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        auto desc = [&](uint32_t addr, bool with_bo) {
            uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            if (with_bo) d |= (uint64_t)((addr >> 7) & 7) << 49;
            return d;
        };
        if (mode == 0) {
            // D[128 x 64] = A[shift : shift+128][0:64] * B[0:64][0:64]^T
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            for (int kk = 0; kk < 4; ++kk) {
                uint64_t ad = desc(s2u(sa) + shift * 128 + kk * 32, use_base_offset);
                uint64_t bd = desc(s2u(sb) + kk * 32, false);
                uint32_t acc = kk > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
        } else {
            // D[64 x 64] = sum_k At[k][m] * Bt[k + shift][n],  k = 0..127; both operands MN-major (bits 15,16 set)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
            for (int kk = 0; kk < 8; ++kk) {   // 16 K-rows (= 2 groups of 8 rows x 128 B) per MMA
                uint64_t ad = desc(s2u(sa) + kk * 16 * 128, false);
                uint64_t bd = desc(s2u(sb) + (kk * 16 + shift) * 128, use_base_offset);
                uint32_t acc = kk > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(&bar_mma)) : "memory");
    }
    __syncwarp();
    mb_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its lane quadrant: 32 lanes x 64 columns
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 8; ++q) out[(warp * 32 + lane) * 64 + c0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    const int ROWS = 320;
    std::vector<__nv_bfloat16> hA(ROWS * 64), hB(ROWS * 64);
    std::vector<float> fA(ROWS * 64), fB(ROWS * 64);
    srand(1);
    for (int i = 0; i < ROWS * 64; ++i) {
        fA[i] = (float)((rand() % 17) - 8) / 8.f;
        fB[i] = (float)((rand() % 13) - 6) / 4.f;
        hA[i] = __float2bfloat16(fA[i]);
        hB[i] = __float2bfloat16(fB[i]);
    }
    __nv_bfloat16 *dA, *dB;
    float* dO;
    CK(cudaMalloc(&dA, ROWS * 64 * 2));
    CK(cudaMalloc(&dB, ROWS * 64 * 2));
    CK(cudaMalloc(&dO, 128 * 64 * 4));
    CK(cudaMemcpy(dA, hA.data(), ROWS * 64 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), ROWS * 64 * 2, cudaMemcpyHostToDevice));
    CUtensorMap ta, tb;
    cuuint64_t dims[2] = {64, (cuuint64_t)ROWS};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
        enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
        printf("encode failed\n");
        return 1;
    }
    const int smem = 64 * 1024 + 48 * 1024 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> hO(128 * 64);
    for (int mode = 0; mode < 2; ++mode)
        for (int bo = 0; bo < 2; ++bo)
            for (int shift : {0, 1, 2, 3, 5, 8, 9, 36, 41}) {
                CK(cudaMemset(dO, 0, 128 * 64 * 4));
                probe_kernel<<<1, 128, smem>>>(ta, tb, mode, shift, bo, 256, dO);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d bo %d shift %d: launch error %s\n", mode, bo, shift, cudaGetErrorString(e)); return 2; }
                CK(cudaMemcpy(hO.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                if (mode == 0) {
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 64; ++n) {
                            double s = 0;
                            for (int k = 0; k < 64; ++k) s += (double)fA[(m + shift) * 64 + k] * fB[n * 64 + k];
                            maxerr = fmax(maxerr, fabs(s - hO[m * 64 + n]));
                        }
                } else {
                    double e_lin = 0, e_q16 = 0;   // candidate TMEM row mappings for M = 64
                    for (int m = 0; m < 64; ++m)
                        for (int n = 0; n < 64; ++n) {
                            double s = 0;
                            for (int k = 0; k < 128; ++k) s += (double)fA[k * 64 + m] * fB[(k + shift) * 64 + n];
                            e_lin = fmax(e_lin, fabs(s - hO[m * 64 + n]));
                            e_q16 = fmax(e_q16, fabs(s - hO[((m / 16) * 32 + (m % 16)) * 64 + n]));
                        }
                    maxerr = fmin(e_lin, e_q16);
                    printf("   [M=64 lane map: linear err %.3f, 16-per-quadrant err %.3f] ", e_lin, e_q16);
                }
                printf("mode %d (%s) base_offset=%d shift=%2d : max abs err %.4f %s\n", mode, mode ? "MN-major" : "K-major", bo,
                       shift, maxerr, maxerr < 1e-2 ? "OK" : "WRONG");
            }
    return 0;
}
