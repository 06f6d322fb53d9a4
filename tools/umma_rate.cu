// Hardware probe: sustained tcgen05.mma issue / execution rate for small-N tiles on B200.
// One CTA, `issuers` warps each issue `iters` x 4 MMAs (M=128, K=16, N in {32,64,128,256}) on smem operands that
// are already resident (garbage data is fine), into `naccs` rotating accumulators.  Reports cycles per MMA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s2u(b)), "r"(parity) : "memory");
        if (spin > (1u << 26)) { printf("rate: wait timeout\n"); __trap(); }
    }
}
template <int N, int M = 128, int MN = 0>
__global__ void __launch_bounds__(128) rate_kernel(int issuers, int iters, int naccs, int same_acc_run, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    long long t0 = 0, t1 = 0;
    if (warp < issuers && lane == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24) | (MN ? ((1u << 15) | (1u << 16)) : 0u);
        // K-major: LBO field 1 (unused), SBO = 1024 B; MN-major: LBO = MN (in 16-byte units: 8 = next 128-byte row), SBO = 1024 B
        const uint64_t base = ((uint64_t)(MN ? (MN == 1 ? 0 : MN) : 1) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const int kstep = MN ? 128 : 2;     // descriptor units (16 B) per K = 16 slice
        const uint64_t ad0 = base | ((s2u(smem) & 0x3FFFF) >> 4);
        const uint64_t bd0 = base | ((s2u(smem + 32 * 1024) & 0x3FFFF) >> 4);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int a = same_acc_run ? (warp * naccs + (it % naccs)) : 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int acc = same_acc_run ? a : (warp * naccs + ((it * 4 + kk) % naccs));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (uint32_t)(acc * N)), "l"(ad0 + kstep * kk), "l"(bd0 + kstep * kk), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(&bar[warp])) : "memory");
        mb_wait(&bar[warp], 0);
        t1 = clock64();
        out[warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int N, int M = 128, int MN = 0>
int run(int issuers, int naccs, int same) {
    long long* d;
    CK(cudaMalloc(&d, 4 * sizeof(long long)));
    CK(cudaMemset(d, 0, 4 * sizeof(long long)));
    const int iters = 2000, smem = 65 * 1024 + 1024;
    CK(cudaFuncSetAttribute(rate_kernel<N, M, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    rate_kernel<N, M, MN><<<1, 128, smem>>>(issuers, iters, naccs, same, d);
    CK(cudaDeviceSynchronize());
    long long h[4];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < issuers; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / (iters * 4.0 * issuers);
    printf("%s M=%3d N=%3d issuers=%d accs/issuer=%d %s: %7.1f cycles per MMA (ideal %d) -> %.0f%% of tensor peak\n", MN ? (MN == 1 ? "MN-major      " : "MN-major LBO=r") : "K-major       ", M, N, issuers, naccs,
           same ? "4 k-slices per acc" : "rotating acc     ", per, N / 2, 100.0 * (N / 2) * (M / 128.0) / per);
    cudaFree(d);
    return 0;
}
int main() {
    for (int issuers : {1, 2, 3, 4}) {
        run<64>(issuers, 1, 1);
        if (issuers * 2 * 64 <= 512) run<64>(issuers, 2, 1);
    }
    run<64>(1, 4, 0);
    run<32>(1, 1, 1); run<32>(3, 1, 1); run<32>(4, 2, 1);
    run<128>(1, 1, 1); run<128>(2, 1, 1); run<128>(3, 1, 1);
    run<256>(1, 1, 1); run<256>(2, 1, 1);
    run<64, 64>(1, 1, 1); run<64, 64>(3, 1, 1); run<128, 64>(1, 1, 1); run<128, 64>(3, 1, 1); run<256, 64>(1, 1, 1); run<256, 64>(2, 1, 1);
    run<64, 64, 1>(1, 1, 1); run<64, 64, 1>(3, 1, 1); run<64, 64, 1>(4, 1, 1); run<64, 128, 8>(2, 1, 1); run<64, 128, 8>(3, 1, 1); run<64, 128, 8>(4, 1, 1); run<32, 128, 4>(3, 1, 1); run<32, 64, 1>(3, 1, 1); run<128, 64, 1>(3, 1, 1); run<128, 128, 8>(3, 1, 1);
    run<160>(1, 1, 1); run<160>(2, 1, 1); run<160>(3, 1, 1); run<192>(1,1,1); run<192>(2,1,1);
    return 0;
}
