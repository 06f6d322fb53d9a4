// Peer-memory exchange for expert parallelism over NVLink / NVSwitch: the equal-split all-to-all of the
// dispatch / combine row buffers and the all-gather of the per-expert counts as plain CUDA kernels over peer-mapped
// buffers (CUDA IPC, one process per GPU), so the expert-parallel step records into a CUDA graph like the data-parallel
// one.  (NCCL all_to_all_single captured in the step's graph deadlocked on replay on this stack; these kernels have no
// host-side state at all.)
//
//   peer_barrier:  every rank writes its epoch into slot [rank] of every peer's flag array (release, system scope) and
//                  spins until its own array shows that epoch from every peer (acquire).  Epochs only grow; the counter
//                  lives on the device and is advanced by the kernel itself, so a graph replay needs no host input.
//                  All ranks must issue the same sequence of barriers (the expert-parallel layer is symmetric).
//   peer_pull:     dst segment g  <-  peer g's staging buffer, segment `rank` (all-to-all) or the whole buffer
//                  (all-gather), 16-byte vectors over NVLink, 4 loads in flight per thread.
#include "common.cuh"

namespace hdmoe {

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// epoch[0] = barriers passed, epoch[1] = sticky failure flag: a peer that does not arrive within `timeout_ns` (it died,
// or the ranks issued different sequences) marks the group dead instead of hanging the GPU; later barriers return at
// once and the host side raises on its next check (peer.PeerGroup.check()).
__global__ void __launch_bounds__(64)
peer_barrier_kernel(int32_t* __restrict__ my_flags, const long long* __restrict__ peer_flags, int32_t* __restrict__ epoch,
                    int rank, int G, unsigned long long timeout_ns) {
    __shared__ int e_s, dead_s;
    if (threadIdx.x == 0) {
        e_s = epoch[0] + 1;
        dead_s = epoch[1];
    }
    __syncthreads();
    if (dead_s) return;
    const int e = e_s;
    const int g = threadIdx.x;
    if (g < G) {
        // kernels issued earlier on this stream have completed; make their writes visible at system scope before the signal
        __threadfence_system();
        int32_t* remote = reinterpret_cast<int32_t*>(peer_flags[g]) + rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
        const unsigned long long t0 = global_ns();
        int v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(my_flags + g) : "memory");
            if (v >= e) break;
            if (global_ns() - t0 > timeout_ns) {
                atomicExch(&dead_s, 1);
                break;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch[0] = e;
        if (dead_s) epoch[1] = 1;
    }
}

__global__ void __launch_bounds__(256)
peer_pull_kernel(int4* __restrict__ dst, const long long* __restrict__ peer_src, long long src_seg16, long long seg16, int G) {
    const long long total = seg16 * G;
    const long long stride = (long long)gridDim.x * 256;
    for (long long i0 = (long long)blockIdx.x * 256 + threadIdx.x; i0 < total; i0 += 4 * stride) {
        int4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = i0 + q * stride;
            if (i < total) {
                const long long g = i / seg16, o = i - g * seg16;
                v[q] = reinterpret_cast<const int4*>(peer_src[g])[src_seg16 + o];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = i0 + q * stride;
            if (i < total) dst[i] = v[q];
        }
    }
}
}  // namespace hdmoe

extern "C" int hdmoe_peer_barrier(int32_t* my_flags, const int64_t* peer_flag_ptrs_dev, int32_t* epoch_dev, int rank,
                                  int world, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(my_flags && peer_flag_ptrs_dev && epoch_dev, "peer_barrier: null pointer");
    HDMOE_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "peer_barrier: 1 <= world <= 64, 0 <= rank < world");
    hdmoe::peer_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(my_flags, (const long long*)peer_flag_ptrs_dev, epoch_dev,
                                                                  rank, world, 8000000000ull /* 8 s */);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_peer_pull(void* dst, const int64_t* peer_src_ptrs_dev, int64_t seg_bytes, int rank_segment, int world,
                               hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(dst && peer_src_ptrs_dev && seg_bytes > 0 && seg_bytes % 16 == 0, "peer_pull: null pointer or seg_bytes %% 16 != 0");
    HDMOE_CHECK_ARG(world >= 1 && world <= 64 && rank_segment >= -1 && rank_segment < world, "peer_pull: bad world / segment");
    HDMOE_CHECK_ARG(((uintptr_t)dst & 15) == 0, "peer_pull: dst must be 16-byte aligned");
    const long long seg16 = seg_bytes / 16;
    const long long src_seg16 = rank_segment < 0 ? 0 : seg16 * rank_segment;
    hdmoe::peer_pull_kernel<<<hdmoe::grid_for(seg16 * world, 1024, 8), 256, 0, (cudaStream_t)stream>>>(
        (int4*)dst, (const long long*)peer_src_ptrs_dev, src_seg16, seg16, world);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
