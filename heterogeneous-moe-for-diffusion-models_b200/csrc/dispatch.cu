// PERMUTE / COMBINE: bit-exact dispatch-plan build, HBM row gather, gate-weighted combine and their
// backward.  Replaces the per-expert boolean-mask loop of router_to_unet_experts
// (models/model_config2.py:23-37): nonzero()/index/index_put_ x E with >= 2E host syncs become a plan
// (integer, bit-exact) plus one gather and one combine launch, no host sync.
//
// Roofline: HBM.  Algorithmic bytes: permute R*rowbytes read + written (+4 B index per
// row); combine reads R*D*s + R*8 and writes T*D*s.
#include "common.cuh"

namespace hdmoe {

// ---------------------------------------------------------------------------------------------------
// Dispatch plan.  Stable counting sort of the (token, expert) pairs with sparse_w > 0 by expert.
//   pass 1 (count)   : per tile (G groups of 256 tokens), per-expert counts -> tilecnt[e][tile]
//   pass 2 (scan)    : exclusive scan over the expert-major (e, tile) array -> tileoff, counts, offsets
//                      (one CTA, coalesced 4096-element blocks with a running carry; <= 1024 tiles)
//   pass 3 (scatter) : ballot ranks inside the tile give each pair its row; writes row_src / row_expert /
//                      row_w / tok_rows.
// One thread owns one token and keeps its E selection flags in a 64-bit register mask.
// ---------------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 256;

// Selection flags of one 256-token tile, built with COALESCED loads: warp-wide 32-element reads of the dense
// [T, E] matrix are turned into bit words with __ballot_sync (sparse_w > 0; NaN > 0 is false, quirk Q3) and
// parked in shared memory; each thread then extracts the E bits of its own token.
__device__ __forceinline__ unsigned long long token_mask(const float* __restrict__ w, int tile, int T, int E,
                                                         uint32_t* bits /* [kPlanThreads*64/32 + 2] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = (long long)tile * kPlanThreads * E;
    const int ntok = min(kPlanThreads, T - tile * kPlanThreads);
    const int nelem = ntok * E;
    const int nwords = (kPlanThreads * E + 31) / 32;
    // 8 independent 128-byte loads in flight per warp before the ballots (the loop is latency-bound otherwise)
    constexpr int kU = 8, kW = kPlanThreads / 32;
    for (int w0 = warp * kU; w0 < nwords + 2; w0 += kW * kU) {
        float v[kU];
#pragma unroll
        for (int q = 0; q < kU; ++q) {
            const int idx = (w0 + q) * 32 + lane;
            v[q] = idx < nelem ? __ldg(w + base + idx) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < kU; ++q) {
            const unsigned b = __ballot_sync(0xffffffffu, v[q] > 0.f);
            if (lane == 0 && w0 + q < nwords + 2) bits[w0 + q] = b;
        }
    }
    __syncthreads();
    const int o = threadIdx.x * E;
    const int wd = o >> 5, sh = o & 31;
    const unsigned long long lo = (unsigned long long)bits[wd] | ((unsigned long long)bits[wd + 1] << 32);
    unsigned long long v = lo >> sh;
    if (sh) v |= (unsigned long long)bits[wd + 2] << (64 - sh);
    if (E < 64) v &= (1ull << E) - 1ull;
    return threadIdx.x < ntok ? v : 0ull;
}

// Selection source of the plan kernels: the dense [T, E] sparse-weight matrix (the reference helper's argument), or the
// router kernel's own top-k output (topk_idx / topk_w [T, K], T*K*8 bytes instead of two passes over T*E*4).
struct PlanSrc {
    const float* dense;          // [T, E] or NULL
    const int32_t* topk_idx;     // [T, K]
    const float* topk_w;         // [T, K]
    int K;
};

// One thread = one token of the current 256-token group: its E selection flags (criterion weight > 0, quirk Q3) and, for
// the top-k source, the (index, weight) pairs in registers.
struct TokSel {
    unsigned long long mask;
    int32_t idx[HDMOE_MAX_TOPK];
    float w[HDMOE_MAX_TOPK];
};

__device__ __forceinline__ void token_select(const PlanSrc& src, int group, int T, int E, uint32_t* bits, TokSel& ts) {
    if (src.dense) {
        ts.mask = token_mask(src.dense, group, T, E, bits);
        return;
    }
    const int t = group * kPlanThreads + threadIdx.x;
    ts.mask = 0ull;
#pragma unroll
    for (int k = 0; k < HDMOE_MAX_TOPK; ++k) {
        ts.idx[k] = -1;
        ts.w[k] = 0.f;
        if (k < src.K && t < T) {
            const int e = src.topk_idx[(size_t)t * src.K + k];
            const float v = src.topk_w[(size_t)t * src.K + k];
            ts.idx[k] = e;
            ts.w[k] = v;
            if (v > 0.f && e >= 0 && e < E) ts.mask |= 1ull << e;      // NaN > 0 is false: all-masked rows dispatch nothing
        }
    }
}

// tile = G consecutive 256-token groups handled by one CTA (G keeps the (expert, tile) scan array small at T ~ 1 M)
// per-expert counts of tile `tile` into cnt[] (shared)
__device__ __forceinline__ void plan_count_tile(const PlanSrc& src, int T, int E, int tile, int G, int* cnt, uint32_t* bits) {
    if (threadIdx.x < E) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int ngroups = (T + kPlanThreads - 1) / kPlanThreads;
    for (int g = tile * G; g < min((tile + 1) * G, ngroups); ++g) {
        TokSel ts;
        token_select(src, g, T, E, bits, ts);
        if (src.K == 1 && !src.dense) {
            // one expert per token: the lanes of a warp that chose the same expert find each other with one match
            const int e = ts.mask ? ts.idx[0] : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, e >= 0 ? e : 64 + lane);
            if (e >= 0 && lane == __ffs(peers) - 1) atomicAdd(&cnt[e], __popc(peers));
        } else {
            // only the experts some lane of this warp selected (<= 32 * K of the E columns)
            unsigned p_lo = __reduce_or_sync(0xffffffffu, (unsigned)ts.mask);
            unsigned p_hi = __reduce_or_sync(0xffffffffu, (unsigned)(ts.mask >> 32));
            for (unsigned long long present = ((unsigned long long)p_hi << 32) | p_lo; present; present &= present - 1) {
                const int e = __ffsll((long long)present) - 1;
                const unsigned b = __ballot_sync(0xffffffffu, (ts.mask >> e) & 1ull);
                if (lane == 0) atomicAdd(&cnt[e], __popc(b));
            }
        }
        __syncthreads();       // bits[] is reused by the next group
    }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_count_kernel(PlanSrc src, int T, int E, int ntiles, int G, int32_t* __restrict__ tilecnt) {
    __shared__ int cnt[HDMOE_MAX_EXPERTS];
    __shared__ uint32_t bits[kPlanThreads * HDMOE_MAX_EXPERTS / 32 + 2];
    plan_count_tile(src, T, E, blockIdx.x, G, cnt, bits);
    if (threadIdx.x < E) tilecnt[(size_t)threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}

// E blocks of 1024 threads: block e turns the per-tile counts of expert e into tile-local exclusive offsets (coalesced
// blocks of 4096 with a running carry) and writes the expert's total to counts[e]; the expert bases (an exclusive scan
// over <= 64 totals) are formed by every scatter block in its prologue.  (Round 1 scanned all E * ntiles values in ONE
// block: 657 us at T = 1 M; the coalesced single-block version still took 38 us of the 68 us plan -- 16 dependent
// global round trips; one block per expert is one round trip.)
__global__ void __launch_bounds__(1024)
plan_scan_kernel(const int32_t* __restrict__ tilecnt, int ntiles, int32_t* __restrict__ tileoff,
                 int32_t* __restrict__ counts, int32_t* __restrict__ status) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int32_t* in = tilecnt + (size_t)blockIdx.x * ntiles;
    int32_t* out = tileoff + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        carry_s = 0;
        if (blockIdx.x == 0) status[0] = 0;
    }
    __syncthreads();
    for (int base = 0; base < ntiles; base += 4096) {
        const int carry = carry_s;       // stable here: written only between the two barriers below
        const int i0 = base + threadIdx.x * 4;
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = i0 + q < ntiles ? in[i0 + q] : 0;
        const int s = v[0] + v[1] + v[2] + v[3];
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = wsum[lane];
            int iw = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, iw, o);
                if (lane >= o) iw += u;
            }
            wsum[lane] = iw - w;                     // exclusive warp offsets
            if (lane == 31) carry_s = carry + iw;    // total so far (read again only after the next barrier)
        }
        __syncthreads();
        int run = carry + wsum[warp] + incl - s;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q;
            if (i < ntiles) {
                out[i] = run;
                run += v[q];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = carry_s;
}

// rows of tile `tile`; run_s[e] (shared) = first free row of expert e for this tile on entry
__device__ __forceinline__ void plan_scatter_tile(const PlanSrc& src, int T, int E, int tile, int G, int cap, int K,
                                                  int32_t* __restrict__ row_src, int32_t* __restrict__ row_expert,
                                                  float* __restrict__ row_w, int32_t* __restrict__ tok_rows,
                                                  int32_t* __restrict__ status, int (*warpoff)[HDMOE_MAX_EXPERTS],
                                                  int* run_s, uint32_t* bits) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ngroups = (T + kPlanThreads - 1) / kPlanThreads;
    for (int g = tile * G; g < min((tile + 1) * G, ngroups); ++g) {
        const int t = g * kPlanThreads + threadIdx.x;
        TokSel ts;
        token_select(src, g, T, E, bits, ts);
        const bool one = src.K == 1 && !src.dense;
        for (int e = lane; e < E; e += 32) warpoff[warp][e] = 0;
        __syncwarp();
        unsigned peers1 = 0;
        unsigned long long present = 0ull;
        if (one) {
            const int e = ts.mask ? ts.idx[0] : -1;
            peers1 = __match_any_sync(0xffffffffu, e >= 0 ? e : 64 + lane);
            if (e >= 0 && lane == __ffs(peers1) - 1) warpoff[warp][e] = __popc(peers1);
        } else {
            const unsigned p_lo = __reduce_or_sync(0xffffffffu, (unsigned)ts.mask);
            const unsigned p_hi = __reduce_or_sync(0xffffffffu, (unsigned)(ts.mask >> 32));
            present = ((unsigned long long)p_hi << 32) | p_lo;
            for (unsigned long long pr = present; pr; pr &= pr - 1) {
                const int e = __ffsll((long long)pr) - 1;
                const unsigned b = __ballot_sync(0xffffffffu, (ts.mask >> e) & 1ull);
                if (lane == 0) warpoff[warp][e] = __popc(b);
            }
        }
        __syncthreads();
        if (threadIdx.x < E) {   // exclusive scan over warps, per expert, seeded with the running offset of the tile
            int run = run_s[threadIdx.x];
            for (int wv = 0; wv < kPlanThreads / 32; ++wv) {
                const int c = warpoff[wv][threadIdx.x];
                warpoff[wv][threadIdx.x] = run;
                run += c;
            }
            run_s[threadIdx.x] = run;
        }
        __syncthreads();
        int slot = 0;
        // ascending expert order over the experts present in this warp (uniform loop: every lane takes part in the ballot)
        unsigned long long walk = one ? 0ull : present;
        bool one_pending = one;
        while (walk || one_pending) {
            int e;
            bool sel;
            unsigned b;
            if (one_pending) {
                one_pending = false;
                e = ts.mask ? ts.idx[0] : 0;
                sel = ts.mask != 0ull;
                b = peers1;
            } else {
                e = __ffsll((long long)walk) - 1;
                walk &= walk - 1;
                sel = (ts.mask >> e) & 1ull;
                b = __ballot_sync(0xffffffffu, sel);
            }
            if (sel) {
                const int pos = warpoff[warp][e] + __popc(b & ((1u << lane) - 1u));
                if (pos < cap) {
                    float wv = 0.f;
                    if (src.dense) {
                        wv = src.dense[(size_t)t * E + e];
                    } else {
#pragma unroll
                        for (int k = 0; k < HDMOE_MAX_TOPK; ++k)
                            if (ts.idx[k] == e) wv = ts.w[k];
                    }
                    row_src[pos] = t;
                    row_expert[pos] = e;
                    row_w[pos] = wv;
                }
                if (slot < K) tok_rows[(size_t)t * K + slot] = pos < cap ? pos : -1;
                else status[0] = 2;
                ++slot;
            }
        }
        if (t < T)
            for (; slot < K; ++slot) tok_rows[(size_t)t * K + slot] = -1;
        __syncthreads();       // warpoff / bits are rewritten by the next group
    }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_scatter_kernel(PlanSrc src, int T, int E, int ntiles, int G, int cap, int K,
                    const int32_t* __restrict__ tileoff, const int32_t* __restrict__ counts,
                    int32_t* __restrict__ offsets, int32_t* __restrict__ row_src,
                    int32_t* __restrict__ row_expert, float* __restrict__ row_w, int32_t* __restrict__ tok_rows,
                    int32_t* __restrict__ status) {
    __shared__ int warpoff[kPlanThreads / 32][HDMOE_MAX_EXPERTS];
    __shared__ int run_s[HDMOE_MAX_EXPERTS];         // next free row of every expert inside this tile
    __shared__ uint32_t bits[kPlanThreads * HDMOE_MAX_EXPERTS / 32 + 2];
    __shared__ int total_s;
    // expert bases = exclusive scan over the E totals the scan kernel left in counts[] (E <= 64: one thread, shared memory)
    if (threadIdx.x < E) run_s[threadIdx.x] = counts[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int e = 0; e < E; ++e) {
            const int c = run_s[e];
            run_s[e] = run;
            run += c;
        }
        total_s = run;
        if (blockIdx.x == 0) {
            for (int e = 0; e < E; ++e) offsets[e] = run_s[e];
            offsets[E] = run;
            if (run > cap) atomicMax(&status[0], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < E) run_s[threadIdx.x] += tileoff[(size_t)threadIdx.x * ntiles + blockIdx.x];
    __syncthreads();
    plan_scatter_tile(src, T, E, blockIdx.x, G, cap, K, row_src, row_expert, row_w, tok_rows, status, warpoff, run_s, bits);
    // the unused tail [R, cap): well-defined values for fixed-size consumers (was a launch of its own)
    for (int i = total_s + blockIdx.x * kPlanThreads + threadIdx.x; i < cap; i += gridDim.x * kPlanThreads) {
        row_src[i] = -1;
        row_expert[i] = -1;
        row_w[i] = 0.f;
    }
}

// T <= kPlanSmallT: the whole plan (count, offsets, scatter, tail fill) in ONE CTA -- at the reference's own sizes
// (T = batch = 256 ... 1024) the four-launch version is pure launch latency (19 us against 30 us of data movement)
// (only up to 1024 tokens: at T = 4096 the single CTA takes 35 us at E = 4 and 147 us at E = 64, the three-launch path 15-24 us
// -- profiles/r2_dispatch_full_sweep.md)
constexpr int kPlanSmallT = 1024;
__global__ void __launch_bounds__(kPlanThreads)
plan_small_kernel(PlanSrc src, int T, int E, int cap, int K, int32_t* __restrict__ counts, int32_t* __restrict__ offsets,
                  int32_t* __restrict__ row_src, int32_t* __restrict__ row_expert, float* __restrict__ row_w,
                  int32_t* __restrict__ tok_rows, int32_t* __restrict__ status) {
    __shared__ int warpoff[kPlanThreads / 32][HDMOE_MAX_EXPERTS];
    __shared__ int run_s[HDMOE_MAX_EXPERTS];
    __shared__ int cnt[HDMOE_MAX_EXPERTS];
    __shared__ uint32_t bits[kPlanThreads * HDMOE_MAX_EXPERTS / 32 + 2];
    __shared__ int total_s;
    const int G = (T + kPlanThreads - 1) / kPlanThreads;
    if (threadIdx.x == 0) status[0] = 0;
    plan_count_tile(src, T, E, 0, G, cnt, bits);
    if (threadIdx.x == 0) {
        int run = 0;
        for (int e = 0; e < E; ++e) {
            offsets[e] = run;
            run_s[e] = run;
            counts[e] = cnt[e];
            run += cnt[e];
        }
        offsets[E] = run;
        total_s = run;
        if (run > cap) status[0] = 1;
    }
    __syncthreads();
    plan_scatter_tile(src, T, E, 0, G, cap, K, row_src, row_expert, row_w, tok_rows, status, warpoff, run_s, bits);
    for (int i = total_s + threadIdx.x; i < cap; i += kPlanThreads) {
        row_src[i] = -1;
        row_expert[i] = -1;
        row_w[i] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------
// Row gather.  Two paths:
//  * bulk: rows >= 4 KiB move as cp.async.bulk global->shared->global (the TMA engine does the copy, one
//    elected thread per CTA drives a 4-deep ring of 16 KiB buffers);
//  * vector: a warp copies a row chunk with 128-bit streaming loads/stores, 4 in flight per lane.
// ---------------------------------------------------------------------------------------------------
struct PermuteArgs {
    const char* src[4];
    char* dst[4];
    long long row_bytes[4];
    long long chunks_per_row[4];   // chunk = kChunk bytes
    long long chunk_base[4];       // first global chunk id of tensor i (for one row-set)
    int n_tensors;
};

constexpr int kChunk = 16384;
constexpr int kBulkStages = 6;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}

// one work item = (row r, tensor i, chunk c); items are enumerated tensor-major per row so that the big
// payload rows dominate and every CTA streams contiguous 16 KiB pieces.
__global__ void __launch_bounds__(32)
permute_bulk_kernel(PermuteArgs a, const int32_t* __restrict__ row_src, const int32_t* __restrict__ n_rows_dev,
                    int cap, long long chunks_per_rowset) {
    extern __shared__ __align__(128) unsigned char sbuf[];
    __shared__ __align__(8) uint64_t full[kBulkStages];
    const int R = min(*n_rows_dev, cap);
    const long long total = (long long)cap * chunks_per_rowset;
    if (threadIdx.x != 0) return;
    for (int s = 0; s < kBulkStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    // software pipeline over this CTA's items k = 0..n-1 (global item = blockIdx.x + k*gridDim.x):
    // loads run kLookahead items ahead of the stores; the buffer a load reuses was last read by store
    // k-2, i.e. everything but the most recent bulk group must have finished reading shared memory.
    constexpr int kLookahead = kBulkStages - 2;
    const long long n = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto decode = [&](long long k, const char*& src, char*& dst, uint32_t& bytes) {
        const long long item = blockIdx.x + k * (long long)gridDim.x;
        const int r = (int)(item / chunks_per_rowset);
        long long c = item - (long long)r * chunks_per_rowset;
        int i = 0;
        while (i + 1 < a.n_tensors && c >= a.chunk_base[i + 1]) ++i;
        c -= a.chunk_base[i];
        const long long off = c * kChunk;
        const long long rem = a.row_bytes[i] - off;
        bytes = (uint32_t)(rem < kChunk ? rem : kChunk);
        dst = a.dst[i] + (long long)r * a.row_bytes[i] + off;
        const int sr = r < R ? row_src[r] : -1;                 // tail rows and holes (row_src < 0): zero-fill
        src = sr >= 0 ? a.src[i] + (long long)sr * a.row_bytes[i] + off : nullptr;
    };
    auto load = [&](long long k) {
        const int s = (int)(k % kBulkStages);
        const char* src;
        char* dst;
        uint32_t bytes;
        decode(k, src, dst, bytes);
        if (src) {
            mbar_expect_tx(&full[s], bytes);
            bulk_g2s(sbuf + (size_t)s * kChunk, src, bytes, &full[s]);
        } else {
            int4* z = reinterpret_cast<int4*>(sbuf + (size_t)s * kChunk);
            for (uint32_t q = 0; q < bytes / 16; ++q) z[q] = make_int4(0, 0, 0, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[s], 0);
        }
    };
    for (long long k = 0; k < kLookahead && k < n; ++k) load(k);
    for (long long k = 0; k < n; ++k) {
        if (k + kLookahead < n) {
            if (k >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(k + kLookahead);
        }
        const int s = (int)(k % kBulkStages);
        const char* src;
        char* dst;
        uint32_t bytes;
        decode(k, src, dst, bytes);
        mbar_wait(&full[s], (uint32_t)((k / kBulkStages) & 1));
        bulk_s2g(dst, sbuf + (size_t)s * kChunk, bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// vector path A (16-byte aligned rows): flat over 16-byte vectors, 4 independent loads in flight per thread
constexpr int kVecPiece = 2048;
__global__ void __launch_bounds__(256)
permute_vec16_kernel(PermuteArgs a, const int32_t* __restrict__ row_src, const int32_t* __restrict__ n_rows_dev,
                     int cap, int vec_per_rowset, int vb1, int vb2, int vb3) {
    const int R = min(*n_rows_dev, cap);
    const long long total = (long long)cap * vec_per_rowset;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += 4 * stride) {
        int4 val[4];
        int4* dstp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long v = v0 + q * stride;
            dstp[q] = nullptr;
            val[q] = make_int4(0, 0, 0, 0);
            if (v < total) {
                const int r = (int)(v / vec_per_rowset);
                int c = (int)(v - (long long)r * vec_per_rowset);
                int i = 0;
                if (a.n_tensors > 3 && c >= vb3) { i = 3; c -= vb3; }
                else if (a.n_tensors > 2 && c >= vb2) { i = 2; c -= vb2; }
                else if (a.n_tensors > 1 && c >= vb1) { i = 1; c -= vb1; }
                dstp[q] = reinterpret_cast<int4*>(a.dst[i] + (long long)r * a.row_bytes[i]) + c;
                const int sr = r < R ? row_src[r] : -1;
                if (sr >= 0)
                    val[q] = ld_stream(reinterpret_cast<const int4*>(a.src[i] + (long long)sr * a.row_bytes[i]) + c);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (dstp[q]) st_stream(dstp[q], val[q]);
    }
}

// vector path B (rows that are only 4-byte aligned): one warp per (row, tensor, 2 KiB piece)
__global__ void __launch_bounds__(256)
permute_vec_kernel(PermuteArgs a, const int32_t* __restrict__ row_src, const int32_t* __restrict__ n_rows_dev,
                   int cap, long long pieces_per_rowset, long long piece_base1, long long piece_base2,
                   long long piece_base3) {
    const int R = min(*n_rows_dev, cap);
    const long long total = (long long)cap * pieces_per_rowset;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long item = warp0; item < total; item += nwarps) {
        const int r = (int)(item / pieces_per_rowset);
        long long c = item - (long long)r * pieces_per_rowset;
        int i = 0;
        if (a.n_tensors > 3 && c >= piece_base3) { i = 3; c -= piece_base3; }
        else if (a.n_tensors > 2 && c >= piece_base2) { i = 2; c -= piece_base2; }
        else if (a.n_tensors > 1 && c >= piece_base1) { i = 1; c -= piece_base1; }
        const long long off = c * kVecPiece;
        const long long rem = a.row_bytes[i] - off;
        const int bytes = (int)(rem < kVecPiece ? rem : kVecPiece);
        char* dst = a.dst[i] + (long long)r * a.row_bytes[i] + off;
        const int sr = r < R ? row_src[r] : -1;
        const bool live = sr >= 0;               // tail rows and holes (row_src < 0) are zero-filled
        const char* src = live ? a.src[i] + (long long)sr * a.row_bytes[i] + off : nullptr;
        const bool al16 = ((((uintptr_t)dst) | ((uintptr_t)(live ? src : dst)) | (uintptr_t)bytes) & 15) == 0;
        if (al16) {
            const int n16 = bytes >> 4;
            int4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lane + 32 * q;
                v[q] = (live && j < n16) ? ld_stream(reinterpret_cast<const int4*>(src) + j) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lane + 32 * q;
                if (j < n16) st_stream(reinterpret_cast<int4*>(dst) + j, v[q]);
            }
        } else {
            const int n4 = bytes >> 2;
            for (int j = lane; j < n4; j += 32)
                reinterpret_cast<int*>(dst)[j] = live ? reinterpret_cast<const int*>(src)[j] : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Combine.  grid-stride over (token, 4-element vector); K <= 8 gathered rows per token.
// ---------------------------------------------------------------------------------------------------
// N-element register vectors: 16 bytes of the ROW dtype per access (8 bf16 / 4 fp32)
template <typename T>
struct VecN;
template <>
struct VecN<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 f = *reinterpret_cast<const float4*>(p);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
};
template <>
struct VecN<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const int4 u = ld_stream(reinterpret_cast<const int4*>(p));
        const uint32_t w[4] = {(uint32_t)u.x, (uint32_t)u.y, (uint32_t)u.z, (uint32_t)u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[2 * q] = __uint_as_float(w[q] << 16);
            v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
        }
    }
};
template <typename TO, int N>
__device__ __forceinline__ void store_n(TO* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store_n<float, 4>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store_n<float, 8>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store_n<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&v)[8]) {
    Vec4<__nv_bfloat16>::store(p, make_float4(v[0], v[1], v[2], v[3]));
}
template <>
__device__ __forceinline__ void store_n<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __nv_bfloat162 o = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
        w[q] = *reinterpret_cast<uint32_t*>(&o);
    }
    st_stream(reinterpret_cast<int4*>(p), make_int4(w[0], w[1], w[2], w[3]));
}
template <typename TO, int N>
__device__ __forceinline__ void load_base(const TO* p, float (&v)[8]) {
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = to_f32<TO>(p[q]);
}

// K = 1 without a residual base (the reference's top-1 configuration): out[t] = rows[tok_rows[t]] * w.  Four independent
// 16-byte row vectors in flight per thread -- index loads first, then the row loads, then the stores -- instead of the
// generic kernel's dependent index -> weight -> row chain per vector.
template <typename TR, typename TO>
__global__ void __launch_bounds__(256)
combine_k1_kernel(const TR* __restrict__ rows, const int32_t* __restrict__ tok_rows, const float* __restrict__ row_w,
                  TO* __restrict__ out, int T, long long D, int vshift) {
    constexpr int N = VecN<TR>::N, U = 4;
    const long long vec_per_row = D / N;
    const long long total = (long long)T * vec_per_row;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
        int t[U], r[U];
        long long d[U];
        float w[U], v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            t[u] = -1;
            r[u] = -1;
            if (i < total) {
                t[u] = vshift >= 0 ? (int)(i >> vshift) : (int)(i / vec_per_row);
                d[u] = (i - (long long)t[u] * vec_per_row) * N;
                r[u] = __ldg(tok_rows + t[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            w[u] = (r[u] >= 0 && row_w) ? __ldg(row_w + r[u]) : 1.f;
#pragma unroll
            for (int q = 0; q < N; ++q) v[u][q] = 0.f;
            if (r[u] >= 0) VecN<TR>::load(rows + (long long)r[u] * D + d[u], v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t[u] < 0) continue;
            float acc[8];
            // multiply, round, then add to zero (the reference does `out_e * w` then `+=` into zeros): no FMA contraction
#pragma unroll
            for (int q = 0; q < N; ++q) acc[q] = r[u] >= 0 ? __fadd_rn(0.f, __fmul_rn(v[u][q], w[u])) : 0.f;
            store_n<TO, N>(out + (long long)t[u] * D + d[u], acc);
        }
    }
}

template <typename TR, typename TO>
__global__ void __launch_bounds__(256)
combine_kernel(const TR* __restrict__ rows, const int32_t* __restrict__ tok_rows, const float* __restrict__ row_w,
               const TO* __restrict__ base, TO* __restrict__ out, int T, int K, long long D, int vshift) {
    constexpr int N = VecN<TR>::N;
    const long long vec_per_row = D / N;
    const long long total = (long long)T * vec_per_row;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long i = i0 + u * stride;
            if (i >= total) break;
            const int t = vshift >= 0 ? (int)(i >> vshift) : (int)(i / vec_per_row);
            const long long d = (i - (long long)t * vec_per_row) * N;
            float acc[8];
            if (base) load_base<TO, N>(base + (long long)t * D + d, acc);
            else {
#pragma unroll
                for (int q = 0; q < N; ++q) acc[q] = 0.f;
            }
            for (int j = 0; j < K; ++j) {
                const int r = __ldg(tok_rows + (long long)t * K + j);
                if (r < 0) continue;
                const float w = row_w ? __ldg(row_w + r) : 1.f;
                float v[8];
                VecN<TR>::load(rows + (long long)r * D + d, v);
                // multiply, round, then add (the reference does `out_e * w` then `+=`): no FMA contraction
#pragma unroll
                for (int q = 0; q < N; ++q) acc[q] = __fadd_rn(acc[q], __fmul_rn(v[q], w));
            }
            store_n<TO, N>(out + (long long)t * D + d, acc);
        }
    }
}

// backward: one CTA per permuted row.  d_rows[r] = w_r * dY[src_r];  d_sparse[src_r, e_r] = <rows[r], dY[src_r]>
template <typename TR, typename TY>
__global__ void __launch_bounds__(256)
combine_bwd_kernel(const TR* __restrict__ rows, const TY* __restrict__ dY, const int32_t* __restrict__ row_src,
                   const int32_t* __restrict__ row_expert, const float* __restrict__ row_w,
                   const int32_t* __restrict__ n_rows_dev, int cap, int E, long long D, TR* __restrict__ d_rows,
                   float* __restrict__ d_sparse_w) {
    __shared__ float red[8];
    const int R = min(*n_rows_dev, cap);
    for (int r = blockIdx.x; r < cap; r += gridDim.x) {
        TR* dr = d_rows + (long long)r * D;
        if (r >= R || row_src[r] < 0) {          // unused tail rows and holes of a spread (expert-parallel) layout
            for (long long d = (long long)threadIdx.x << 2; d < D; d += (long long)blockDim.x << 2)
                Vec4<TR>::store(dr + d, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const int t = row_src[r];
        const float w = row_w ? row_w[r] : 1.f;
        const TY* gy = dY + (long long)t * D;
        const TR* xr = rows ? rows + (long long)r * D : nullptr;
        float dot = 0.f;
        for (long long d = (long long)threadIdx.x << 2; d < D; d += (long long)blockDim.x << 2) {
            const float4 g = Vec4<TY>::load(gy + d);
            Vec4<TR>::store(dr + d, make_float4(g.x * w, g.y * w, g.z * w, g.w * w));
            if (xr) {
                const float4 x = Vec4<TR>::load(xr + d);
                dot += g.x * x.x + g.y * x.y + g.z * x.z + g.w * x.w;
            }
        }
        if (d_sparse_w) {
            dot = warp_sum(dot);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
            __syncthreads();
            if (threadIdx.x == 0) {
                float s = 0.f;
                for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += red[q];
                d_sparse_w[(long long)t * E + row_expert[r]] = s;
            }
            __syncthreads();
        }
    }
}

}  // namespace hdmoe

using namespace hdmoe;

// groups of 256 tokens per CTA: at most 4096 tiles, so that one block of the per-expert scan covers an expert's tile
// counts in a single coalesced pass (T <= 1 M: one group per tile, the count / scatter blocks have no serial loop)
static int plan_groups_per_tile(int T) {
    const int ngroups = (T + kPlanThreads - 1) / kPlanThreads;
    return (ngroups + 4095) / 4096;
}

extern "C" size_t hdmoe_dispatch_plan_workspace_bytes(int T, int E) {
    const size_t ntiles = (size_t)(T + kPlanThreads - 1) / kPlanThreads + 1;
    return 2 * ntiles * (size_t)E * sizeof(int32_t) + 64;
}

static int dispatch_plan_impl(PlanSrc src, int T, int E, int cap, int K, int32_t* counts, int32_t* offsets,
                              int32_t* row_src, int32_t* row_expert, float* row_w, int32_t* tok_rows, int32_t* status,
                              void* workspace, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(T >= 1 && E >= 1 && E <= HDMOE_MAX_EXPERTS, "dispatch_plan: need T >= 1, 1 <= E <= %d",
                    HDMOE_MAX_EXPERTS);
    HDMOE_CHECK_ARG(cap >= 1 && K >= 1 && K <= E, "dispatch_plan: need cap >= 1 and 1 <= K <= E");
    HDMOE_CHECK_ARG(counts && offsets && row_src && row_expert && row_w && tok_rows && status && workspace,
                    "dispatch_plan: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (T <= kPlanSmallT) {
        plan_small_kernel<<<1, kPlanThreads, 0, st>>>(src, T, E, cap, K, counts, offsets, row_src, row_expert, row_w, tok_rows,
                                                      status);
        HDMOE_CHECK_LAUNCH();
        return HDMOE_OK;
    }
    const int G = plan_groups_per_tile(T);
    const int ngroups = (T + kPlanThreads - 1) / kPlanThreads;
    const int ntiles = (ngroups + G - 1) / G;
    int32_t* tilecnt = (int32_t*)workspace;
    int32_t* tileoff = tilecnt + (size_t)E * ntiles;
    plan_count_kernel<<<ntiles, kPlanThreads, 0, st>>>(src, T, E, ntiles, G, tilecnt);
    HDMOE_CHECK_LAUNCH();
    plan_scan_kernel<<<E, 1024, 0, st>>>(tilecnt, ntiles, tileoff, counts, status);
    HDMOE_CHECK_LAUNCH();
    plan_scatter_kernel<<<ntiles, kPlanThreads, 0, st>>>(src, T, E, ntiles, G, cap, K, tileoff, counts, offsets, row_src,
                                                         row_expert, row_w, tok_rows, status);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_dispatch_plan(const float* sparse_w, int T, int E, int cap, int K, int32_t* counts,
                                   int32_t* offsets, int32_t* row_src, int32_t* row_expert, float* row_w,
                                   int32_t* tok_rows, int32_t* status, void* workspace, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(sparse_w, "dispatch_plan: null pointer");
    PlanSrc src{sparse_w, nullptr, nullptr, 0};
    return dispatch_plan_impl(src, T, E, cap, K, counts, offsets, row_src, row_expert, row_w, tok_rows, status, workspace,
                              stream);
}

extern "C" int hdmoe_dispatch_plan_topk(const int32_t* topk_idx, const float* topk_w, int T, int E, int K, int cap,
                                        int32_t* counts, int32_t* offsets, int32_t* row_src, int32_t* row_expert,
                                        float* row_w, int32_t* tok_rows, int32_t* status, void* workspace,
                                        hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(topk_idx && topk_w, "dispatch_plan_topk: null pointer");
    HDMOE_CHECK_ARG(K >= 1 && K <= HDMOE_MAX_TOPK, "dispatch_plan_topk: 1 <= K <= %d", HDMOE_MAX_TOPK);
    PlanSrc src{nullptr, topk_idx, topk_w, K};
    return dispatch_plan_impl(src, T, E, cap, K, counts, offsets, row_src, row_expert, row_w, tok_rows, status, workspace,
                              stream);
}

extern "C" int hdmoe_permute_rows(const void* const* srcs, void* const* dsts, const int64_t* row_bytes, int n_tensors,
                                  const int32_t* row_src, const int32_t* n_rows_dev, int cap, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(n_tensors >= 1 && n_tensors <= 4, "permute_rows: 1..4 tensors per call");
    HDMOE_CHECK_ARG(row_src && n_rows_dev && cap >= 1, "permute_rows: null index / cap < 1");
    cudaStream_t st = (cudaStream_t)stream;
    PermuteArgs a{};
    a.n_tensors = n_tensors;
    bool bulk_ok = true;
    long long max_row = 0;
    for (int i = 0; i < n_tensors; ++i) {
        HDMOE_CHECK_ARG(srcs[i] && dsts[i] && row_bytes[i] > 0 && row_bytes[i] % 4 == 0,
                        "permute_rows: tensor %d: null pointer or row_bytes %% 4 != 0", i);
        a.src[i] = (const char*)srcs[i];
        a.dst[i] = (char*)dsts[i];
        a.row_bytes[i] = row_bytes[i];
        if (row_bytes[i] % 16 || ((uintptr_t)srcs[i] & 15) || ((uintptr_t)dsts[i] & 15)) bulk_ok = false;
        if (row_bytes[i] > max_row) max_row = row_bytes[i];
    }
    if (bulk_ok && max_row >= 4096) {
        long long base = 0;
        for (int i = 0; i < n_tensors; ++i) {
            a.chunk_base[i] = base;
            a.chunks_per_row[i] = (row_bytes[i] + kChunk - 1) / kChunk;
            base += a.chunks_per_row[i];
        }
        const long long total = (long long)cap * base;
        const int smem = kBulkStages * kChunk;
        // per (device, function) attribute: set on every call (cheap), a process may drive several GPUs
        HDMOE_CHECK_CUDA(cudaFuncSetAttribute(permute_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        long long g = total < (long long)kNumSMs * 2 ? total : (long long)kNumSMs * 2;
        permute_bulk_kernel<<<(int)g, 32, smem, st>>>(a, row_src, n_rows_dev, cap, base);
        HDMOE_CHECK_LAUNCH();
    } else if (bulk_ok && max_row < 1024) {
        int vb[4] = {0, 0, 0, 0}, base = 0;
        for (int i = 0; i < n_tensors; ++i) {
            vb[i] = base;
            base += (int)(row_bytes[i] / 16);
        }
        const long long total = (long long)cap * base;
        permute_vec16_kernel<<<grid_for(total, 1024, 16), 256, 0, st>>>(a, row_src, n_rows_dev, cap, base, vb[1], vb[2], vb[3]);
        HDMOE_CHECK_LAUNCH();
    } else {
        long long pb[4] = {0, 0, 0, 0}, base = 0;
        for (int i = 0; i < n_tensors; ++i) {
            pb[i] = base;
            base += (row_bytes[i] + kVecPiece - 1) / kVecPiece;
        }
        const long long total = (long long)cap * base;
        permute_vec_kernel<<<grid_for(total, 8, 16), 256, 0, st>>>(a, row_src, n_rows_dev, cap, base, pb[1], pb[2], pb[3]);
        HDMOE_CHECK_LAUNCH();
    }
    return HDMOE_OK;
}

template <typename TR, typename TO>
static int launch_combine(const void* rows, const int32_t* tok_rows, const float* row_w, const void* base, void* out,
                          int T, int K, int64_t D, cudaStream_t st) {
    constexpr int N = VecN<TR>::N;
    const long long vpr = D / N;
    const long long total = (long long)T * vpr;
    int vshift = -1;
    for (int b = 0; b < 31; ++b)
        if ((1ll << b) == vpr) vshift = b;
    if (K == 1 && !base)
        combine_k1_kernel<TR, TO><<<grid_for(total, 1024, 8), 256, 0, st>>>((const TR*)rows, tok_rows, row_w, (TO*)out, T, D,
                                                                            vshift);
    else
        combine_kernel<TR, TO><<<grid_for(total, 512, 16), 256, 0, st>>>((const TR*)rows, tok_rows, row_w, (const TO*)base,
                                                                        (TO*)out, T, K, D, vshift);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_combine_rows(const void* rows, int rows_dtype, const int32_t* tok_rows, const float* row_w,
                                  const void* base, void* out, int out_dtype, int T, int K, int64_t D,
                                  hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(rows && tok_rows && out && T >= 1 && K >= 1 && K <= HDMOE_MAX_EXPERTS, "combine_rows: bad args");
    const int vecn = rows_dtype == HDMOE_BF16 ? 8 : 4;
    HDMOE_CHECK_ARG(D >= vecn && D % vecn == 0, "combine_rows: row width must be a multiple of %d elements (got %lld)",
                    vecn, (long long)D);
    HDMOE_CHECK_ARG((((uintptr_t)rows | (uintptr_t)out | (uintptr_t)base) & 15) == 0, "combine_rows: 16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows_dtype == HDMOE_F32 && out_dtype == HDMOE_F32)
        return launch_combine<float, float>(rows, tok_rows, row_w, base, out, T, K, D, st);
    if (rows_dtype == HDMOE_BF16 && out_dtype == HDMOE_BF16)
        return launch_combine<__nv_bfloat16, __nv_bfloat16>(rows, tok_rows, row_w, base, out, T, K, D, st);
    if (rows_dtype == HDMOE_BF16 && out_dtype == HDMOE_F32)
        return launch_combine<__nv_bfloat16, float>(rows, tok_rows, row_w, base, out, T, K, D, st);
    if (rows_dtype == HDMOE_F32 && out_dtype == HDMOE_BF16)
        return launch_combine<float, __nv_bfloat16>(rows, tok_rows, row_w, base, out, T, K, D, st);
    HDMOE_CHECK_ARG(false, "combine_rows: unsupported dtype pair (%d, %d)", rows_dtype, out_dtype);
}

template <typename TR, typename TY>
static int launch_combine_bwd(const void* rows, const void* dY, const int32_t* row_src, const int32_t* row_expert,
                              const float* row_w, const int32_t* n_rows_dev, int cap, int E, int64_t D, void* d_rows,
                              float* d_sparse_w, cudaStream_t st) {
    int grid = cap < kNumSMs * 8 ? cap : kNumSMs * 8;
    combine_bwd_kernel<TR, TY><<<grid, 256, 0, st>>>((const TR*)rows, (const TY*)dY, row_src, row_expert, row_w,
                                                     n_rows_dev, cap, E, D, (TR*)d_rows, d_sparse_w);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_combine_rows_bwd(const void* rows, int rows_dtype, const void* dY, int dy_dtype,
                                      const int32_t* row_src, const int32_t* row_expert, const float* row_w,
                                      const int32_t* n_rows_dev, int cap, int T, int E, int64_t D, void* d_rows,
                                      float* d_sparse_w, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(dY && row_src && row_expert && n_rows_dev && d_rows && cap >= 1, "combine_rows_bwd: bad args");
    HDMOE_CHECK_ARG(D >= 4 && D % 4 == 0, "combine_rows_bwd: row width must be a multiple of 4 elements");
    HDMOE_CHECK_ARG(!d_sparse_w || rows, "combine_rows_bwd: d_sparse_w needs the forward rows");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_sparse_w) HDMOE_CHECK_CUDA(cudaMemsetAsync(d_sparse_w, 0, (size_t)T * E * sizeof(float), st));
    if (rows_dtype == HDMOE_F32 && dy_dtype == HDMOE_F32)
        return launch_combine_bwd<float, float>(rows, dY, row_src, row_expert, row_w, n_rows_dev, cap, E, D, d_rows,
                                                d_sparse_w, st);
    if (rows_dtype == HDMOE_BF16 && dy_dtype == HDMOE_BF16)
        return launch_combine_bwd<__nv_bfloat16, __nv_bfloat16>(rows, dY, row_src, row_expert, row_w, n_rows_dev, cap,
                                                                E, D, d_rows, d_sparse_w, st);
    if (rows_dtype == HDMOE_BF16 && dy_dtype == HDMOE_F32)
        return launch_combine_bwd<__nv_bfloat16, float>(rows, dY, row_src, row_expert, row_w, n_rows_dev, cap, E, D,
                                                        d_rows, d_sparse_w, st);
    if (rows_dtype == HDMOE_F32 && dy_dtype == HDMOE_BF16)
        return launch_combine_bwd<float, __nv_bfloat16>(rows, dY, row_src, row_expert, row_w, n_rows_dev, cap, E, D,
                                                        d_rows, d_sparse_w, st);
    HDMOE_CHECK_ARG(false, "combine_rows_bwd: unsupported dtype pair");
}
