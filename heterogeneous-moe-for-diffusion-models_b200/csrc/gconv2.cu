// G-CONV v2: grouped implicit-GEMM convolution with HALO REUSE (tcgen05 + TMEM + TMA).
//
// v1 (gconv.cu) re-loads the shifted 128-pixel input tile from L2 once per filter tap: 9x / 25x re-reads make
// it L2-bandwidth-bound (~42 B/clk/SM) at 11-20 % of tensor peak.  v2 loads the input ONCE per tile:
//
//   * a tile is a strip of SH output rows of one sample; its zero-padded input window
//     [SH + k - 1] x [W + k - 1] x KC is ONE 4-D TMA box (out-of-bounds zero fill = 'same' padding) landing
//     in shared memory as a dense array of "positions" p = row * Wp + col, KC*2 bytes each (SW64 / SW128);
//   * in that flattened padded image the neighbour of output position q under tap (r, s) is simply
//     q + r*Wp + s, so every tap's A operand is the SAME buffer with a different start address:
//     adesc.start = base + (r*Wp + s + 128*mt) * rowbytes.  Hardware fact (tools/umma_probe.cu, B200): a K-major
//     swizzled operand may start at any 128-byte row with base_offset = 0 -- the swizzle acts on absolute
//     shared-memory address bits;
//   * M-tiles cover 128 consecutive positions of the flattened padded image [0, H*Wp); positions in the padding
//     columns produce garbage rows that the epilogue drops.  A tile is a run of <= 3 M-tiles that need NOT start on
//     an image row (32x32, k=5: 1152 positions = exactly 9 M-tiles = 3 tiles), so every issuer warp is busy on
//     every tile -- a lone M-tile would run at the single-thread issue rate (~103 cycles / MMA, tools/umma_rate.cu);
//   * weights stream through a small ring, one [Cout x KC] box per (tap, chunk), reused by all M-tiles;
//   * the epilogue goes TMEM -> registers -> global with 32-byte vector stores (a thread owns one position's
//     channel vector = one contiguous NHWC line).  It must NOT touch shared memory: the SS-mode MMAs at N <= 64
//     already consume the full 128 B/clk of shared-memory bandwidth, and an smem-staged epilogue ran 3x slower
//     under them (15 k cycles per tile, trace in profiles/) while slowing the MMAs from 48 to 60 cycles;
//   * tiles are handed out dynamically (heavy kernel sizes first) from a self-resetting global counter: a 5x5 tile
//     costs 2.8x a 3x3 tile, static striding left 16 % of the SM-cycles idle.
//
// Everything else follows v1: per-tile expert lookup (kernel size, weight block) for the grouped /
// heterogeneous case, warp-specialised roles (TMA producer, MMA issuers, 4 epilogue warps),
// double-buffered TMEM accumulators, fused epilogue (scale, mp_silu, mp_sum residual), NHWC bf16 output.
#include "tc.cuh"
#include "../../include/hdmoe_gemm.h"

namespace hdmoe {

#ifdef HDMOE_G2_TRACE
__device__ long long g2_trace[148 * 64];
__device__ long long g2_span[148 * 2];          // globaltimer at kernel entry / exit of every CTA
__device__ __forceinline__ long long g2_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define G2T(slot) do { if (blockIdx.x < 148 && tcount < 8) { g2_trace[blockIdx.x * 64 + tcount * 8 + (slot)] = clock64(); \
    if ((slot) == 0) g2_trace[blockIdx.x * 64 + tcount * 8 + 6] = g2_gtime(); if ((slot) == 5) g2_trace[blockIdx.x * 64 + tcount * 8 + 7] = g2_gtime(); } } while (0)
#else
#define G2T(slot) do { } while (0)
#endif

constexpr int kG2Issuers = 3;      // MMA-issuer warps, one per M-tile accumulator (a single thread cannot issue
                                   // 32-cycle N=64 MMAs fast enough: ~12 uniform-datapath instructions per UTCHMMA)
constexpr int kG2Threads = 32 * (1 + kG2Issuers + 4);
constexpr int kG2MaxE = HDMOE_MAX_EXPERTS;
constexpr int kG2Classes = 4;     // distinct kernel sizes per launch
constexpr int kG2RowCache = 4096; // rows whose expert id is staged in shared memory (larger launches read global memory)
constexpr int kG2Queue = 16;      // depth of the tile-id queue between the scheduler (producer) and the other roles

struct GConv2Params {
    int n_tiles, smax;                // tiles = cap_rows * smax (strips per sample, max over classes)
    int H, W;
    int upt;                          // channel chunks of KC per tap
    int n_experts;
    int a_stage_bytes;                // bytes of one halo buffer
    int b_stages;                     // depth of the weight ring (4 .. kG2BStagesMax)
    const int32_t* row_expert;
    const int32_t* n_rows_dev;
    __nv_bfloat16* Y;
    const float* scale;
    const __nv_bfloat16* res;
    float res_a, res_b;
    int act;
    int32_t wrow[kG2MaxE];
    uint8_t kclass[kG2MaxE];
    int32_t* sched;                   // [0] next tile, [1] finished CTAs (self-resetting)
    // per kernel-size class: tile j covers M-tiles [j*mt_base + min(j, mt_extra), +mt_base + (j < mt_extra))
    int32_t ksize[kG2Classes], wp[kG2Classes], box_bytes[kG2Classes];
    int32_t ntile[kG2Classes], mt_base[kG2Classes], mt_extra[kG2Classes];
};

struct G2Tile {
    int r, j, e, kc;                  // row (sample), tile of the sample, expert, kernel-size class
    int mt_n, p0, h0, c0;             // M-tiles, first position, its image row / offset inside that row's box
};

// weight ring: four stages, deepened (launch parameter b_stages) until it holds HDMOE_G2_RING_BYTES when the halo
// buffers leave room -- the small stages of Cout = 32 / KC = 32 otherwise cover less than one TMA round trip of MMAs
#ifndef HDMOE_G2_RING_BYTES
#define HDMOE_G2_RING_BYTES 65536
#endif
constexpr int kG2BStagesMax = 16;
constexpr int g2_b_stages(int b_stage_bytes) {
    int n = HDMOE_G2_RING_BYTES / b_stage_bytes;
    return n < 4 ? 4 : (n > kG2BStagesMax ? kG2BStagesMax : n);
}

// filter taps per weight stage (one TMA box of TPS * N rows): stages of ~16 KiB, so that the producer's per-stage work
// (barrier wait, expect_tx, TMA issue from a single lane) is spread over >= 8 MMAs per issuer also at Cout = 32 / KC = 32
constexpr int g2_tps(int kc, int n) {
    int t = 16384 / (n * kc * 2);
    t = t < 1 ? 1 : (t > 4 ? 4 : t);
    while (t > 1 && t * n > 256) --t;
    return t;
}

template <int KC, int N>
struct Conv2Cfg {
    static constexpr int MT_MAX = N <= 64 ? 3 : 2;             // M-tiles (of 128 positions) per strip (<= kG2Issuers)
    static constexpr int ROWB = KC * 2;                        // bytes per position
    static constexpr int TPS = g2_tps(KC, N);
    static constexpr int B_TAP = N * KC * 2;                   // bytes of one tap's [N x KC] weight tile
    static constexpr int B_STAGE = TPS * B_TAP;
    static constexpr int A_STAGES = 2;
    static constexpr int B_STAGES_MIN = 4;
    static constexpr int EPI_NB = (N % 64 == 0) ? 2 : 1;      // 32-column accumulator loads in flight per epilogue thread
    static constexpr int TMEM_NEED = 2 * MT_MAX * N;
    static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512));
};

// 32-byte global accesses (sm_100: LDG/STG.256)
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t* v) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue of one tile for one thread: position q0 + 128*mt of the flattened padded image, accumulator columns
// [tc0 + mt*N, +N).  TMEM -> registers -> (scale, mp_silu, mp_sum residual) -> bf16 -> 32-byte global stores.
template <int N, bool SC, bool ACT, bool RES>
__device__ __forceinline__ void g2_epilogue(const GConv2Params& p, const G2Tile& t, int q0, uint32_t tc0) {
    const int Wp = p.wp[t.kc];
    const float* sc = SC ? p.scale + (size_t)t.r * N : nullptr;
    for (int mt = 0; mt < t.mt_n; ++mt) {
        // this thread's position -> output pixel (or a padding column / the tail past the image)
        const int pa = q0 + mt * 128;
        const int hl = pa / Wp, w = pa - hl * Wp;
        const bool valid = hl < p.H && w < p.W;
        const size_t go = (((size_t)t.r * p.H + (size_t)hl) * p.W + w) * N;
        const uint32_t tcol = tc0 + (uint32_t)(mt * N);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32_issue(tcol + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 2; ++g) {                      // 16 channels = one 32-byte store
                uint32_t packed[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float a = __uint_as_float(v[g * 16 + 2 * u]), b = __uint_as_float(v[g * 16 + 2 * u + 1]);
                    if (SC) {
                        const float2 s2 = __ldg(reinterpret_cast<const float2*>(sc + c0 + g * 16 + 2 * u));
                        a *= s2.x;
                        b *= s2.y;
                    }
                    if (ACT) {
                        a = __fdividef(a, 1.f + __expf(-a)) * (1.f / 0.596f);
                        b = __fdividef(b, 1.f + __expf(-b)) * (1.f / 0.596f);
                    }
                    __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
                    packed[u] = *reinterpret_cast<uint32_t*>(&o);
                }
                if (valid) {
                    if (RES) {   // mp_sum folded: out = res_a * residual + res_b * value (value rounded to bf16 first)
                        uint32_t rr[8];
                        ld_global_v8(p.res + go + c0 + g * 16, rr);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float v0 = __uint_as_float(packed[u] << 16), v1 = __uint_as_float(packed[u] & 0xffff0000u);
                            const float r0 = __uint_as_float(rr[u] << 16), r1 = __uint_as_float(rr[u] & 0xffff0000u);
                            __nv_bfloat162 o = __floats2bfloat162_rn(p.res_a * r0 + p.res_b * v0, p.res_a * r1 + p.res_b * v1);
                            packed[u] = *reinterpret_cast<uint32_t*>(&o);
                        }
                    }
                    st_global_v8(p.Y + go + c0 + g * 16, packed);
                }
            }
        }
    }
}

template <int KC, int N>
__global__ void __launch_bounds__(kG2Threads, 1)
gconv2_fwd_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
                  const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap ta3,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ GConv2Params p) {
    using Cfg = Conv2Cfg<KC, N>;
    extern __shared__ uint8_t smem_raw[];
#ifdef HDMOE_G2_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 148) g2_span[blockIdx.x * 2] = g2_gtime();
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;                                              // A_STAGES halo buffers
    uint8_t* b_buf = smem + (size_t)Cfg::A_STAGES * p.a_stage_bytes;    // weight ring
    __shared__ __align__(8) uint64_t a_full[Cfg::A_STAGES], a_empty[Cfg::A_STAGES], b_full[kG2BStagesMax],
        b_empty[kG2BStagesMax], t_full[2], t_empty[2], q_full[kG2Queue], q_empty[kG2Queue];
    const int BS = p.b_stages;
    __shared__ int32_t tile_q[kG2Queue];
    // what the MMA issuers need of a tile, decoded once by the producer: kernel size (0 = not a compute tile), padded
    // width, offset of the first position inside its box row, M-tiles
    __shared__ int4 tile_geo[kG2Queue];
    __shared__ uint32_t tmem_base_s;
    // expert of every row, staged once: each role decodes every tile, and a global load per tile (~700 cycles of L2
    // latency in front of the first MMA of the tile) showed up as a 1 100-cycle gap between tiles
    __shared__ int8_t row_e_s[kG2RowCache];
    const int cap_rows = p.n_tiles / p.smax;
    const bool rows_cached = cap_rows <= kG2RowCache;
    const int n_rows = *p.n_rows_dev;                  // requested first: its L2 round trip overlaps the whole set-up
    if (rows_cached)
        for (int r = threadIdx.x; r < cap_rows; r += kG2Threads) {
            const int e = p.row_expert[r];
            row_e_s[r] = (int8_t)((e < 0 || e >= p.n_experts) ? -1 : e);
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 32) {   // descriptor fetches overlap the barrier / TMEM / row-cache set-up
        asm volatile("prefetch.tensormap [%0];" ::"l"(&ta0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&ta1) : "memory");
    }
    // barrier set-up spread over the threads of warps 0, 2 and 3 (48+ serial mbarrier.init by one thread were ~1 000
    // cycles of every CTA's prologue)
    {
        const int x = threadIdx.x;
        bool did = false;
        if (x < kG2Queue) {
            mb_init(&q_full[x], 1);
            mb_init(&q_empty[x], kG2Issuers + 4);
            did = true;
        } else if (x >= 64 && x < 64 + BS) {
            mb_init(&b_full[x - 64], 1);
            mb_init(&b_empty[x - 64], kG2Issuers);
            did = true;
        } else if (x >= 96 && x < 96 + Cfg::A_STAGES) {
            mb_init(&a_full[x - 96], 1);
            mb_init(&a_empty[x - 96], kG2Issuers);
            did = true;
        } else if (x >= 100 && x < 102) {
            mb_init(&t_full[x - 100], kG2Issuers);
            mb_init(&t_empty[x - 100], 4);
            did = true;
        }
        if (did) {
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&tmem_base_s)),
                     "n"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // tile id -> geometry.  Row-major with the rows reversed: the heavy (large-kernel) experts sit at the end of the
    // expert-major row order and are handed out first (longest-processing-time-first for the dynamic scheduler).
    auto tile_at = [&](int i, G2Tile& t) -> bool {
        const int q = i / p.smax;
        t.j = i - q * p.smax;
        t.r = cap_rows - 1 - q;
        t.e = -1;
        t.kc = 0;
        if (t.r >= n_rows) return false;
        t.e = rows_cached ? (int)row_e_s[t.r] : p.row_expert[t.r];
        if (t.e < 0 || t.e >= p.n_experts) return false;
        t.kc = p.kclass[t.e];
        if (t.j >= p.ntile[t.kc]) return false;
        const int base = p.mt_base[t.kc], extra = p.mt_extra[t.kc], Wp = p.wp[t.kc];
        t.mt_n = base + (t.j < extra ? 1 : 0);
        t.p0 = 128 * (t.j * base + min(t.j, extra));
        t.h0 = t.p0 / Wp;
        t.c0 = t.p0 - t.h0 * Wp;
        return true;
    };
    // consumer side of the tile queue (one elected lane per consumer warp arrives on q_empty)
    int qs = 0;
    uint32_t qph = 0;
    auto next_tile = [&](bool whole_warp) -> int {
        mb_wait(&q_full[qs], qph);
        const int i = tile_q[qs];
        if (whole_warp) __syncwarp();
        if (lane == 0) mb_arrive(&q_empty[qs]);
        if (++qs == kG2Queue) {
            qs = 0;
            qph ^= 1;
        }
        return i;
    };

    if (warp == 0) {
        // ============================== scheduler + TMA producer ==============================
        if (lane == 0) {
            const CUtensorMap* maps[kG2Classes] = {&ta0, &ta1, &ta2, &ta3};
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            // The producer walks the sequence of A items = (tile, channel chunk).  The halo box of item m+1 is
            // requested while the weights of item m are still streaming (after B_STAGES weight stages of item m, when
            // item m-1 has provably drained and its halo buffer is free), so the MMAs never wait for it.
            G2Tile cur, nxt;
            int cur_c = 0, nxt_c = 0;
            bool more = true;                       // the scheduler still has tiles
            bool first_draw = true;                 // tile blockIdx.x is this CTA's without asking (saves one L2 round trip)
            auto advance = [&](const G2Tile& from, int from_c, bool first, G2Tile& to, int& to_c) -> bool {
                if (!first && from_c + 1 < p.upt) {
                    to = from;
                    to_c = from_c + 1;
                    return true;
                }
                while (more) {
                    int i = first_draw ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(p.sched, 1);
                    first_draw = false;
                    if (i >= p.n_tiles) i = -1;
                    const bool ok = i >= 0 && tile_at(i, to);
                    // the other roles see valid tiles, the zero-fill tiles of unused tail rows and the end marker; the
                    // empty tile slots of a kernel-size class with fewer tiles than smax are dropped here.  (The
                    // producer may block on a full queue while weight stages of the current tile are outstanding, so a
                    // run of non-compute entries between two compute tiles must stay below the queue depth: unused
                    // rows come first in the reversed row order, before any tile is in flight.)
                    if (ok || i < 0 || to.r >= n_rows || to.e < 0) {
                        mb_wait(&q_empty[qs], qph ^ 1);
                        tile_q[qs] = i;
                        tile_geo[qs] = ok ? make_int4(p.ksize[to.kc], p.wp[to.kc], to.c0, to.mt_n) : make_int4(0, 0, 0, 0);
                        mb_arrive(&q_full[qs]);
                        if (++qs == kG2Queue) {
                            qs = 0;
                            qph ^= 1;
                        }
                    }
                    if (i < 0) {
                        more = false;
                        // this CTA draws no more tiles: the last CTA to get here re-arms the scheduler for the next
                        // launch on this stream (done now, under the remaining tiles, not in the kernel's tail)
                        __threadfence();
                        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
                            p.sched[0] = 0;
                            p.sched[1] = 0;
                            __threadfence();
                        }
                        break;
                    }
                    if (ok) {
                        to_c = 0;
                        return true;
                    }
                }
                return false;
            };
            auto load_a = [&](const G2Tile& t, int c) {
                const int pad = (p.ksize[t.kc] - 1) >> 1;
                mb_wait(&a_empty[as], aph ^ 1);
                mb_expect_tx(&a_full[as], (uint32_t)p.box_bytes[t.kc]);
                tma_load_4d(a_buf + (size_t)as * p.a_stage_bytes, maps[t.kc], &a_full[as], c * KC, -pad, t.h0 - pad, t.r);
                if (++as == Cfg::A_STAGES) {
                    as = 0;
                    aph ^= 1;
                }
            };
            bool have = advance(cur, 0, true, cur, cur_c);
            if (have) load_a(cur, cur_c);
            while (have) {
                const int k = p.ksize[cur.kc], taps = k * k;
                const int wrow = p.wrow[cur.e];
                const int nst = (taps + Cfg::TPS - 1) / Cfg::TPS;
                const int pre = nst - 1 < BS ? nst - 1 : BS;
                bool have_next = false;
                for (int s = 0; s < nst; ++s) {
                    if (s == pre) {
                        have_next = advance(cur, cur_c, false, nxt, nxt_c);
                        if (have_next) load_a(nxt, nxt_c);
                    }
                    mb_wait(&b_empty[bs], bph ^ 1);
                    mb_expect_tx(&b_full[bs], (uint32_t)Cfg::B_STAGE);
                    // TPS consecutive taps are TPS*N consecutive rows of the tap-major weight block: one box
                    // (a trailing odd tap drags in N rows of the next block / OOB zeros; they are not used)
                    tma_load_2d(b_buf + (size_t)bs * Cfg::B_STAGE, &tmap_b, &b_full[bs], cur_c * KC, wrow + s * Cfg::TPS * N);
                    if (++bs == BS) {
                        bs = 0;
                        bph ^= 1;
                    }
                }
                have = have_next;
                cur = nxt;
                cur_c = nxt_c;
            }
        }
    } else if (warp <= kG2Issuers) {
        // ============================== MMA issuers (warp w owns M-tile w-1) ==============================
        // The WHOLE warp runs this loop converged and one elected lane issues.  Tile geometry is made warp-uniform with
        // redux (`uni`), so descriptors and barrier addresses live in uniform registers: a loop entered by a single
        // lane makes the compiler wrap every UTCHMMA in an ELECT + 3 x R2UR + branch waterfall and keeps the
        // descriptor arithmetic on the vector datapath (~135-160 cycles per MMA per issuing thread, measured: the
        // kernel ran at 59 / 68 / 82 cycles per MMA at (N, KC) = (64, 64) / (32, 64) / (32, 32), i.e. issue-bound,
        // against the 48 / 40 / 40 cycles of the shared-memory operand feed).
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const int mt = uni(warp) - 1;
        int tcount = 0;
        (void)tcount;
        int as = 0, bs = 0, acc = 0;
        uint32_t aph = 0, bph = 0, acc_ph = 0;
        // The producer publishes every tile with its geometry already decoded (tile_geo): re-deriving it here (constant-
        // bank table reads, two divisions) left the tensor pipe idle for ~1 000 cycles between tiles.
        struct TileU { int i, k, Wp, c0, active; };
        auto fetch = [&]() -> TileU {
            TileU u{-1, 0, 0, 0, 0};
            for (;;) {
                mb_wait(&q_full[qs], qph);
                const int i = tile_q[qs];
                const int4 g = tile_geo[qs];
                __syncwarp();
                if (lane == 0) mb_arrive(&q_empty[qs]);
                if (++qs == kG2Queue) {
                    qs = 0;
                    qph ^= 1;
                }
                u.i = uni(i);
                if (u.i < 0) return u;
                u.k = uni(g.x);
                if (u.k == 0) continue;                    // zero-fill entries of unused rows: the epilogue's business
                u.Wp = uni(g.y);
                u.c0 = uni(g.z);
                u.active = uni((int)(mt < g.w));
                return u;
            }
        };
        TileU nx = fetch();
        for (;;) {
            if (nx.i < 0) break;
            const int k = nx.k, Wp = nx.Wp, taps = k * k, c0u = nx.c0;
            const bool active = nx.active != 0;
            if (mt == 0 && lane == 0) G2T(0);
            mb_wait(&t_empty[acc], acc_ph ^ 1);
            tc_fence_after();
            if (mt == 0 && lane == 0) G2T(1);
            const uint32_t d = tmem_base + (uint32_t)((acc * Cfg::MT_MAX + mt) * N);
            for (int c = 0; c < p.upt; ++c) {
                mb_wait(&a_full[as], aph);
                tc_fence_after();
                if (mt == 0 && c == 0 && lane == 0) G2T(2);
                // descriptor of this M-tile's first position; taps / k-slices only add to the 14-bit address field
                const uint64_t a_desc0 =
                    umma_desc<KC>(s2u(a_buf + (size_t)as * p.a_stage_bytes) + (uint32_t)(c0u + mt * 128) * Cfg::ROWB);
                int tr = 0, ts = 0;                              // tap (tr, ts) of the next MMA, walked without division
                for (int t0 = 0; t0 < taps; t0 += Cfg::TPS) {
                    mb_wait(&b_full[bs], bph);
                    tc_fence_after();
                    const uint64_t bd0 = umma_desc<KC>(s2u(b_buf + (size_t)bs * Cfg::B_STAGE));
                    uint64_t ad[Cfg::TPS];
                    bool live[Cfg::TPS];
#pragma unroll
                    for (int q = 0; q < Cfg::TPS; ++q) {
                        live[q] = t0 + q < taps;
                        ad[q] = a_desc0 + (uint64_t)(((uint32_t)(tr * Wp + ts) * Cfg::ROWB) >> 4);
                        if (++ts == k) {
                            ts = 0;
                            ++tr;
                        }
                    }
                    if (elect_one()) {
                        if (active) {
#pragma unroll
                            for (int q = 0; q < Cfg::TPS; ++q)
                                if (live[q]) {
                                    const uint64_t bd = bd0 + (uint64_t)((q * Cfg::B_TAP) >> 4);
#pragma unroll
                                    for (int kk = 0; kk < KC / 16; ++kk)
                                        tc_mma(d, ad[q] + 2 * kk, bd + 2 * kk, idesc, (uint32_t)(c | (t0 + q) | kk));
                                }
                            tc_commit(&b_empty[bs]);
                        } else {
                            mb_arrive(&b_empty[bs]);
                        }
                    }
                    __syncwarp();
                    if (++bs == BS) {
                        bs = 0;
                        bph ^= 1;
                    }
                }
                if (elect_one()) {
                    if (active) tc_commit(&a_empty[as]);
                    else mb_arrive(&a_empty[as]);
                }
                __syncwarp();
                if (++as == Cfg::A_STAGES) {
                    as = 0;
                    aph ^= 1;
                }
            }
            if (elect_one()) {
                if (active) tc_commit(&t_full[acc]);
                else mb_arrive(&t_full[acc]);
            }
            __syncwarp();
            if (mt == 0 && lane == 0) G2T(3);
            ++tcount;
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
            nx = fetch();            // published by the producer before the last weight stage of the tile just issued
        }
    } else {
        // ============================== epilogue (4 warps): TMEM -> registers -> global ==============================
        const int quad = warp & 3;
        int acc = 0;
        int tcount = 0;
        (void)tcount;
        uint32_t acc_ph = 0;
        for (;;) {
            const int i = next_tile(true);
            if (i < 0) break;
            G2Tile t;
            if (!tile_at(i, t)) {
                // tiles of an unused tail row: zero-fill so downstream elementwise ops stay finite
                if (t.r >= n_rows || t.e < 0) {
                    const int rows_per = (p.H + p.smax - 1) / p.smax;
                    const int hs = t.j * rows_per, he = min(p.H, hs + rows_per);
                    const long long n16 = (long long)(he - hs) * p.W * N / 8;
                    int4* o = reinterpret_cast<int4*>(p.Y + ((size_t)t.r * p.H + hs) * p.W * N);
                    for (long long q = quad * 32 + lane; q < n16; q += 128) o[q] = make_int4(0, 0, 0, 0);
                }
                continue;
            }
            const int Wp = p.wp[t.kc];
            mb_wait(&t_full[acc], acc_ph);
            tc_fence_after();
            if (quad == 0 && lane == 0) G2T(4);
            // one compact, branch-free instantiation per (scale, activation, residual) combination: the fully unrolled
            // generic body was ~2 100 instructions per M-tile and ran at instruction-fetch speed (12 k cycles per tile)
            const int fl = (p.scale ? 1 : 0) | (p.act == 1 ? 2 : 0) | (p.res ? 4 : 0);
            const int q0 = t.p0 + quad * 32 + lane;
            const uint32_t tc0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * Cfg::MT_MAX * N);
            switch (fl) {
                case 0: g2_epilogue<N, false, false, false>(p, t, q0, tc0); break;
                case 1: g2_epilogue<N, true, false, false>(p, t, q0, tc0); break;
                case 2: g2_epilogue<N, false, true, false>(p, t, q0, tc0); break;
                case 3: g2_epilogue<N, true, true, false>(p, t, q0, tc0); break;
                case 4: g2_epilogue<N, false, false, true>(p, t, q0, tc0); break;
                case 5: g2_epilogue<N, true, false, true>(p, t, q0, tc0); break;
                case 6: g2_epilogue<N, false, true, true>(p, t, q0, tc0); break;
                default: g2_epilogue<N, true, true, true>(p, t, q0, tc0); break;
            }
            tc_fence_before();
            __syncwarp();
            if (quad == 0 && lane == 0) G2T(5);
            ++tcount;
            if (lane == 0) mb_arrive(&t_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
    }
#ifdef HDMOE_G2_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 148) g2_span[blockIdx.x * 2 + 1] = g2_gtime();
#endif
}

template <int KC, int N>
static int launch_gconv2(const CUtensorMap* ta, const CUtensorMap& tb, const GConv2Params& p, cudaStream_t st) {
    using Cfg = Conv2Cfg<KC, N>;
    auto kfn = gconv2_fwd_kernel<KC, N>;
    const int smem = Cfg::A_STAGES * p.a_stage_bytes + p.b_stages * Cfg::B_STAGE + 1024;
    HDMOE_CHECK_ARG(smem <= 227 * 1024, "gconv2: tile does not fit shared memory (%d bytes)", smem);
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
    kfn<<<grid, kG2Threads, smem, st>>>(ta[0], ta[1], ta[2], ta[3], tb, p);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

}  // namespace hdmoe
using namespace hdmoe;

#ifdef HDMOE_G2_TRACE
extern "C" int hdmoe_g2_trace_read(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g2_trace, sizeof(long long) * 148 * 64);
}
extern "C" int hdmoe_g2_span_read(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g2_span, sizeof(long long) * 148 * 2);
}
#endif

extern "C" int hdmoe_gconv2_fwd(const void* X, const void* Wt, void* Y, int cap_rows, int H, int W, int Cin_pad,
                                int Cout, int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev,
                                int n_experts, const int32_t* ksize_host, const int32_t* wrow_host, const float* scale,
                                int act, const void* residual, float res_a, float res_b, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(X && Wt && Y && row_expert && n_rows_dev && ksize_host && wrow_host, "gconv2_fwd: null pointer");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= kG2MaxE, "gconv2_fwd: 1 <= n_experts <= %d", kG2MaxE);
    HDMOE_CHECK_ARG(Cout == 32 || Cout == 64 || Cout == 96 || Cout == 128, "gconv2_fwd: Cout must be 32, 64, 96 or 128");
    HDMOE_CHECK_ARG(Cin_pad >= 32 && Cin_pad % 32 == 0, "gconv2_fwd: Cin_pad must be a multiple of 32 (got %d)", Cin_pad);
    HDMOE_CHECK_ARG(H >= 1 && H <= 255 && W >= 1 && W <= 248, "gconv2_fwd: H <= 255, W <= 248");
    HDMOE_CHECK_ARG((((uintptr_t)X | (uintptr_t)Wt) & 15) == 0, "gconv2_fwd: 16-byte alignment required");
    HDMOE_CHECK_ARG((((uintptr_t)Y | (uintptr_t)residual) & 31) == 0, "gconv2_fwd: Y / residual need 32-byte alignment");
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) {
        set_error("gconv2_fwd: cuTensorMapEncodeTiled not available from the driver");
        return HDMOE_ERR_CUDA;
    }
    const int KC = (Cin_pad % 64 == 0) ? 64 : 32;
    const int mt_max = Cout <= 64 ? 3 : 2;
    GConv2Params p{};
    p.H = H;
    p.W = W;
    p.upt = Cin_pad / KC;
    p.n_experts = n_experts;
    p.row_expert = row_expert;
    p.n_rows_dev = n_rows_dev;
    p.Y = (__nv_bfloat16*)Y;
    p.scale = scale;
    p.act = act;
    p.res = (const __nv_bfloat16*)residual;
    p.res_a = res_a;
    p.res_b = res_b;
    // kernel-size classes and their strip tables
    int ncls = 0, cls_k[kG2Classes];
    for (int e = 0; e < n_experts; ++e) {
        const int k = ksize_host[e];
        HDMOE_CHECK_ARG(k >= 1 && k <= 7 && (k & 1), "gconv2_fwd: odd kernel sizes 1..7");
        int c = -1;
        for (int q = 0; q < ncls; ++q)
            if (cls_k[q] == k) c = q;
        if (c < 0) {
            HDMOE_CHECK_ARG(ncls < kG2Classes, "gconv2_fwd: at most %d distinct kernel sizes per launch", kG2Classes);
            c = ncls++;
            cls_k[c] = k;
        }
        p.kclass[e] = (uint8_t)c;
        p.wrow[e] = wrow_host[e];
    }
    int smax = 0, box_bytes_max = 0;
    int cls_box_rows[kG2Classes];
    const int tps = g2_tps(KC, Cout);
    const int b_stage_bytes = tps * Cout * KC * 2;                            // Conv2Cfg::B_STAGE
    const int b_ring_bytes = Conv2Cfg<32, 32>::B_STAGES_MIN * b_stage_bytes;  // the minimum ring decides the tile size
    // M-tiles per tile: as many as there are issuer warps, fewer only if two halo buffers would not fit shared memory
    for (int mt_try = mt_max; mt_try >= 1; --mt_try) {
        smax = 0;
        box_bytes_max = 0;
        for (int c = 0; c < ncls; ++c) {
            const int k = cls_k[c], Wp = W + k - 1;
            // M-tiles of 128 positions over the flattened padded image [0, H*Wp), split as evenly as possible into
            // tiles of <= mt_try M-tiles (tiles need not start on an image row)
            const int m_total = (H * Wp + 127) / 128;
            const int n = (m_total + mt_try - 1) / mt_try;
            const int base = m_total / n, extra = m_total % n;
            // input rows one tile's TMA box must hold: the tile's positions plus the largest tap offset
            // (k-1)*(Wp+1), counted from the start of the image row that contains the tile's first position
            int box_rows = 0;
            for (int j = 0; j < n; ++j) {
                const int m0 = j * base + (j < extra ? j : extra), mt_n = base + (j < extra ? 1 : 0);
                const int c0 = (128 * m0) % Wp;
                const int rows = (c0 + 128 * mt_n - 1 + (k - 1) * (Wp + 1)) / Wp + 1;
                if (rows > box_rows) box_rows = rows;
            }
            HDMOE_CHECK_ARG(box_rows <= 256 && Wp <= 256, "gconv2_fwd: TMA box too large");
            p.ksize[c] = k;
            p.wp[c] = Wp;
            p.ntile[c] = n;
            p.mt_base[c] = base;
            p.mt_extra[c] = extra;
            cls_box_rows[c] = box_rows;
            p.box_bytes[c] = box_rows * Wp * KC * 2;
            if (n > smax) smax = n;
            if (p.box_bytes[c] > box_bytes_max) box_bytes_max = p.box_bytes[c];
        }
        if (2 * (((box_bytes_max + 1023) / 1024) * 1024) + b_ring_bytes + 1024 <= 227 * 1024) break;
    }
    p.smax = smax;
    p.n_tiles = cap_rows * smax;
    p.a_stage_bytes = ((box_bytes_max + 1023) / 1024) * 1024;
    p.b_stages = g2_b_stages(b_stage_bytes);          // deeper ring where the halo buffers leave room
    while (p.b_stages > 4 && 2 * p.a_stage_bytes + p.b_stages * b_stage_bytes + 1024 > 227 * 1024) --p.b_stages;
    cudaStream_t st = (cudaStream_t)stream;
    p.sched = sched_slot(st);
    HDMOE_CHECK_ARG(p.sched != nullptr, "gconv2_fwd: more than %d distinct streams in use", kSchedSlots);
    CUtensorMap ta[kG2Classes], tb;
    const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int c = 0; c < kG2Classes; ++c) {
        const int cc = c < ncls ? c : 0;
        const int k = cls_k[cc], Wp = W + k - 1;
        cuuint64_t dims[4] = {(cuuint64_t)Cin_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap_rows};
        cuuint64_t strides[3] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)W * Cin_pad * 2, (cuuint64_t)H * W * Cin_pad * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)Wp, (cuuint32_t)cls_box_rows[cc], 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&ta[c], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("gconv2_fwd: cuTensorMapEncodeTiled(A, class %d) failed with %d", c, (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin_pad, (cuuint64_t)w_rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)Cin_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(tps * Cout)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Wt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("gconv2_fwd: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
#define GC2(KCV, NV) \
    if (KC == KCV && Cout == NV) return launch_gconv2<KCV, NV>(ta, tb, p, st);
    GC2(32, 32) GC2(32, 64) GC2(32, 96) GC2(32, 128) GC2(64, 32) GC2(64, 64) GC2(64, 96) GC2(64, 128)
#undef GC2
    HDMOE_CHECK_ARG(false, "gconv2_fwd: unsupported (KC=%d, Cout=%d)", KC, Cout);
}
