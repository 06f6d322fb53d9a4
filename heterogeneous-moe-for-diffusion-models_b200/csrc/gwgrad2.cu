// G-WGRAD v2: grouped convolution weight gradient on tcgen05 -- taps stacked in M, double-buffered accumulators.
//
//   dW_e[tap][o][c] = sum over rows r of expert e, output pixels q:  dY[r, q, o] * Xpad[r, q + delta_tap, c]
//
// v1 (gwgrad.cu) runs one M = 64 (= Cout) MMA per (tap, 16 positions): half the tensor data path, 34.6 cycles each,
// and its single accumulator set makes the MMAs wait while the epilogue flushes (30 % of the kernel at 32x32, 57 %
// at 16x16, cycle accounting in profiles/).  v2 changes two things:
//
//   * TAPS STACKED IN M.  Both operands are MN-major (pixel-major NHWC): an M = 128 A operand is TPM = 128 / Cout
//     swizzle atoms of Cout elements, LBO bytes apart.  With LBO = ONE POSITION ROW the atoms are the same dY strip
//     read at consecutive position offsets, and since
//         sum_q dY[q] X[q + off + s]  =  sum_q' dY[q' - s] X[q' + off]
//     consecutive taps of one kernel row share ONE B operand (the input window at the row's first tap):
//     atom a of A = dY shifted by a - (TPM-1) positions  <->  tap s0 + TPM-1 - a.  One M = 128 x N = Cin-chunk MMA per
//     TPM taps, 48 cycles at N = 64 (tools/umma_rate.cu, "MN-major LBO=r") instead of 2 x 34.6.  The positions a
//     shifted atom drops at the end of a strip are padding columns of dY (zero), the ones it adds at the front are a
//     zeroed lead-in in front of the A box.
//   * KERNEL ROWS STACKED IN N (round 2).  The B operand is MN-major as well, so an N = nr * KC operand is nr swizzle
//     atoms LBO bytes apart; with LBO = ONE PADDED IMAGE ROW (Wp positions) atom j is the same input window read one
//     kernel row further down, i.e. column block j of the accumulator is kernel row tr0 + j.  One M = 128 x N = 128 MMA
//     then covers up to 128/Cout x 128/KC taps at the full tensor rate (64 cycles) instead of one N = KC MMA per kernel
//     row at the operand-feed-bound 40 / 48 cycles.  Unlike the forward's tap stacking (tools/legacy/gconv3.cu) nothing
//     has to be shifted or added afterwards: every (lane block, column block) pair is its own tap of dW.  Only whole
//     kernel rows are stacked (N = exactly the rows that exist), so no operand read leaves the staged window.
//   * MMAs DEALT TO THE ISSUER WARPS BY (unit, K slice), all accumulating.  With rows stacked in N a tap group holds
//     only 2-3 wide units, fewer than issuer warps, and one thread sustains one MMA per ~100 cycles; so every issuer
//     walks every unit and takes every 4th position slice.  Several warps then feed one accumulator, which forbids a
//     "first MMA overwrites" flag (no order between warps): every MMA accumulates, and the accumulators are zeroed by
//     their reader -- once in the prologue, then by the epilogue right after it has read them.
//   * DOUBLE-BUFFERED ACCUMULATORS.  A tap group is sized to 256 TMEM columns; consecutive items alternate between
//     the two halves, so the flush (TMEM -> vector atomics, all 128 lanes useful now) of item i runs under the MMAs
//     of item i + 1.
//
// Everything else follows v1: flattened halo formulation (A = dY strip box with zero-filled surplus columns, B = the
// zero-padded input window, tap (r, s) reads it at start + (r * Wp + s) rows), item = (row chunk, tap group), split-K
// over row chunks with `red.global.add.v4.f32`, dynamic item scheduler (last rows first), 4 MMA-issuer warps.
#include <algorithm>
#include <type_traits>

#include "tc.cuh"
#include "../../include/hdmoe_gemm.h"

namespace hdmoe {

// cycle accounting of issuer warp 0 (-DHDMOE_WG_TRACE, read with tools/dbg_wgrad.py)
#ifdef HDMOE_WG_TRACE
__device__ long long wg2_trace[148 * 16];
#define WGT_DECL long long wt_[7] = {0, 0, 0, 0, 0, 0, 0}; long long wt0_ = clock64(); (void)wt0_
#define WGT_LAP(i) do { const long long n_ = clock64(); wt_[i] += n_ - wt0_; wt0_ = n_; } while (0)
#define WGT_CNT(i) do { wt_[i] += 1; } while (0)
#define WGT_OUT(base) do { for (int q_ = 0; q_ < 7; ++q_) wg2_trace[blockIdx.x * 16 + (base) + q_] = wt_[q_]; } while (0)
#else
#define WGT_DECL do { } while (0)
#define WGT_LAP(i) do { } while (0)
#define WGT_CNT(i) do { } while (0)
#define WGT_OUT(base) do { } while (0)
#endif

constexpr int kW2Issuers = 4;
constexpr int kW2Threads = 32 * (1 + kW2Issuers + 4);
constexpr int kW2Classes = 4;
constexpr int kW2MaxE = HDMOE_MAX_EXPERTS;
constexpr int kW2Stages = 2;
constexpr int kW2Queue = 8;
constexpr int kW2Lead = 1024;          // zeroed bytes in front of the A box (>= (TPM-1) position rows)
constexpr int kW2BufCols = 256;        // TMEM columns per accumulator buffer
constexpr int kW2MaxUnits = 32;        // units (column-tap block x kernel-row block) per kernel-size class (k = 7, Cout = 64,
                                       // Cin_pad >= 192: 4 column blocks x 7 single rows = 28)

struct WGrad2Params {
    int n_items, gmax, rows_per_item, cap_rows;
    int H, W, SH;                      // strip height (rows), H % SH == 0
    int nchunks;                       // Cin_pad / KC
    int cin_pad;
    int n_experts;
    int a_stage_bytes, b_stage_bytes;  // a_stage_bytes includes the lead-in
    int cout_total, o_off;             // dY / dW channel count and this launch's first output channel (Cout = 128: 2 passes)
    const int32_t* row_expert;
    const int32_t* n_rows_dev;
    float* dW;                         // fp32 [w_rows_total][cin_pad], tap-major blocks per expert (accumulated)
    int32_t* sched;                    // [0] next item, [1] finished CTAs (self-resetting, core.cu)
    int32_t wrow[kW2MaxE];
    uint8_t kclass[kW2MaxE];
    int32_t ksize[kW2Classes], wp[kW2Classes], ngroups[kW2Classes];
    int32_t a_box_bytes[kW2Classes], b_box_bytes[kW2Classes];
    // unit i of class c: TPM consecutive taps (columns s0 = u_cu * TPM ...) of the u_nr kernel rows starting at u_tr0;
    // its accumulators are [chunk][row][KC] columns starting at u_col inside the group's TMEM buffer;
    // tap group g = units [g_lo[g], g_lo[g + 1])
    uint8_t u_cu[kW2Classes][kW2MaxUnits], u_tr0[kW2Classes][kW2MaxUnits], u_nr[kW2Classes][kW2MaxUnits];
    uint16_t u_col[kW2Classes][kW2MaxUnits];
    uint8_t g_lo[kW2Classes][kW2MaxUnits + 1];
};

// MN-major operand descriptor: rows are K (positions) of ROWB bytes (64 -> SW64, 128 -> SW128); LBO = byte distance
// between consecutive swizzle atoms along M / N (0 when the operand is a single atom wide)
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_mn2(uint32_t smem_addr, uint32_t lbo_bytes) {
    constexpr uint64_t sbo = (8 * ROWB) >> 4;
    constexpr uint64_t layout = ROWB == 128 ? 2 : 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (sbo << 32) | (1ull << 46) |
           (layout << 61);
}
__device__ __forceinline__ void red_add_v4_2(float* addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// COUT in {32, 64}: A rows are COUT*2 bytes, TPM = 128 / COUT taps per MMA; KC in {32, 64}: B rows are KC*2 bytes
template <int COUT, int KC>
__global__ void __launch_bounds__(kW2Threads, 1)
gwgrad2_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
               const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap ta3,
               const __grid_constant__ CUtensorMap tb0, const __grid_constant__ CUtensorMap tb1,
               const __grid_constant__ CUtensorMap tb2, const __grid_constant__ CUtensorMap tb3,
               const __grid_constant__ WGrad2Params p) {
    constexpr int ROWA = COUT * 2, ROWB_ = KC * 2, TPM = 128 / COUT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[kW2Stages], empty[kW2Stages], t_full[2], t_empty[2], q_full[kW2Queue],
        q_empty[kW2Queue];
    __shared__ int32_t item_q[kW2Queue];
    __shared__ uint32_t tmem_base_s;
    const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // operand buffers must never hold NaN/Inf garbage (positions past a box meet zeros of dY), and the lead-in in
    // front of every A box must be zero
    for (int i = threadIdx.x; i < kW2Stages * stage_bytes / 16; i += kW2Threads)
        reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kW2Stages; ++s) {
            mb_init(&full[s], 1);
            mb_init(&empty[s], kW2Issuers);
        }
        for (int b = 0; b < 2; ++b) {
            mb_init(&t_full[b], kW2Issuers);
            mb_init(&t_empty[b], 4);
        }
        for (int q = 0; q < kW2Queue; ++q) {
            mb_init(&q_full[q], 1);
            mb_init(&q_empty[q], kW2Issuers + 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp > kW2Issuers) {            // epilogue warps: clear both accumulator buffers (all MMAs accumulate)
        for (int c0 = 0; c0 < 2 * kW2BufCols; c0 += 32) tmem_st32_zero(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c0);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int n_rows = min(*p.n_rows_dev, p.cap_rows);
    const int nstrips = p.H / p.SH;

    // Walk of one item, identical in every role.  Calls stage_fn for every pipeline stage and flush_fn whenever the
    // accumulators must be written out.  Returns the number of flushes (= accumulator buffers consumed).
    // `converged` (a whole warp walks together): the expert ids are made warp-uniform with redux so that everything
    // derived from them stays on the uniform datapath (see the issuer role).
    auto walk = [&](int item, auto converged, auto&& stage_fn, auto&& flush_fn) {
        const int g = item % p.gmax, rc = p.n_items / p.gmax - 1 - item / p.gmax;       // last row chunks first
        const int r0 = rc * p.rows_per_item, r1 = min(r0 + p.rows_per_item, n_rows);
        int cur_e = -1, cur_kc = 0;
        bool fresh = true;
        for (int r = r0; r < r1; ++r) {
            const int e = decltype(converged)::value ? uni(p.row_expert[r]) : p.row_expert[r];
            if (e < 0 || e >= p.n_experts) continue;
            const int kc = p.kclass[e];
            if (g >= p.ngroups[kc]) continue;
            if (cur_e >= 0 && e != cur_e) {
                flush_fn(cur_e, cur_kc, g);
                fresh = true;
            }
            cur_e = e;
            cur_kc = kc;
            for (int st = 0; st < nstrips; ++st)
                for (int c = 0; c < p.nchunks; ++c) stage_fn(r, e, kc, g, st, c, fresh && st == 0);
            fresh = false;
        }
        if (cur_e >= 0) flush_fn(cur_e, cur_kc, g);
    };
    // item queue: the producer lane draws item ids from the global counter and publishes them (-1 = end)
    int qs = 0;
    uint32_t qph = 0;
    auto next_item = [&](bool whole_warp) -> int {
        mb_wait(&q_full[qs], qph);
        const int i = item_q[qs];
        if (whole_warp) __syncwarp();
        if (lane == 0) mb_arrive(&q_empty[qs]);
        if (++qs == kW2Queue) {
            qs = 0;
            qph ^= 1;
        }
        return i;
    };

    if (warp == 0) {
        // ============================== scheduler + TMA producer ==============================
        if (lane == 0) {
            const CUtensorMap* ma[kW2Classes] = {&ta0, &ta1, &ta2, &ta3};
            const CUtensorMap* mb[kW2Classes] = {&tb0, &tb1, &tb2, &tb3};
            int s = 0;
            uint32_t ph = 0;
            bool first_draw = true;
            for (;;) {
                int item = first_draw ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(p.sched, 1);
                first_draw = false;
                if (item >= p.n_items) item = -1;
                mb_wait(&q_empty[qs], qph ^ 1);
                item_q[qs] = item;
                mb_arrive(&q_full[qs]);
                if (++qs == kW2Queue) {
                    qs = 0;
                    qph ^= 1;
                }
                if (item < 0) break;
                walk(item, std::false_type{},
                     [&](int r, int e, int kc, int g, int st, int c, bool) {
                         const int pad = (p.ksize[kc] - 1) >> 1;
                         mb_wait(&empty[s], ph ^ 1);
                         mb_expect_tx(&full[s], (uint32_t)(p.a_box_bytes[kc] + p.b_box_bytes[kc]));
                         uint8_t* base = smem + (size_t)s * stage_bytes;
                         tma_load_4d(base + kW2Lead, ma[kc], &full[s], 0, 0, st * p.SH, r);
                         tma_load_4d(base + p.a_stage_bytes, mb[kc], &full[s], c * KC, -pad, st * p.SH - pad, r);
                         if (++s == kW2Stages) {
                             s = 0;
                             ph ^= 1;
                         }
                     },
                     [&](int, int, int) {});
            }
        }
    } else if (warp <= kW2Issuers) {
        // ============================== MMA issuers (units of the group dealt round-robin) ==============================
        // The whole warp walks the item converged and one elected lane issues (descriptors and barrier addresses in
        // uniform registers; a single-lane loop costs an ELECT + R2UR waterfall per UTCHMMA, see gconv2.cu).
        {
            // D[128 x KC] (+)= A^T B : A, B MN-major (bits 15, 16), M = 128, N = KC, bf16 -> fp32
            constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                        ((uint32_t)(128 >> 4) << 24);            // N = nr * KC is added per unit
            const int me = uni(warp) - 1;
            int s = 0, buf = 0;
            uint32_t ph = 0, tph[2] = {0, 0};
            bool need_buf = true;                       // the next stage is the first of a fresh accumulator buffer
            WGT_DECL;
            for (;;) {
                const int item = uni(next_item(true));
                WGT_LAP(0);                             // waiting for an item
                if (item < 0) break;
                WGT_CNT(5);
                walk(item, std::true_type{},
                     [&](int r, int e, int kc, int g, int st, int c, bool first) {
                         const int Wp = uni(p.wp[kc]);
                         const int u_lo = uni(p.g_lo[kc][g]), u_hi = uni(p.g_lo[kc][g + 1]);
                         const int nslice = (p.SH * Wp) >> 4;
                         WGT_LAP(1);                             // walk / decode
                         if (need_buf) {
                             mb_wait(&t_empty[buf], tph[buf] ^ 1);      // the epilogue has drained this buffer
                             tc_fence_after();
                             need_buf = false;
                             WGT_LAP(4);                         // waiting for the accumulator buffer (flush not hidden)
                         }
                         mb_wait(&full[s], ph);
                         tc_fence_after();
                         WGT_LAP(2);                             // waiting for the stage's TMA loads
                         const uint32_t a0 = s2u(smem + (size_t)s * stage_bytes) + kW2Lead - (TPM - 1) * ROWA;
                         const uint32_t b0 = s2u(smem + (size_t)s * stage_bytes) + p.a_stage_bytes;
                         const uint64_t ad0 = umma_desc_mn2<ROWA>(a0, ROWA);     // atoms one position row apart
                         const uint64_t bd0 = umma_desc_mn2<ROWB_>(b0, (uint32_t)(Wp * ROWB_));   // atoms one image row apart
                         const bool leader = elect_one();
                         (void)first;
                         int j0 = me;                                   // this warp's first slice of the current unit
                         for (int u = u_lo; u < u_hi; ++u) {
                             const int tr0 = uni(p.u_tr0[kc][u]), nr = uni(p.u_nr[kc][u]), s0 = uni(p.u_cu[kc][u]) * TPM;
                             const uint32_t d = tmem_base + (uint32_t)(buf * kW2BufCols + uni(p.u_col[kc][u]) + c * nr * KC);
                             const uint64_t bd = bd0 + (uint64_t)(((uint32_t)(tr0 * Wp + s0) * ROWB_) >> 4);
                             const uint32_t idesc = idesc0 | ((uint32_t)((nr * KC) >> 3) << 17);
                             if (leader) {
#pragma unroll 4
                                 for (int j = j0; j < nslice; j += kW2Issuers)
                                     tc_mma(d, ad0 + (uint64_t)((j * 16 * ROWA) >> 4), bd + (uint64_t)((j * 16 * ROWB_) >> 4),
                                            idesc, 1u);
                             }
                             // keep the deal round-robin across units: (unit index * nslice + j) % issuers == me
                             j0 = (j0 + kW2Issuers - (nslice % kW2Issuers)) % kW2Issuers;
                         }
                         if (leader) tc_commit(&empty[s]);
                         __syncwarp();
                         WGT_LAP(3);                             // issuing MMAs
                         if (++s == kW2Stages) {
                             s = 0;
                             ph ^= 1;
                         }
                     },
                     [&](int, int, int) {
                         WGT_CNT(6);
                         if (elect_one()) tc_commit(&t_full[buf]);           // all accumulators of the group are final
                         __syncwarp();
                         tph[buf] ^= 1;
                         buf ^= 1;
                         need_buf = true;
                     });
            }
            if (me == 0 && lane == 0) WGT_OUT(0);
        }
    } else {
        // ============================== epilogue: TMEM -> vector atomics ==============================
        const int quad = warp & 3;
        int buf = 0;
        uint32_t tph[2] = {0, 0};
        const int L = quad * 32 + lane;                   // accumulator row: atom a = L / COUT, channel o = L % COUT
        const int a = L / COUT, o = L - a * COUT;
        for (;;) {
            const int item = next_item(true);
            if (item < 0) break;
            walk(item, std::false_type{}, [&](int, int, int, int, int, int, bool) {},
                 [&](int e, int kc, int g) {
                     const int k = p.ksize[kc];
                     const int u_lo = p.g_lo[kc][g], u_hi = p.g_lo[kc][g + 1];
                     mb_wait(&t_full[buf], tph[buf]);
                     tph[buf] ^= 1;
                     tc_fence_after();
                     for (int u = u_lo; u < u_hi; ++u) {
                         const int tr0 = p.u_tr0[kc][u], nr = p.u_nr[kc][u], s0 = p.u_cu[kc][u] * TPM;
                         const int ts = s0 + TPM - 1 - a;               // this lane's tap (may lie past the kernel row)
                         const uint32_t ucol = (uint32_t)(buf * kW2BufCols + p.u_col[kc][u]);
                         for (int c = 0; c < p.nchunks; ++c) {
                             for (int j = 0; j < nr; ++j) {             // column block j = kernel row tr0 + j
                                 float* drow = p.dW + ((size_t)p.wrow[e] + (size_t)((tr0 + j) * k + ts) * p.cout_total + p.o_off + o) * p.cin_pad;
#pragma unroll 1
                                 for (int c0 = 0; c0 < KC; c0 += 32) {
                                     uint32_t v[32];
                                     const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + ucol + (uint32_t)((c * nr + j) * KC + c0);
                                     tmem_ld32(ta, v);
                                     tmem_st32_zero(ta);                 // the next item accumulates into a clean buffer
                                     if (ts < k) {
                                         float* dst = drow + c * KC + c0;
#pragma unroll
                                         for (int x = 0; x < 8; ++x)
                                             red_add_v4_2(dst + 4 * x, __uint_as_float(v[4 * x]), __uint_as_float(v[4 * x + 1]),
                                                          __uint_as_float(v[4 * x + 2]), __uint_as_float(v[4 * x + 3]));
                                     }
                                 }
                             }
                         }
                     }
                     tmem_wait_st();
                     tc_fence_before();
                     __syncwarp();
                     if (lane == 0) mb_arrive(&t_empty[buf]);
                     buf ^= 1;
                 });
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
    if (threadIdx.x == 0) {
        // the last CTA to finish re-arms the scheduler for the next launch on this stream
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
}

template <int COUT, int KC>
static int launch_wgrad2(const CUtensorMap* ta, const CUtensorMap* tb, const WGrad2Params& p, cudaStream_t st) {
    auto kfn = gwgrad2_kernel<COUT, KC>;
    const int smem = kW2Stages * (p.a_stage_bytes + p.b_stage_bytes) + 1024;
    HDMOE_CHECK_ARG(smem <= 227 * 1024, "gwgrad: strip does not fit shared memory (%d bytes)", smem);
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    kfn<<<grid, kW2Threads, smem, st>>>(ta[0], ta[1], ta[2], ta[3], tb[0], tb[1], tb[2], tb[3], p);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

}  // namespace hdmoe
using namespace hdmoe;

#ifdef HDMOE_WG_TRACE
extern "C" int hdmoe_wg_trace_read(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, wg2_trace, sizeof(long long) * 148 * 16);
}
#endif

// one pass over output channels [o_off, o_off + Cout) of a dY / dW that hold cout_total channels
static int gconv_wgrad_pass(const void* X, const void* dY, float* dW, int cap_rows, int H, int W, int Cin_pad, int Cout,
                            int cout_total, int o_off, int64_t w_rows_total, const int32_t* row_expert,
                            const int32_t* n_rows_dev, int n_experts, const int32_t* ksize_host, const int32_t* wrow_host,
                            hdmoe_stream_t stream);

extern "C" int hdmoe_gconv_wgrad(const void* X, const void* dY, float* dW, int cap_rows, int H, int W, int Cin_pad,
                                 int Cout, int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev,
                                 int n_experts, const int32_t* ksize_host, const int32_t* wrow_host,
                                 hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(Cout == 32 || Cout == 64 || Cout == 128, "gconv_wgrad: Cout must be 32, 64 or 128 (got %d)", Cout);
    if (Cout == 128) {     // the router-trunk widths: two 64-channel passes over the same input windows
        for (int h = 0; h < 2; ++h) {
            const int rc = gconv_wgrad_pass(X, dY, dW, cap_rows, H, W, Cin_pad, 64, 128, 64 * h, w_rows_total, row_expert,
                                            n_rows_dev, n_experts, ksize_host, wrow_host, stream);
            if (rc != HDMOE_OK) return rc;
        }
        return HDMOE_OK;
    }
    return gconv_wgrad_pass(X, dY, dW, cap_rows, H, W, Cin_pad, Cout, Cout, 0, w_rows_total, row_expert, n_rows_dev,
                            n_experts, ksize_host, wrow_host, stream);
}

static int gconv_wgrad_pass(const void* X, const void* dY_base, float* dW, int cap_rows, int H, int W, int Cin_pad, int Cout,
                            int cout_total, int o_off, int64_t w_rows_total, const int32_t* row_expert,
                            const int32_t* n_rows_dev, int n_experts, const int32_t* ksize_host, const int32_t* wrow_host,
                            hdmoe_stream_t stream) {
    const void* dY = (const void*)((const __nv_bfloat16*)dY_base + o_off);
    HDMOE_CHECK_ARG(X && dY_base && dW && row_expert && n_rows_dev && ksize_host && wrow_host, "gconv_wgrad: null pointer");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= kW2MaxE, "gconv_wgrad: 1 <= n_experts <= %d", kW2MaxE);
    HDMOE_CHECK_ARG(Cout == 32 || Cout == 64, "gconv_wgrad: Cout must be 32 or 64 (got %d)", Cout);
    HDMOE_CHECK_ARG(Cin_pad >= 32 && Cin_pad % 32 == 0 && Cin_pad <= 256, "gconv_wgrad: Cin_pad in 32..256, multiple of 32");
    HDMOE_CHECK_ARG(H % 4 == 0 && H <= 248 && W <= 232, "gconv_wgrad: need H %% 4 == 0, H <= 248, W <= 232");
    HDMOE_CHECK_ARG((((uintptr_t)X | (uintptr_t)dY | (uintptr_t)dW) & 15) == 0, "gconv_wgrad: 16-byte alignment required");
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) {
        set_error("gconv_wgrad: cuTensorMapEncodeTiled not available from the driver");
        return HDMOE_ERR_CUDA;
    }
    const int KC = (Cin_pad % 64 == 0) ? 64 : 32;
    const int TPM = 128 / Cout;
    WGrad2Params p{};
    p.H = H;
    p.W = W;
    p.cap_rows = cap_rows;
    p.nchunks = Cin_pad / KC;
    p.cin_pad = Cin_pad;
    p.n_experts = n_experts;
    p.row_expert = row_expert;
    p.n_rows_dev = n_rows_dev;
    p.dW = dW;
    p.cout_total = cout_total;
    p.o_off = o_off;
    int ncls = 0, cls_k[kW2Classes], kmax = 1;
    for (int e = 0; e < n_experts; ++e) {
        const int k = ksize_host[e];
        HDMOE_CHECK_ARG(k >= 1 && k <= 7 && (k & 1), "gconv_wgrad: odd kernel sizes 1..7");
        int c = -1;
        for (int q = 0; q < ncls; ++q)
            if (cls_k[q] == k) c = q;
        if (c < 0) {
            HDMOE_CHECK_ARG(ncls < kW2Classes, "gconv_wgrad: at most %d distinct kernel sizes per launch", kW2Classes);
            c = ncls++;
            cls_k[c] = k;
        }
        p.kclass[e] = (uint8_t)c;
        p.wrow[e] = wrow_host[e];
        if (k > kmax) kmax = k;
    }
    // strip height: largest candidate dividing H whose two stages fit shared memory for the widest kernel.  The padded
    // row pitch Wp is rounded up so that a strip is a whole number of 16-position MMA slices; the surplus columns are
    // out of bounds for both boxes and zero-filled by TMA like the k-1 halo columns.
    auto pitch = [&](int k, int sh) {
        const int m = sh % 16 == 0 ? 1 : (sh % 8 == 0 ? 2 : 4);
        return ((W + k - 1 + m - 1) / m) * m;
    };
    int SH = 0;
    for (int cand : {32, 16, 8, 4}) {
        if (H % cand) continue;
        const int Wp = pitch(kmax, cand);
        const long long a = kW2Lead + (long long)cand * Wp * Cout * 2;
        const long long b = ((long long)(cand + kmax - 1) * Wp + (kmax - 1) + 16) * KC * 2;
        if (kW2Stages * (((a + 1023) / 1024 + (b + 1023) / 1024) * 1024) + 1024 <= 220 * 1024) {
            SH = cand;
            break;
        }
    }
    HDMOE_CHECK_ARG(SH > 0, "gconv_wgrad: no strip height fits shared memory for %dx%d, k=%d", H, W, kmax);
    p.SH = SH;
    // a unit = TPM column taps x nr kernel rows needs nr * Cin_pad TMEM columns; a tap group fills one 256-column buffer
    HDMOE_CHECK_ARG(Cin_pad <= kW2BufCols, "gconv_wgrad: Cin_pad too large for a TMEM accumulator buffer");
    const int nr_max = std::max(1, std::min(128 / KC, kW2BufCols / Cin_pad));      // kernel rows stacked in N (N <= 128)
    int gmax = 1, a_max = 0, b_max = 0;
    for (int c = 0; c < ncls; ++c) {
        const int k = cls_k[c], Wp = pitch(k, SH);
        HDMOE_CHECK_ARG((SH * Wp) % 16 == 0 && Wp <= 256, "gconv_wgrad: strip of %d rows x %d padded columns is not a multiple of 16", SH, Wp);
        const int upr = (k + TPM - 1) / TPM;                    // column-tap blocks per kernel row
        // units = (column block, row block); packed into 256-column tap groups first-fit by decreasing width (every
        // group is one more pass over the rows, so as few groups as the TMEM budget allows)
        struct Unit { int cu, tr0, nr, grp, col; };
        Unit un[kW2MaxUnits];
        int nu = 0, ng = 0, fill[kW2MaxUnits] = {0};
        for (int cu = 0; cu < upr; ++cu)
            for (int tr0 = 0; tr0 < k; tr0 += nr_max) {
                HDMOE_CHECK_ARG(nu < kW2MaxUnits, "gconv_wgrad: kernel size %d needs more than %d units", k, kW2MaxUnits);
                un[nu++] = Unit{cu, tr0, std::min(nr_max, k - tr0), -1, 0};
            }
        for (int width = nr_max; width >= 1; --width)               // decreasing row count = decreasing column width
            for (int i = 0; i < nu; ++i) {
                if (un[i].nr != width) continue;
                int gsel = 0;
                while (gsel < ng && fill[gsel] + width * Cin_pad > kW2BufCols) ++gsel;
                if (gsel == ng) fill[ng++] = 0;
                un[i].grp = gsel;
                un[i].col = fill[gsel];
                fill[gsel] += width * Cin_pad;
            }
        int slot = 0;
        for (int gi = 0; gi < ng; ++gi) {                            // the kernel wants a group's units contiguous
            p.g_lo[c][gi] = (uint8_t)slot;
            for (int i = 0; i < nu; ++i)
                if (un[i].grp == gi) {
                    p.u_cu[c][slot] = (uint8_t)un[i].cu;
                    p.u_tr0[c][slot] = (uint8_t)un[i].tr0;
                    p.u_nr[c][slot] = (uint8_t)un[i].nr;
                    p.u_col[c][slot] = (uint16_t)un[i].col;
                    ++slot;
                }
        }
        p.g_lo[c][ng] = (uint8_t)slot;
        p.ksize[c] = k;
        p.wp[c] = Wp;
        p.ngroups[c] = ng;
        p.a_box_bytes[c] = SH * Wp * Cout * 2;
        p.b_box_bytes[c] = (SH + k - 1) * Wp * KC * 2;
        // box + the reach of the last unit's B start ((k-1) rows, (upr-1)*TPM columns) + one slice
        const int b_need = ((SH + k - 1) * Wp + (k - 1) + 16) * KC * 2;
        if (ng > gmax) gmax = ng;
        if (kW2Lead + p.a_box_bytes[c] > a_max) a_max = kW2Lead + p.a_box_bytes[c];
        if (b_need > b_max) b_max = b_need;
    }
    p.a_stage_bytes = ((a_max + 1023) / 1024) * 1024;
    p.b_stage_bytes = ((b_max + 1023) / 1024) * 1024;
    p.gmax = gmax;
    // rows per item: >= 8 items per SM for the dynamic scheduler; at least 2 pipeline stages per item so that the
    // flush of one item fits under the MMAs of the next
    int rpi = cap_rows * gmax / (8 * kNumSMs);
    const int stages_per_row = (H / SH) * p.nchunks;
    if (rpi * stages_per_row < 2) rpi = (2 + stages_per_row - 1) / stages_per_row;
    if (rpi < 1) rpi = 1;
    p.rows_per_item = rpi;
    p.n_items = ((cap_rows + rpi - 1) / rpi) * gmax;
    CUtensorMap ta[kW2Classes], tb[kW2Classes];
    const CUtensorMapSwizzle swa = Cout == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUtensorMapSwizzle swb = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int c = 0; c < kW2Classes; ++c) {
        const int cc = c < ncls ? c : 0;
        const int k = cls_k[cc], Wp = pitch(k, SH);
        cuuint32_t es[4] = {1, 1, 1, 1};
        {
            cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap_rows};
            cuuint64_t strides[3] = {(cuuint64_t)cout_total * 2, (cuuint64_t)W * cout_total * 2,
                                     (cuuint64_t)H * W * cout_total * 2};
            cuuint32_t box[4] = {(cuuint32_t)Cout, (cuuint32_t)Wp, (cuuint32_t)SH, 1};
            CUresult r = enc(&ta[c], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dY), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swa, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("gconv_wgrad: cuTensorMapEncodeTiled(dY) failed with %d", (int)r);
                return HDMOE_ERR_CUDA;
            }
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)Cin_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap_rows};
            cuuint64_t strides[3] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)W * Cin_pad * 2, (cuuint64_t)H * W * Cin_pad * 2};
            cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)Wp, (cuuint32_t)(SH + k - 1), 1};
            CUresult r = enc(&tb[c], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swb, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("gconv_wgrad: cuTensorMapEncodeTiled(X) failed with %d", (int)r);
                return HDMOE_ERR_CUDA;
            }
        }
    }
    (void)w_rows_total;
    cudaStream_t st = (cudaStream_t)stream;
    p.sched = sched_slot(st);
    HDMOE_CHECK_ARG(p.sched != nullptr, "gconv_wgrad: more than %d distinct streams in use", kSchedSlots);
    if (Cout == 64 && KC == 64) return launch_wgrad2<64, 64>(ta, tb, p, st);
    if (Cout == 64 && KC == 32) return launch_wgrad2<64, 32>(ta, tb, p, st);
    if (Cout == 32 && KC == 64) return launch_wgrad2<32, 64>(ta, tb, p, st);
    return launch_wgrad2<32, 32>(ta, tb, p, st);
}
