// G-WGRAD v1: grouped convolution weight gradient on tcgen05 (MN-major operands, halo reuse, split-K).
// Superseded by gwgrad2.cu (taps stacked in M, double-buffered accumulators); kept for A/B measurement.
//
//   dW_e[tap][o][c] = sum over rows r of expert e, output pixels q:  dY[r, q, o] * Xpad[r, q + delta_tap, c]
//
// GEMM view per (expert, tap): D[M = Cout][N = Cin chunk] with the reduction K running over PIXELS.  Both
// operands are stored pixel-major (NHWC), i.e. with their M / N index contiguous: "MN-major" UMMA operands
// (instruction-descriptor bits 15/16).  Hardware facts established with tools/umma_probe.cu on B200:
// MN-major SWIZZLE_128B tiles [K rows][64 elements] written by TMA work with SBO = 8 rows, a row-shifted start
// address needs base_offset = 0, and an M = 64 accumulator lives in TMEM lanes (m / 16) * 32 + m % 16.
//
// Flattened halo formulation (as in gconv2.cu): a strip of SH output rows is a run of P = SH * Wp "positions"
// (Wp = W + k - 1).  A = the dY strip loaded as a [Cout, Wp, SH] box -- the k-1 surplus columns are out of bounds
// and zero-filled, so padding positions contribute nothing; B = the zero-padded input window of the strip, ONE
// box per channel chunk; tap (r, s) reads it at start + (r * Wp + s) rows.  One MMA consumes 16 positions.
//
// Work decomposition (no host knowledge of the routing): item = (chunk of rows_per_item consecutive rows, tap
// group g).  Rows are expert-major, so an item sees at most a few expert changes; accumulators are flushed (vector
// atomics into the fp32 tap-major gradient buffer) at each change and at the end: split-K over row chunks.  Tap
// groups exist because all taps of a group keep their [Cout x Cin_chunk] accumulators in the 512 TMEM columns at
// once.  Items are handed out dynamically, last rows (the large-kernel experts) first, from a self-resetting global
// counter: a 5x5 item costs 1.5-2x a 3x3 item and tap groups beyond a small kernel's are empty, so the static
// stride left CTAs with 4 032 MMAs next to an average of 2 440 (65 % imbalance at B = 256).
//
// Roles: warp 0 TMA producer, warps 1-3 MMA issuers (taps of the group are dealt round-robin; one thread
// sustains only ~1 MMA / 100 cycles, tools/umma_rate.cu), warps 4-7 epilogue.
#include "tc.cuh"
#include "../../include/hdmoe_gemm.h"

namespace hdmoe {

#ifdef HDMOE_WG_TRACE
__device__ long long wg_trace[148 * 16];
#define WGT_DECL long long wt_[6] = {0, 0, 0, 0, 0, 0}; long long wt0_ = clock64(); (void)wt0_
#define WGT_LAP(i) do { const long long n_ = clock64(); wt_[i] += n_ - wt0_; wt0_ = n_; } while (0)
#define WGT_CNT(i) do { wt_[i] += 1; } while (0)
#define WGT_OUT(base) do { for (int q_ = 0; q_ < 6; ++q_) wg_trace[blockIdx.x * 16 + (base) + q_] = wt_[q_]; } while (0)
#else
#define WGT_DECL do { } while (0)
#define WGT_LAP(i) do { } while (0)
#define WGT_CNT(i) do { } while (0)
#define WGT_OUT(base) do { } while (0)
#endif

constexpr int kWgIssuers = 4;
constexpr int kWgThreads = 32 * (1 + kWgIssuers + 4);
constexpr int kWgClasses = 4;
constexpr int kWgMaxE = HDMOE_MAX_EXPERTS;
constexpr int kWgStages = 2;
constexpr int kWgQueue = 8;            // item-id queue between the scheduler (producer lane) and the other roles

struct WGradParams {
    int n_items, gmax, rows_per_item, cap_rows;
    int H, W, SH;                      // strip height (rows), H % SH == 0
    int nchunks;                       // Cin_pad / KC
    int cout, cin_pad;
    int n_experts;
    int a_stage_bytes, b_stage_bytes;
    const int32_t* row_expert;
    const int32_t* n_rows_dev;
    float* dW;                         // fp32 [w_rows_total][cin_pad], tap-major blocks per expert (accumulated)
    int32_t* sched;                    // [0] next item, [1] finished CTAs (self-resetting, core.cu)
    int32_t wrow[kWgMaxE];
    uint8_t kclass[kWgMaxE];
    int32_t ksize[kWgClasses], wp[kWgClasses], ngroups[kWgClasses], tg[kWgClasses];
    int32_t a_box_bytes[kWgClasses], b_box_bytes[kWgClasses];
};

// MN-major operand descriptor: rows are K (positions), ROWB bytes each (64 -> SW64, 128 -> SW128)
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
    constexpr uint64_t sbo = (8 * ROWB) >> 4;
    constexpr uint64_t layout = ROWB == 128 ? 2 : 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (0ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// COUT in {32, 64}: A rows are COUT*2 bytes; KC in {32, 64}: B rows are KC*2 bytes
template <int COUT, int KC>
__global__ void __launch_bounds__(kWgThreads, 1)
gwgrad_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
              const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap ta3,
              const __grid_constant__ CUtensorMap tb0, const __grid_constant__ CUtensorMap tb1,
              const __grid_constant__ CUtensorMap tb2, const __grid_constant__ CUtensorMap tb3,
              const __grid_constant__ WGradParams p) {
    constexpr int ROWA = COUT * 2, ROWB_ = KC * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[kWgStages], empty[kWgStages], t_full, t_empty, q_full[kWgQueue], q_empty[kWgQueue];
    __shared__ int32_t item_q[kWgQueue];
    __shared__ uint32_t tmem_base_s;
    const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // operand buffers must never hold NaN/Inf garbage: positions past a box are multiplied by zeros of dY
    for (int i = threadIdx.x; i < kWgStages * stage_bytes / 16; i += kWgThreads)
        reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            mb_init(&full[s], 1);
            mb_init(&empty[s], kWgIssuers);
        }
        mb_init(&t_full, kWgIssuers);
        mb_init(&t_empty, 4);
        for (int q = 0; q < kWgQueue; ++q) {
            mb_init(&q_full[q], 1);
            mb_init(&q_empty[q], kWgIssuers + 4);
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
    const int n_rows = min(*p.n_rows_dev, p.cap_rows);
    const int nstrips = p.H / p.SH;

    // Walk of one item, identical in every role.  Calls stage(r, e, kc, strip, chunk, first) for every pipeline
    // stage and flush(e, kc) whenever the accumulators must be written out.
    auto walk = [&](int item, auto&& stage_fn, auto&& flush_fn) {
        const int g = item % p.gmax, rc = p.n_items / p.gmax - 1 - item / p.gmax;       // last row chunks first
        const int r0 = rc * p.rows_per_item, r1 = min(r0 + p.rows_per_item, n_rows);
        int cur_e = -1, cur_kc = 0;
        bool fresh = true;
        for (int r = r0; r < r1; ++r) {
            const int e = p.row_expert[r];
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
                for (int c = 0; c < p.nchunks; ++c) {
                    stage_fn(r, e, kc, g, st, c, fresh && st == 0);
                }
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
        if (++qs == kWgQueue) {
            qs = 0;
            qph ^= 1;
        }
        return i;
    };

    if (warp == 0) {
        // ============================== scheduler + TMA producer ==============================
        if (lane == 0) {
            const CUtensorMap* ma[kWgClasses] = {&ta0, &ta1, &ta2, &ta3};
            const CUtensorMap* mb[kWgClasses] = {&tb0, &tb1, &tb2, &tb3};
            int s = 0;
            uint32_t ph = 0;
            for (;;) {
                int item = atomicAdd(p.sched, 1);
                if (item >= p.n_items) item = -1;
                mb_wait(&q_empty[qs], qph ^ 1);
                item_q[qs] = item;
                mb_arrive(&q_full[qs]);
                if (++qs == kWgQueue) {
                    qs = 0;
                    qph ^= 1;
                }
                if (item < 0) break;
                walk(item,
                     [&](int r, int e, int kc, int g, int st, int c, bool) {
                         const int pad = (p.ksize[kc] - 1) >> 1;
                         mb_wait(&empty[s], ph ^ 1);
                         mb_expect_tx(&full[s], (uint32_t)(p.a_box_bytes[kc] + p.b_box_bytes[kc]));
                         uint8_t* base = smem + (size_t)s * stage_bytes;
                         tma_load_4d(base, ma[kc], &full[s], 0, 0, st * p.SH, r);
                         tma_load_4d(base + p.a_stage_bytes, mb[kc], &full[s], c * KC, -pad, st * p.SH - pad, r);
                         if (++s == kWgStages) {
                             s = 0;
                             ph ^= 1;
                         }
                     },
                     [&](int, int, int) {});
            }
        }
    } else if (warp <= kWgIssuers) {
        // ============================== MMA issuers ==============================
        if (lane == 0) {
            // D[64 x KC] (+)= A^T B : A, B MN-major (bits 15, 16), M = 64, N = KC, bf16 -> fp32
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(KC >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            const int me = warp - 1;
            int s = 0;
            uint32_t ph = 0, tph = 0;
            WGT_DECL;
            for (;;) {
                const int item = next_item(false);
                WGT_LAP(0);                       // waiting for an item
                if (item < 0) break;
                WGT_CNT(5);
                walk(item,
                     [&](int r, int e, int kc, int g, int st, int c, bool first) {
                         const int k = p.ksize[kc], Wp = p.wp[kc];
                         const int t_lo = g * p.tg[kc], t_hi = min(k * k, t_lo + p.tg[kc]);
                         const int nslice = (p.SH * Wp) >> 4;
                         WGT_LAP(1);                 // walk / decode
                         mb_wait(&full[s], ph);
                         tc_fence_after();
                         WGT_LAP(2);                 // waiting for the stage's TMA loads
                         const uint32_t a0 = s2u(smem + (size_t)s * stage_bytes);
                         const uint32_t b0 = a0 + p.a_stage_bytes;
                         const uint64_t ad0 = umma_desc_mn<ROWA>(a0);
                         const uint64_t bd0 = umma_desc_mn<ROWB_>(b0);
                         for (int t = t_lo + me; t < t_hi; t += kWgIssuers) {
                             const int tr = t / k, ts = t - tr * k;
                             const uint32_t d = tmem_base + (uint32_t)(((t - t_lo) * p.nchunks + c) * KC);
                             const uint64_t bd = bd0 + (uint64_t)(((uint32_t)(tr * Wp + ts) * ROWB_) >> 4);
                             for (int j = 0; j < nslice; ++j)
                                 tc_mma(d, ad0 + (uint64_t)((j * 16 * ROWA) >> 4), bd + (uint64_t)((j * 16 * ROWB_) >> 4), idesc,
                                        !(first && j == 0));
                         }
                         tc_commit(&empty[s]);
                         WGT_LAP(3);                 // issuing MMAs
                         if (++s == kWgStages) {
                             s = 0;
                             ph ^= 1;
                         }
                     },
                     [&](int, int, int) {
                         WGT_LAP(1);
                         tc_commit(&t_full);               // all accumulators of the group are final
                         mb_wait(&t_empty, tph);           // epilogue has read them
                         tph ^= 1;
                         tc_fence_after();
                         WGT_LAP(4);                 // flush: MMAs drain + epilogue reads the accumulators
                     });
            }
            if (me == 0) WGT_OUT(0);
        }
    } else {
        // ============================== epilogue: TMEM -> vector atomics ==============================
        const int quad = warp & 3;
        uint32_t tph = 0;
        for (;;) {
            const int item = next_item(true);
            if (item < 0) break;
            walk(item, [&](int, int, int, int, int, int, bool) {},
                 [&](int e, int kc, int g) {
                     const int k = p.ksize[kc];
                     const int t_lo = g * p.tg[kc], t_hi = min(k * k, t_lo + p.tg[kc]);
                     mb_wait(&t_full, tph);
                     tph ^= 1;
                     tc_fence_after();
                     const int o = quad * 16 + lane;              // M = 64: 16 accumulator rows per lane quadrant
                     for (int t = t_lo; t < t_hi; ++t)
                         for (int c = 0; c < p.nchunks; ++c) {
#pragma unroll
                             for (int c0 = 0; c0 < KC; c0 += 32) {
                                 uint32_t v[32];
                                 tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) +
                                               (uint32_t)(((t - t_lo) * p.nchunks + c) * KC + c0), v);
                                 if (lane < 16 && o < COUT) {
                                     float* dst = p.dW + ((size_t)p.wrow[e] + (size_t)t * COUT + o) * p.cin_pad + c * KC + c0;
#pragma unroll
                                     for (int u = 0; u < 8; ++u)
                                         red_add_v4(dst + 4 * u, __uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]),
                                                    __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
                                 }
                             }
                         }
                     tc_fence_before();
                     __syncwarp();
                     if (lane == 0) mb_arrive(&t_empty);
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
static int launch_wgrad(const CUtensorMap* ta, const CUtensorMap* tb, const WGradParams& p, cudaStream_t st) {
    auto kfn = gwgrad_kernel<COUT, KC>;
    const int smem = kWgStages * (p.a_stage_bytes + p.b_stage_bytes) + 1024;
    HDMOE_CHECK_ARG(smem <= 227 * 1024, "gwgrad: strip does not fit shared memory (%d bytes)", smem);
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    kfn<<<grid, kWgThreads, smem, st>>>(ta[0], ta[1], ta[2], ta[3], tb[0], tb[1], tb[2], tb[3], p);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

}  // namespace hdmoe
using namespace hdmoe;

#ifdef HDMOE_WG_TRACE
extern "C" int hdmoe_wg_trace_read(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, wg_trace, sizeof(long long) * 148 * 16);
}
#endif

extern "C" int hdmoe_gconv_wgrad_v1(const void* X, const void* dY, float* dW, int cap_rows, int H, int W, int Cin_pad,
                                 int Cout, int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev,
                                 int n_experts, const int32_t* ksize_host, const int32_t* wrow_host,
                                 hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(X && dY && dW && row_expert && n_rows_dev && ksize_host && wrow_host, "gconv_wgrad: null pointer");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= kWgMaxE, "gconv_wgrad: 1 <= n_experts <= %d", kWgMaxE);
    HDMOE_CHECK_ARG(Cout == 32 || Cout == 64, "gconv_wgrad: Cout must be 32 or 64 (got %d)", Cout);
    HDMOE_CHECK_ARG(Cin_pad >= 32 && Cin_pad % 32 == 0 && Cin_pad <= 256, "gconv_wgrad: Cin_pad in 32..256, multiple of 32");
    HDMOE_CHECK_ARG(H % 8 == 0 && W % 2 == 0 && H <= 248 && W <= 240, "gconv_wgrad: need H %% 8 == 0 and even W");
    HDMOE_CHECK_ARG((((uintptr_t)X | (uintptr_t)dY | (uintptr_t)dW) & 15) == 0, "gconv_wgrad: 16-byte alignment required");
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) {
        set_error("gconv_wgrad: cuTensorMapEncodeTiled not available from the driver");
        return HDMOE_ERR_CUDA;
    }
    const int KC = (Cin_pad % 64 == 0) ? 64 : 32;
    WGradParams p{};
    p.H = H;
    p.W = W;
    p.cap_rows = cap_rows;
    p.nchunks = Cin_pad / KC;
    p.cout = Cout;
    p.cin_pad = Cin_pad;
    p.n_experts = n_experts;
    p.row_expert = row_expert;
    p.n_rows_dev = n_rows_dev;
    p.dW = dW;
    int ncls = 0, cls_k[kWgClasses], kmax = 1;
    for (int e = 0; e < n_experts; ++e) {
        const int k = ksize_host[e];
        HDMOE_CHECK_ARG(k >= 1 && k <= 7 && (k & 1), "gconv_wgrad: odd kernel sizes 1..7");
        int c = -1;
        for (int q = 0; q < ncls; ++q)
            if (cls_k[q] == k) c = q;
        if (c < 0) {
            HDMOE_CHECK_ARG(ncls < kWgClasses, "gconv_wgrad: at most %d distinct kernel sizes per launch", kWgClasses);
            c = ncls++;
            cls_k[c] = k;
        }
        p.kclass[e] = (uint8_t)c;
        p.wrow[e] = wrow_host[e];
        if (k > kmax) kmax = k;
    }
    // strip height: largest multiple of 8 dividing H whose two stages fit shared memory for the widest kernel
    int SH = 0;
    for (int cand : {32, 16, 8}) {
        if (H % cand) continue;
        const int Wp = W + kmax - 1;
        const long long a = (long long)cand * Wp * Cout * 2;
        const long long b = ((long long)(cand + kmax - 1) * Wp + (kmax - 1) + 16) * KC * 2;
        if (kWgStages * (((a + 1023) / 1024 + (b + 1023) / 1024) * 1024) + 1024 <= 220 * 1024) {
            SH = cand;
            break;
        }
    }
    HDMOE_CHECK_ARG(SH > 0, "gconv_wgrad: no strip height fits shared memory for %dx%d, k=%d", H, W, kmax);
    p.SH = SH;
    int gmax = 1, a_max = 0, b_max = 0;
    for (int c = 0; c < ncls; ++c) {
        const int k = cls_k[c], Wp = W + k - 1, taps = k * k;
        HDMOE_CHECK_ARG((SH * Wp) % 16 == 0, "gconv_wgrad: strip of %d rows x %d padded columns is not a multiple of 16", SH, Wp);
        const int tg_cap = 512 / Cin_pad;                       // taps whose accumulators fit TMEM together
        HDMOE_CHECK_ARG(tg_cap >= 1, "gconv_wgrad: Cin_pad too large for TMEM");
        const int ng = (taps + tg_cap - 1) / tg_cap;
        p.ksize[c] = k;
        p.wp[c] = Wp;
        p.ngroups[c] = ng;
        p.tg[c] = (taps + ng - 1) / ng;
        p.a_box_bytes[c] = SH * Wp * Cout * 2;
        p.b_box_bytes[c] = (SH + k - 1) * Wp * KC * 2;
        const int b_need = ((SH + k - 1) * Wp + (k - 1) + 16) * KC * 2;   // box + the reach of the last tap
        if (ng > gmax) gmax = ng;
        if (p.a_box_bytes[c] > a_max) a_max = p.a_box_bytes[c];
        if (b_need > b_max) b_max = b_need;
    }
    p.a_stage_bytes = ((a_max + 1023) / 1024) * 1024;
    p.b_stage_bytes = ((b_max + 1023) / 1024) * 1024;
    p.gmax = gmax;
    // rows per item: >= 8 items per SM for the dynamic scheduler (an item ends with one flush of the group's
    // accumulators, ~2 k cycles during which the MMAs of the CTA wait, so items should not be smaller than needed)
    int rpi = cap_rows * gmax / (8 * kNumSMs);
    if (rpi < 1) rpi = 1;
    p.rows_per_item = rpi;
    p.n_items = ((cap_rows + rpi - 1) / rpi) * gmax;
    CUtensorMap ta[kWgClasses], tb[kWgClasses];
    const CUtensorMapSwizzle swa = Cout == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUtensorMapSwizzle swb = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int c = 0; c < kWgClasses; ++c) {
        const int cc = c < ncls ? c : 0;
        const int k = cls_k[cc], Wp = W + k - 1;
        cuuint32_t es[4] = {1, 1, 1, 1};
        {
            cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap_rows};
            cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
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
    if (Cout == 64 && KC == 64) return launch_wgrad<64, 64>(ta, tb, p, st);
    if (Cout == 64 && KC == 32) return launch_wgrad<64, 32>(ta, tb, p, st);
    if (Cout == 32 && KC == 64) return launch_wgrad<32, 64>(ta, tb, p, st);
    return launch_wgrad<32, 32>(ta, tb, p, st);
}
