// G-CONV v3 (EXPERIMENTAL, opt-in through ops.set_gconv_impl(3); not yet validated on hardware -- the core
// formulation is, see tools/umma_tpair_probe.cu): grouped implicit-GEMM convolution with TAP GROUPS STACKED IN N.
//
// gconv2 issues one M128 x N=Cout x K16 MMA per filter tap.  At Cout <= 64 those MMAs are bound by the shared-memory
// operand feed (4 KiB of A per MMA: 48 cycles for 32 of math at Cout = 64, 40 for 16 at Cout = 32).  v3 keeps gconv2's
// operand roles, buffers and weight layout and changes the MMA shape: TPM = 128 / Cout consecutive taps of one kernel
// row are ONE MMA with N = TPM * Cout = 128 -- their weight tiles are consecutive row blocks of the tap-major operand,
// i.e. one [128 x KC] B tile -- using the A start of the group's FIRST tap.  Column block j of the accumulator then
// holds tap j of the group evaluated one position too far to the left per step of j:
//
//     D[q][j * Cout + co] = sum_ci W[tr, s0 + j][co][ci] * Xpad[q + tr * Wp + s0][ci]      (wanted: ... + s0 + j)
//     Y[q][co] = sum_j D[q + j][j * Cout + co]
//
// so the epilogue adds column block j of the accumulator row j lanes further down: a warp shuffle, plus a tiny
// shared-memory exchange for the lanes whose partner lives in the next warp.  M-tiles overlap by TPM - 1 positions
// (stride 127 / 125), so no row needs a partner from another accumulator.  The last, incomplete tap group of a kernel
// row is simply an MMA with a smaller N (N is an instruction field), so the weight layout is unchanged.
// Per kernel row, k = 5, Cout = 64: 64 + 64 + 48 cycles instead of 5 x 48; Cout = 32: 64 + 40 instead of 5 x 40.
// Accumulators are 128 columns per M-tile: 2 M-tiles x 2 buffers fill TMEM, two issuer warps.
#include "tc.cuh"
#include "../../include/hdmoe_gemm.h"

namespace hdmoe {

#ifdef HDMOE_G3_TRACE
__device__ long long g3_trace[148 * 64];
__device__ __forceinline__ long long g3_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define G3T(slot) do { if (blockIdx.x < 148 && tcount < 8 && lane == 0) { g3_trace[blockIdx.x * 64 + tcount * 8 + (slot)] = clock64(); \
    if ((slot) == 0) g3_trace[blockIdx.x * 64 + tcount * 8 + 6] = g3_gtime(); if ((slot) == 5) g3_trace[blockIdx.x * 64 + tcount * 8 + 7] = g3_gtime(); } } while (0)
#else
#define G3T(slot) do { } while (0)
#endif

constexpr int kG3Issuers = 2;      // one per M-tile accumulator (N = 128 MMAs take 64 cycles; one thread issues one per ~103)
constexpr int kG3Threads = 32 * (1 + kG3Issuers + 4);
constexpr int kG3MaxE = HDMOE_MAX_EXPERTS;
constexpr int kG3Classes = 4;
constexpr int kG3RowCache = 4096;
constexpr int kG3Queue = 16;
constexpr int kG3AStages = 2, kG3BStages = 4;

struct GConv3Params {
    int n_tiles, smax;                // tiles = cap_rows * smax (tiles per sample, max over classes)
    int H, W;
    int upt;                          // channel chunks of KC per tap
    int n_experts;
    int a_stage_bytes;                // bytes of one halo buffer
    int mt_per_tile;                  // M-tiles per tile (2; 1 if two halo buffers of a 2-M-tile tile do not fit)
    const int32_t* row_expert;
    const int32_t* n_rows_dev;
    __nv_bfloat16* Y;
    const float* scale;
    const __nv_bfloat16* res;
    float res_a, res_b;
    int act;
    int32_t wrow[kG3MaxE];
    uint8_t kclass[kG3MaxE];
    int32_t* sched;                   // [0] next tile, [1] finished CTAs (self-resetting, core.cu)
    int32_t ksize[kG3Classes], wp[kG3Classes], box_bytes[kG3Classes], ntile[kG3Classes], npos[kG3Classes];
};

struct G3Tile {
    int r, j, e, kc;                  // row (sample), tile of the sample, expert, kernel-size class
    int mt_n, p0, h0, c0;             // M-tiles, first position, its image row / offset inside that row's box
};

__device__ __forceinline__ void g3_st_v8(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void g3_ld_v8(const void* p, uint32_t* v) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void g3_epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// Epilogue of one tile for one epilogue thread.  xch: shared [2][4 quads][TPM-1 blocks][TPM-1 lanes][COUT] floats.
template <int COUT, bool SC, bool ACT, bool RES>
__device__ __forceinline__ void g3_epilogue(const GConv3Params& p, const G3Tile& t, int quad, int lane, uint32_t tmem_acc,
                                            float* xch, int& mcount) {
    constexpr int TPM = 128 / COUT, MS = 128 - (TPM - 1), XL = TPM - 1;
    const int k = p.ksize[t.kc], Wp = p.wp[t.kc];
    const int nblk = k < TPM ? k : TPM;                        // column blocks the MMAs wrote
    const float* sc = SC ? p.scale + (size_t)t.r * COUT : nullptr;
    for (int mt = 0; mt < t.mt_n; ++mt) {
        const uint32_t tcol = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * 128);
        // exchange buffers alternate over the running M-tile count (not mt: a one-M-tile tile followed by another tile
        // would reuse a buffer with no barrier in between); M-tile n writes buffer n & 1 after barrier n-1, which every
        // warp reaches only after it finished reading that buffer for M-tile n-2
        float* xb = xch + (size_t)(mcount & 1) * 4 * XL * XL * COUT;
        ++mcount;
        // ---- publish: the first j lanes of this warp hold the block-j partners of the previous warp's last j lanes
        for (int j = 1; j < nblk; ++j)
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tcol + (uint32_t)(j * COUT + c0), v);
                if (lane < j) {
                    float* dst = xb + (((size_t)quad * XL + (j - 1)) * XL + lane) * COUT + c0;
#pragma unroll
                    for (int u = 0; u < 32; ++u) dst[u] = __uint_as_float(v[u]);
                }
            }
        g3_epi_barrier();
        // ---- combine + write: this thread's position
        const int L = quad * 32 + lane;
        const int pa = t.p0 + mt * MS + L;
        const int hl = pa / Wp, w = pa - hl * Wp;
        const bool valid = L < MS && hl < p.H && w < p.W;
        const size_t go = (((size_t)t.r * p.H + (size_t)hl) * p.W + w) * COUT;
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t v[32];
            float acc[32];
            tmem_ld32(tcol + (uint32_t)c0, v);
#pragma unroll
            for (int u = 0; u < 32; ++u) acc[u] = __uint_as_float(v[u]);
            for (int j = 1; j < nblk; ++j) {
                tmem_ld32(tcol + (uint32_t)(j * COUT + c0), v);
                const bool cross = lane + j >= 32;               // partner row lives in the next warp's first lanes
                const float* src = xb + (((size_t)((quad + 1) & 3) * XL + (j - 1)) * XL + (cross ? lane + j - 32 : 0)) * COUT + c0;
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const float s = __uint_as_float(__shfl_down_sync(0xffffffffu, v[u], j));
                    acc[u] += (cross && quad < 3) ? src[u] : (cross ? 0.f : s);
                }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {                      // 16 channels = one 32-byte store
                uint32_t packed[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float a = acc[g * 16 + 2 * u], b = acc[g * 16 + 2 * u + 1];
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
                        g3_ld_v8(p.res + go + c0 + g * 16, rr);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float v0 = __uint_as_float(packed[u] << 16), v1 = __uint_as_float(packed[u] & 0xffff0000u);
                            const float r0 = __uint_as_float(rr[u] << 16), r1 = __uint_as_float(rr[u] & 0xffff0000u);
                            __nv_bfloat162 o = __floats2bfloat162_rn(p.res_a * r0 + p.res_b * v0, p.res_a * r1 + p.res_b * v1);
                            packed[u] = *reinterpret_cast<uint32_t*>(&o);
                        }
                    }
                    g3_st_v8(p.Y + go + c0 + g * 16, packed);
                }
            }
        }
    }
}

template <int KC, int COUT>
__global__ void __launch_bounds__(kG3Threads, 1)
gconv3_fwd_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
                  const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap ta3,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ GConv3Params p) {
    constexpr int TPM = 128 / COUT, MS = 128 - (TPM - 1), XL = TPM - 1;
    constexpr int ROWB = KC * 2, B_STAGE = 128 * KC * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;                                              // halo buffers
    uint8_t* b_buf = smem + (size_t)kG3AStages * p.a_stage_bytes;       // weight ring: one tap group per stage
    __shared__ __align__(8) uint64_t a_full[kG3AStages], a_empty[kG3AStages], b_full[kG3BStages], b_empty[kG3BStages],
        t_full[2], t_empty[2], q_full[kG3Queue], q_empty[kG3Queue];
    __shared__ int32_t tile_q[kG3Queue];
    __shared__ uint32_t tmem_base_s;
    __shared__ int8_t row_e_s[kG3RowCache];
    __shared__ __align__(16) float xch[2 * 4 * XL * XL * COUT];
    const int cap_rows = p.n_tiles / p.smax;
    const bool rows_cached = cap_rows <= kG3RowCache;
    if (rows_cached)
        for (int r = threadIdx.x; r < cap_rows; r += kG3Threads) {
            const int e = p.row_expert[r];
            row_e_s[r] = (int8_t)((e < 0 || e >= p.n_experts) ? -1 : e);
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kG3AStages; ++s) {
            mb_init(&a_full[s], 1);
            mb_init(&a_empty[s], kG3Issuers);
        }
        for (int s = 0; s < kG3BStages; ++s) {
            mb_init(&b_full[s], 1);
            mb_init(&b_empty[s], kG3Issuers);
        }
        for (int a = 0; a < 2; ++a) {
            mb_init(&t_full[a], kG3Issuers);
            mb_init(&t_empty[a], 4);
        }
        for (int s = 0; s < kG3Queue; ++s) {
            mb_init(&q_full[s], 1);
            mb_init(&q_empty[s], kG3Issuers + 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int n_rows = *p.n_rows_dev;
    const int PT = p.mt_per_tile * MS;                                   // valid positions per tile

    // tile id -> geometry (row-major, rows reversed: heavy experts first, as gconv2)
    auto tile_at = [&](int i, G3Tile& t) -> bool {
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
        const int Wp = p.wp[t.kc];
        t.p0 = t.j * PT;
        const int remaining = p.npos[t.kc] - t.p0;
        t.mt_n = (remaining + MS - 1) / MS;
        if (t.mt_n > p.mt_per_tile) t.mt_n = p.mt_per_tile;
        t.h0 = t.p0 / Wp;
        t.c0 = t.p0 - t.h0 * Wp;
        return true;
    };
    int qs = 0;
    uint32_t qph = 0;
    auto next_tile = [&](bool whole_warp) -> int {
        mb_wait(&q_full[qs], qph);
        const int i = tile_q[qs];
        if (whole_warp) __syncwarp();
        if (lane == 0) mb_arrive(&q_empty[qs]);
        if (++qs == kG3Queue) {
            qs = 0;
            qph ^= 1;
        }
        return i;
    };

    if (warp == 0) {
        // ============================== scheduler + TMA producer ==============================
        if (lane == 0) {
            const CUtensorMap* maps[kG3Classes] = {&ta0, &ta1, &ta2, &ta3};
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            G3Tile cur, nxt;
            int cur_c = 0, nxt_c = 0;
            bool more = true, first_draw = true;
            auto advance = [&](const G3Tile& from, int from_c, bool first, G3Tile& to, int& to_c) -> bool {
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
                    if (ok || i < 0 || to.r >= n_rows || to.e < 0) {
                        mb_wait(&q_empty[qs], qph ^ 1);
                        tile_q[qs] = i;
                        mb_arrive(&q_full[qs]);
                        if (++qs == kG3Queue) {
                            qs = 0;
                            qph ^= 1;
                        }
                    }
                    if (i < 0) {
                        more = false;
                        break;
                    }
                    if (ok) {
                        to_c = 0;
                        return true;
                    }
                }
                return false;
            };
            auto load_a = [&](const G3Tile& t, int c) {
                const int pad = (p.ksize[t.kc] - 1) >> 1;
                mb_wait(&a_empty[as], aph ^ 1);
                mb_expect_tx(&a_full[as], (uint32_t)p.box_bytes[t.kc]);
                tma_load_4d(a_buf + (size_t)as * p.a_stage_bytes, maps[t.kc], &a_full[as], c * KC, -pad, t.h0 - pad, t.r);
                if (++as == kG3AStages) {
                    as = 0;
                    aph ^= 1;
                }
            };
            bool have = advance(cur, 0, true, cur, cur_c);
            if (have) load_a(cur, cur_c);
            while (have) {
                const int k = p.ksize[cur.kc], G = (k + TPM - 1) / TPM;
                const int wrow = p.wrow[cur.e];
                const int nst = k * G;                           // weight stages of this item: (kernel row, tap group)
                const int pre = nst - 1 < kG3BStages ? nst - 1 : kG3BStages;
                bool have_next = false;
                for (int s = 0; s < nst; ++s) {
                    if (s == pre) {
                        have_next = advance(cur, cur_c, false, nxt, nxt_c);
                        if (have_next) load_a(nxt, nxt_c);
                    }
                    const int tr = s / G, g = s - tr * G;
                    mb_wait(&b_empty[bs], bph ^ 1);
                    mb_expect_tx(&b_full[bs], (uint32_t)B_STAGE);
                    // one tap group = 128 consecutive rows of the tap-major weight block (an incomplete last group drags
                    // in rows of the next kernel row / expert / OOB zeros: the MMA's N leaves them unused)
                    tma_load_2d(b_buf + (size_t)bs * B_STAGE, &tmap_b, &b_full[bs], cur_c * KC, wrow + (tr * k + g * TPM) * COUT);
                    if (++bs == kG3BStages) {
                        bs = 0;
                        bph ^= 1;
                    }
                }
                have = have_next;
                cur = nxt;
                cur_c = nxt_c;
            }
        }
    } else if (warp <= kG3Issuers) {
        // ============================== MMA issuers (warp w owns M-tile w-1) ==============================
        // The WHOLE warp runs this loop converged and one elected lane issues: tile geometry is made warp-uniform with
        // redux, so descriptors / barrier addresses stay in uniform registers (a loop entered by a single lane makes
        // the compiler wrap every UTCHMMA in an ELECT + R2UR waterfall: ~100 cycles per MMA per issuing thread).
        constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
        const int mt = warp - 1;
        int as = 0, bs = 0, acc = 0, tcount = 0;
        (void)tcount;
        uint32_t aph = 0, bph = 0, acc_ph = 0;
        for (;;) {
            const int i = uni(next_tile(true));
            if (i < 0) break;
            G3Tile t;
            const bool ok = tile_at(i, t);
            if (!uni((int)ok)) continue;
            const int k = uni(p.ksize[t.kc]), Wp = uni(p.wp[t.kc]), G = (k + TPM - 1) / TPM;
            const int c0u = uni(t.c0);
            const bool active = uni((int)(mt < t.mt_n)) != 0;
            if (mt == 0) G3T(0);
            mb_wait(&t_empty[acc], acc_ph ^ 1);
            tc_fence_after();
            if (mt == 0) G3T(1);
            const uint32_t d = tmem_base + (uint32_t)((acc * kG3Issuers + mt) * 128);
            for (int c = 0; c < p.upt; ++c) {
                mb_wait(&a_full[as], aph);
                tc_fence_after();
                if (mt == 0 && c == 0) G3T(2);
                const uint64_t a_desc0 =
                    umma_desc<KC>(s2u(a_buf + (size_t)as * p.a_stage_bytes) + (uint32_t)(c0u + mt * MS) * ROWB);
                for (int tr = 0; tr < k; ++tr)
                    for (int g = 0; g < G; ++g) {
                        mb_wait(&b_full[bs], bph);
                        tc_fence_after();
                        const int ntaps = (k - g * TPM) < TPM ? (k - g * TPM) : TPM;
                        const uint32_t idesc = idesc0 | ((uint32_t)((ntaps * COUT) >> 3) << 17);
                        const uint64_t bd = umma_desc<KC>(s2u(b_buf + (size_t)bs * B_STAGE));
                        const uint64_t ad = a_desc0 + (uint64_t)(((uint32_t)(tr * Wp + g * TPM) * ROWB) >> 4);
                        const uint32_t first = (uint32_t)(c | tr | g);
                        if (elect_one()) {
                            if (active) {
#pragma unroll
                                for (int kk = 0; kk < KC / 16; ++kk) tc_mma(d, ad + 2 * kk, bd + 2 * kk, idesc, first | kk);
                                tc_commit(&b_empty[bs]);
                            } else {
                                mb_arrive(&b_empty[bs]);
                            }
                        }
                        __syncwarp();
                        if (++bs == kG3BStages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                if (elect_one()) {
                    if (active) tc_commit(&a_empty[as]);
                    else mb_arrive(&a_empty[as]);
                }
                __syncwarp();
                if (++as == kG3AStages) {
                    as = 0;
                    aph ^= 1;
                }
            }
            if (elect_one()) {
                if (active) tc_commit(&t_full[acc]);
                else mb_arrive(&t_full[acc]);
            }
            __syncwarp();
            if (mt == 0) G3T(3);
            ++tcount;
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
        }
    } else {
        // ============================== epilogue (4 warps): TMEM -> shift-add -> registers -> global ==============
        const int quad = warp & 3;
        int acc = 0, mcount = 0, tcount = 0;
        (void)tcount;
        uint32_t acc_ph = 0;
        for (;;) {
            const int i = next_tile(true);
            if (i < 0) break;
            G3Tile t;
            if (!tile_at(i, t)) {
                // tiles of an unused tail row: zero-fill so downstream elementwise ops stay finite
                if (t.r >= n_rows || t.e < 0) {
                    const int rows_per = (p.H + p.smax - 1) / p.smax;
                    const int hs = t.j * rows_per, he = min(p.H, hs + rows_per);
                    const long long n16 = (long long)(he - hs) * p.W * COUT / 8;
                    int4* o = reinterpret_cast<int4*>(p.Y + ((size_t)t.r * p.H + hs) * p.W * COUT);
                    for (long long q = quad * 32 + lane; q < n16; q += 128) o[q] = make_int4(0, 0, 0, 0);
                }
                continue;
            }
            mb_wait(&t_full[acc], acc_ph);
            tc_fence_after();
            if (quad == 0) G3T(4);
            const int fl = (p.scale ? 1 : 0) | (p.act == 1 ? 2 : 0) | (p.res ? 4 : 0);
            const uint32_t tacc = tmem_base + (uint32_t)(acc * kG3Issuers * 128);
            switch (fl) {
                case 0: g3_epilogue<COUT, false, false, false>(p, t, quad, lane, tacc, xch, mcount); break;
                case 1: g3_epilogue<COUT, true, false, false>(p, t, quad, lane, tacc, xch, mcount); break;
                case 2: g3_epilogue<COUT, false, true, false>(p, t, quad, lane, tacc, xch, mcount); break;
                case 3: g3_epilogue<COUT, true, true, false>(p, t, quad, lane, tacc, xch, mcount); break;
                case 4: g3_epilogue<COUT, false, false, true>(p, t, quad, lane, tacc, xch, mcount); break;
                case 5: g3_epilogue<COUT, true, false, true>(p, t, quad, lane, tacc, xch, mcount); break;
                case 6: g3_epilogue<COUT, false, true, true>(p, t, quad, lane, tacc, xch, mcount); break;
                default: g3_epilogue<COUT, true, true, true>(p, t, quad, lane, tacc, xch, mcount); break;
            }
            tc_fence_before();
            __syncwarp();
            if (quad == 0) G3T(5);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
}

template <int KC, int COUT>
static int launch_gconv3(const CUtensorMap* ta, const CUtensorMap& tb, const GConv3Params& p, cudaStream_t st) {
    auto kfn = gconv3_fwd_kernel<KC, COUT>;
    const int smem = kG3AStages * p.a_stage_bytes + kG3BStages * 128 * KC * 2 + 1024;
    HDMOE_CHECK_ARG(smem <= 210 * 1024, "gconv3: tile does not fit shared memory (%d bytes)", smem);
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
    kfn<<<grid, kG3Threads, smem, st>>>(ta[0], ta[1], ta[2], ta[3], tb, p);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

}  // namespace hdmoe
using namespace hdmoe;

#ifdef HDMOE_G3_TRACE
extern "C" int hdmoe_g3_trace_read(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g3_trace, sizeof(long long) * 148 * 64);
}
#endif

extern "C" int hdmoe_gconv3_fwd(const void* X, const void* Wt, void* Y, int cap_rows, int H, int W, int Cin_pad,
                                int Cout, int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev,
                                int n_experts, const int32_t* ksize_host, const int32_t* wrow_host, const float* scale,
                                int act, const void* residual, float res_a, float res_b, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(X && Wt && Y && row_expert && n_rows_dev && ksize_host && wrow_host, "gconv3_fwd: null pointer");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= kG3MaxE, "gconv3_fwd: 1 <= n_experts <= %d", kG3MaxE);
    HDMOE_CHECK_ARG(Cout == 32 || Cout == 64, "gconv3_fwd: Cout must be 32 or 64 (use hdmoe_gconv2_fwd otherwise)");
    HDMOE_CHECK_ARG(Cin_pad >= 32 && Cin_pad % 32 == 0, "gconv3_fwd: Cin_pad must be a multiple of 32 (got %d)", Cin_pad);
    HDMOE_CHECK_ARG(H >= 1 && H <= 255 && W >= 1 && W <= 248, "gconv3_fwd: H <= 255, W <= 248");
    HDMOE_CHECK_ARG((((uintptr_t)X | (uintptr_t)Wt) & 15) == 0, "gconv3_fwd: 16-byte alignment required");
    HDMOE_CHECK_ARG((((uintptr_t)Y | (uintptr_t)residual) & 31) == 0, "gconv3_fwd: Y / residual need 32-byte alignment");
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) {
        set_error("gconv3_fwd: cuTensorMapEncodeTiled not available from the driver");
        return HDMOE_ERR_CUDA;
    }
    const int KC = (Cin_pad % 64 == 0) ? 64 : 32;
    const int TPM = 128 / Cout, MS = 128 - (TPM - 1);
    GConv3Params p{};
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
    int ncls = 0, cls_k[kG3Classes];
    for (int e = 0; e < n_experts; ++e) {
        const int k = ksize_host[e];
        HDMOE_CHECK_ARG(k >= 1 && k <= 7 && (k & 1), "gconv3_fwd: odd kernel sizes 1..7");
        int c = -1;
        for (int q = 0; q < ncls; ++q)
            if (cls_k[q] == k) c = q;
        if (c < 0) {
            HDMOE_CHECK_ARG(ncls < kG3Classes, "gconv3_fwd: at most %d distinct kernel sizes per launch", kG3Classes);
            c = ncls++;
            cls_k[c] = k;
        }
        p.kclass[e] = (uint8_t)c;
        p.wrow[e] = wrow_host[e];
    }
    int smax = 0, box_bytes_max = 0, cls_box_rows[kG3Classes];
    const int b_ring_bytes = kG3BStages * 128 * KC * 2;
    for (int mt_try = kG3Issuers; mt_try >= 1; --mt_try) {
        smax = 0;
        box_bytes_max = 0;
        const int PT = mt_try * MS;
        for (int c = 0; c < ncls; ++c) {
            const int k = cls_k[c], Wp = W + k - 1, npos = H * Wp;
            const int n = (npos + PT - 1) / PT;
            // input rows one tile's TMA box must hold: the tile's M-tiles (128 positions each, stride MS) plus the largest
            // tap offset (k-1)*(Wp+1), counted from the start of the image row that contains the tile's first position
            int box_rows = 0;
            for (int j = 0; j < n; ++j) {
                const int p0 = j * PT, c0 = p0 % Wp;
                int mt_n = (npos - p0 + MS - 1) / MS;
                if (mt_n > mt_try) mt_n = mt_try;
                const int rows = (c0 + (mt_n - 1) * MS + 127 + (k - 1) * (Wp + 1)) / Wp + 1;
                if (rows > box_rows) box_rows = rows;
            }
            HDMOE_CHECK_ARG(box_rows <= 256 && Wp <= 256, "gconv3_fwd: TMA box too large");
            p.ksize[c] = k;
            p.wp[c] = Wp;
            p.npos[c] = npos;
            p.ntile[c] = n;
            cls_box_rows[c] = box_rows;
            p.box_bytes[c] = box_rows * Wp * KC * 2;
            if (n > smax) smax = n;
            if (p.box_bytes[c] > box_bytes_max) box_bytes_max = p.box_bytes[c];
        }
        p.mt_per_tile = mt_try;
        if (kG3AStages * (((box_bytes_max + 1023) / 1024) * 1024) + b_ring_bytes + 1024 <= 210 * 1024) break;
    }
    p.smax = smax;
    p.n_tiles = cap_rows * smax;
    p.a_stage_bytes = ((box_bytes_max + 1023) / 1024) * 1024;
    cudaStream_t st = (cudaStream_t)stream;
    p.sched = sched_slot(st);
    HDMOE_CHECK_ARG(p.sched != nullptr, "gconv3_fwd: more than %d distinct streams in use", kSchedSlots);
    CUtensorMap ta[kG3Classes], tb;
    const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int c = 0; c < kG3Classes; ++c) {
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
            set_error("gconv3_fwd: cuTensorMapEncodeTiled(A, class %d) failed with %d", c, (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin_pad, (cuuint64_t)w_rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)Cin_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, 128};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Wt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("gconv3_fwd: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
    if (KC == 64 && Cout == 64) return launch_gconv3<64, 64>(ta, tb, p, st);
    if (KC == 64 && Cout == 32) return launch_gconv3<64, 32>(ta, tb, p, st);
    if (KC == 32 && Cout == 64) return launch_gconv3<32, 64>(ta, tb, p, st);
    return launch_gconv3<32, 32>(ta, tb, p, st);
}
