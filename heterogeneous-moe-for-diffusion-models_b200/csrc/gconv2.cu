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
//   * M-tiles cover 128 consecutive positions; positions in the padding columns produce garbage rows that the
//     epilogue drops (cost: Wp/W and 128-rounding, 67-92 % MMA efficiency, far cheaper than the re-reads);
//   * weights stream through a small ring, one [Cout x KC] box per (tap, chunk), reused by all M-tiles.
//
// Everything else follows v1: per-tile expert lookup (kernel size, weight block) for the grouped /
// heterogeneous case, warp-specialised roles (TMA producer, single-thread MMA issuer, 4 epilogue warps),
// double-buffered TMEM accumulators, fused epilogue (scale, mp_silu, mp_sum residual), NHWC bf16 output.
#include "tc.cuh"
#include "../../include/hdmoe_gemm.h"

namespace hdmoe {

#ifdef HDMOE_G2_TRACE
__device__ long long g2_trace[148 * 64];
#define G2T(slot) do { if (blockIdx.x < 148 && tcount < 8) g2_trace[blockIdx.x * 64 + tcount * 8 + (slot)] = clock64(); } while (0)
#else
#define G2T(slot) do { } while (0)
#endif

constexpr int kG2Issuers = 3;      // MMA-issuer warps, one per M-tile accumulator (a single thread cannot issue
                                   // 32-cycle N=64 MMAs fast enough: ~12 uniform-datapath instructions per UTCHMMA)
constexpr int kG2Threads = 32 * (1 + kG2Issuers + 4);
constexpr int kG2MaxE = HDMOE_MAX_EXPERTS;
constexpr int kG2Classes = 4;     // distinct kernel sizes per launch
constexpr int kG2MaxStrips = 32;

struct GConv2Params {
    int n_tiles, smax;                // tiles = cap_rows * smax (strips per sample, max over classes)
    int H, W;
    int upt;                          // channel chunks of KC per tap
    int n_experts;
    int a_stage_bytes;                // bytes of one halo buffer
    const int32_t* row_expert;
    const int32_t* n_rows_dev;
    __nv_bfloat16* Y;
    const float* scale;
    const __nv_bfloat16* res;
    float res_a, res_b;
    int act;
    int32_t wrow[kG2MaxE];
    uint8_t kclass[kG2MaxE];
    // per kernel-size class
    int32_t ksize[kG2Classes], wp[kG2Classes], box_bytes[kG2Classes];
    uint8_t nstrips[kG2Classes];
    uint8_t strip_h0[kG2Classes][kG2MaxStrips], strip_sh[kG2Classes][kG2MaxStrips];
};

template <int KC, int N>
struct Conv2Cfg {
    static constexpr int MT_MAX = N <= 64 ? 3 : 2;             // M-tiles (of 128 positions) per strip (<= kG2Issuers)
    static constexpr int ROWB = KC * 2;                        // bytes per position
    static constexpr int TPS = N <= 64 ? 2 : 1;                // filter taps per weight stage (one TMA box of TPS*N rows)
    static constexpr int B_TAP = N * KC * 2;                   // bytes of one tap's [N x KC] weight tile
    static constexpr int B_STAGE = TPS * B_TAP;
    static constexpr int A_STAGES = 2;
    static constexpr int B_STAGES = 4;
    static constexpr int TMEM_NEED = 2 * MT_MAX * N;
    static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512));
};

template <int KC, int N>
__global__ void __launch_bounds__(kG2Threads, 1)
gconv2_fwd_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
                  const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap ta3,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ GConv2Params p) {
    using Cfg = Conv2Cfg<KC, N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;                                              // A_STAGES halo buffers
    uint8_t* b_buf = smem + (size_t)Cfg::A_STAGES * p.a_stage_bytes;    // weight ring
    uint8_t* stage_buf = b_buf + (size_t)Cfg::B_STAGES * Cfg::B_STAGE;  // epilogue staging: 4 warps x 32 rows x N bf16
    __shared__ int32_t stage_off[4 * 32];
    __shared__ __align__(8) uint64_t a_full[Cfg::A_STAGES], a_empty[Cfg::A_STAGES], b_full[Cfg::B_STAGES],
        b_empty[Cfg::B_STAGES], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::A_STAGES; ++s) {
            mb_init(&a_full[s], 1);
            mb_init(&a_empty[s], kG2Issuers);
        }
        for (int s = 0; s < Cfg::B_STAGES; ++s) {
            mb_init(&b_full[s], 1);
            mb_init(&b_empty[s], kG2Issuers);
        }
        for (int a = 0; a < 2; ++a) {
            mb_init(&t_full[a], kG2Issuers);
            mb_init(&t_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
    const int n_rows = *p.n_rows_dev;

    // tile -> (row, strip); every role walks the same sequence and skips the same tiles
    auto tile_at = [&](int i, int& r, int& j, int& e, int& kc) -> bool {
        // strip-major, rows reversed inside a strip: the heavy (large-kernel) experts sit at the end of the row
        // order and are scheduled first, and a CTA's static stride (gridDim = 148 = 4*37) does not alias with
        // the strip index (a row-major order would hand some CTAs only the small last strips)
        const int cap = p.n_tiles / p.smax;
        j = i / cap;
        r = cap - 1 - (i - j * cap);
        e = -1;
        kc = 0;
        if (r >= n_rows) return false;
        e = p.row_expert[r];
        if (e < 0 || e >= p.n_experts) return false;
        kc = p.kclass[e];
        return j < p.nstrips[kc];
    };

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            const CUtensorMap* maps[kG2Classes] = {&ta0, &ta1, &ta2, &ta3};
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            for (int i = blockIdx.x; i < p.n_tiles; i += gridDim.x) {
                int r, j, e, kc;
                if (!tile_at(i, r, j, e, kc)) continue;
                const int k = p.ksize[kc], pad = (k - 1) >> 1, taps = k * k;
                const int h0 = p.strip_h0[kc][j];
                const int wrow = p.wrow[e];
                for (int c = 0; c < p.upt; ++c) {
                    mb_wait(&a_empty[as], aph ^ 1);
                    mb_expect_tx(&a_full[as], (uint32_t)p.box_bytes[kc]);
                    tma_load_4d(a_buf + (size_t)as * p.a_stage_bytes, maps[kc], &a_full[as], c * KC, -pad, h0 - pad, r);
                    if (++as == Cfg::A_STAGES) {
                        as = 0;
                        aph ^= 1;
                    }
                    for (int t = 0; t < taps; t += Cfg::TPS) {
                        mb_wait(&b_empty[bs], bph ^ 1);
                        mb_expect_tx(&b_full[bs], (uint32_t)Cfg::B_STAGE);
                        // TPS consecutive taps are TPS*N consecutive rows of the tap-major weight block: one box
                        // (a trailing odd tap drags in N rows of the next block / OOB zeros; they are not used)
                        tma_load_2d(b_buf + (size_t)bs * Cfg::B_STAGE, &tmap_b, &b_full[bs], c * KC, wrow + t * N);
                        if (++bs == Cfg::B_STAGES) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp <= kG2Issuers) {
        // ============================== MMA issuers (warp w owns M-tile w-1) ==============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                                       ((uint32_t)(128 >> 4) << 24);
            const int mt = warp - 1;
            int tcount = 0;
            (void)tcount;
            int as = 0, bs = 0, acc = 0;
            uint32_t aph = 0, bph = 0, acc_ph = 0;
            for (int i = blockIdx.x; i < p.n_tiles; i += gridDim.x) {
                int r, j, e, kc;
                if (!tile_at(i, r, j, e, kc)) continue;
                const int k = p.ksize[kc], Wp = p.wp[kc], taps = k * k;
                const int sh = p.strip_sh[kc][j];
                const bool active = mt < ((sh * Wp + 127) >> 7);
                if (mt == 0) G2T(0);
                mb_wait(&t_empty[acc], acc_ph ^ 1);
                tc_fence_after();
                if (mt == 0) G2T(1);
                const uint32_t d = tmem_base + (uint32_t)((acc * Cfg::MT_MAX + mt) * N);
                for (int c = 0; c < p.upt; ++c) {
                    mb_wait(&a_full[as], aph);
                    tc_fence_after();
                    if (mt == 0 && c == 0) G2T(2);
                    // descriptor of this M-tile's first position; taps / k-slices only add to the 14-bit address field
                    const uint64_t a_desc0 = umma_desc<KC>(s2u(a_buf + (size_t)as * p.a_stage_bytes) + (uint32_t)mt * 128u * Cfg::ROWB);
                    for (int t0 = 0; t0 < taps; t0 += Cfg::TPS) {
                        mb_wait(&b_full[bs], bph);
                        if (active) {
                            tc_fence_after();
                            const uint64_t bd0 = umma_desc<KC>(s2u(b_buf + (size_t)bs * Cfg::B_STAGE));
#pragma unroll
                            for (int q = 0; q < Cfg::TPS; ++q) {
                                const int t = t0 + q;
                                if (t < taps) {
                                    const int tr = t / k, ts = t - tr * k;
                                    const uint64_t ad = a_desc0 + (uint64_t)(((uint32_t)(tr * Wp + ts) * Cfg::ROWB) >> 4);
                                    const uint64_t bd = bd0 + (uint64_t)((q * Cfg::B_TAP) >> 4);
#pragma unroll
                                    for (int kk = 0; kk < KC / 16; ++kk)
                                        tc_mma(d, ad + 2 * kk, bd + 2 * kk, idesc, (c | t | kk) != 0);
                                }
                            }
                            tc_commit(&b_empty[bs]);
                        } else {
                            mb_arrive(&b_empty[bs]);
                        }
                        if (++bs == Cfg::B_STAGES) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                    if (active) tc_commit(&a_empty[as]);
                    else mb_arrive(&a_empty[as]);
                    if (++as == Cfg::A_STAGES) {
                        as = 0;
                        aph ^= 1;
                    }
                }
                if (active) tc_commit(&t_full[acc]);
                else mb_arrive(&t_full[acc]);
                if (mt == 0) G2T(3);
                ++tcount;
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
        }
    } else {
        // ============================== epilogue (4 warps) ==============================
        const int quad = warp & 3;
        int acc = 0;
        int tcount = 0;
        (void)tcount;
        uint32_t acc_ph = 0;
        for (int i = blockIdx.x; i < p.n_tiles; i += gridDim.x) {
            int r, j, e, kc;
            if (!tile_at(i, r, j, e, kc)) {
                // strips of an unused tail row: zero-fill so downstream elementwise ops stay finite
                if (r >= n_rows || e < 0) {
                    const int rows_per = (p.H + p.smax - 1) / p.smax;
                    const int hs = j * rows_per, he = min(p.H, hs + rows_per);
                    const long long n16 = (long long)(he - hs) * p.W * N / 8;
                    int4* o = reinterpret_cast<int4*>(p.Y + ((size_t)r * p.H + hs) * p.W * N);
                    for (long long q = quad * 32 + lane; q < n16; q += 128) o[q] = make_int4(0, 0, 0, 0);
                }
                continue;
            }
            const int Wp = p.wp[kc];
            const int h0 = p.strip_h0[kc][j], sh = p.strip_sh[kc][j];
            const int mt_n = (sh * Wp + 127) >> 7;
            mb_wait(&t_full[acc], acc_ph);
            tc_fence_after();
            if (quad == 0 && lane == 0) G2T(4);
            const float* sc = p.scale ? p.scale + (size_t)r * N : nullptr;
            // Each thread owns one position's channel vector; writing it straight to global memory would touch 32
            // cache lines per warp store.  Stage the warp's 32 x N tile in shared memory (16-byte chunks, XOR
            // swizzled against bank conflicts) and write it out with consecutive lanes on consecutive chunks.
            constexpr int RB = N * 2, CPR = RB / 16;                    // row bytes, 16-byte chunks per row
            uint8_t* stg = stage_buf + (size_t)quad * 32 * RB;
            int32_t* poff = stage_off + quad * 32;
            const int my_swz = (RB % 128 == 0) ? (lane & 7) : ((lane >> 1) & 3);
            for (int mt = 0; mt < mt_n; ++mt) {
                const int q = mt * 128 + quad * 32 + lane;
                const int hl = q / Wp, w = q - hl * Wp;
                const bool valid = hl < sh && w < p.W;
                poff[lane] = valid ? (int32_t)((((size_t)r * p.H + (size_t)(h0 + hl)) * p.W + w)) : -1;
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((acc * Cfg::MT_MAX + mt) * N + c0), v);
                    uint32_t packed[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        float a = __uint_as_float(v[2 * u]), b = __uint_as_float(v[2 * u + 1]);
                        if (sc) {
                            a *= sc[c0 + 2 * u];
                            b *= sc[c0 + 2 * u + 1];
                        }
                        if (p.act == 1) {
                            a = a / (1.f + __expf(-a)) * (1.f / 0.596f);
                            b = b / (1.f + __expf(-b)) * (1.f / 0.596f);
                        }
                        __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
                        packed[u] = *reinterpret_cast<uint32_t*>(&o);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = c0 / 8 + u;
                        *reinterpret_cast<int4*>(stg + lane * RB + ((j ^ my_swz) << 4)) =
                            make_int4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < CPR; ++it) {
                    const int f = it * 32 + lane;
                    const int row = f / CPR, ch = f - row * CPR;
                    const int pix = poff[row];
                    if (pix >= 0) {
                        const int swz = (RB % 128 == 0) ? (row & 7) : ((row >> 1) & 3);
                        int4 val = *reinterpret_cast<const int4*>(stg + row * RB + ((ch ^ swz) << 4));
                        const size_t go = (size_t)pix * N + ch * 8;
                        if (p.res) {   // mp_sum folded: out = res_a * residual + res_b * value
                            const int4 rr = *reinterpret_cast<const int4*>(p.res + go);
                            const uint32_t* vi = reinterpret_cast<const uint32_t*>(&val);
                            const uint32_t* ri = reinterpret_cast<const uint32_t*>(&rr);
                            uint32_t oo[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float v0 = __uint_as_float(vi[u] << 16), v1 = __uint_as_float(vi[u] & 0xffff0000u);
                                const float r0 = __uint_as_float(ri[u] << 16), r1 = __uint_as_float(ri[u] & 0xffff0000u);
                                __nv_bfloat162 o = __floats2bfloat162_rn(p.res_a * r0 + p.res_b * v0, p.res_a * r1 + p.res_b * v1);
                                oo[u] = *reinterpret_cast<uint32_t*>(&o);
                            }
                            val = make_int4(oo[0], oo[1], oo[2], oo[3]);
                        }
                        *reinterpret_cast<int4*>(p.Y + go) = val;
                    }
                }
                __syncwarp();
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
}

template <int KC, int N>
static int launch_gconv2(const CUtensorMap* ta, const CUtensorMap& tb, const GConv2Params& p, cudaStream_t st) {
    using Cfg = Conv2Cfg<KC, N>;
    auto kfn = gconv2_fwd_kernel<KC, N>;
    const int smem = Cfg::A_STAGES * p.a_stage_bytes + Cfg::B_STAGES * Cfg::B_STAGE + 4 * 32 * N * 2 + 1024;
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
    HDMOE_CHECK_ARG((((uintptr_t)X | (uintptr_t)Wt | (uintptr_t)Y) & 15) == 0, "gconv2_fwd: 16-byte alignment required");
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
    int ncls = 0, cls_k[kG2Classes], cls_shmax[kG2Classes];
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
    int smax = 0, a_rows_max = 0;
    for (int c = 0; c < ncls; ++c) {
        const int k = cls_k[c], Wp = W + k - 1;
        // strip height: the candidate (1..mt_max M-tiles) that needs the fewest M-tiles for the whole image
        int best_sh = 0, best_cost = 1 << 30, best_n = 0;
        for (int mt = mt_max; mt >= 1; --mt) {
            int sh = (mt * 128) / Wp;
            if (sh > H) sh = H;
            if (sh < 1) continue;
            const int n = (H + sh - 1) / sh;
            if (n > kG2MaxStrips) continue;
            int cost = 0;
            for (int j = 0; j < n; ++j) {
                const int s_ = (j == n - 1) ? H - sh * (n - 1) : sh;
                cost += (s_ * Wp + 127) / 128;
            }
            if (cost < best_cost) {
                best_cost = cost;
                best_sh = sh;
                best_n = n;
            }
        }
        HDMOE_CHECK_ARG(best_sh > 0, "gconv2_fwd: image %dx%d with kernel %d does not tile", H, W, k);
        p.ksize[c] = k;
        p.wp[c] = Wp;
        p.nstrips[c] = (uint8_t)best_n;
        for (int j = 0; j < best_n; ++j) {
            p.strip_h0[c][j] = (uint8_t)(j * best_sh);
            p.strip_sh[c][j] = (uint8_t)((j == best_n - 1) ? H - best_sh * (best_n - 1) : best_sh);
        }
        cls_shmax[c] = best_sh;
        p.box_bytes[c] = (best_sh + k - 1) * Wp * KC * 2;
        if (best_n > smax) smax = best_n;
        const int a_rows = mt_max * 128 + (k - 1) * (Wp + 1);
        const int box_rows = (best_sh + k - 1) * Wp;
        if (a_rows > a_rows_max) a_rows_max = a_rows;
        if (box_rows > a_rows_max) a_rows_max = box_rows;
        HDMOE_CHECK_ARG(best_sh + k - 1 <= 256 && Wp <= 256, "gconv2_fwd: TMA box too large");
    }
    p.smax = smax;
    p.n_tiles = cap_rows * smax;
    p.a_stage_bytes = ((a_rows_max * KC * 2 + 1023) / 1024) * 1024;
    CUtensorMap ta[kG2Classes], tb;
    const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int c = 0; c < kG2Classes; ++c) {
        const int cc = c < ncls ? c : 0;
        const int k = cls_k[cc], Wp = W + k - 1;
        cuuint64_t dims[4] = {(cuuint64_t)Cin_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap_rows};
        cuuint64_t strides[3] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)W * Cin_pad * 2, (cuuint64_t)H * W * Cin_pad * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)Wp, (cuuint32_t)(cls_shmax[cc] + k - 1), 1};
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
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(Cout <= 64 ? 2 * Cout : Cout)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Wt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("gconv2_fwd: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
#define GC2(KCV, NV) \
    if (KC == KCV && Cout == NV) return launch_gconv2<KCV, NV>(ta, tb, p, st);
    GC2(32, 32) GC2(32, 64) GC2(32, 96) GC2(32, 128) GC2(64, 32) GC2(64, 64) GC2(64, 96) GC2(64, 128)
#undef GC2
    HDMOE_CHECK_ARG(false, "gconv2_fwd: unsupported (KC=%d, Cout=%d)", KC, Cout);
}
