// G-CONV: grouped implicit-GEMM convolution on tcgen05 tensor cores with TMEM accumulators, fed by TMA.
//
// Replaces the F.conv2d / F.linear calls of MP_Conv inside the experts (models/model_internals.py:261-271)
// for ALL experts of one layer in a single persistent launch: "group" = expert, the per-group problem is
// M_e = (rows routed to e) * H * W, N = Cout, K_e = Cin * k_e^2 with a per-expert kernel size k_e
// (Utils/configs.py:23: 3,3,5,5) -- variable width AND variable token count.
//
// Formulation.  Activations are NHWC bf16 in the expert-major permuted row order of the dispatch plan, so
// a tile of 128 output pixels (BH x BW block of one sample) never straddles experts.  For filter tap (r,s)
// and a chunk of KC input channels the A operand is the BH x BW x KC box of the input shifted by
// (r-pad, s-pad): ONE 4-D TMA load, and the TMA's out-of-bounds zero fill implements the 'same' padding
// (models/model_internals.py:268-271) with no halo logic.  The box lands in shared memory pixel-major with
// KC*2-byte rows in the 64B / 128B swizzle, which is exactly the canonical K-major UMMA operand layout.
// B is the [Cout x KC] slice of the tap-major prepared weights (W-PREP layout [tap][Cout][Cin_pad]) of the
// tile's expert.  D[128 x Cout] accumulates in TMEM over taps x chunks; 4 epilogue warps read it back with
// tcgen05.ld, apply the fused epilogue and store NHWC bf16 (each thread owns one pixel's channel vector, so
// stores are 16-byte vectors and per-pixel ops such as pixel-norm are thread-local).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2-5 = epilogue (TMEM lane quadrant = warp_id % 4).  Pipelines: smem full/empty ring (kStages),
// TMEM full/empty (2 accumulators), static persistent tile schedule, heavy (5x5) tiles first.
#include "tc.cuh"
#include "../../include/hdmoe_gemm.h"

namespace hdmoe {

constexpr int kTileM = 128;
constexpr int kStageK = 64;   // channels-worth of K per pipeline stage (one 64-ch unit or two 32-ch units)
constexpr int kMaxE = HDMOE_MAX_EXPERTS;
constexpr int kConvThreads = 192;

struct GConvParams {
    int n_tiles, tiles_per_sample;
    int H, W, BW, BH;
    int upt;                 // units (channel chunks of KC) per tap
    int n_experts;
    int reverse;             // schedule tiles last-to-first (heavy experts sit at the end)
    const int32_t* row_expert;
    const int32_t* n_rows_dev;
    __nv_bfloat16* Y;
    const float* scale;      // optional [rows][Cout] per-sample channel gain (1 + emb), applied before `act`
    const __nv_bfloat16* res;  // optional residual, NHWC [rows][H][W][Cout]
    float res_a, res_b;      // out = res_a * res + res_b * f(acc)      (mp_sum folded: lerp/sqrt)
    int act;                 // 0 = none, 1 = mp_silu
    int32_t ksize[kMaxE];
    int32_t wrow[kMaxE];
};

template <int KC, int N>
struct ConvCfg {
    static constexpr int UPS = kStageK / KC;                  // units per stage
    static constexpr int A_UNIT = kTileM * KC * 2;            // bytes
    static constexpr int B_UNIT = N * KC * 2;
    static constexpr int STAGE = UPS * (A_UNIT + B_UNIT);
    static constexpr int STAGES = (200 * 1024) / STAGE > 8 ? 8 : (200 * 1024) / STAGE;
    static constexpr int TMEM_COLS = 2 * N <= 64 ? 64 : (2 * N <= 128 ? 128 : 256);
    static constexpr int SMEM = STAGES * STAGE + 1024;        // + alignment slack
};

// ------------------------------------------------------------------------------------------------- kernel
template <int KC, int N>
__global__ void __launch_bounds__(kConvThreads, 1)
gconv_fwd_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ GConvParams p) {
    using Cfg = ConvCfg<KC, N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES], empty_bar[Cfg::STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mb_init(&full_bar[s], 1);
            mb_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mb_init(&tfull_bar[a], 1);
            mb_init(&tempty_bar[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-collective; this warp also frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&tmem_base_s)),
                     "n"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int n_rows = *p.n_rows_dev;

    auto a_unit = [&](int stage, int j) { return smem + (size_t)stage * Cfg::STAGE + (size_t)j * Cfg::A_UNIT; };
    auto b_unit = [&](int stage, int j) {
        return smem + (size_t)stage * Cfg::STAGE + (size_t)Cfg::UPS * Cfg::A_UNIT + (size_t)j * Cfg::B_UNIT;
    };
    // every role walks the same tile sequence; a tile of an empty / tail row is skipped by all of them
    auto tile_at = [&](int i, int& r, int& pb, int& e) -> bool {
        const int t = p.reverse ? p.n_tiles - 1 - i : i;
        r = t / p.tiles_per_sample;
        pb = t - r * p.tiles_per_sample;
        if (r >= n_rows) return false;
        e = p.row_expert[r];
        return e >= 0 && e < p.n_experts;
    };

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int i = blockIdx.x; i < p.n_tiles; i += gridDim.x) {
                int r, pb, e;
                if (!tile_at(i, r, pb, e)) continue;
                const int k = p.ksize[e], pad = (k - 1) >> 1;
                const int nu = k * k * p.upt;
                const int h0 = pb * p.BH;
                const int wrow = p.wrow[e];
                for (int u0 = 0; u0 < nu; u0 += Cfg::UPS) {
                    const int nun = min(Cfg::UPS, nu - u0);
                    mb_wait(&empty_bar[stage], phase ^ 1);
                    mb_expect_tx(&full_bar[stage], (uint32_t)nun * (Cfg::A_UNIT + Cfg::B_UNIT));
                    for (int j = 0; j < nun; ++j) {
                        const int u = u0 + j;
                        const int tap = u / p.upt, ch = u - tap * p.upt;
                        const int tr = tap / k, ts = tap - tr * k;
                        tma_load_4d(a_unit(stage, j), &tmap_a, &full_bar[stage], ch * KC, ts - pad, h0 + tr - pad, r);
                        tma_load_2d(b_unit(stage, j), &tmap_b, &full_bar[stage], ch * KC, wrow + tap * N);
                    }
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A=B=bf16 [7,10)/[10,13),
            // K-major A and B (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                                       ((uint32_t)(kTileM >> 4) << 24);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int i = blockIdx.x; i < p.n_tiles; i += gridDim.x) {
                int r, pb, e;
                if (!tile_at(i, r, pb, e)) continue;
                const int k = p.ksize[e];
                const int nu = k * k * p.upt;
                mb_wait(&tempty_bar[acc], acc_phase ^ 1);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                uint32_t accumulate = 0;
                for (int u0 = 0; u0 < nu; u0 += Cfg::UPS) {
                    const int nun = min(Cfg::UPS, nu - u0);
                    mb_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    for (int j = 0; j < nun; ++j) {
                        const uint32_t a0 = s2u(a_unit(stage, j)), b0 = s2u(b_unit(stage, j));
#pragma unroll
                        for (int kk = 0; kk < KC / 16; ++kk) {
                            tc_mma(d_tmem, umma_desc<KC>(a0 + kk * 32), umma_desc<KC>(b0 + kk * 32), idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    tc_commit(&empty_bar[stage]);             // smem slot free once these MMAs retire
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&tfull_bar[acc]);                   // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ============================== epilogue (4 warps) ==============================
        const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
        const int pix = quad * 32 + lane;                     // pixel inside the tile == TMEM lane
        const int hh = pix / p.BW, ww = pix - hh * p.BW;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int i = blockIdx.x; i < p.n_tiles; i += gridDim.x) {
            int r, pb, e;
            if (!tile_at(i, r, pb, e)) {
                // unused tail row: define the output (zeros) so that downstream elementwise ops stay finite
                __nv_bfloat16* o = p.Y + (((size_t)r * p.H + (size_t)(pb * p.BH + hh)) * p.W + ww) * N;
#pragma unroll
                for (int q = 0; q < N / 8; ++q) *reinterpret_cast<int4*>(o + 8 * q) = make_int4(0, 0, 0, 0);
                continue;
            }
            mb_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const size_t pix_off = (((size_t)r * p.H + (size_t)(pb * p.BH + hh)) * p.W + ww) * N;
            __nv_bfloat16* out = p.Y + pix_off;
            const float* sc = p.scale ? p.scale + (size_t)r * N : nullptr;
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * N + c0), v);
                uint32_t packed[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float a = __uint_as_float(v[2 * q]), b = __uint_as_float(v[2 * q + 1]);
                    if (sc) {
                        a *= sc[c0 + 2 * q];
                        b *= sc[c0 + 2 * q + 1];
                    }
                    if (p.act == 1) {   // mp_silu (models/model_internals.py:47)
                        a = a / (1.f + __expf(-a)) * (1.f / 0.596f);
                        b = b / (1.f + __expf(-b)) * (1.f / 0.596f);
                    }
                    if (p.res) {
                        const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(p.res + pix_off + c0 + 2 * q);
                        a = p.res_a * __low2float(rr) + p.res_b * a;
                        b = p.res_a * __high2float(rr) + p.res_b * b;
                    }
                    __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
                    packed[q] = *reinterpret_cast<uint32_t*>(&o);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<int4*>(out + c0 + 8 * q) =
                        make_int4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mb_arrive(&tempty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
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

// ----------------------------------------------------------------------------------------------- host side
template <int KC, int N>
static int launch_gconv(const CUtensorMap& ta, const CUtensorMap& tb, const GConvParams& p, cudaStream_t st) {
    using Cfg = ConvCfg<KC, N>;
    auto kfn = gconv_fwd_kernel<KC, N>;
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
    kfn<<<grid, kConvThreads, Cfg::SMEM, st>>>(ta, tb, p);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_gconv_fwd(const void* X, const void* Wt, void* Y, int cap_rows, int H, int W, int Cin_pad,
                               int Cout, int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev,
                               int n_experts, const int32_t* ksize_host, const int32_t* wrow_host, const float* scale,
                               int act, const void* residual, float res_a, float res_b, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(X && Wt && Y && row_expert && n_rows_dev && ksize_host && wrow_host, "gconv_fwd: null pointer");
    HDMOE_CHECK_ARG(n_experts >= 1 && n_experts <= kMaxE, "gconv_fwd: 1 <= n_experts <= %d", kMaxE);
    HDMOE_CHECK_ARG(Cout == 32 || Cout == 64 || Cout == 96 || Cout == 128, "gconv_fwd: Cout must be 32, 64, 96 or 128 (got %d)", Cout);
    HDMOE_CHECK_ARG(Cin_pad >= 32 && Cin_pad % 32 == 0, "gconv_fwd: Cin_pad must be a multiple of 32 (got %d)", Cin_pad);
    HDMOE_CHECK_ARG(W >= 1 && W <= 128 && 128 % W == 0 && (H * W) % 128 == 0 && H % (128 / W) == 0,
                    "gconv_fwd: need W | 128 and 128 | H*W (got %dx%d)", H, W);
    HDMOE_CHECK_ARG((((uintptr_t)X | (uintptr_t)Wt | (uintptr_t)Y) & 15) == 0, "gconv_fwd: 16-byte alignment required");
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) {
        set_error("gconv_fwd: cuTensorMapEncodeTiled not available from the driver");
        return HDMOE_ERR_CUDA;
    }
    const int KC = (Cin_pad % 64 == 0) ? 64 : 32;
    GConvParams p{};
    p.BW = W;
    p.BH = 128 / W;
    p.H = H;
    p.W = W;
    p.tiles_per_sample = H * W / 128;
    p.n_tiles = cap_rows * p.tiles_per_sample;
    p.upt = Cin_pad / KC;
    p.n_experts = n_experts;
    p.reverse = 1;
    p.row_expert = row_expert;
    p.n_rows_dev = n_rows_dev;
    p.Y = (__nv_bfloat16*)Y;
    p.scale = scale;
    p.act = act;
    p.res = (const __nv_bfloat16*)residual;
    p.res_a = res_a;
    p.res_b = res_b;
    for (int e = 0; e < n_experts; ++e) {
        HDMOE_CHECK_ARG(ksize_host[e] >= 1 && ksize_host[e] <= 7 && (ksize_host[e] & 1), "gconv_fwd: odd kernel sizes 1..7");
        p.ksize[e] = ksize_host[e];
        p.wrow[e] = wrow_host[e];
    }
    CUtensorMap ta, tb;
    const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap_rows};
        cuuint64_t strides[3] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)W * Cin_pad * 2, (cuuint64_t)H * W * Cin_pad * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("gconv_fwd: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin_pad, (cuuint64_t)w_rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)Cin_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)Cout};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Wt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("gconv_fwd: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
            return HDMOE_ERR_CUDA;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
#define GC(KCV, NV) \
    if (KC == KCV && Cout == NV) return launch_gconv<KCV, NV>(ta, tb, p, st);
    GC(32, 32) GC(32, 64) GC(32, 96) GC(32, 128) GC(64, 32) GC(64, 64) GC(64, 96) GC(64, 128)
#undef GC
    HDMOE_CHECK_ARG(false, "gconv_fwd: unsupported (KC=%d, Cout=%d)", KC, Cout);
}
