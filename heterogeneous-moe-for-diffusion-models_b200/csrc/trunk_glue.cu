// HDMOEM.forward glue between the two MoE layers and output_proj (models/model_config2.py:276-302, the cfg1 soft
// query / context swap of models/model_config1.py:277-283) as fused per-pixel kernels on channels-last fp32 activations
// [P = B*H*W, C]: the reference runs ~25 elementwise / 1x1-convolution launches forward and ~50 backward over
// activation-sized tensors here (25 % of the device time of the round-1 train step).
//
//   swap   q = w*v + (1-w)*u,  ctx = w*u + (1-w)*v           w[b] = sigmoid(alpha_routing * (s_vit - s_unet))   (cfg1)
//   gate   fin  = a + alpha_txt * (b - a)                                                       (:291)
//          xcat = mp_cat(u, fin)  ->  h = W1 xcat  ->  mp_silu  ->  l = W2 (.)  ->  g = softmax_2(l)          (:293-297)
//          mix  = mp_sum(u, g0*u + g1*fin, 0.5)                                                 (:298-301)
//
// One thread owns one pixel's channel vector (C = 32: 128 contiguous bytes); W1 / W2 are broadcast from shared
// memory.  The backward kernel recomputes the forward per pixel, writes du / da / db, and accumulates the gate weight
// gradients in registers across all pixels a persistent CTA visits (per-thread 2 x 4 tile of dW1 fed from a shared-
// memory staging of the CTA's pixel batch), so they cost one vector atomic set per CTA.  HBM-bound: forward 520 B,
// backward 900 B per pixel.
#include "common.cuh"

namespace hdmoe {

constexpr int kGC = 32;            // internal channels (Utils/configs.py:8)
constexpr int kGT = 256;           // threads per CTA = pixels per batch

__device__ __forceinline__ float silu_mp(float x) { return x / (1.f + __expf(-x)) * (1.f / 0.596f); }
__device__ __forceinline__ float dsilu_mp(float x) {
    const float s = 1.f / (1.f + __expf(-x));
    return s * (1.f + x * (1.f - s)) * (1.f / 0.596f);
}

// ------------------------------------------------------------------------------------------------ swap
__global__ void __launch_bounds__(256)
swap_fwd_kernel(const float4* __restrict__ u, const float4* __restrict__ v, const float* __restrict__ w, float4* __restrict__ q,
                float4* __restrict__ ctx, long long n4, long long per_sample4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float wb = w[i / per_sample4], wc = 1.f - wb;
        const float4 a = u[i], b = v[i];
        q[i] = make_float4(wb * b.x + wc * a.x, wb * b.y + wc * a.y, wb * b.z + wc * a.z, wb * b.w + wc * a.w);
        ctx[i] = make_float4(wb * a.x + wc * b.x, wb * a.y + wc * b.y, wb * a.z + wc * b.z, wb * a.w + wc * b.w);
    }
}

// one CTA per (sample, slice): du, dv and the partial of dw[b] = sum (v - u) * (dq - dctx)
__global__ void __launch_bounds__(256)
swap_bwd_kernel(const float4* __restrict__ u, const float4* __restrict__ v, const float* __restrict__ w,
                const float4* __restrict__ dq, const float4* __restrict__ dctx, float4* __restrict__ du, float4* __restrict__ dv,
                float* __restrict__ dw_part, long long per_sample4, int slices) {
    __shared__ float red[8];
    const int b = blockIdx.x / slices, sl = blockIdx.x - b * slices;
    const float wb = w[b], wc = 1.f - wb;
    const long long lo = (long long)b * per_sample4 + per_sample4 * sl / slices;
    const long long hi = (long long)b * per_sample4 + per_sample4 * (sl + 1) / slices;
    float acc = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        const float4 a = u[i], c = v[i], gq = dq[i], gc = dctx[i];
        du[i] = make_float4(wc * gq.x + wb * gc.x, wc * gq.y + wb * gc.y, wc * gq.z + wb * gc.z, wc * gq.w + wb * gc.w);
        dv[i] = make_float4(wb * gq.x + wc * gc.x, wb * gq.y + wc * gc.y, wb * gq.z + wc * gc.z, wb * gq.w + wc * gc.w);
        acc += (c.x - a.x) * (gq.x - gc.x) + (c.y - a.y) * (gq.y - gc.y) + (c.z - a.z) * (gq.z - gc.z) +
               (c.w - a.w) * (gq.w - gc.w);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        dw_part[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------ gate
struct GateConsts {
    float c1, c2;          // mp_cat weights of u and fin
    float ms_a, ms_b;      // mp_sum(u, gated, t): (1-t)/norm, t/norm
};

__device__ __forceinline__ void load_vec32(const float* __restrict__ p, float (&x)[kGC]) {
#pragma unroll
    for (int i = 0; i < kGC / 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store_vec32(float* __restrict__ p, const float (&x)[kGC]) {
#pragma unroll
    for (int i = 0; i < kGC / 4; ++i)
        reinterpret_cast<float4*>(p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

// W1 in shared memory as [i][j] (input-major) so that a thread's loop over j for fixed i reads consecutive words that
// every thread of the warp reads identically (broadcast)
__global__ void __launch_bounds__(kGT)
gate_fwd_kernel(const float* __restrict__ u, const float* __restrict__ a, const float* __restrict__ b,
                const float* __restrict__ alpha_p, const float* __restrict__ W1, const float* __restrict__ W2,
                float* __restrict__ mix, float* __restrict__ g_out, long long P, long long HW, GateConsts k) {
    extern __shared__ float sm[];
    float* w1s = sm;                               // [i (64)][j (32)]
    float* w2s = w1s + 2 * kGC * kGC;              // [2][32]
    float* s_xc = w2s + 2 * kGC;                   // [kGT][65]: this thread's xcat column (runtime-indexed in the matvec)
    for (int t = threadIdx.x; t < 2 * kGC * kGC; t += kGT) {
        const int j = t / (2 * kGC), i = t - j * 2 * kGC;      // W1 is [j][i] row-major
        w1s[i * kGC + j] = W1[t];
    }
    if (threadIdx.x < 2 * kGC) w2s[threadIdx.x] = W2[threadIdx.x];
    __syncthreads();
    const float alpha = *alpha_p;
    float* my_xc = s_xc + threadIdx.x * (2 * kGC + 1);
    for (long long p = (long long)blockIdx.x * kGT + threadIdx.x; p < P; p += (long long)gridDim.x * kGT) {
        float x[kGC], f[kGC], h[kGC];
        load_vec32(u + p * kGC, x);
        {
            float ta[kGC];
            load_vec32(a + p * kGC, ta);
            load_vec32(b + p * kGC, f);
#pragma unroll
            for (int c = 0; c < kGC; ++c) f[c] = ta[c] + alpha * (f[c] - ta[c]);
        }
#pragma unroll
        for (int c = 0; c < kGC; ++c) {
            my_xc[c] = k.c1 * x[c];
            my_xc[kGC + c] = k.c2 * f[c];
        }
#pragma unroll
        for (int j = 0; j < kGC; ++j) h[j] = 0.f;
#pragma unroll 2
        for (int i = 0; i < 2 * kGC; ++i) {
            const float xi = my_xc[i];
            const float4* wr = reinterpret_cast<const float4*>(w1s + i * kGC);
#pragma unroll
            for (int j4 = 0; j4 < kGC / 4; ++j4) {
                const float4 w = wr[j4];
                h[4 * j4] = fmaf(w.x, xi, h[4 * j4]);
                h[4 * j4 + 1] = fmaf(w.y, xi, h[4 * j4 + 1]);
                h[4 * j4 + 2] = fmaf(w.z, xi, h[4 * j4 + 2]);
                h[4 * j4 + 3] = fmaf(w.w, xi, h[4 * j4 + 3]);
            }
        }
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < kGC; ++j) {
            const float s = silu_mp(h[j]);
            l0 = fmaf(w2s[j], s, l0);
            l1 = fmaf(w2s[kGC + j], s, l1);
        }
        const float m = fmaxf(l0, l1), e0 = __expf(l0 - m), e1 = __expf(l1 - m), inv = 1.f / (e0 + e1);
        const float g0 = e0 * inv, g1 = e1 * inv;
#pragma unroll
        for (int c = 0; c < kGC; ++c) x[c] = k.ms_a * x[c] + k.ms_b * (g0 * x[c] + g1 * f[c]);
        store_vec32(mix + p * kGC, x);
        // out_gate in the reference's [B, 2, H, W] layout
        const long long bi = p / HW, px = p - bi * HW;
        g_out[(bi * 2) * HW + px] = g0;
        g_out[(bi * 2 + 1) * HW + px] = g1;
    }
}

// Backward.  dW1 [32][64] and dW2 [2][32] are accumulated per CTA: every batch of kGT pixels is staged in shared memory
// (d_h [kGT][32], xcat [kGT][64], dl [kGT][2], hs [kGT][32]) and thread t adds its 2 x 4 tile of dW1 (j in {2*(t/16),
// +1}, i in {4*(t%16) .. +3}) over the batch; dW2 is owned by threads 0..63.  d_alpha is reduced per CTA.
__global__ void __launch_bounds__(kGT)
gate_bwd_kernel(const float* __restrict__ u, const float* __restrict__ a, const float* __restrict__ b,
                const float* __restrict__ alpha_p, const float* __restrict__ W1, const float* __restrict__ W2,
                const float* __restrict__ d_mix, const float* __restrict__ d_g, float* __restrict__ du, float* __restrict__ da,
                float* __restrict__ db, float* __restrict__ dW1, float* __restrict__ dW2, float* __restrict__ d_alpha,
                long long P, long long HW, GateConsts k) {
    extern __shared__ float sm[];
    float* w1s = sm;                               // [64][32]  (input-major, as forward)
    float* w1t = w1s + 2 * kGC * kGC;              // [32][64]  (output-major: d_xcat = W1^T d_h)
    float* w2s = w1t + 2 * kGC * kGC;              // [2][32]
    float* s_dh = w2s + 2 * kGC;                   // [kGT][33]  (padded rows: column reads without bank conflicts)
    float* s_xc = s_dh + kGT * (kGC + 1);          // [kGT][65]
    float* s_dl = s_xc + kGT * (2 * kGC + 1);      // [kGT][2]
    float* s_hs = s_dl + kGT * 2;                  // [kGT][33]
    __shared__ float red[8];
    for (int t = threadIdx.x; t < 2 * kGC * kGC; t += kGT) {
        const int j = t / (2 * kGC), i = t - j * 2 * kGC;
        w1s[i * kGC + j] = W1[t];
        w1t[t] = W1[t];
    }
    if (threadIdx.x < 2 * kGC) w2s[threadIdx.x] = W2[threadIdx.x];
    __syncthreads();
    const float alpha = *alpha_p;
    const int tj = 2 * (threadIdx.x >> 4), ti = 4 * (threadIdx.x & 15);
    float acc1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float acc2 = 0.f, acc_alpha = 0.f;
    const long long nbatch = (P + kGT - 1) / kGT;
    for (long long bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
        const long long p = bt * kGT + threadIdx.x;
        const bool live = p < P;
        float dh[kGC];
        float dl0 = 0.f, dl1 = 0.f;
        if (live) {
            float x[kGC], f[kGC], h[kGC], diff[kGC];          // diff = b - a (d_alpha needs it; a, b themselves do not stay live)
            load_vec32(u + p * kGC, x);
            load_vec32(a + p * kGC, f);
            load_vec32(b + p * kGC, diff);
#pragma unroll
            for (int c = 0; c < kGC; ++c) {
                diff[c] -= f[c];
                f[c] = fmaf(alpha, diff[c], f[c]);
            }
            float* my_xc = s_xc + threadIdx.x * (2 * kGC + 1);
#pragma unroll
            for (int c = 0; c < kGC; ++c) {
                my_xc[c] = k.c1 * x[c];
                my_xc[kGC + c] = k.c2 * f[c];
            }
#pragma unroll
            for (int j = 0; j < kGC; ++j) h[j] = 0.f;
#pragma unroll 2
            for (int i = 0; i < 2 * kGC; ++i) {
                const float xi = my_xc[i];
                const float4* wr = reinterpret_cast<const float4*>(w1s + i * kGC);
#pragma unroll
                for (int j4 = 0; j4 < kGC / 4; ++j4) {
                    const float4 w = wr[j4];
                    h[4 * j4] = fmaf(w.x, xi, h[4 * j4]);
                    h[4 * j4 + 1] = fmaf(w.y, xi, h[4 * j4 + 1]);
                    h[4 * j4 + 2] = fmaf(w.z, xi, h[4 * j4 + 2]);
                    h[4 * j4 + 3] = fmaf(w.w, xi, h[4 * j4 + 3]);
                }
            }
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int j = 0; j < kGC; ++j) {
                const float s = silu_mp(h[j]);
                s_hs[threadIdx.x * (kGC + 1) + j] = s;
                l0 = fmaf(w2s[j], s, l0);
                l1 = fmaf(w2s[kGC + j], s, l1);
            }
            const float m = fmaxf(l0, l1), e0 = __expf(l0 - m), e1 = __expf(l1 - m), inv = 1.f / (e0 + e1);
            const float g0 = e0 * inv, g1 = e1 * inv;
            // mix = ms_a * u + ms_b * (g0 * u + g1 * f)
            float gm[kGC];
            load_vec32(d_mix + p * kGC, gm);
            float dg0 = 0.f, dg1 = 0.f;
#pragma unroll
            for (int c = 0; c < kGC; ++c) {
                const float dgt = k.ms_b * gm[c];          // d gated
                dg0 = fmaf(dgt, x[c], dg0);
                dg1 = fmaf(dgt, f[c], dg1);
            }
            if (d_g) {
                const long long bi = p / HW, px = p - bi * HW;
                dg0 += d_g[(bi * 2) * HW + px];
                dg1 += d_g[(bi * 2 + 1) * HW + px];
            }
            const float dot = g0 * dg0 + g1 * dg1;
            dl0 = g0 * (dg0 - dot);
            dl1 = g1 * (dg1 - dot);
#pragma unroll
            for (int j = 0; j < kGC; ++j) dh[j] = (w2s[j] * dl0 + w2s[kGC + j] * dl1) * dsilu_mp(h[j]);
            // d xcat = W1^T d_h ; du = ms_a*gm + ms_b*g0*gm + c1*dxc_u ; dfin = ms_b*g1*gm + c2*dxc_f
            float dfin[kGC];
#pragma unroll
            for (int i = 0; i < kGC; ++i) {
                float su = 0.f, sf = 0.f;
#pragma unroll
                for (int j = 0; j < kGC; ++j) {
                    su = fmaf(w1t[j * 2 * kGC + i], dh[j], su);
                    sf = fmaf(w1t[j * 2 * kGC + kGC + i], dh[j], sf);
                }
                x[i] = (k.ms_a + k.ms_b * g0) * gm[i] + k.c1 * su;          // x now holds du
                dfin[i] = k.ms_b * g1 * gm[i] + k.c2 * sf;
                acc_alpha = fmaf(dfin[i], diff[i], acc_alpha);
            }
            store_vec32(du + p * kGC, x);
#pragma unroll
            for (int c = 0; c < kGC; ++c) diff[c] = (1.f - alpha) * dfin[c];
            store_vec32(da + p * kGC, diff);
#pragma unroll
            for (int c = 0; c < kGC; ++c) dfin[c] *= alpha;
            store_vec32(db + p * kGC, dfin);
        } else {
#pragma unroll
            for (int j = 0; j < kGC; ++j) dh[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * kGC; ++i) s_xc[threadIdx.x * (2 * kGC + 1) + i] = 0.f;
#pragma unroll
            for (int j = 0; j < kGC; ++j) s_hs[threadIdx.x * (kGC + 1) + j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < kGC; ++j) s_dh[threadIdx.x * (kGC + 1) + j] = dh[j];
        s_dl[threadIdx.x * 2] = dl0;
        s_dl[threadIdx.x * 2 + 1] = dl1;
        __syncthreads();
        // weight-gradient tiles over the batch
#pragma unroll 4
        for (int q = 0; q < kGT; ++q) {
            const float d0 = s_dh[q * (kGC + 1) + tj], d1 = s_dh[q * (kGC + 1) + tj + 1];
            const float* xr = s_xc + q * (2 * kGC + 1) + ti;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc1[0][c] = fmaf(d0, xr[c], acc1[0][c]);
                acc1[1][c] = fmaf(d1, xr[c], acc1[1][c]);
            }
        }
        if (threadIdx.x < 2 * kGC) {
            const int kk = threadIdx.x >> 5, j = threadIdx.x & 31;
            for (int q = 0; q < kGT; ++q) acc2 = fmaf(s_dl[q * 2 + kk], s_hs[q * (kGC + 1) + j], acc2);
        }
        __syncthreads();
    }
    // flush the CTA's partials
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(dW1 + (tj + r) * 2 * kGC + ti + c, acc1[r][c]);
    if (threadIdx.x < 2 * kGC) atomicAdd(dW2 + threadIdx.x, acc2);
    acc_alpha = warp_sum(acc_alpha);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc_alpha;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kGT / 32; ++i) t += red[i];
        atomicAdd(d_alpha, t);
    }
}

static constexpr int kGateFwdSmem = (2 * kGC * kGC + 2 * kGC + kGT * (2 * kGC + 1)) * (int)sizeof(float);
static constexpr int kGateBwdSmem =
    (2 * (2 * kGC * kGC) + 2 * kGC + kGT * (kGC + 1) + kGT * (2 * kGC + 1) + kGT * 2 + kGT * (kGC + 1)) * (int)sizeof(float);

// ------------------------------------------------------------------------------------------------ branch scaling
// models/model_config2.py:244-251 (analytic sigma-sigmoid path scaling) and models/model_config1.py:246-252 (learned
// Scaling_router gains): in_vit = s_vit[b] * feats, in_unet = s_unet[b] * feats.  ONE pass over the features writes
// both products (fp32 NCHW, the dispatch payload source) AND the channels-last bf16 copy [2B, HW, C] the tcgen05 router
// trunk reads (rows [0, B) = ViT router input, [B, 2B) = U-Net router input) instead of 2 broadcast multiplies + 2 casts
// + a concatenation + a layout kernel.  Tile = 32 pixels x 32 channels, transposed through shared memory.
__global__ void __launch_bounds__(256)
analytic_scaling_kernel(const float* __restrict__ time_vec, float tp, float soft, float* __restrict__ scaling, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float w = 1.f / (1.f + expf(-(time_vec[b] * 4.f - tp) / soft));
    scaling[2 * b] = (w + 1e-2f) * 2.f;                  // ViT path
    scaling[2 * b + 1] = ((1.f - w) + 1e-2f) * 2.f;      // U-Net path
}

__global__ void __launch_bounds__(256)
scale_pair_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ scaling, float* __restrict__ in_v,
                      float* __restrict__ in_u, __nv_bfloat16* __restrict__ trunk, int B, int HW) {
    __shared__ float tile[kGC][33];
    const int b = blockIdx.y, p0 = blockIdx.x * 32;
    const float sv = scaling[2 * b], su = scaling[2 * b + 1];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;          // 8 warps: warp wq handles channels wq, wq+8, ...
    const size_t base = (size_t)b * kGC * HW;
#pragma unroll
    for (int cc = 0; cc < kGC / 8; ++cc) {
        const int c = wq + 8 * cc, p = p0 + lane;
        float f = 0.f;
        if (p < HW) {
            f = feats[base + (size_t)c * HW + p];
            in_v[base + (size_t)c * HW + p] = sv * f;
            in_u[base + (size_t)c * HW + p] = su * f;
        }
        tile[c][lane] = f;
    }
    if (!trunk) return;
    __syncthreads();
    // channels-last: thread (pixel = threadIdx.x / 8, channel quad = threadIdx.x % 8) writes 4 bf16 (8 bytes)
    const int px = threadIdx.x >> 3, cq = (threadIdx.x & 7) * 4, p = p0 + px;
    if (p < HW) {
        const float f0 = tile[cq][px], f1 = tile[cq + 1][px], f2 = tile[cq + 2][px], f3 = tile[cq + 3][px];
        Vec4<__nv_bfloat16>::store(trunk + ((size_t)b * HW + p) * kGC + cq, make_float4(sv * f0, sv * f1, sv * f2, sv * f3));
        Vec4<__nv_bfloat16>::store(trunk + ((size_t)(B + b) * HW + p) * kGC + cq, make_float4(su * f0, su * f1, su * f2, su * f3));
    }
}

// d_feats = s_v * (g_v + g_trunk_v) + s_u * (g_u + g_trunk_u);  ds_part[b][tile][2] = sum over the tile of feats * g
__global__ void __launch_bounds__(256)
scale_pair_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ scaling, const float* __restrict__ g_v,
                      const float* __restrict__ g_u, const __nv_bfloat16* __restrict__ g_trunk, float* __restrict__ d_feats,
                      float* __restrict__ ds_part, int B, int HW) {
    __shared__ float tv[kGC][33], tu[kGC][33];
    __shared__ float red[2][8];
    const int b = blockIdx.y, p0 = blockIdx.x * 32;
    const float sv = scaling[2 * b], su = scaling[2 * b + 1];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const size_t base = (size_t)b * kGC * HW;
    if (g_trunk) {
        const int px = threadIdx.x >> 3, cq = (threadIdx.x & 7) * 4, p = p0 + px;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
        if (p < HW) {
            a = Vec4<__nv_bfloat16>::load(g_trunk + ((size_t)b * HW + p) * kGC + cq);
            c = Vec4<__nv_bfloat16>::load(g_trunk + ((size_t)(B + b) * HW + p) * kGC + cq);
        }
        tv[cq][px] = a.x; tv[cq + 1][px] = a.y; tv[cq + 2][px] = a.z; tv[cq + 3][px] = a.w;
        tu[cq][px] = c.x; tu[cq + 1][px] = c.y; tu[cq + 2][px] = c.z; tu[cq + 3][px] = c.w;
        __syncthreads();
    }
    float av = 0.f, au = 0.f;
#pragma unroll
    for (int cc = 0; cc < kGC / 8; ++cc) {
        const int c = wq + 8 * cc, p = p0 + lane;
        if (p < HW) {
            const size_t i = base + (size_t)c * HW + p;
            float gv = g_v ? g_v[i] : 0.f, gu = g_u ? g_u[i] : 0.f;
            if (g_trunk) {
                gv += tv[c][lane];
                gu += tu[c][lane];
            }
            const float f = feats[i];
            d_feats[i] = sv * gv + su * gu;
            av = fmaf(f, gv, av);
            au = fmaf(f, gu, au);
        }
    }
    av = warp_sum(av);
    au = warp_sum(au);
    if (lane == 0) {
        red[0][wq] = av;
        red[1][wq] = au;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
        ds_part[((size_t)b * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = t;
    }
}

}  // namespace hdmoe
using namespace hdmoe;

extern "C" int hdmoe_trunk_swap_fwd(const float* u, const float* v, const float* w, float* q, float* ctx, int B,
                                    int64_t per_sample, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(u && v && w && q && ctx && B >= 1 && per_sample >= 4 && per_sample % 4 == 0, "trunk_swap_fwd: bad args");
    const long long n4 = (long long)B * per_sample / 4;
    swap_fwd_kernel<<<grid_for(n4, 256, 16), 256, 0, (cudaStream_t)stream>>>((const float4*)u, (const float4*)v, w, (float4*)q,
                                                                             (float4*)ctx, n4, per_sample / 4);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_trunk_swap_slices(void) { return 8; }

extern "C" int hdmoe_trunk_swap_bwd(const float* u, const float* v, const float* w, const float* dq, const float* dctx,
                                    float* du, float* dv, float* dw_part, int B, int64_t per_sample, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(u && v && w && dq && dctx && du && dv && dw_part && B >= 1 && per_sample >= 4 && per_sample % 4 == 0,
                    "trunk_swap_bwd: bad args");
    const int slices = hdmoe_trunk_swap_slices();
    swap_bwd_kernel<<<B * slices, 256, 0, (cudaStream_t)stream>>>((const float4*)u, (const float4*)v, w, (const float4*)dq,
                                                                  (const float4*)dctx, (float4*)du, (float4*)dv, dw_part,
                                                                  per_sample / 4, slices);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_trunk_gate_fwd(const float* u, const float* a, const float* b, const float* alpha_txt, const float* W1,
                                    const float* W2, float* mix, float* g_out, int64_t P, int64_t HW, int C, float c1,
                                    float c2, float ms_a, float ms_b, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(u && a && b && alpha_txt && W1 && W2 && mix && g_out && P >= 1 && HW >= 1 && P % HW == 0, "trunk_gate_fwd: bad args");
    HDMOE_CHECK_ARG(C == kGC, "trunk_gate_fwd: internal_channels must be %d (got %d)", kGC, C);
    GateConsts k{c1, c2, ms_a, ms_b};
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(gate_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGateFwdSmem));
    gate_fwd_kernel<<<grid_for(P, kGT, 3), kGT, kGateFwdSmem, (cudaStream_t)stream>>>(u, a, b, alpha_txt, W1, W2, mix, g_out, P, HW, k);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_trunk_gate_bwd(const float* u, const float* a, const float* b, const float* alpha_txt, const float* W1,
                                    const float* W2, const float* d_mix, const float* d_g, float* du, float* da, float* db,
                                    float* dW1, float* dW2, float* d_alpha, int64_t P, int64_t HW, int C, float c1, float c2,
                                    float ms_a, float ms_b, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(u && a && b && alpha_txt && W1 && W2 && d_mix && du && da && db && dW1 && dW2 && d_alpha && P >= 1 &&
                        HW >= 1 && P % HW == 0, "trunk_gate_bwd: bad args");
    HDMOE_CHECK_ARG(C == kGC, "trunk_gate_bwd: internal_channels must be %d (got %d)", kGC, C);
    GateConsts k{c1, c2, ms_a, ms_b};
    HDMOE_CHECK_CUDA(cudaFuncSetAttribute(gate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGateBwdSmem));
    const long long nbatch = (P + kGT - 1) / kGT;
    const int grid = (int)(nbatch < kNumSMs ? nbatch : kNumSMs);      // persistent: dW partials stay in registers
    gate_bwd_kernel<<<grid, kGT, kGateBwdSmem, (cudaStream_t)stream>>>(u, a, b, alpha_txt, W1, W2, d_mix, d_g, du, da, db, dW1, dW2,
                                                                      d_alpha, P, HW, k);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_analytic_scaling(const float* time_vec, float transition_point, float softness, float* scaling, int B,
                                      hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(time_vec && scaling && B >= 1 && softness != 0.f, "analytic_scaling: bad args");
    analytic_scaling_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(time_vec, transition_point, softness, scaling, B);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_scale_pair_fwd(const float* feats, const float* scaling, float* in_vit, float* in_unet, void* trunk_bf16,
                                    int B, int C, int64_t HW, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(feats && scaling && in_vit && in_unet && B >= 1 && HW >= 1 && B <= 65535, "scale_pair_fwd: bad args");
    HDMOE_CHECK_ARG(C == kGC, "scale_pair_fwd: internal_channels must be %d (got %d)", kGC, C);
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)B);
    scale_pair_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feats, scaling, in_vit, in_unet, (__nv_bfloat16*)trunk_bf16, B,
                                                                  (int)HW);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_scale_pair_tiles(int64_t HW) { return (int)((HW + 31) / 32); }

extern "C" int hdmoe_scale_pair_bwd(const float* feats, const float* scaling, const float* g_vit, const float* g_unet,
                                    const void* g_trunk_bf16, float* d_feats, float* ds_part, int B, int C, int64_t HW,
                                    hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(feats && scaling && d_feats && ds_part && B >= 1 && HW >= 1 && B <= 65535, "scale_pair_bwd: bad args");
    HDMOE_CHECK_ARG(C == kGC, "scale_pair_bwd: internal_channels must be %d (got %d)", kGC, C);
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)B);
    scale_pair_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feats, scaling, g_vit, g_unet,
                                                                  (const __nv_bfloat16*)g_trunk_bf16, d_feats, ds_part, B, (int)HW);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

// ------------------------------------------------------------------------------------------------ Scaling_router
// models/model_components.py:41-66 (cfg1): x [B, D] -> W1 (D -> 2D) -> GroupNorm(1) -> ReLU -> W2 (2D -> 4D) ->
// GroupNorm(1) -> ReLU -> Dropout -> W3 (4D -> 2) -> + zeta * noise -> softmax * 2.  ~25 small launches forward and ~50
// backward in the op-by-op version; here one CTA per sample does the whole chain (D = 64: 41 k weights, L2 resident),
// the backward recomputes it and accumulates the weight / affine gradients with atomics.
namespace hdmoe {
constexpr int kSrD = 64, kSrH1 = 128, kSrH2 = 256, kSrT = 256;

__device__ __forceinline__ float sr_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kSrT / 32; ++i) t += red[i];
    return t;
}

struct SrShared {
    float x[kSrD], h1[kSrH1], y1[kSrH1], h2[kSrH2], y2[kSrH2], red[8];
    float mu1, rs1, mu2, rs2, l[2];
};

// forward of one sample into shared memory; returns the two logits in s.l (before noise)
__device__ __forceinline__ void sr_forward(SrShared& s, const float* __restrict__ xb, const float* __restrict__ W1,
                                           const float* __restrict__ g1, const float* __restrict__ b1,
                                           const float* __restrict__ W2, const float* __restrict__ g2,
                                           const float* __restrict__ b2, const float* __restrict__ W3,
                                           const float* __restrict__ keep, float eps) {
    const int t = threadIdx.x;
    if (t < kSrD) s.x[t] = xb[t];
    __syncthreads();
    float h = 0.f;
    if (t < kSrH1) {
        const float4* w = reinterpret_cast<const float4*>(W1 + (size_t)t * kSrD);
#pragma unroll 4
        for (int k = 0; k < kSrD / 4; ++k) {
            const float4 q = w[k];
            h += q.x * s.x[4 * k] + q.y * s.x[4 * k + 1] + q.z * s.x[4 * k + 2] + q.w * s.x[4 * k + 3];
        }
        s.h1[t] = h;
    }
    float m = sr_block_sum(t < kSrH1 ? h : 0.f, s.red) / kSrH1;
    float d = t < kSrH1 ? h - m : 0.f;
    float var = sr_block_sum(d * d, s.red) / kSrH1;
    if (t == 0) {
        s.mu1 = m;
        s.rs1 = rsqrtf(var + eps);
    }
    __syncthreads();
    if (t < kSrH1) s.y1[t] = fmaxf((h - s.mu1) * s.rs1 * g1[t] + b1[t], 0.f);
    __syncthreads();
    {
        const float4* w = reinterpret_cast<const float4*>(W2 + (size_t)t * kSrH1);
        h = 0.f;
#pragma unroll 4
        for (int k = 0; k < kSrH1 / 4; ++k) {
            const float4 q = w[k];
            h += q.x * s.y1[4 * k] + q.y * s.y1[4 * k + 1] + q.z * s.y1[4 * k + 2] + q.w * s.y1[4 * k + 3];
        }
        s.h2[t] = h;
    }
    m = sr_block_sum(h, s.red) / kSrH2;
    d = h - m;
    var = sr_block_sum(d * d, s.red) / kSrH2;
    if (t == 0) {
        s.mu2 = m;
        s.rs2 = rsqrtf(var + eps);
    }
    __syncthreads();
    float y = fmaxf((h - s.mu2) * s.rs2 * g2[t] + b2[t], 0.f);
    if (keep) y *= keep[t];
    s.y2[t] = y;
    const float l0 = sr_block_sum(W3[t] * y, s.red), l1 = sr_block_sum(W3[kSrH2 + t] * y, s.red);
    if (t == 0) {
        s.l[0] = l0;
        s.l[1] = l1;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSrT)
scaling_router_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ g1,
                          const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ g2,
                          const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ noise,
                          float zeta, const float* __restrict__ keep, float eps, float* __restrict__ out) {
    __shared__ SrShared s;
    const int b = blockIdx.x;
    sr_forward(s, x + (size_t)b * kSrD, W1, g1, b1, W2, g2, b2, W3, keep ? keep + (size_t)b * kSrH2 : nullptr, eps);
    if (threadIdx.x == 0) {
        float l0 = s.l[0], l1 = s.l[1];
        if (noise) {
            l0 += noise[2 * b] * zeta;
            l1 += noise[2 * b + 1] * zeta;
        }
        const float m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m), inv = 2.f / (e0 + e1);
        out[2 * b] = e0 * inv;
        out[2 * b + 1] = e1 * inv;
    }
}

// GroupNorm(1) backward of one layer held in shared memory: dy (gradient w.r.t. the post-ReLU output, dropout already
// applied) -> dh in place; accumulates dgamma / dbeta
__device__ __forceinline__ float sr_gn_bwd(float dy, float h, float y_pos, float mu, float rs, float gamma, int n, bool live,
                                           float* __restrict__ dgamma, float* __restrict__ dbeta, float* red) {
    const float xh = (h - mu) * rs;
    const float dyn = (live && y_pos > 0.f) ? dy : 0.f;          // ReLU
    if (live) {
        atomicAdd(dgamma, dyn * xh);
        atomicAdd(dbeta, dyn);
    }
    const float g = dyn * gamma;
    const float m1 = sr_block_sum(g, red) / n, m2 = sr_block_sum(g * xh, red) / n;
    return live ? rs * (g - m1 - xh * m2) : 0.f;
}

__global__ void __launch_bounds__(kSrT)
scaling_router_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ g1,
                          const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ g2,
                          const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ noise,
                          float zeta, const float* __restrict__ keep, float eps, const float* __restrict__ d_out,
                          float* __restrict__ dx, float* __restrict__ dW1, float* __restrict__ dg1, float* __restrict__ db1,
                          float* __restrict__ dW2, float* __restrict__ dg2, float* __restrict__ db2, float* __restrict__ dW3) {
    __shared__ SrShared s;
    __shared__ float dh2[kSrH2], dh1[kSrH1];
    const int b = blockIdx.x, t = threadIdx.x;
    const float* kp = keep ? keep + (size_t)b * kSrH2 : nullptr;
    sr_forward(s, x + (size_t)b * kSrD, W1, g1, b1, W2, g2, b2, W3, kp, eps);
    float l0 = s.l[0], l1 = s.l[1];
    if (noise) {
        l0 += noise[2 * b] * zeta;
        l1 += noise[2 * b + 1] * zeta;
    }
    const float m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m), inv = 1.f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    const float go0 = d_out[2 * b], go1 = d_out[2 * b + 1];
    const float dot = p0 * go0 + p1 * go1;
    const float dl0 = 2.f * p0 * (go0 - dot), dl1 = 2.f * p1 * (go1 - dot);     // out = 2 * softmax(l)
    // W3 and the input of W3 (y2 includes the dropout scaling)
    atomicAdd(dW3 + t, dl0 * s.y2[t]);
    atomicAdd(dW3 + kSrH2 + t, dl1 * s.y2[t]);
    float dy2 = W3[t] * dl0 + W3[kSrH2 + t] * dl1;
    if (kp) dy2 *= kp[t];
    // undo the dropout factor for the ReLU test: y2 > 0 iff the pre-dropout activation > 0 (or the unit was dropped: dy2 = 0 then)
    const float pre2 = (s.h2[t] - s.mu2) * s.rs2 * g2[t] + b2[t];
    dh2[t] = sr_gn_bwd(dy2, s.h2[t], pre2, s.mu2, s.rs2, g2[t], kSrH2, true, dg2 + t, db2 + t, s.red);
    __syncthreads();
    // dW2[j][k] += dh2[j] * y1[k]: thread t owns column block: loop over rows j, coalesced over k
    for (int j = 0; j < kSrH2; ++j) {
        const float dj = dh2[j];
        if (t < kSrH1 && dj != 0.f) atomicAdd(dW2 + (size_t)j * kSrH1 + t, dj * s.y1[t]);
    }
    float dy1 = 0.f;
    if (t < kSrH1)
        for (int j = 0; j < kSrH2; ++j) dy1 = fmaf(W2[(size_t)j * kSrH1 + t], dh2[j], dy1);
    const bool live1 = t < kSrH1;
    const float pre1 = live1 ? (s.h1[t] - s.mu1) * s.rs1 * g1[t] + b1[t] : 0.f;
    const float d1 = sr_gn_bwd(dy1, live1 ? s.h1[t] : 0.f, pre1, s.mu1, s.rs1, live1 ? g1[t] : 0.f, kSrH1, live1,
                               dg1 + (live1 ? t : 0), db1 + (live1 ? t : 0), s.red);
    if (live1) dh1[t] = d1;
    __syncthreads();
    for (int j = 0; j < kSrH1; ++j) {
        const float dj = dh1[j];
        if (t < kSrD && dj != 0.f) atomicAdd(dW1 + (size_t)j * kSrD + t, dj * s.x[t]);
    }
    if (t < kSrD) {
        float a = 0.f;
        for (int j = 0; j < kSrH1; ++j) a = fmaf(W1[(size_t)j * kSrD + t], dh1[j], a);
        dx[(size_t)b * kSrD + t] = a;
    }
}
}  // namespace hdmoe

extern "C" int hdmoe_scaling_router_fwd(const float* x, const float* W1, const float* g1, const float* b1, const float* W2,
                                        const float* g2, const float* b2, const float* W3, const float* noise, float zeta,
                                        const float* keep, float eps, float* out, int B, int D, hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x && W1 && g1 && b1 && W2 && g2 && b2 && W3 && out && B >= 1, "scaling_router_fwd: null pointer");
    HDMOE_CHECK_ARG(D == hdmoe::kSrD, "scaling_router_fwd: emb_dim must be %d (got %d)", hdmoe::kSrD, D);
    hdmoe::scaling_router_fwd_kernel<<<B, hdmoe::kSrT, 0, (cudaStream_t)stream>>>(x, W1, g1, b1, W2, g2, b2, W3, noise, zeta, keep,
                                                                                 eps, out);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}

extern "C" int hdmoe_scaling_router_bwd(const float* x, const float* W1, const float* g1, const float* b1, const float* W2,
                                        const float* g2, const float* b2, const float* W3, const float* noise, float zeta,
                                        const float* keep, float eps, const float* d_out, float* dx, float* dW1, float* dg1,
                                        float* db1, float* dW2, float* dg2, float* db2, float* dW3, int B, int D,
                                        hdmoe_stream_t stream) {
    HDMOE_CHECK_ARG(x && W1 && g1 && b1 && W2 && g2 && b2 && W3 && d_out && dx && dW1 && dg1 && db1 && dW2 && dg2 && db2 && dW3 &&
                        B >= 1, "scaling_router_bwd: null pointer");
    HDMOE_CHECK_ARG(D == hdmoe::kSrD, "scaling_router_bwd: emb_dim must be %d (got %d)", hdmoe::kSrD, D);
    hdmoe::scaling_router_bwd_kernel<<<B, hdmoe::kSrT, 0, (cudaStream_t)stream>>>(x, W1, g1, b1, W2, g2, b2, W3, noise, zeta, keep,
                                                                                 eps, d_out, dx, dW1, dg1, db1, dW2, dg2, db2, dW3);
    HDMOE_CHECK_LAUNCH();
    return HDMOE_OK;
}
