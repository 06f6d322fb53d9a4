/*
 * hdmoe_gemm.h -- tcgen05 / TMEM / TMA entry points of libhdmoe_b200.so (part of the C ABI; see
 * hdmoe_b200.h for the conventions).
 */
#ifndef HDMOE_GEMM_H_
#define HDMOE_GEMM_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Grouped implicit-GEMM convolution, forward (and, with transposed + flipped weights, the data gradient).
 * Replaces F.conv2d / F.linear inside MP_Conv (models/model_internals.py:261-271) for every expert of one
 * layer in ONE persistent launch.
 *   X   bf16 NHWC [cap_rows, H, W, Cin_pad]   rows in the dispatch plan's expert-major order
 *   Wt  bf16 [w_rows_total, Cin_pad]          expert e occupies rows wrow[e] .. wrow[e] + k_e^2*Cout, laid out
 *                                             [tap][Cout][Cin_pad] (hdmoe_wprep_fwd, HDMOE_WLAYOUT_TAPS)
 *   Y   bf16 NHWC [cap_rows, H, W, Cout]      'same' zero padding, stride 1, odd k_e
 *   row_expert[cap_rows] (device, -1 = unused row), *n_rows_dev = number of live rows (device),
 *   ksize_host / wrow_host: HOST arrays of length n_experts.
 * Fused epilogue:  v = acc * scale[row, c] (scale may be NULL);  v = mp_silu(v) if act == 1;
 *                  out = res_a * residual + res_b * v  if residual != NULL  (mp_sum folded).
 * Constraints: Cout in {32, 64, 96, 128}; Cin_pad % 32 == 0. */
/* Halo-reuse implementation of the contract above: the zero-padded input window of a tile (a run of up to
 * three 128-position M-tiles of the flattened padded image) is loaded ONCE and every filter tap reads it through a
 * shifted UMMA descriptor, removing the k^2-fold L2 re-reads of the per-tap loader.  Any H <= 255, W <= 248 (no
 * 128-pixel divisibility requirement); at most 4 distinct kernel sizes per launch; Y and residual 32-byte aligned.
 * Preconditions: row_expert[r] in [0, n_experts) for r < *n_rows_dev.  Tiles are handed out by a device-side counter
 * that belongs to (device, stream): launches that may run concurrently must be issued on different streams. */
int hdmoe_gconv2_fwd(const void* X, const void* Wt, void* Y, int cap_rows, int H, int W, int Cin_pad, int Cout,
                     int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev, int n_experts,
                     const int32_t* ksize_host, const int32_t* wrow_host, const float* scale, int act,
                     const void* residual, float res_a, float res_b, void* stream);

/* Grouped convolution weight gradient (tcgen05, MN-major operands, split-K with vector atomics):
 *   dW[wrow[e] + tap*Cout + o, c] += sum over rows r of expert e and pixels q of dY[r,q,o] * Xpad[r, q+delta_tap, c]
 * dW is fp32 [w_rows_total, Cin_pad] in the tap-major block layout of the forward operand and must be zeroed by
 * the caller before the first accumulation of a step.  X / dY are NHWC bf16 as in hdmoe_gconv2_fwd.
 * Constraints: Cout in {32, 64, 128} (128 runs as two 64-channel passes); Cin_pad % 32 == 0, <= 256; H % 4 == 0 (the largest of 32 / 16 / 8 / 4 strip rows that
 * divides H and fits shared memory is used); W <= 232.  Same stream rule as hdmoe_gconv2_fwd. */
int hdmoe_gconv_wgrad(const void* X, const void* dY, float* dW, int cap_rows, int H, int W, int Cin_pad, int Cout,
                      int64_t w_rows_total, const int32_t* row_expert, const int32_t* n_rows_dev, int n_experts,
                      const int32_t* ksize_host, const int32_t* wrow_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif
