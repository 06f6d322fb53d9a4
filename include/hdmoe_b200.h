/*
 * hdmoe_b200.h -- C ABI of libhdmoe_b200.so (hand-written sm_100a kernels for the heterogeneous-MoE
 * hot path of the EDM denoiser).
 *
 * The reference (cs2mosa/Heterogeneous-MOE-for-Diffusion-models) is 100 % Python/PyTorch and has no
 * FFI of its own; the drop-in boundary is its nn.Module surface, which the Python
 * package next to this header mirrors.  This header is the boundary *below* those modules: each entry
 * point replaces the chain of ATen calls the cited reference lines launch.  All `file:line` citations
 * are relative to the reference repository root.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless it says "host";
 *   - no allocation inside: outputs and workspaces are caller-provided;
 *   - kernels are enqueued on `stream` (a cudaStream_t); nothing synchronises the host;
 *   - return 0 on success, a negative HDMOE_ERR_* otherwise; hdmoe_last_error() gives the text;
 *   - dtype codes: HDMOE_F32 = 0, HDMOE_BF16 = 1.
 */
#ifndef HDMOE_B200_H_
#define HDMOE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hdmoe_stream_t; /* cudaStream_t */

#define HDMOE_F32 0
#define HDMOE_BF16 1

#define HDMOE_OK 0
#define HDMOE_ERR_ARG (-1)      /* bad argument (shape / alignment / unsupported size)     */
#define HDMOE_ERR_CUDA (-2)     /* a CUDA runtime / driver call failed                      */
#define HDMOE_ERR_NO_DEVICE (-3)/* no sm_100 device present                                 */

#define HDMOE_MAX_EXPERTS 64
#define HDMOE_MAX_TOPK 8

int hdmoe_version(void);
const char* hdmoe_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches counter) */
int64_t hdmoe_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (1) Router gate -- replaces Router.forward's tail, models/model_components.py:148-168, and emits the
 *     load-balance / z-loss statistics of Utils/utils.py:158-172.
 *
 *     xmod   = pooled * (1 + gamma) + beta            (cond = [gamma | beta], NULL -> no modulation)
 *     logits = xmod . w_hat^T  (+ zeta * noise)       (w_hat already weight-normalised, [E, C])
 *     logits = -inf where mask == 0
 *     gate_probs = softmax(logits);  (vals, idx) = top_k(logits)  [ties: lowest index];
 *     topk_w = softmax(vals);  sparse_w = scatter(idx, topk_w)
 *     stats[0:E]   = sum_t gate_probs[t, e]           (load_balance = E * sum_e (stats[e]/T)^2)
 *     stats[E:2E]  = #tokens with sparse_w[t, e] > 0  (dispatch counts, as float)
 *     stats[2E]    = sum_t min(logsumexp(clamp(logits, -50, 50))^2, 100)     (z_loss * T)
 *     workspace: word 0 is a zero-initialised ticket counter (the kernel resets it), per-CTA partials follow;
 *     size from hdmoe_router_gate_workspace_bytes().
 *     `logits_in` != NULL skips the linear part and gates the given (already masked) logits: the
 *     teacher-forced entry used for bit-exact index parity.
 * ---------------------------------------------------------------------------------------------- */
size_t hdmoe_router_gate_workspace_bytes(int T, int E);
int hdmoe_router_gate_fwd(const float* pooled, const float* cond, const float* w_hat, const float* noise,
                          float zeta, const float* mask, const float* logits_in, int T, int C, int E, int top_k,
                          float* logits, float* gate_probs, float* sparse_w, int32_t* topk_idx, float* topk_w,
                          float* stats, void* workspace, hdmoe_stream_t stream);
/* Backward of the above.  Any of the g_* may be NULL.  d_w_hat [E,C] is overwritten (not accumulated).
 * g_stats has the layout of `stats`; only [0:E] and [2E] carry gradient. */
int hdmoe_router_gate_bwd(const float* pooled, const float* cond, const float* w_hat, const float* logits,
                          const int32_t* topk_idx, const float* g_sparse, const float* g_probs,
                          const float* g_logits, const float* g_stats, int T, int C, int E, int top_k,
                          float* d_pooled, float* d_cond, float* d_w_hat, float* d_logits_out,
                          hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Dispatch plan + permute -- replaces the boolean-mask gathers of router_to_unet_experts,
 *     models/model_config2.py:25-33 (== models/model_config1.py:25-33).
 *
 *     Row order is the reference's: expert-major, ascending token index inside an expert; an entry is
 *     dispatched iff sparse_w[t, e] > 0 (NaN is not).  Integer outputs are bit-exact.
 *       counts[E], offsets[E+1]                (offsets[E] = number of rows R)
 *       row_src[cap], row_expert[cap], row_w[cap]   (valid for r < R; the tail is set to -1 / -1 / 0)
 *       tok_rows[T, K]  row index of the j-th dispatched expert of token t (ascending e), -1 if none
 *     cap >= R is required (cap = T * top_k always suffices); K >= max dispatched experts per token.
 *     status[0] (device int) is set non-zero on overflow of cap or K.
 * ---------------------------------------------------------------------------------------------- */
size_t hdmoe_dispatch_plan_workspace_bytes(int T, int E);
int hdmoe_dispatch_plan(const float* sparse_w, int T, int E, int cap, int K, int32_t* counts, int32_t* offsets,
                        int32_t* row_src, int32_t* row_expert, float* row_w, int32_t* tok_rows,
                        int32_t* status, void* workspace, hdmoe_stream_t stream);
/* Same plan from the router kernel's own top-k output (hdmoe_router_gate_fwd: topk_idx int32 [T, K], topk_w fp32
 * [T, K]; entry (t, j) is dispatched to expert topk_idx[t, j] iff topk_w[t, j] > 0 -- the same criterion on the same
 * values as sparse_w > 0, so the plan is bit-identical) without reading the dense [T, E] matrix: T*K*8 bytes in. */
int hdmoe_dispatch_plan_topk(const int32_t* topk_idx, const float* topk_w, int T, int E, int K, int cap, int32_t* counts,
                             int32_t* offsets, int32_t* row_src, int32_t* row_expert, float* row_w, int32_t* tok_rows,
                             int32_t* status, void* workspace, hdmoe_stream_t stream);

/* Gather rows: for up to 4 tensors at once, dst_i[r, :] = src_i[row_src[r], :] for r < *n_rows_dev
 * (rows r in [*n_rows_dev, cap) are zero-filled).  Rows are raw bytes (row_bytes[i] % 16 == 0 and
 * 16-byte-aligned bases take the 128-bit / bulk-copy paths; otherwise % 4 == 0 is required).
 * srcs/dsts/row_bytes are HOST arrays of length n_tensors.  models/model_config2.py:31-33. */
int hdmoe_permute_rows(const void* const* srcs, void* const* dsts, const int64_t* row_bytes, int n_tensors,
                       const int32_t* row_src, const int32_t* n_rows_dev, int cap, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Combine -- replaces `output = zeros_like(x); output[mask] += out_e * w[mask, e]`,
 *     models/model_config2.py:23,35-37.  out[t,:] = (base ? base[t,:] : 0) + sum_j w_j * rows[r_j,:]
 *     with r_j = tok_rows[t, j] >= 0 taken in ascending j (= ascending expert: the reference's
 *     summation order; fp32 multiply then add, no FMA contraction, so fp32 results are bit-exact).
 *     row_w == NULL means weight 1 (the backward of the permute).  `base` is the optional residual of
 *     north_star item (4); the reference passes none.  D must be a multiple of 16 bytes of the row dtype.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_combine_rows(const void* rows, int rows_dtype, const int32_t* tok_rows, const float* row_w,
                       const void* base, void* out, int out_dtype, int T, int K, int64_t D,
                       hdmoe_stream_t stream);
/* Backward of combine w.r.t. the expert rows and the gate weights:
 *   d_rows[r,:] = row_w[r] * dY[row_src[r],:]            (r < R; tail rows zero-filled)
 *   d_sparse_w[row_src[r], row_expert[r]] = <rows[r,:], dY[row_src[r],:]>   (d_sparse_w pre-zeroed by callee) */
int hdmoe_combine_rows_bwd(const void* rows, int rows_dtype, const void* dY, int dy_dtype, const int32_t* row_src,
                           const int32_t* row_expert, const float* row_w, const int32_t* n_rows_dev, int cap,
                           int T, int E, int64_t D, void* d_rows, float* d_sparse_w, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (4) EDM preconditioning and Heun step -- replaces models/model_config2.py:431-449 and
 *     Utils/EDM_sampler.py:98-107 (+ the CFG lerp of :70).
 *     sigma: device pointer to 1 value (sampler, 0-dim sigma) or B values (training).
 *     Quirk Q1 is preserved: the skip connection uses the already scaled x_in = c_in * x.
 * ---------------------------------------------------------------------------------------------- */
/* x_in = x * c_in(sigma)  (x fp32 [B, n_per_sample]) */
int hdmoe_edm_precond_in(const float* x, const float* sigma, int n_sigma, float sigma_data, void* x_in,
                         int x_in_dtype, int64_t B, int64_t n_per_sample, hdmoe_stream_t stream);
/* D = c_skip * x_in + c_out * F   (fp32 out) */
int hdmoe_edm_precond_out(const void* x_in, int x_in_dtype, const void* F, int f_dtype, const float* sigma,
                          int n_sigma, float sigma_data, float* D, int64_t B, int64_t n_per_sample,
                          hdmoe_stream_t stream);
/* backward of precond_out:  d_F = c_out * dD ;  d_x_in = c_skip * dD   (either output may be NULL) */
int hdmoe_edm_precond_out_bwd(const float* dD, const float* sigma, int n_sigma, float sigma_data, void* dF,
                              int df_dtype, void* d_xin, int dxin_dtype, int64_t B, int64_t n_per_sample,
                              hdmoe_stream_t stream);
/* backward of precond_in:   d_x = c_in * d_x_in   (fp32 out) */
int hdmoe_edm_precond_in_bwd(const void* d_xin, int dxin_dtype, const float* sigma, int n_sigma, float sigma_data,
                             float* dx, int64_t B, int64_t n_per_sample, hdmoe_stream_t stream);
/* Heun step, stage A (Utils/EDM_sampler.py:98-99 + model_config2.py:440):
 *   x_hat = x_cur + noise_scale * eps   (eps may be NULL when noise_scale == 0)
 *   x_in  = x_hat * c_in(t_hat)                                                                   */
int hdmoe_edm_heun_pre(const float* x_cur, const float* eps, float noise_scale, float t_hat, float sigma_data,
                       float* x_hat, void* x_in, int x_in_dtype, int64_t n, hdmoe_stream_t stream);
/* stage B (Euler, :100-102):  D = c_skip*x_in + c_out*F  [guided: D = Dg + guidance*(D - Dg)]
 *   d_cur = (x_hat - D)/t_hat ; x_next = x_hat + (t_next - t_hat)*d_cur ; x_in_next = x_next*c_in(t_next)
 * F_guide == NULL means guidance == 1.  t_next == 0 (last step) skips x_in_next.
 * x_in == NULL: F (and F_guide) already are denoised estimates D of a foreign model.                */
int hdmoe_edm_heun_euler(const float* x_hat, const void* x_in, int x_in_dtype, const void* F, const void* F_guide,
                         int f_dtype, float guidance, float t_hat, float t_next, float sigma_data, float* d_cur,
                         float* x_next, void* x_in_next, int64_t n, hdmoe_stream_t stream);
/* stage C (2nd-order correction, :104-107):  D' from (x_in_next, F', t_next);
 *   d' = (x_next - D')/t_next ; x_out = x_hat + (t_next - t_hat)*(0.5*d_cur + 0.5*d')             */
int hdmoe_edm_heun_correct(const float* x_hat, const float* x_next, const void* x_in_next, int x_in_dtype,
                           const void* F, const void* F_guide, int f_dtype, float guidance, float t_hat,
                           float t_next, float sigma_data, const float* d_cur, float* x_out, int64_t n,
                           hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5) W-PREP -- multi-tensor magnitude-preserving weight preparation; replaces the per-call weight
 *     math of MP_Conv.forward, models/model_internals.py:253-260 (+ normalize, :26-30), for MANY
 *     layers in one launch.  For row o of tensor i (fan_in = elements per row):
 *        force != 0 (training):  w[o,:] <- w[o,:] / (eps + ||w[o,:]|| / sqrt(fan_in))     (in place, Q6)
 *        w_hat[o,:] = w[o,:] / (eps + ||w[o,:]|| / sqrt(fan_in)) * gain / sqrt(fan_in)
 *     Output layout HDMOE_WLAYOUT_SAME keeps [rows][fan_in]; HDMOE_WLAYOUT_TAPS writes the implicit-GEMM
 *     layout [tap][rows][cin_pad] (K-major, zero padded) consumed by hdmoe_gconv2_fwd; HDMOE_WLAYOUT_TAPS_T the
 *     transposed, tap-flipped operand with which the same kernel computes the data gradient.
 * ---------------------------------------------------------------------------------------------- */
#define HDMOE_WLAYOUT_SAME 0
#define HDMOE_WLAYOUT_TAPS 1
#define HDMOE_WLAYOUT_TAPS_T 2  /* data-gradient operand: [taps-1-tap][cin_rows][cout_pad] (transposed, flipped)    */
typedef struct {
    float* w;               /* [rows][fan_in] fp32 master weights (mutated when force != 0)                   */
    void* w_hat;            /* output 1                                                                        */
    void* w_hat2;           /* optional output 2 (layout2), NULL to skip                                       */
    const float* gain_ptr;  /* optional device scalar gain (e.g. Unet_expert.out_gain); NULL -> gain           */
    const int32_t* active;  /* optional device flag: the in-place rewrite is skipped when *active == 0 (an     */
                            /* expert that received no rows does not run MP_Conv.forward in the reference)     */
    float gain;
    int32_t rows, fan_in;   /* fan_in = cin * taps                                                             */
    int32_t cin, taps, cin_pad;   /* TAPS: inner dim padded to cin_pad                                         */
    int32_t cin_rows, cout_pad;   /* TAPS_T: cin_rows (<= cin) rows per tap, inner dim padded to cout_pad       */
    int32_t out_dtype, layout, layout2;
    int32_t block_start;    /* filled by the host wrapper: first CTA of this tensor                            */
} hdmoe_wprep_desc;
/* descs: HOST array of n descriptors (copied to `descs_dev`, a device buffer of n*sizeof(desc) bytes). */
int hdmoe_wprep_fwd(hdmoe_wprep_desc* descs_host, void* descs_dev, int n, int force, hdmoe_stream_t stream);
/* d_w[o,:] from d_w_hat[o,:] (layout SAME, fp32), through ONE normalisation (the gradient path of
 * models/model_internals.py:258-259).  d_gain (device scalar, accumulated) may be NULL. */
int hdmoe_wprep_bwd(const float* w, const float* d_w_hat, const float* gain_ptr, float gain, int rows, int fan_in,
                    float* d_w, float* d_gain, hdmoe_stream_t stream);
/* Multi-tensor backward: one launch for n weights.  d_w_hat is fp32 in layout SAME ([rows][fan_in]) or TAPS
 * ([tap][rows][cin_pad], what the weight-gradient kernel accumulates); d_w is written in the master layout. */
typedef struct {
    const float* w;
    const float* d_w_hat;
    const float* gain_ptr;
    float* d_w;
    float* d_gain;          /* optional, accumulated with atomicAdd */
    float gain;
    int32_t rows, fan_in, cin, taps, cin_pad, layout;
    int32_t block_start;
} hdmoe_wprep_bwd_desc;
int hdmoe_wprep_bwd_multi(hdmoe_wprep_bwd_desc* descs_host, void* descs_dev, int n, hdmoe_stream_t stream);
/* Launch-only variants for a descriptor table that is already resident on the device (block_start filled,
 * total_rows = sum of rows): no host->device copy, so the launch can be captured into a CUDA graph. */
int hdmoe_wprep_fwd_resident(const void* descs_dev, int n, int total_rows, int force, hdmoe_stream_t stream);
int hdmoe_wprep_bwd_multi_resident(const void* descs_dev, int n, int total_rows, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (7) Trunk attention, head_dim = 4 -- replaces the matmul / softmax / matmul chain of MP_Attention.forward,
 *     models/model_internals.py:380-404, when there is no rel_pos_bias (cross_attn, cross_attn_text of
 *     models/model_config2.py:279-289).  q [B,Sq,heads*4], k / v [B,Sk,heads*4], o like q, fp32, contiguous.
 *     lse [B,heads,Sq] (log2-domain log-sum-exp of the scaled logits) is saved for the backward;
 *     Dbuf [B,heads,Sq] is backward scratch.
 * ---------------------------------------------------------------------------------------------- */
/* csrc/attention_tc.cu: warp-level m16n8k8 TF32 MMAs with split (hi + lo) operands for the
 * logits and for V / K / Q / dO, probabilities kept in registers.  split_p != 0 also splits p and dS into hi + lo
 * (fp32-grade results, used when TF32 matmuls are disabled); split_p == 0 rounds them to TF32 (the setting of the
 * reference's own run with torch.backends.cuda.matmul.allow_tf32).  Any number of heads. */
int hdmoe_attn_d4_tc_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int B, int Sq, int Sk,
                         int heads, float scale, int split_p, hdmoe_stream_t stream);
int hdmoe_attn_d4_tc_bwd(const float* q, const float* k, const float* v, const float* o, const float* dO,
                         const float* lse, float* dq, float* dk, float* dv, float* Dbuf, int B, int Sq, int Sk,
                         int heads, float scale, int split_p, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (10) Fused DiffiT block of the ViT experts: Vit_block.forward (models/model_components.py:525-562) including its
 *      MP_Attention self-attention (models/model_internals.py:354-409), for emb 32, 8 heads, time_dim 64, hidden
 *      128, <= 64 tokens.  One launch runs every expert of the layer: row r uses expert row_expert[r] (< 0: row
 *      unused, output zero).  tok_in / tok_out [rows][64][32] fp32 (tokens beyond the expert's count ignored / zero),
 *      time [rows][64].  w_hat: prepared (normalised, gain-scaled) MP_Conv weights, expert e's block at float offset
 *      w_off[e] in the order linear1, q, k, v, out_proj [32x32 each], q_time, k_time, v_time [32x64], linear2
 *      [128x32], linear3 [32x128].  aux: expert e's block at a_off[e]: GN (w, b), norm1 (w, b), norm2 (w, b), final
 *      LayerNorm (w, b) [32 each], rel_pos_bias [8][S_e][S_e].  final_ln: also apply the final LayerNorm
 *      (Vit_expert.norm, :698).  w_off / a_off / tokens are HOST arrays of n_experts (<= 8) entries.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_vit_block_fwd(const float* tok_in, const float* time, const int32_t* row_expert, const float* w_hat,
                        const float* aux, const int64_t* w_off, const int64_t* a_off, const int32_t* tokens, int n_experts,
                        int64_t rows, int final_ln, float* tok_out, hdmoe_stream_t stream);
/* backward of the same block: recomputes it from tok_in.  d_out [rows][64][32] -> d_tok [rows][64][32] and d_time
 * [rows][64] (written); d_w / d_aux: fp32 buffers laid out like w_hat / aux, ZEROED BY THE CALLER, accumulated with
 * atomics (the rows of one expert meet there; the summation order is not deterministic). */
int hdmoe_vit_block_bwd(const float* tok_in, const float* time, const int32_t* row_expert, const float* w_hat,
                        const float* aux, const int64_t* w_off, const int64_t* a_off, const int32_t* tokens, int n_experts,
                        int64_t rows, int final_ln, const float* d_out, float* d_tok, float* d_time, float* d_w,
                        float* d_aux, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (11) Weight gradient of the thin trunk projections (the 1x1 q / k / v / out projections of MP_Attention._proj,
 *      models/model_internals.py:364-372,407): dW[32][32] += dY^T X, dY and X [rows][32] fp32, rows = B * S.
 *      dW must be zeroed by the caller (CTA partials are added with vector atomics).
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_lin32_wgrad(const float* dY, const float* X, float* dW, int64_t rows, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (9) Router trunk normalisation: GroupNorm(num_groups = 1, C) + ReLU [+ AdaptiveAvgPool2d((1,1))] of
 *     Router.hard_route (models/model_components.py:92-103) on channels-last fp32 activations x [B, HW, C].
 *     fwd: y (may be NULL) = relu(gn(x)), pooled (may be NULL) [B, C] = mean over HW of y, stats [B, 2] = (mean, rstd).
 *     bwd: exactly one of dy [B, HW, C] / dpooled [B, C]; dx [B, HW, C]; per-sample partials dgamma_part, dbeta_part
 *     [B, C] (the caller sums over B: deterministic).  C / 4 must divide 1024.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_gn1_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* pooled, float* stats, int B,
                       int HW, int C, float eps, hdmoe_stream_t stream);
int hdmoe_gn1_relu_bwd(const float* x, const float* gamma, const float* beta, const float* stats, const float* dy,
                       const float* dpooled, float* dx, float* dgamma_part, float* dbeta_part, int B, int HW, int C,
                       hdmoe_stream_t stream);
/* Same operation with the activation tensors (x, y, dy, dx) in `dtype` (HDMOE_F32 or HDMOE_BF16; NHWC bf16 is the layout
 * between the tcgen05 router-trunk convolutions); gamma, beta, stats, pooled and the partials stay fp32.  Rows
 * [g * rows_per_group, (g + 1) * rows_per_group) use gamma[g], beta[g] ([G, C] tables: several routers in one launch);
 * rows_per_group <= 0 means one group. */
int hdmoe_gn1_relu_fwd_t(const void* x, int dtype, const float* gamma, const float* beta, void* y, float* pooled,
                         float* stats, int B, int HW, int C, float eps, int rows_per_group, hdmoe_stream_t stream);
int hdmoe_gn1_relu_bwd_t(const void* x, int dtype, const float* gamma, const float* beta, const float* stats,
                         const void* dy, const float* dpooled, void* dx, float* dgamma_part, float* dbeta_part, int B,
                         int HW, int C, int rows_per_group, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (8) Fused NHWC bf16 elementwise kernels of the U-Net expert block -- replace the elementwise ATen chains of
 *     Unet_block.forward (models/model_components.py:232-253) and of Unet_expert.forward (:416, :428):
 *       pixnorm_silu : xn = x / (1e-4 + ||x||_C / sqrt(C)),  a = mp_silu(xn)            (:238, :240)
 *       gain_silu    : y = mp_silu(z * gain[row, c])   (gain NULL -> plain mp_silu)      (:242-243)
 *       axpby/scale2 : mp_sum(x, y, t) = ca*x + cb*y and its backward                    (:253)
 *       cat/split    : mp_cat along channels and its backward                            (:428)
 *       nchw_to_nhwc / nhwc_to_nchw : layout change at the dispatch / combine boundary (+ ones channel, :416)
 *     All tensors bf16, channel counts multiples of 8, 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_nhwc_pixnorm_silu_fwd(const void* x, void* xn, void* a, int64_t npix, int C, hdmoe_stream_t stream);
int hdmoe_nhwc_pixnorm_silu_bwd(const void* x, const void* g_xn, const void* g_a, void* dx, int64_t npix, int C,
                                hdmoe_stream_t stream);
int hdmoe_nhwc_gain_silu_fwd(const void* z, const float* gain, void* y, int64_t rows, int64_t pix_per_row, int C,
                             hdmoe_stream_t stream);
int hdmoe_nhwc_gain_silu_bwd(const void* z, const float* gain, const void* dy, void* dz, float* dgain, int64_t rows,
                             int64_t pix_per_row, int C, hdmoe_stream_t stream);
int hdmoe_nhwc_axpby(const void* x, const void* y, float ca, float cb, void* out, int64_t n, hdmoe_stream_t stream);
int hdmoe_nhwc_scale2(const void* g, float ca, float cb, void* gx, void* gy, int64_t n, hdmoe_stream_t stream);
int hdmoe_nhwc_cat(const void* a, const void* b, float wa, float wb, int Ca, int Cb, void* out, int64_t npix,
                   hdmoe_stream_t stream);
int hdmoe_nhwc_split(const void* g, float wa, float wb, int Ca, int Cb, void* ga, void* gb, int64_t npix,
                     hdmoe_stream_t stream);
int hdmoe_nchw_to_nhwc(const void* src, void* dst, int64_t rows, int Cs, int Cd, int64_t HW, int one_channel,
                       hdmoe_stream_t stream);
int hdmoe_nhwc_to_nchw(const void* src, void* dst, int64_t rows, int Cs, int Cd, int64_t HW, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (6) Grouped implicit-GEMM convolution / GEMM on tcgen05 + TMEM + TMA -- replaces the F.conv2d /
 *     F.linear calls of MP_Conv inside the experts (models/model_internals.py:261-271), for ALL
 *     experts of one layer in a single persistent launch.  Declared in hdmoe_gemm.h.
 * ---------------------------------------------------------------------------------------------- */

/* ------------------------------------------------------------------------------------------------
 * (10) Optimizer side of the train step -- replaces torch.nn.utils.clip_grad_norm_(parameters, max_norm) followed by
 *      AdamW.step() (Utils/training.py:195-197) over all parameter tensors with three launches.
 *      table_dev: device array of n_tensors descriptors
 *          { float* p, g, m, v; int64 numel; float lr, weight_decay; int32 chunk_start, pad }   (56 bytes)
 *      where tensor i owns chunks [chunk_start, chunk_start + ceil(numel / hdmoe_optim_chunk_elems())); g == NULL skips
 *      the tensor (a parameter without a gradient, as torch does).  partial: n_chunks floats of scratch.
 *      state[3] (device, fp32): [0] number of calls (incremented), [1] total gradient norm (output),
 *      [2] clip coefficient min(1, max_norm / (norm + 1e-6)) (output; max_norm <= 0: no clipping).
 *      steps[n_tensors] (device, fp32): per-tensor update count (torch.optim.AdamW semantics), incremented for every
 *      tensor with a gradient and used for its bias correction.
 *      write_back_grad != 0 stores the clipped gradients like clip_grad_norm_ does in place.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_optim_chunk_elems(void);
int hdmoe_adamw_step(const void* table_dev, int n_tensors, int n_chunks, float* partial, float* state, float* steps,
                     float max_norm, float beta1, float beta2, float eps, int write_back_grad, hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (11) HDMOEM.forward glue on channels-last fp32 activations [P = B*H*W, C = 32] -- replaces the elementwise / 1x1
 *      convolution chains of models/model_config2.py:291-301 (text blend, mp_cat -> gate1 -> mp_silu -> gate2 -> pixel
 *      softmax -> blend -> mp_sum) and the cfg1 soft query / context swap of models/model_config1.py:277-283.
 *      swap:  q = w[b] * v + (1 - w[b]) * u,  ctx = w[b] * u + (1 - w[b]) * v     (per_sample = H*W*C elements per b)
 *             bwd writes du, dv and dw_part [B * hdmoe_trunk_swap_slices()] (sum over the slices of b = d w[b]).
 *      gate:  fin = a + *alpha_txt * (b - a);  h = W1 [c1*u ; c2*fin]  (W1 [32, 64] prepared gate1 weight);
 *             g = softmax(W2 mp_silu(h))  (W2 [2, 32]);  mix = ms_a*u + ms_b*(g0*u + g1*fin);
 *             g_out is [B, 2, H, W] (the reference's out_gate).  bwd: d_g may be NULL; dW1 / dW2 / d_alpha are
 *             ACCUMULATED (zero them first).
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_trunk_swap_fwd(const float* u, const float* v, const float* w, float* q, float* ctx, int B, int64_t per_sample,
                         hdmoe_stream_t stream);
int hdmoe_trunk_swap_slices(void);
int hdmoe_trunk_swap_bwd(const float* u, const float* v, const float* w, const float* dq, const float* dctx, float* du,
                         float* dv, float* dw_part, int B, int64_t per_sample, hdmoe_stream_t stream);
int hdmoe_trunk_gate_fwd(const float* u, const float* a, const float* b, const float* alpha_txt, const float* W1,
                         const float* W2, float* mix, float* g_out, int64_t P, int64_t HW, int C, float c1, float c2,
                         float ms_a, float ms_b, hdmoe_stream_t stream);
int hdmoe_trunk_gate_bwd(const float* u, const float* a, const float* b, const float* alpha_txt, const float* W1,
                         const float* W2, const float* d_mix, const float* d_g, float* du, float* da, float* db,
                         float* dW1, float* dW2, float* d_alpha, int64_t P, int64_t HW, int C, float c1, float c2,
                         float ms_a, float ms_b, hdmoe_stream_t stream);

/* Branch scaling (models/model_config2.py:244-251, models/model_config1.py:246-252).
 *   analytic_scaling: w = sigmoid((4 * time_vec[b] - transition_point) / softness);
 *                     scaling[b] = ((w + 0.01) * 2, (1 - w + 0.01) * 2)                 (ViT gain, U-Net gain)
 *   scale_pair_fwd:   in_vit = scaling[b, 0] * feats, in_unet = scaling[b, 1] * feats  (fp32 [B, C, HW], C = 32) and,
 *                     when trunk_bf16 != NULL, the channels-last bf16 copy [2B, HW, C] of both (ViT rows first) that the
 *                     tcgen05 router trunk reads.
 *   scale_pair_bwd:   d_feats from the gradients of the three outputs (each may be NULL); ds_part
 *                     [B, hdmoe_scale_pair_tiles(HW), 2]: the caller sums over the tiles (deterministic) -> d scaling. */
int hdmoe_analytic_scaling(const float* time_vec, float transition_point, float softness, float* scaling, int B,
                           hdmoe_stream_t stream);
int hdmoe_scale_pair_fwd(const float* feats, const float* scaling, float* in_vit, float* in_unet, void* trunk_bf16, int B,
                         int C, int64_t HW, hdmoe_stream_t stream);
int hdmoe_scale_pair_tiles(int64_t HW);
int hdmoe_scale_pair_bwd(const float* feats, const float* scaling, const float* g_vit, const float* g_unet,
                         const void* g_trunk_bf16, float* d_feats, float* ds_part, int B, int C, int64_t HW,
                         hdmoe_stream_t stream);

/* Scaling_router.forward (models/model_components.py:41-66, model_config1) as one kernel per direction:
 *   x [B, 64] -> W1 [128, 64] -> GroupNorm(1, 128) (g1, b1) -> ReLU -> W2 [256, 128] -> GroupNorm(1, 256) (g2, b2) -> ReLU
 *   -> * keep [B, 256] (dropout mask already divided by 1 - p; NULL = no dropout) -> W3 [2, 256] -> + zeta * noise [B, 2]
 *   (NULL in eval) -> softmax * 2 -> out [B, 2].  W1..W3 are the PREPARED MP_Conv weights.  The backward recomputes
 *   the forward and ACCUMULATES dW1, dg1, db1, dW2, dg2, db2, dW3 (zero them first); dx [B, 64] is written. */
int hdmoe_scaling_router_fwd(const float* x, const float* W1, const float* g1, const float* b1, const float* W2,
                             const float* g2, const float* b2, const float* W3, const float* noise, float zeta,
                             const float* keep, float eps, float* out, int B, int D, hdmoe_stream_t stream);
int hdmoe_scaling_router_bwd(const float* x, const float* W1, const float* g1, const float* b1, const float* W2,
                             const float* g2, const float* b2, const float* W3, const float* noise, float zeta,
                             const float* keep, float eps, const float* d_out, float* dx, float* dW1, float* dg1,
                             float* db1, float* dW2, float* dg2, float* db2, float* dW3, int B, int D,
                             hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (12) EDM_LOSS data term (Utils/utils.py:135-146): per-sample squared error se[b] = sum_i (D[b,i] - x0[b,i])^2 and its
 *      backward dD[b,i] = 2 (D - x0) g_se[b]; every image-dependent loss term is a function of se and the per-sample
 *      log-variance, so the rest of the loss runs on [B]-sized vectors.  fp32, row length a multiple of 4.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_sqerr_rows(const float* d, const float* x, float* se, int B, int64_t per, hdmoe_stream_t stream);
int hdmoe_sqerr_rows_bwd(const float* d, const float* x, const float* g_se, float* dd, int B, int64_t per,
                         hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (13) On-device producers of the train step's inputs, one launch:
 *      x[b,i] = x0[b,i] + eps[b,i] * sigma[b]            (Utils/training.py:133-134; fp32, product rounded, then sum)
 *      mask_g[b,e] = |pct(sigma[b]) - centers_g[e]| <= bandwidth_g, and the min_active nearest experts forced live,
 *      pct = clamp(0.5 (1 + erf((log sigma - p_mean) / (p_std sqrt 2))), 0, 1)   (MaskGenerator, Utils/utils.py:281-309)
 *      for up to two generators (U-Net and ViT router masks).  `bandwidth` is the value of the host-side
 *      bandwidth_scheduler at the current step.  Either generator may be NULL (with its mask).
 * ---------------------------------------------------------------------------------------------- */
#define HDMOE_MAX_MASK_EXPERTS 16
typedef struct hdmoe_maskgen {
    float centers[HDMOE_MAX_MASK_EXPERTS];
    float p_mean, p_std, bandwidth;
    int32_t n_experts, min_active;
} hdmoe_maskgen_t;
int hdmoe_train_inputs(const float* x0, const float* eps, const float* sigma, float* x, int B, int64_t per,
                       const hdmoe_maskgen_t* gen_a, float* mask_a, const hdmoe_maskgen_t* gen_b, float* mask_b,
                       hdmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (14) Peer-memory exchange for expert parallelism (the reference has no distributed code).  Buffers are
 *      mapped into every rank's address space by the host side (CUDA IPC); the pointer tables are DEVICE arrays of
 *      `world` addresses.  Both calls are plain kernel launches: CUDA-graph capturable, no host state.
 *   hdmoe_peer_barrier: device-side barrier over all ranks (flag arrays of `world` int32 per rank, zero-initialised;
 *      `epoch_dev` = 2 int32 of this rank: [0] barriers passed, advanced by the kernel; [1] sticky failure flag, set when
 *      a peer did not arrive within 8 s -- later barriers then return at once instead of hanging the GPU).  Every rank
 *      must issue the same sequence.
 *   hdmoe_peer_pull: dst[g * seg_bytes ...] <- peer g's buffer at byte offset rank_segment * seg_bytes (equal-split
 *      all-to-all: rank_segment = this rank) or 0 (all-gather: rank_segment = -1), g = 0 .. world-1.
 * ---------------------------------------------------------------------------------------------- */
int hdmoe_peer_barrier(int32_t* my_flags, const int64_t* peer_flag_ptrs_dev, int32_t* epoch_dev, int rank, int world,
                       hdmoe_stream_t stream);
int hdmoe_peer_pull(void* dst, const int64_t* peer_src_ptrs_dev, int64_t seg_bytes, int rank_segment, int world,
                    hdmoe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HDMOE_B200_H_ */
