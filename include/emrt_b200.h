/* emrt_b200 — C ABI of the B200-native (sm_100a) EMRT hot path.
 *
 * The reference (peach-xiao/EMRT) has no FFI / custom-op interface: its hot path is a composition of
 * PaddlePaddle Python ops.  Each entry point below names the reference code it replaces; the Python shims in
 * emrt_b200/{msda,infer,paddle_shim}.py bind them with ctypes.
 * Paths are relative to /root/reference/semantic_segmentation/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - the library allocates nothing the caller can see; scratch is passed in (query *_workspace_bytes);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value: 0 on success, negative emrt_status otherwise; emrt_last_error() gives the text
 *     (thread-local);
 *   - shapes / level tables are HOST int32 arrays, copied by value into the launch (no device->host sync);
 *   - no CPU fallback and no other architecture: emrt_device_check() fails unless the current device is
 *     compute capability 10.x.
 */
#ifndef EMRT_B200_H_
#define EMRT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMRT_ABI_VERSION 1
#define EMRT_MAX_LEVELS 8

typedef enum emrt_status {
  EMRT_OK = 0,
  EMRT_ERR_INVALID_ARGUMENT = -1,
  EMRT_ERR_CUDA = -2,
  EMRT_ERR_UNSUPPORTED = -3,
  EMRT_ERR_ARCH = -4,
  EMRT_ERR_WORKSPACE = -5
} emrt_status;

typedef enum emrt_dtype { EMRT_F32 = 0, EMRT_BF16 = 1, EMRT_F16 = 2, EMRT_I32 = 3, EMRT_U8 = 4 } emrt_dtype;

/* How sampling positions are given to the gather kernels. */
typedef enum emrt_loc_mode {
  /* `loc` holds normalised absolute locations [B,Lq,M,L,P,2] (x,y) in [0,1] — exactly the
   * `sampling_locations` argument of deformable_attention_core_func (src/models/EMRT_utils/utils.py:64). */
  EMRT_LOC_NORMALIZED = 0,
  /* `loc` holds the raw sampling_offsets output in PIXELS [B,Lq,M,L,P,2] and `ref` the reference points
   * [Bref,Lq,L,2] (Bref = B or 1) in [0,1]; the kernel forms x = ref_x*W_l + off_x - 0.5, which is
   * transformer_encoder_decoder.py:98-102 folded into utils.py:79,87. */
  EMRT_LOC_PIXEL_OFFSET = 1,
  /* flag, OR-ed into `mode`: `value` is laid out head-major [B,M,Lv,D] (what emrt_linear_fwd writes with
   * EMRT_EPI_HEAD_MAJOR) instead of the reference's [B,Lv,M,D].  Forward only; bf16, D=32, L=3, P=6. */
  EMRT_VALUE_HEAD_MAJOR = 2,
  /* flag, OR-ed into `mode`: a locality promise, never a correctness condition.  The Lq == Lv queries are the
   * pixels of the value pyramid in level-major raster order and their reference points lie near their own pixel
   * centres (TransformerEncoder.get_reference_points, transformer_encoder_decoder.py:213-228).  Selects the
   * window-staged kernels (TMA -> shared-memory windows), forward and backward; the forward takes either value
   * layout (4-D tensor maps over the head-major tensor, 5-D over the reference's pixel-major one). */
  EMRT_QUERY_PIXEL_GRID = 4
} emrt_loc_mode;

/* Epilogues of emrt_linear_fwd (bit flags). */
typedef enum emrt_epilogue {
  EMRT_EPI_NONE = 0,
  EMRT_EPI_ROW_MASK = 1,      /* y[r,:] *= row_scale[r]      (value_mask, t_e_d.py:84-86)              */
  EMRT_EPI_RELU = 2,          /* y = max(y,0)                 (FFN activation, t_e_d.py:158)            */
  EMRT_EPI_RESIDUAL_LN = 4,   /* y = LayerNorm(residual + y)  (t_e_d.py:199-200; N must be 256)         */
  EMRT_EPI_MSDA_QPROJ = 8,    /* N = M*L*P*3: [offsets | logits] -> pixel offsets + softmax(L*P)         */
  EMRT_EPI_HEAD_MAJOR = 16    /* store y[b, n/hm_D, r % hm_rows, n % hm_D] (r = b*hm_rows + pix): the value    *
                               * tensor laid out [B,M,Lv,D] for the gather; needs hm_rows, hm_D (tcgen05 only) */
} emrt_epilogue;

int emrt_version(void);
const char* emrt_last_error(void);
/* 0 iff the current CUDA device is a compute-capability-10.x part (B200). */
int emrt_device_check(void);
/* Number of kernels this library has launched since load / last reset (bench.py's gpu_launches). */
int64_t emrt_launch_count(void);
void emrt_reset_launch_count(void);

/* ---- a3: deformable_attention_core_func (src/models/EMRT_utils/utils.py:64-97) -------------------------
 * value [B,Lv,M,D] (value_dtype: F32|BF16), loc/attn per `mode` (loc_dtype: F32|F16|BF16),
 * out [B,Lq,M*D] (value_dtype).  shapes_hw_host = {H0,W0,H1,W1,...}, level_start_host = {0,H0*W0,...}.
 * ref (mode PIXEL_OFFSET only) is F32 [Bref,Lq,L,2], ref_batch_stride = 0 (shared) or Lq*L*2.            */
int emrt_msda_gather_fwd(const void* value, const void* loc, const void* attn, const float* ref,
                         int64_t ref_batch_stride, void* out, int B, int Lq, int Lv, int M, int D, int L,
                         int P, const int32_t* shapes_hw_host, const int32_t* level_start_host,
                         int value_dtype, int loc_dtype, int mode, void* stream);

/* Same, with a locality hint for the window-staged kernel (mode EMRT_QUERY_PIXEL_GRID): window_center_host is a host
 * int32 [M,L,2] (x, y) in pixels of level l — where head m's samples of level l lie relative to the reference point
 * on average, i.e. the rounded mid-range of the `sampling_offsets` bias (transformer_encoder_decoder.py:47-55 initialises
 * it to one direction per head).  The staged windows are centred there.  NULL = centred on the reference points.
 * Results never depend on the hint: a sample outside the staged window is read from global memory.               */
int emrt_msda_gather_fwd_hint(const void* value, const void* loc, const void* attn, const float* ref,
                              int64_t ref_batch_stride, void* out, int B, int Lq, int Lv, int M, int D, int L,
                              int P, const int32_t* shapes_hw_host, const int32_t* level_start_host,
                              int value_dtype, int loc_dtype, int mode, const int32_t* window_center_host,
                              void* stream);

/* Backward of the above (what Paddle autograd derives through F.grid_sample, utils.py:87-94).
 * grad_out [B,Lq,M*D] (value_dtype); grad_value F32 [B,Lv,M,D] MUST be zeroed by the caller (accumulated);
 * grad_loc F32 [B,Lq,M,L,P,2] (d/d loc in the units of `mode`), grad_attn F32 [B,Lq,M,L,P].              */
int emrt_msda_gather_bwd(const void* grad_out, const void* value, const void* loc, const void* attn,
                         const float* ref, int64_t ref_batch_stride, float* grad_value, float* grad_loc,
                         float* grad_attn, int B, int Lq, int Lv, int M, int D, int L, int P,
                         const int32_t* shapes_hw_host, const int32_t* level_start_host, int value_dtype,
                         int loc_dtype, int mode, void* stream);

/* Same, selecting the windowed backward when `mode` carries EMRT_QUERY_PIXEL_GRID (bf16 pixel-major value, D=32, L=3,
 * P=6, Lq == Lv on a regular pyramid): grad_value is accumulated in fixed point in shared-memory windows with integer
 * shared-memory reductions and sent to L2 once per window pixel (no float atomics on the hot path); window_center_host
 * is the same optional hint as in emrt_msda_gather_fwd_hint.  Other shapes run the generic backward.               */
int emrt_msda_gather_bwd_hint(const void* grad_out, const void* value, const void* loc, const void* attn,
                              const float* ref, int64_t ref_batch_stride, float* grad_value, float* grad_loc,
                              float* grad_attn, int B, int Lq, int Lv, int M, int D, int L, int P,
                              const int32_t* shapes_hw_host, const int32_t* level_start_host, int value_dtype,
                              int loc_dtype, int mode, const int32_t* window_center_host, void* stream);

/* ---- nn.Linear (transformer_encoder_decoder.py:36-42,83,89,92,106,118,121) ------------------------------
 * y[rows,N] = epilogue(x[rows,K] @ W + bias).  W is given either in Paddle's own layout [K,N]
 * (w_transposed = 0) or pre-packed [N,K] (w_transposed = 1, see emrt_pack_weight).  x_dtype F32 runs the
 * fp32 SIMT path (parity mode); BF16 runs the tcgen05/TMEM/TMA path (w must then be BF16 [N,K]).
 * bias is always F32 [N].  y_dtype: F32|BF16|F16.  Extra operands by epilogue flag:
 *   ROW_MASK: row_scale F32 [rows];  RESIDUAL_LN: residual (x_dtype) [rows,N], ln_gamma/ln_beta F32 [N] (BF16 /
 *               tcgen05 only, N = 256, K % 64 == 0; y may alias residual or x);
 *   MSDA_QPROJ: y = pixel offsets [rows, 2N/3] (y_dtype), y2 = softmax'd weights [rows, N/3] (y_dtype),
 *               softmax group = qproj_group (= L*P).                                                      */
/* Optional operand of the RESIDUAL_LN epilogue: the encoder layer's conv branch added BEHIND the LayerNorm,
 *   y = LayerNorm(residual + x W + b) * gamma + beta + GELU(GroupNorm_l(conv)) + skip     (t_e_d.py:159-160,187-189,203)
 * conv / skip BF16 [rows, N]; stats = emrt_groupnorm_stats' sums [B, L, groups, 2] of conv; gamma / beta F32 [L, N];
 * rows = B * Lv tokens of the level pyramid shapes_hw = {H0,W0,H1,W1,...}.  N = 256, groups = 32, K > 256 (linear2).  */
typedef struct emrt_gn_branch {
  const void* conv; const void* skip; const float* stats; const float* gamma; const float* beta;
  int32_t L; int32_t groups; int32_t Lv; float eps;
  int32_t shapes_hw[2 * EMRT_MAX_LEVELS];
} emrt_gn_branch;

typedef struct emrt_linear_args {
  const void* x; const void* w; const float* bias; void* y;
  int64_t rows; int32_t K; int32_t N;
  int32_t x_dtype; int32_t w_dtype; int32_t y_dtype; int32_t w_transposed;
  int32_t epilogue;
  const float* row_scale;
  const void* residual; const float* ln_gamma; const float* ln_beta; float ln_eps;
  void* y2; int32_t qproj_group;
  int32_t impl;               /* 0 = auto, 1 = force SIMT, 2 = force tcgen05 */
  int32_t hm_rows; int32_t hm_D;   /* HEAD_MAJOR: rows per batch element (Lv) and head dim */
  /* Optional broadcast addend of x (with_pos_embed, transformer_encoder_decoder.py:154-155,198,283,288):
   * y = epilogue((x[r,:] + x2[r % x2_period,:]) @ W + bias), evaluated as x W + x2 W in the fp32 accumulator (the sum is
   * never rounded to bf16 and never written).  x2 is BF16 [x2_period + 127, K]: the x2_period rows of the addend
   * continued cyclically for 127 more rows, so that any 128-row tile reads one contiguous box.  tcgen05 path only. */
  const void* x2; int32_t x2_period;
  /* The same addend for a projection whose weights are fixed between calls, precomputed: row_bias F16
   * [row_bias_period + 127, N] = x2 W + bias (cyclic like x2; pass bias = NULL), added to the fp32 accumulator in the
   * epilogue — (x + pos) W + b = x W + (pos W + b) with no extra MMA work.  MSDA_QPROJ epilogue only.            */
  const void* row_bias; int32_t row_bias_period;
  const emrt_gn_branch* gn;   /* RESIDUAL_LN only; NULL = plain LayerNorm epilogue */
  /* x given channel-major: x is [B, K, x_nchw_hw] (an NCHW feature map, x_nchw_hw = H * W pixels per image) and the GEMM
   * row r = b * x_nchw_hw + pixel reads x[b, :, pixel] — the 1x1 convolutions of input_proj (t_e_d.py:417-419) without a
   * transposed copy of the feature maps: the tensor pipe takes A pixel-contiguous (MN-major).  0 = x is [rows, K].
   * tcgen05 path only; x_nchw_hw % 128 == 0, K % 64 == 0, bias-only epilogue.                                            */
  int32_t x_nchw_hw;
} emrt_linear_args;
int emrt_linear_fwd(const emrt_linear_args* args, void* stream);

/* ---- TransformerEncoderLayer.forward_ffn in one kernel (transformer_encoder_decoder.py:157-160, + :187-189,203) ------
 * y = LayerNorm(x + linear2(relu(linear1(x)))) * ln_gamma + ln_beta  (+ GELU(GroupNorm_l(conv)) + skip when gn != NULL).
 * The hidden activations [rows, d_ff] never reach HBM: each CTA walks d_ff in 128-unit chunks, the first GEMM's
 * accumulator chunk is converted (bias, ReLU, bf16) into the shared-memory A operand of the second, whose accumulator
 * (one 256-column TMEM row per token) feeds the LayerNorm epilogue.  BF16 / tcgen05 only: x, y BF16 [rows, 256];
 * w1 BF16 [d_ff, 256] and w2 BF16 [256, d_ff] pre-packed (emrt_pack_weight); b1 F32 [d_ff], b2 / ln_gamma / ln_beta F32
 * [256]; d_model = 256, d_ff % 128 == 0; every pointer 16-byte aligned.  y must not alias x (x is re-read as the
 * residual while other tiles are being written).                                                                    */
typedef struct emrt_ffn_args {
  const void* x; const void* w1; const float* b1; const void* w2; const float* b2;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  void* y; int64_t rows; int32_t d_model; int32_t d_ff;
  const emrt_gn_branch* gn;   /* NULL = plain LayerNorm epilogue */
} emrt_ffn_args;
int emrt_ffn_fused_fwd(const emrt_ffn_args* args, void* stream);

/* ---- backward of nn.Linear (what Paddle autograd derives for transformer_encoder_decoder.py:83,89,92,106) --------
 * Data gradient dx = dy W^T needs no entry point of its own: call emrt_linear_fwd with x = dy, K = N_fwd,
 * N = K_fwd, bias = NULL and w = the Paddle-layout weight [in,out] passed as w_transposed = 1.
 * Weight / bias gradient: dw F32 [K,N] += x[rows,K]^T dy[rows,N]; db F32 [N] += column sums of dy (db may be NULL).
 * dw / db are ACCUMULATED (caller zeroes them or keeps accumulating over micro-batches).  x, dy: F32 | BF16.       */
int emrt_linear_bwd_weight(const void* x, const void* dy, float* dw, float* db, int64_t rows, int K, int N,
                           int x_dtype, int dy_dtype, void* stream);

/* Backward of softmax(L*P) + sampling-location arithmetic (transformer_encoder_decoder.py:92-102):
 * grad_loc F32 [rows,M,L,P,2] and grad_attn F32 [rows,M,L,P] (from emrt_msda_gather_bwd), attn [rows,M,L,P] (the
 * saved softmax output; attn_dtype F32|F16|BF16) -> dq [rows, 3*M*L*P] = [offset grads | logit grads] (dq_dtype
 * F32|BF16), the column layout of the fused [sampling_offsets | attention_weights] projection.  `mode` is the loc
 * mode the gather ran in (NORMALIZED divides the offset grads by (W_l, H_l)).                                       */
int emrt_msda_qproj_bwd(const float* grad_loc, const float* grad_attn, const void* attn, void* dq, int64_t rows,
                        int M, int L, int P, const int32_t* shapes_hw_host, int attn_dtype, int dq_dtype, int mode,
                        void* stream);

/* Gradient of the reference points (sampling_locations = reference_points[:, :, None, :, None, :] + offsets / (W_l, H_l),
 * transformer_encoder_decoder.py:98-102; the decoder's reference points are a trained Linear + sigmoid, :466-467):
 * grad_ref F32 [ref_batches, Lq, L, 2] = sum over heads and points of grad_loc F32 [B, Lq, M, L, P, 2] (times (W_l, H_l)
 * in PIXEL_OFFSET mode, where grad_loc is per pixel), summed over the batch in a fixed order when ref_batches == 1.
 * Overwrites grad_ref.                                                                                               */
int emrt_msda_ref_bwd(const float* grad_loc, float* grad_ref, int B, int ref_batches, int Lq, int M, int L, int P,
                      const int32_t* shapes_hw_host, int mode, void* stream);

/* dst[r,c] = src[r,c] * row_scale[r] (row_scale may be NULL): value_mask backward + cast of the fp32 grad_value
 * (transformer_encoder_decoder.py:84-86).  dst_dtype F32|BF16; cols % 4 == 0.                                       */
int emrt_scale_rows_cast(const float* src, const float* row_scale, void* dst, int64_t rows, int cols, int dst_dtype,
                         void* stream);

/* Pack Paddle-layout weights: dst[n, k] (BF16, [N_total,K]) = src[k, n] (F32|BF16, [K,N]) for n in
 * [0,N), written at row offset dst_row0 — lets sampling_offsets.weight and attention_weights.weight be
 * concatenated into one [M*L*P*3, K] operand.                                                             */
int emrt_pack_weight(const void* src, int src_dtype, void* dst_bf16, int K, int N, int dst_row0,
                     void* stream);

/* softmax(L*P) + location arithmetic on raw projections (transformer_encoder_decoder.py:92-102), for the
 * un-fused path.  off_raw F32 [rows, M*L*P*2] (+bias already applied), logit_raw F32 [rows, M*L*P].
 * mode NORMALIZED: loc_out = ref + off/(W_l,H_l) (out_dtype), needs ref [Bref,Lq,L,2];
 * mode PIXEL_OFFSET: loc_out = off_raw converted.  attn_out = softmax (out_dtype).                         */
int emrt_msda_softmax_loc(const float* off_raw, int64_t off_ld, const float* logit_raw, int64_t logit_ld,
                          const float* ref, int64_t ref_batch_stride, void* loc_out, void* attn_out, int B,
                          int Lq, int M, int L, int P, const int32_t* shapes_hw_host, int out_dtype,
                          int mode, void* stream);

/* y = LayerNorm(residual + x) * gamma + beta over the last dim N (nn.LayerNorm eps 1e-5;
 * transformer_encoder_decoder.py:199-200, 159-160).  residual may be NULL.  dtype F32|BF16 (x, residual, y). */
int emrt_add_layernorm(const void* x, const void* residual, const float* gamma, const float* beta, void* y,
                       int64_t rows, int N, float eps, int dtype, void* stream);

/* ---- TransformerEncoderLayer glue (transformer_encoder_decoder.py:184-204) -------------
 * y = LayerNorm(x + residual) * gamma + beta (+ post_add): norm1 / norm2 (:199-200, :159-160) with the layer's final
 * `src + src_flatten` (:203) folded in as post_add.  residual / post_add may be NULL.  dtype F32|BF16 (x, residual,
 * post_add, y); N a multiple of 128 (F32) / 256 (BF16), <= 8 16-byte vectors per lane.                                      */
int emrt_residual_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                            const void* post_add, void* y, int64_t rows, int N, float eps, int dtype, void* stream);

/* Pack one level's Conv2D weight (Paddle layout [Cout, Cin, 3, 3], F32) into the conv operand
 * dst[level][tap = ky*3+kx][Cout][Cin] (dst_dtype F32|BF16; dst holds all levels, 9*C*C elements each).             */
int emrt_pack_conv3x3_weight(const float* src, void* dst, int C, int level, int dst_dtype, void* stream);

/* The conv branch's convolutions (conv{l}: 3x3, stride 1, zero padding 1, no bias, :125-144) applied per level directly
 * on the token layout: x, y [B, Lv, C] (dtype F32|BF16), level l using w_packed[l].  BF16 runs the tcgen05 implicit
 * GEMM (nine shifted TMA loads per tile; C = 256); impl: 0 = auto, 1 = force SIMT, 2 = force tcgen05.                  */
int emrt_conv3x3_tokens_fwd(const void* x, const void* w_packed, void* y, int B, int Lv, int C, int L,
                            const int32_t* shapes_hw_host, int dtype, int w_dtype, int impl, void* stream);

/* The same convolutions (BF16, tcgen05 path only) whose epilogue also leaves the GroupNorm(32) statistics of the stored
 * output: stats_workspace F32 [emrt_conv3x3_stats_workspace_floats(B, Lv, L)], its first 2 * B * L * 32 floats the sums
 * [B, L, 32, 2] that emrt_groupnorm_stats would compute (per-slab partial sums combined in a fixed order: bit-reproducible
 * and independent of the batch a tile is processed in).  The separate statistics pass over the conv output disappears.  */
long long emrt_conv3x3_stats_workspace_floats(int B, int Lv, int L);
int emrt_conv3x3_tokens_stats_fwd(const void* x, const void* w_packed, void* y, float* stats_workspace, int B, int Lv, int C,
                                  int L, const int32_t* shapes_hw_host, int groups, void* stream);
/* The same on at most max_ctas SMs (0 = all; same results bit for bit: the tile loop is persistent and the statistics are
 * per-tile partial sums).  For a caller that runs the convolution on a second stream BESIDE the sampling gather of the same
 * layer (the conv branch t_e_d.py:185-196 does not depend on the attention branch :198-200): the tensor-pipe kernel is
 * power-limited when it owns the whole GPU and costs 28 % less SM time on a share of it, while the gather, bound by the
 * shared-memory pipe, takes the other SMs (emrt_msda_args.gather_start_event marks the moment).                           */
int emrt_conv3x3_tokens_stats_part_fwd(const void* x, const void* w_packed, void* y, float* stats_workspace, int B, int Lv,
                                       int C, int L, const int32_t* shapes_hw_host, int groups, int max_ctas, void* stream);

/* y = GELU(GroupNorm_l(conv)) + x per level (GroupNorm(groups, C) eps, exact erf GELU, :187-189); conv, x, y
 * [B, Lv, C] (dtype F32|BF16); gamma, beta F32 [L, C]; stats_workspace F32
 * [emrt_groupnorm_workspace_floats(B, L, groups)] (scratch: sums, per-CTA partials, ticket counters — the statistics are
 * reduced in a fixed order, no floating-point atomics, so results are bit-reproducible run to run).                  */
long long emrt_groupnorm_workspace_floats(int B, int L, int groups);
int emrt_groupnorm_gelu_residual(const void* conv, const void* x, const float* gamma, const float* beta, void* y,
                                 float* stats_workspace, int B, int Lv, int C, int L, int groups, float eps,
                                 const int32_t* shapes_hw_host, int dtype, void* stream);

/* The same conv branch without ever materialising it: emrt_groupnorm_stats leaves the per (batch, level, group) sums
 * (sum, sum of squares; F32 [B, L, groups, 2] at the start of a buffer of emrt_groupnorm_workspace_floats(B, L, groups)
 * floats) of x [B, Lv, C]; emrt_residual_layernorm_gn then evaluates
 *   y = LayerNorm(x + residual) * ln_gamma + ln_beta + GELU(GroupNorm_l(conv)) + skip
 * i.e. norm2 (:159-160) plus the layer's final `src + src_flatten` (:203) with src_flatten = conv branch (:187-196)
 * computed on the fly from the conv output, the layer input (skip) and those sums.  All tensors [B, Lv, C].            */
int emrt_groupnorm_stats(const void* x, float* stats, int B, int Lv, int C, int L, int groups,
                         const int32_t* shapes_hw_host, int dtype, void* stream);
int emrt_residual_layernorm_gn(const void* x, const void* residual, const float* ln_gamma, const float* ln_beta,
                               const void* conv, const void* skip, const float* gn_stats, const float* gn_gamma,
                               const float* gn_beta, void* y, int B, int Lv, int C, int L, int groups, float ln_eps,
                               float gn_eps, const int32_t* shapes_hw_host, int dtype, void* stream);

/* ---- input_proj glue (EncoderDecoder.forward, transformer_encoder_decoder.py:417-436,469) --------------------------
 * y[b, p, c] = x[b, c, p]: NCHW feature maps [B, C, P = H*W] (or src_psp [B, 256, 110]) -> tokens [B, P, C].            */
int emrt_nchw_to_tokens(const void* x, void* y, int B, int C, int P, int dtype, void* stream);

/* GroupNorm(groups, C) of one level's projected tokens x [B, P, C] (input_proj[i][1], no activation) written into that
 * level's slot of the concatenated token tensor: y + b * y_batch_stride + p * C (y_batch_stride in elements = Lv * C).
 * stats_workspace F32 [emrt_groupnorm_workspace_floats(B, 1, groups)].                                                */
int emrt_groupnorm_tokens(const void* x, const float* gamma, const float* beta, void* y, int64_t y_batch_stride,
                          float* stats_workspace, int B, int P, int C, int groups, float eps, int dtype, void* stream);

/* ---- TransformerDecoderLayer self-attention core (MultiHeadAttention, src/models/EMRT_utils/layers.py:282-301) ------
 * out[b,q,m*D+d] = sum_k softmax_k(scale * <Q[b,q,m,:], K[b,k,m,:]>) V[b,k,m,d] on the projected q / k / v in the token
 * layout [B, L, M*D] with row strides q_ld / k_ld / v_ld (elements), so a fused [q | k] projection is read in place.
 * out [B, Lq, M*D] contiguous.  Built for the decoder's 110 query tokens: D = 32, Lk <= 256.  dtype F32|BF16.        */
int emrt_mha_small(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld, void* out,
                   int B, int Lq, int Lk, int M, int D, float scale, int dtype, void* stream);

/* out[i] = a[i] + b[i % b_period]: with_pos_embed (transformer_encoder_decoder.py:154-155,198,283,288);
 * b_period = n for a plain add, Lq*C for a batch-shared positional embedding.  dtype F32|BF16.             */
int emrt_add_bcast(const void* a, const void* b, void* out, int64_t n, int64_t b_period, int dtype,
                   void* stream);

/* ---- a5: UpHead tail (src/models/paddle_EMRT.py:178-180): x2 bilinear, align_corners=False ------------
 * in [n,nc,h,w] (in_dtype F32|BF16) -> out F32 [n,nc,2h,2w].                                              */
int emrt_upsample2x(const void* in, float* out, int n, int nc, int h, int w, int in_dtype, void* stream);

/* ---- a6: slide_inference accumulation (src/api/infer.py:69-73) -----------------------------------------
 * canvas F32 [n_img,nc,H,W] += window logits F32 [n_win,nc,hc,wc] at (win_y0,win_x0) of image win_img;
 * count F32 [n_img,1,H,W] += 1.  Deterministic: every canvas pixel sums its covering windows in window-index
 * order (the reference's r-major, c order when windows are listed that way).  win_* are DEVICE int32 arrays. */
int emrt_window_accumulate(const float* win_logits, float* canvas, float* count, int n_win, int n_img,
                           int nc, int hc, int wc, int H, int W, const int32_t* win_img,
                           const int32_t* win_y0, const int32_t* win_x0, void* stream);

/* ---- a6 (divide) + a7: ss_inference tail (src/api/infer.py:75-79,150-154; predict.py:162-166) -----------
 * logits = canvas/count (count may be NULL => 1); optional bilinear resize to (Ho,Wo) (align_corners=False);
 * softmax(axis=1); argmax (first max) -> labels (I32 or U8) [n_img,1,Ho,Wo]; probs_out (F32, optional)
 * [n_img,nc,Ho,Wo]; logits_out (F32, optional, only when Ho==H && Wo==W) = canvas/count.                  */
int emrt_finalize_argmax(const float* canvas, const float* count, void* labels, int label_dtype,
                         float* probs_out, float* logits_out, int n_img, int nc, int H, int W, int Ho, int Wo,
                         void* stream);

/* ---- a5+a6+a7 fused: half-resolution window logits -> label map, no canvas ------------------------------
 * half_logits [n_win,nc,hc/2,wc/2] (in_dtype) are the class logits BEFORE UpHead's last x2 upsample.
 * For every pixel of every image the kernel upsamples each covering window on the fly, sums them in window
 * index order, divides by the cover count, and takes softmax+argmax.  win_* are DEVICE int32 arrays sorted by
 * (image, r, c) — the reference order.  labels [n_img,1,H,W] (I32|U8); logits_out optional F32 [n_img,nc,H,W]. */
int emrt_stitch_argmax_fused(const void* half_logits, int in_dtype, void* labels, int label_dtype,
                             float* logits_out, int n_win, int n_img, int nc, int hc, int wc, int H, int W,
                             const int32_t* win_img, const int32_t* win_y0, const int32_t* win_x0,
                             void* stream);

/* The same kernel with the evaluation fused behind the argmax:
 *   gt [n_img, H, W] (gt_dtype I32|U8) + areas I64 [n_img, 3, nc] (zeroed by the caller; accumulated):
 *     metrics.calculate_area (src/utils/metrics.py:20-69) per image — rows intersect / pred / label, pixels whose
 *     ground truth equals ignore_index are dropped from all three;
 *   palette U8 [nc, 3] + color U8 [n_img, H, W, 3]: predict.py:171-174's colour image (color = palette[class]).
 * Either pair may be NULL (not both).  Needs even H, W and nc <= 8 (EMRT_ERR_UNSUPPORTED otherwise: run
 * emrt_stitch_argmax_fused + emrt_calculate_area instead).                                                      */
int emrt_stitch_argmax_eval(const void* half_logits, int in_dtype, void* labels, int label_dtype, int n_win, int n_img,
                            int nc, int hc, int wc, int H, int W, const int32_t* win_img, const int32_t* win_y0,
                            const int32_t* win_x0, const void* gt, int gt_dtype, int ignore_index, long long* areas,
                            const uint8_t* palette, uint8_t* color, void* stream);

/* ---- a8: metrics.calculate_area (src/utils/metrics.py:20-69) --------------------------------------------
 * pred I32 [n], label I32 [n]; areas I64 [3*nc] = {intersect[nc], pred[nc], label[nc]} MUST be zeroed.     */
int emrt_calculate_area(const int32_t* pred, const int32_t* label, int64_t n, int nc, int ignore_index,
                        long long* areas, void* stream);


/* ---- a2 in one call: MSDeformableAttention.forward and its backward (transformer_encoder_decoder.py:65-107) -------------------
 * emrt_msda_fused_fwd composes, on `stream`: value_proj (+ value_mask, :83-86) -> [sampling_offsets | attention_weights]
 * projection with softmax over L*P and the location arithmetic (:89-102) -> sampling gather (:104, utils.py:64-97) ->
 * output_proj (:106), optionally followed by LayerNorm(residual + .) (:199-200) in the same epilogue.
 * dtype F32: the parity path — Paddle-layout fp32 weights w_* [in,out], SIMT GEMMs, normalised locations.
 * dtype BF16: the B200 path — packed bf16 [out,in] operands (emrt_pack_weight; wq_packed = [sampling_offsets ; attention_weights],
 *   b_query = their concatenated biases), tcgen05 GEMMs, fp16 pixel offsets / softmax weights, window-staged gather when
 *   flags carries EMRT_QUERY_PIXEL_GRID (queries = the pyramid's own pixels; window_center = the optional hint of
 *   emrt_msda_gather_fwd_hint).  with_pos_embed (:154-155): query_pos BF16 cyclic rows [query_pos_rows + 127, C] (see x2 of
 *   emrt_linear_args) or, better, row_bias F16 [Lq + 127, 3*MLP] = query_pos W_q + b_query; F32 path: query_pos F32
 *   [query_pos_rows, C] and query_scratch F32 [B, Lq, C] for the sum.
 * workspace (emrt_msda_fused_workspace_bytes) receives, 256-byte aligned and in this order, the tensors the backward needs:
 *   projected value [B, Lv, C] | offsets (BF16 path: F16 pixels; F32 path: F32 normalised locations) [B, Lq, 2*MLP] |
 *   softmax weights [B, Lq, MLP] | gathered tokens [B, Lq, C] | (F32 path) raw projections F32 [B, Lq, 3*MLP].
 * keep_pixel_major = 1 keeps the projected value [B, Lv, M, D] (required when emrt_msda_fused_bwd follows); otherwise the
 * BF16 path writes it head-major [B, M, Lv, D] for the specialised gathers.                                                   */
typedef struct emrt_msda_args {
  const void* query; const void* value; const float* ref; int32_t ref_batches;      /* ref F32 [ref_batches, Lq, L, 2] */
  const void* query_pos; int32_t query_pos_rows; void* query_scratch; const void* query_eff;
  const float* value_mask;                                                          /* F32 [B * Lv] or NULL */
  const void* w_value; const float* b_value; const void* w_offsets; const float* b_offsets;
  const void* w_attn; const float* b_attn; const void* w_out; const float* b_out;  /* Paddle layout [in, out] */
  const void* wv_packed; const void* wq_packed; const void* wo_packed; const float* b_query; const void* row_bias;
  const void* residual; const float* ln_gamma; const float* ln_beta; float ln_eps;
  void* out; void* workspace;
  int32_t B, Lq, Lv, C, M, L, P;
  int32_t shapes_hw[2 * EMRT_MAX_LEVELS];
  int32_t dtype; int32_t flags; int32_t keep_pixel_major;
  const int32_t* window_center;
  /* measurement hook (bench.py): when non-NULL, cudaEvent_t handles recorded on `stream` before / after the value projection
   * [0,1], the query projection [2,3], the gather [4,5] and the output projection [6,7]; equal neighbouring handles
   * ([1] == [2], ...) are recorded once                                                                                    */
  void* timing_events[8];
  /* optional cudaEvent_t recorded on `stream` right before the gather is launched: a second stream that waits for it starts its
   * work (the layer's 3x3 convolution, emrt_conv3x3_tokens_stats_part_fwd) together with the gather                          */
  void* gather_start_event;
} emrt_msda_args;
int64_t emrt_msda_fused_workspace_bytes(int B, int Lq, int Lv, int C, int M, int L, int P, int dtype);
int emrt_msda_fused_fwd(const emrt_msda_args* args, void* stream);

/* Backward of the same call (what Paddle autograd derives through :83-106): `args` as given to the forward (same workspace,
 * keep_pixel_major = 1, no residual; query_eff = the tensor the query projection actually multiplied, i.e. query + pos when
 * the caller formed it), plus the Paddle-layout weights in the activation dtype as the K-major operands of dx = dy W^T:
 * wq_cat [C, 3*MLP] = [sampling_offsets.weight | attention_weights.weight], w_value_cast, w_out_cast (NULL = args' w_*).
 * Outputs: d_query [B, Lq, C], d_value [B, Lv, C] (activation dtype), d_ref F32 [ref_batches, Lq, L, 2] or NULL; parameter
 * gradients F32, ACCUMULATED: dw_query [C, 3*MLP], db_query [3*MLP], dw_value / dw_out [C, C], db_value / db_out [C].
 * workspace: emrt_msda_fused_bwd_workspace_bytes.                                                                            */
typedef struct emrt_msda_grads {
  const void* d_out; void* d_query; void* d_value; float* d_ref;
  float* dw_query; float* db_query; float* dw_value; float* db_value; float* dw_out; float* db_out;
  const void* wq_cat; const void* w_value_cast; const void* w_out_cast;
  void* workspace;
} emrt_msda_grads;
int64_t emrt_msda_fused_bwd_workspace_bytes(int B, int Lq, int Lv, int C, int M, int L, int P, int dtype);
int emrt_msda_fused_bwd(const emrt_msda_args* args, const emrt_msda_grads* grads, void* stream);

/* ---- cfg 4: backward of the encoder / decoder glue (the reference trains the whole EncoderDecoder: train.py:146-159 drives
 * Paddle autograd through transformer_encoder_decoder.py:184-204,282-295 and layers.py:236-311).  Parameter gradients are
 * ACCUMULATED into fp32 tensors (the all-reduce buckets) and reduced in a fixed order: no floating-point atomics except the
 * split-K adds of the two tcgen05 weight-gradient kernels.  dtype F32|BF16 everywhere.
 *
 * LayerNorm backward (nn.LayerNorm, t_e_d.py:116,123,199-200,159-160): y = LN(a + b) * gamma + beta (b may be NULL);
 * dz [rows, N] (the gradient of a and of b), dgamma / dbeta F32 [N] accumulated.  N in {64,128,256,512}.
 * workspace: emrt_layernorm_bwd_workspace_floats(rows, N) + 2 * N floats.                                                */
int64_t emrt_layernorm_bwd_workspace_floats(int64_t rows, int N);
int emrt_layernorm_bwd(const void* a, const void* b, const float* gamma, const void* dy, void* dz, float* dgamma,
                       float* dbeta, float* workspace, int64_t rows, int N, float eps, int dtype, void* stream);

/* GroupNorm (+ exact GELU when gelu != 0) backward on tokens (conv{l}.1 + GELU, t_e_d.py:125-144,187-189; input_proj.{l}.1,
 * :417-419 with L = 1): x, dy, dx [B, Lv, C]; stats = emrt_groupnorm_stats' sums [B, L, groups, 2] of x; gamma / beta
 * F32 [L, C]; dgamma / dbeta F32 [L, C] accumulated.  workspace: emrt_groupnorm_bwd_workspace_floats floats.            */
int64_t emrt_groupnorm_bwd_workspace_floats(int B, int L, int C, int groups);
int emrt_groupnorm_bwd(const void* x, const void* dy, const float* stats, const float* gamma, const float* beta, void* dx,
                       float* dgamma, float* dbeta, float* workspace, int B, int Lv, int C, int L, int groups, float eps,
                       const int32_t* shapes_hw_host, int gelu, int dtype, void* stream);

/* dx = y > 0 ? dy : 0 (F.relu of the FFN, t_e_d.py:157,293); in place allowed.                                           */
int emrt_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, int dtype, void* stream);

/* out F32 [n] = sum_b x[b, :] (gradient of a batch-shared addend: with_pos_embed, t_e_d.py:154-155).                     */
int emrt_batch_sum(const void* x, float* out, int B, int64_t n, int dtype, void* stream);

/* out F32 [N] += column sums of x [rows, N] (bias / level_embed gradients).                                               */
int emrt_column_sum(const void* x, float* out, int64_t rows, int N, int dtype, void* stream);

/* F.sigmoid of the decoder's reference-point head (t_e_d.py:466) and its backward dx = dy * y * (1 - y); F32.            */
int emrt_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream);
int emrt_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);

/* Backward of emrt_mha_small (layers.py:282-301): dq / dk / dv with row strides (elements), same q / k / v views as the
 * forward, d_out [B, Lq, M*D] contiguous.  D = 32; one head's Q, K, V, dO and P must fit 200 KB of shared memory.        */
int emrt_mha_small_bwd(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld,
                       const void* d_out, void* dq, int64_t dq_ld, void* dk, int64_t dk_ld, void* dv, int64_t dv_ld, int B,
                       int Lq, int Lk, int M, int D, float scale, int dtype, void* stream);

/* Weight gradient of emrt_conv3x3_tokens_fwd: dw F32 [L, Cout, Cin, 3, 3] (Paddle Conv2D layout per level) +=
 * sum_{b, pixel} x[b, pixel + tap, ci] dy[b, pixel, co].  bf16: tcgen05 (the conv's shifted 4-D TMA boxes as the MN-major
 * A operand, split-K); otherwise / impl = 1: SIMT.  workspace: L * 9 * C * C floats.  (The data gradient is
 * emrt_conv3x3_tokens_fwd itself on dy with the flipped, transposed weights.)                                            */
int emrt_conv3x3_tokens_bwd_weight(const void* x, const void* dy, float* dw, float* workspace, int B, int Lv, int C, int L,
                                   const int32_t* shapes_hw_host, int dtype, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMRT_B200_H_ */
