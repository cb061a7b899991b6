// MSDeformableAttention.forward / its backward as ONE C call each (emrt_msda_fused_fwd / _bwd;
// transformer_encoder_decoder.py:65-107 is one call in the reference).  The entry points own the composition the Python mirror
// used to do — value projection, fused [sampling_offsets | attention_weights] projection with the softmax and the pixel-offset
// epilogue, the sampling gather, output projection (optionally with the residual + LayerNorm of :199-200 in its epilogue) —
// and the choice between the fp32 parity path and the bf16 B200 path, the window-staged gather and the row-bias form of
// with_pos_embed.  Intermediates live in a caller-provided workspace whose layout is part of the contract (they are what the
// backward needs: projected value, offsets, softmax weights, gathered tokens).
#include <cstring>

#include "common.cuh"

using namespace emrt;

namespace {

inline int64_t align256(int64_t n) { return (n + 255) / 256 * 256; }

struct FusedLayout {
  int64_t v, loc, attn, g, raw, total;     // byte offsets
};

// bf16 path: v bf16 [B*Lv, C], offsets f16 [rows, 2tp], weights f16 [rows, tp], gathered bf16 [rows, C]
// fp32 path: v f32, loc f32 [rows, 2tp] (normalised locations), weights f32 [rows, tp], gathered f32, + raw projections f32 [rows, 3tp]
FusedLayout fused_layout(int64_t B, int64_t Lq, int64_t Lv, int C, int tp, int dtype) {
  const int64_t rows = B * Lq;
  const int64_t sa = dtype == EMRT_BF16 ? 2 : 4, sl = dtype == EMRT_BF16 ? 2 : 4;
  FusedLayout f;
  f.v = 0;
  f.loc = f.v + align256(B * Lv * C * sa);
  f.attn = f.loc + align256(rows * 2 * tp * sl);
  f.g = f.attn + align256(rows * tp * sl);
  f.raw = f.g + align256(rows * C * sa);
  f.total = f.raw + (dtype == EMRT_BF16 ? 0 : align256(rows * 3 * tp * 4));
  return f;
}

int check_dims(const emrt_msda_args* a) {
  EMRT_REQUIRE(a && a->query && a->value && a->ref && a->out && a->workspace, "NULL pointer");
  EMRT_REQUIRE(a->B > 0 && a->Lq > 0 && a->Lv > 0 && a->C > 0 && a->M > 0 && a->L > 0 && a->L <= EMRT_MAX_LEVELS && a->P > 0,
               "non-positive dimension");
  EMRT_REQUIRE(a->C % a->M == 0, "embed_dim must be divisible by num_heads");                   // t_e_d.py:34
  int64_t acc = 0;
  for (int l = 0; l < a->L; ++l) acc += (int64_t)a->shapes_hw[2 * l] * a->shapes_hw[2 * l + 1];
  EMRT_REQUIRE(acc == a->Lv, "sum(H*W) != Len_v");                                              // t_e_d.py:81
  EMRT_REQUIRE(a->dtype == EMRT_F32 || a->dtype == EMRT_BF16, "dtype must be F32 or BF16");
  return EMRT_OK;
}

void level_start(const emrt_msda_args* a, int32_t* start) {
  int32_t acc = 0;
  for (int l = 0; l < a->L; ++l) { start[l] = acc; acc += a->shapes_hw[2 * l] * a->shapes_hw[2 * l + 1]; }
}

// bench.py's per-kernel CUDA events (emrt_msda_args.timing_events), recorded on the launching stream
struct Tick {
  const emrt_msda_args* a; cudaStream_t st;
  // (the end of one sub-launch and the start of the next may be the SAME event handle: recorded once)
  void operator()(int i) const {
    if (a->timing_events[i] && (i == 0 || a->timing_events[i] != a->timing_events[i - 1]))
      cudaEventRecord(reinterpret_cast<cudaEvent_t>(a->timing_events[i]), st);
  }
};

emrt_linear_args lin(const void* x, const void* w, const float* b, void* y, int64_t rows, int K, int N, int xd, int wd, int yd,
                     int wt, int impl) {
  emrt_linear_args l;
  memset(&l, 0, sizeof(l));
  l.x = x; l.w = w; l.bias = b; l.y = y; l.rows = rows; l.K = K; l.N = N;
  l.x_dtype = xd; l.w_dtype = wd; l.y_dtype = yd; l.w_transposed = wt; l.impl = impl;
  return l;
}

}  // namespace

extern "C" int64_t emrt_msda_fused_workspace_bytes(int B, int Lq, int Lv, int C, int M, int L, int P, int dtype) {
  return fused_layout(B, Lq, Lv, C, M * L * P, dtype).total;
}

extern "C" int emrt_msda_fused_fwd(const emrt_msda_args* a, void* stream) {
  if (int e = check_dims(a)) return e;
  const int tp = a->M * a->L * a->P, D = a->C / a->M;
  const int64_t rows = (int64_t)a->B * a->Lq, vrows = (int64_t)a->B * a->Lv;
  const FusedLayout f = fused_layout(a->B, a->Lq, a->Lv, a->C, tp, a->dtype);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  void *v = ws + f.v, *loc = ws + f.loc, *attn = ws + f.attn, *g = ws + f.g;
  int32_t start[EMRT_MAX_LEVELS];
  level_start(a, start);
  const int64_t rbs = a->ref_batches == 1 ? 0 : (int64_t)a->Lq * a->L * 2;
  const bool ln = a->residual != nullptr;
  if (ln) EMRT_REQUIRE(a->ln_gamma && a->ln_beta, "residual given without ln_gamma / ln_beta");

  if (a->dtype == EMRT_F32) {
    // ---- parity path: Paddle-layout fp32 weights, SIMT GEMMs, normalised locations (t_e_d.py:83-106 step by step) ----
    EMRT_REQUIRE(a->w_value && a->w_offsets && a->w_attn && a->w_out, "fp32 path needs the four Paddle-layout weights");
    const void* q = a->query;
    if (a->query_pos) {                       // with_pos_embed (:154-155): q = query + pos, written to the raw buffer's tail
      EMRT_REQUIRE(a->query_scratch, "fp32 path with query_pos needs query_scratch [B, Lq, C]");
      if (int e = emrt_add_bcast(a->query, a->query_pos, a->query_scratch, rows * a->C, (int64_t)a->query_pos_rows * a->C, EMRT_F32, stream)) return e;
      q = a->query_scratch;
    }
    emrt_linear_args l = lin(a->value, a->w_value, a->b_value, v, vrows, a->C, a->C, EMRT_F32, EMRT_F32, EMRT_F32, 0, 1);
    if (a->value_mask) { l.epilogue = EMRT_EPI_ROW_MASK; l.row_scale = a->value_mask; }
    if (int e = emrt_linear_fwd(&l, stream)) return e;
    float* raw = reinterpret_cast<float*>(ws + f.raw);
    l = lin(q, a->w_offsets, a->b_offsets, raw, rows, a->C, 2 * tp, EMRT_F32, EMRT_F32, EMRT_F32, 0, 1);
    if (int e = emrt_linear_fwd(&l, stream)) return e;
    l = lin(q, a->w_attn, a->b_attn, raw + rows * 2 * tp, rows, a->C, tp, EMRT_F32, EMRT_F32, EMRT_F32, 0, 1);
    if (int e = emrt_linear_fwd(&l, stream)) return e;
    if (int e = emrt_msda_softmax_loc(raw, 2 * tp, raw + rows * 2 * tp, tp, a->ref, rbs, loc, attn, a->B, a->Lq, a->M, a->L, a->P,
                                      a->shapes_hw, EMRT_F32, EMRT_LOC_NORMALIZED, stream)) return e;
    if (int e = emrt_msda_gather_fwd(v, loc, attn, nullptr, 0, g, a->B, a->Lq, a->Lv, a->M, D, a->L, a->P, a->shapes_hw, start,
                                     EMRT_F32, EMRT_F32, EMRT_LOC_NORMALIZED, stream)) return e;
    l = lin(g, a->w_out, a->b_out, a->out, rows, a->C, a->C, EMRT_F32, EMRT_F32, EMRT_F32, 0, 1);
    if (int e = emrt_linear_fwd(&l, stream)) return e;
    if (ln) return emrt_residual_layernorm(a->out, a->residual, a->ln_gamma, a->ln_beta, nullptr, a->out, rows, a->C, a->ln_eps, EMRT_F32, stream);
    return EMRT_OK;
  }

  // ---- B200 path: packed bf16 [out, in] operands, tcgen05 GEMMs, fp16 pixel offsets / softmax weights ----------------
  EMRT_REQUIRE(a->wv_packed && a->wq_packed && a->wo_packed, "bf16 path needs the packed operands (emrt_pack_weight)");
  const bool fusedq = tp == 144 && a->L * a->P == 18 && a->C % 8 == 0;             // the MSDA_QPROJ epilogue's configuration
  const bool special = D == 32 && a->L == 3 && a->P == 6;
  const bool head_major = special && !a->keep_pixel_major;
  const bool grid = (a->flags & EMRT_QUERY_PIXEL_GRID) && a->Lq == a->Lv && special;
  emrt_linear_args l = lin(a->value, a->wv_packed, a->b_value, v, vrows, a->C, a->C, EMRT_BF16, EMRT_BF16, EMRT_BF16, 1, 0);
  if (a->value_mask) { l.epilogue |= EMRT_EPI_ROW_MASK; l.row_scale = a->value_mask; }
  if (head_major) { l.epilogue |= EMRT_EPI_HEAD_MAJOR; l.hm_rows = a->Lv; l.hm_D = D; }
  const Tick tick{a, as_stream(stream)};
  tick(0);
  if (int e = emrt_linear_fwd(&l, stream)) return e;
  tick(1); tick(2);
  if (fusedq) {
    l = lin(a->query, a->wq_packed, a->row_bias ? nullptr : a->b_query, loc, rows, a->C, 3 * tp, EMRT_BF16, EMRT_BF16, EMRT_F16, 1, 0);
    l.epilogue = EMRT_EPI_MSDA_QPROJ; l.qproj_group = a->L * a->P; l.y2 = attn;
    if (a->row_bias) { l.row_bias = a->row_bias; l.row_bias_period = a->Lq; }
    else if (a->query_pos) { l.x2 = a->query_pos; l.x2_period = a->query_pos_rows; }
    if (int e = emrt_linear_fwd(&l, stream)) return e;
  } else {
    // generic point counts: raw fp32 projection (+ folded position embedding), then softmax + offsets
    EMRT_REQUIRE(a->query_scratch, "this (levels, points) configuration needs query_scratch for the raw projections [rows, 3*MLP] f32");
    l = lin(a->query, a->wq_packed, a->b_query, a->query_scratch, rows, a->C, 3 * tp, EMRT_BF16, EMRT_BF16, EMRT_F32, 1, 0);
    if (a->query_pos) { l.x2 = a->query_pos; l.x2_period = a->query_pos_rows; }
    if (int e = emrt_linear_fwd(&l, stream)) return e;
    const float* raw = reinterpret_cast<const float*>(a->query_scratch);
    if (int e = emrt_msda_softmax_loc(raw, 3 * tp, raw + 2 * tp, 3 * tp, nullptr, 0, loc, attn, a->B, a->Lq, a->M, a->L, a->P, a->shapes_hw,
                                      EMRT_F16, EMRT_LOC_PIXEL_OFFSET, stream)) return e;
  }
  const int mode = EMRT_LOC_PIXEL_OFFSET | (head_major ? EMRT_VALUE_HEAD_MAJOR : 0) | (grid ? EMRT_QUERY_PIXEL_GRID : 0);
  tick(3);
  if (a->gather_start_event) EMRT_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(a->gather_start_event), as_stream(stream)));
  tick(4);
  if (int e = emrt_msda_gather_fwd_hint(v, loc, attn, a->ref, rbs, g, a->B, a->Lq, a->Lv, a->M, D, a->L, a->P, a->shapes_hw, start, EMRT_BF16,
                                        EMRT_F16, mode, grid ? a->window_center : nullptr, stream)) return e;
  tick(5); tick(6);
  l = lin(g, a->wo_packed, a->b_out, a->out, rows, a->C, a->C, EMRT_BF16, EMRT_BF16, EMRT_BF16, 1, 0);
  if (ln && a->C == 256) {
    l.epilogue = EMRT_EPI_RESIDUAL_LN; l.residual = a->residual; l.ln_gamma = a->ln_gamma; l.ln_beta = a->ln_beta; l.ln_eps = a->ln_eps;
    const int e = emrt_linear_fwd(&l, stream);
    tick(7);
    return e;
  }
  if (int e = emrt_linear_fwd(&l, stream)) return e;
  tick(7);
  if (ln) return emrt_residual_layernorm(a->out, a->residual, a->ln_gamma, a->ln_beta, nullptr, a->out, rows, a->C, a->ln_eps, EMRT_BF16, stream);
  return EMRT_OK;
}

extern "C" int64_t emrt_msda_fused_bwd_workspace_bytes(int B, int Lq, int Lv, int C, int M, int L, int P, int dtype) {
  const int64_t rows = (int64_t)B * Lq, tp = (int64_t)M * L * P, s = dtype == EMRT_BF16 ? 2 : 4;
  // d_g [rows, C] | grad_value f32 [B*Lv, C] | grad_loc f32 [rows, 2tp] | grad_attn f32 [rows, tp] | dq [rows, 3tp] | d_v [B*Lv, C]
  return align256(rows * C * s) + align256((int64_t)B * Lv * C * 4) + align256(rows * 2 * tp * 4) + align256(rows * tp * 4) +
         align256(rows * 3 * tp * s) + align256((int64_t)B * Lv * C * s);
}

extern "C" int emrt_msda_fused_bwd(const emrt_msda_args* a, const emrt_msda_grads* gr, void* stream) {
  if (int e = check_dims(a)) return e;
  EMRT_REQUIRE(gr && gr->d_out && gr->workspace && gr->d_query && gr->d_value && gr->dw_query && gr->dw_value && gr->dw_out,
               "NULL pointer in emrt_msda_grads");
  EMRT_REQUIRE(a->w_value && a->w_offsets && a->w_attn && a->w_out, "the backward needs the Paddle-layout weights (K-major operands of dx = dy W^T)");
  EMRT_REQUIRE(!a->residual, "the backward differentiates the plain module (no residual + LayerNorm epilogue)");
  const int tp = a->M * a->L * a->P, D = a->C / a->M, C = a->C;
  const int64_t rows = (int64_t)a->B * a->Lq, vrows = (int64_t)a->B * a->Lv;
  const int dt = a->dtype, s = dt == EMRT_BF16 ? 2 : 4;
  const FusedLayout f = fused_layout(a->B, a->Lq, a->Lv, C, tp, dt);
  const uint8_t* ws = reinterpret_cast<const uint8_t*>(a->workspace);
  const void *v = ws + f.v, *loc = ws + f.loc, *attn = ws + f.attn, *g = ws + f.g;
  uint8_t* bw = reinterpret_cast<uint8_t*>(gr->workspace);
  void* d_g = bw; bw += align256(rows * C * s);
  float* gv = reinterpret_cast<float*>(bw); bw += align256(vrows * C * 4);
  float* gl = reinterpret_cast<float*>(bw); bw += align256(rows * 2 * tp * 4);
  float* ga = reinterpret_cast<float*>(bw); bw += align256(rows * tp * 4);
  void* dq = bw; bw += align256(rows * 3 * tp * s);
  void* d_v = bw;
  int32_t start[EMRT_MAX_LEVELS];
  level_start(a, start);
  const int64_t rbs = a->ref_batches == 1 ? 0 : (int64_t)a->Lq * a->L * 2;
  cudaStream_t st = as_stream(stream);
  const int impl = dt == EMRT_BF16 ? 0 : 1;
  const int ldt = dt == EMRT_BF16 ? EMRT_F16 : EMRT_F32;
  const int mode = dt == EMRT_BF16 ? EMRT_LOC_PIXEL_OFFSET : EMRT_LOC_NORMALIZED;
  // the training forward keeps the projected value pixel-major (the layout the gather backward scatters into)
  EMRT_REQUIRE(dt == EMRT_F32 || a->keep_pixel_major, "run the forward with keep_pixel_major = 1 when a backward follows");
  const bool grid = dt == EMRT_BF16 && (a->flags & EMRT_QUERY_PIXEL_GRID) && a->Lq == a->Lv && D == 32 && a->L == 3 && a->P == 6;

  // output_proj: d_g = d_out W_out^T;  dW_out += g^T d_out
  emrt_linear_args l = lin(gr->d_out, gr->w_out_cast ? gr->w_out_cast : a->w_out, nullptr, d_g, rows, C, C, dt, dt, dt, 1, impl);
  if (int e = emrt_linear_fwd(&l, stream)) return e;
  if (int e = emrt_linear_bwd_weight(g, gr->d_out, gr->dw_out, gr->db_out, rows, C, C, dt, dt, stream)) return e;
  // gather
  EMRT_CUDA_CHECK(cudaMemsetAsync(gv, 0, sizeof(float) * vrows * C, st));
  if (int e = emrt_msda_gather_bwd_hint(d_g, v, loc, attn, mode == EMRT_LOC_PIXEL_OFFSET ? a->ref : nullptr, rbs, gv, gl, ga, a->B, a->Lq, a->Lv,
                                        a->M, D, a->L, a->P, a->shapes_hw, start, dt, ldt, mode | (grid ? EMRT_QUERY_PIXEL_GRID : 0),
                                        grid ? a->window_center : nullptr, stream)) return e;
  if (gr->d_ref)
    if (int e = emrt_msda_ref_bwd(gl, gr->d_ref, a->B, a->ref_batches, a->Lq, a->M, a->L, a->P, a->shapes_hw, mode, stream)) return e;
  // softmax + offsets -> the fused query projection: d_query = dq [W_off | W_attn]^T;  dW_q += query^T dq
  if (int e = emrt_msda_qproj_bwd(gl, ga, attn, dq, rows, a->M, a->L, a->P, a->shapes_hw, ldt, dt, mode, stream)) return e;
  EMRT_REQUIRE(gr->wq_cat, "emrt_msda_grads.wq_cat ([C, 3*MLP] = [sampling_offsets.weight | attention_weights.weight] in the activation dtype) is required");
  l = lin(dq, gr->wq_cat, nullptr, gr->d_query, rows, 3 * tp, C, dt, dt, dt, 1, impl);
  if (int e = emrt_linear_fwd(&l, stream)) return e;
  if (int e = emrt_linear_bwd_weight(a->query_eff ? a->query_eff : a->query, dq, gr->dw_query, gr->db_query, rows, C, 3 * tp, dt, dt, stream)) return e;
  // value_proj (mask backward folded into the cast of the fp32 scatter buffer)
  if (int e = emrt_scale_rows_cast(gv, a->value_mask, d_v, vrows, C, dt, stream)) return e;
  l = lin(d_v, gr->w_value_cast ? gr->w_value_cast : a->w_value, nullptr, gr->d_value, vrows, C, C, dt, dt, dt, 1, impl);
  if (int e = emrt_linear_fwd(&l, stream)) return e;
  return emrt_linear_bwd_weight(a->value, d_v, gr->dw_value, gr->db_value, vrows, C, C, dt, dt, stream);
}
