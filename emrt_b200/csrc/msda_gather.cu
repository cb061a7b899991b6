// Multiscale bilinear sampling-gather, forward and backward.
// Replaces deformable_attention_core_func, src/models/EMRT_utils/utils.py:64-97 (3x F.grid_sample + stack +
// multiply + sum) and the autograd Paddle derives through it, with one kernel each.
//
// Layout: value [B,Lv,M,D] (pixel-major, the reference's own layout after the reshape at
// transformer_encoder_decoder.py:88); loc [B,Lq,M,L,P,2]; attn [B,Lq,M,L,P]; out [B,Lq,M*D].
// Mapping: one (b,q,m) item per group of G = D/VEC lanes, each lane owning one 16-byte channel slice
// (bf16x8 / f32x4) of all four bilinear corners, so every corner fetch is one LDG.128 and a warp's 32 lanes
// cover 32/G items.  Corners outside the map get weight 0 and a clamped (always valid, nearby) address: no
// divergent branches, loads of consecutive points can overlap.
#include <cstdlib>

#include "msda_common.cuh"

namespace emrt {

template <typename TV, typename TL, int MODE, int D>
__global__ void __launch_bounds__(256)
msda_gather_fwd_kernel(const TV* __restrict__ value, const TL* __restrict__ loc, const TL* __restrict__ attn,
                       const float* __restrict__ ref, int64_t ref_bs, TV* __restrict__ out, int Lq, int Lv, int M,
                       int L, int P, const __grid_constant__ LevelTable lv, int64_t n_items) {
  constexpr int VEC = Vec16<TV>::N;
  constexpr int G = D / VEC;
  static_assert(D % VEC == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "unsupported head dim");
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t item = tid / G;
  if (item >= n_items) return;
  const int sub = (int)(tid % G);
  const int m = (int)(item % M);
  const int64_t bq = item / M;
  const int q = (int)(bq % Lq);
  const int64_t b = bq / Lq;

  const int64_t pix_stride = (int64_t)M * D;
  const TV* vbase = value + (b * Lv * M + m) * D + sub * VEC;
  const TL* lp = loc + item * (int64_t)L * P * 2;
  const TL* ap = attn + item * (int64_t)L * P;
  const float* rp = (MODE == EMRT_LOC_PIXEL_OFFSET) ? ref + b * ref_bs + (int64_t)q * L * 2 : nullptr;

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const TV* vl = vbase + (int64_t)lv.start[l] * pix_stride;
    float rx = 0.f, ry = 0.f;
    if (MODE == EMRT_LOC_PIXEL_OFFSET) {
      const float2 r = __ldg(reinterpret_cast<const float2*>(rp + 2 * l));
      rx = r.x * (float)W - 0.5f;
      ry = r.y * (float)H - 0.5f;
    }
#pragma unroll 2
    for (int p = 0; p < P; ++p) {
      const float2 xy = Pair<TL>::load(lp + (l * P + p) * 2);
      const float aw = load1<TL>(ap + l * P + p);
      float x, y;
      if (MODE == EMRT_LOC_PIXEL_OFFSET) { x = rx + xy.x; y = ry + xy.y; }
      else { x = xy.x * (float)W - 0.5f; y = xy.y * (float)H - 0.5f; }
      const Footprint f = make_footprint(x, y, H, W);
      float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
      Vec16<TV>::load(vl + f.i00 * pix_stride, v00);
      Vec16<TV>::load(vl + f.i01 * pix_stride, v01);
      Vec16<TV>::load(vl + f.i10 * pix_stride, v10);
      Vec16<TV>::load(vl + f.i11 * pix_stride, v11);
      const float a00 = f.w00 * aw, a01 = f.w01 * aw, a10 = f.w10 * aw, a11 = f.w11 * aw;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        acc[i] = fmaf(a00, v00[i], acc[i]);
        acc[i] = fmaf(a01, v01[i], acc[i]);
        acc[i] = fmaf(a10, v10[i], acc[i]);
        acc[i] = fmaf(a11, v11[i], acc[i]);
      }
    }
  }
  Vec16<TV>::store(out + item * D + sub * VEC, acc);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int VEC>
__device__ __forceinline__ void scatter_add(float* p, float w, const float (&g)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i += 4) red_add_v4(p + i, w * g[i], w * g[i + 1], w * g[i + 2], w * g[i + 3]);
}

template <int VEC>
__device__ __forceinline__ float dot(const float (&a)[VEC], const float (&b)[VEC]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s = fmaf(a[i], b[i], s);
  return s;
}

// Backward: grad_value via vectorised REDG (red.global.add.v4.f32, 16 B per atomic, no return value);
// grad_loc / grad_attn are reduced over the G lanes of an item with warp shuffles and written once
// (no atomics).
template <typename TV, typename TL, int MODE, int D>
__global__ void __launch_bounds__(256)
msda_gather_bwd_kernel(const TV* __restrict__ grad_out, const TV* __restrict__ value, const TL* __restrict__ loc,
                       const TL* __restrict__ attn, const float* __restrict__ ref, int64_t ref_bs,
                       float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn,
                       int Lq, int Lv, int M, int L, int P, const __grid_constant__ LevelTable lv, int64_t n_items) {
  constexpr int VEC = Vec16<TV>::N;
  constexpr int G = D / VEC;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t item = tid / G;
  const bool active = item < n_items;      // keep whole warps alive for the shuffles
  if (!active) item = n_items - 1;
  const int sub = (int)(tid % G);
  const int m = (int)(item % M);
  const int64_t bq = item / M;
  const int q = (int)(bq % Lq);
  const int64_t b = bq / Lq;

  const int64_t pix_stride = (int64_t)M * D;
  const int64_t voff = (b * Lv * M + m) * D + sub * VEC;
  const TL* lp = loc + item * (int64_t)L * P * 2;
  const TL* ap = attn + item * (int64_t)L * P;
  const float* rp = (MODE == EMRT_LOC_PIXEL_OFFSET) ? ref + b * ref_bs + (int64_t)q * L * 2 : nullptr;

  float g[VEC];
  Vec16<TV>::load(grad_out + item * D + sub * VEC, g);

  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const int64_t lbase = voff + (int64_t)lv.start[l] * pix_stride;
    float rx = 0.f, ry = 0.f;
    if (MODE == EMRT_LOC_PIXEL_OFFSET) {
      const float2 r = __ldg(reinterpret_cast<const float2*>(rp + 2 * l));
      rx = r.x * (float)W - 0.5f;
      ry = r.y * (float)H - 0.5f;
    }
    const float sx = (MODE == EMRT_LOC_PIXEL_OFFSET) ? 1.f : (float)W;
    const float sy = (MODE == EMRT_LOC_PIXEL_OFFSET) ? 1.f : (float)H;
    for (int p = 0; p < P; ++p) {
      const float2 xy = Pair<TL>::load(lp + (l * P + p) * 2);
      const float aw = load1<TL>(ap + l * P + p);
      float x, y;
      if (MODE == EMRT_LOC_PIXEL_OFFSET) { x = rx + xy.x; y = ry + xy.y; }
      else { x = xy.x * (float)W - 0.5f; y = xy.y * (float)H - 0.5f; }
      const Footprint f = make_footprint(x, y, H, W);
      float v[VEC];
      float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
      if (f.v00) { Vec16<TV>::load(value + lbase + f.i00 * pix_stride, v); d00 = dot<VEC>(v, g);
                   if (active) scatter_add<VEC>(grad_value + lbase + f.i00 * pix_stride, f.w00 * aw, g); }
      if (f.v01) { Vec16<TV>::load(value + lbase + f.i01 * pix_stride, v); d01 = dot<VEC>(v, g);
                   if (active) scatter_add<VEC>(grad_value + lbase + f.i01 * pix_stride, f.w01 * aw, g); }
      if (f.v10) { Vec16<TV>::load(value + lbase + f.i10 * pix_stride, v); d10 = dot<VEC>(v, g);
                   if (active) scatter_add<VEC>(grad_value + lbase + f.i10 * pix_stride, f.w10 * aw, g); }
      if (f.v11) { Vec16<TV>::load(value + lbase + f.i11 * pix_stride, v); d11 = dot<VEC>(v, g);
                   if (active) scatter_add<VEC>(grad_value + lbase + f.i11 * pix_stride, f.w11 * aw, g); }
      float ga = f.w00 * d00 + f.w01 * d01 + f.w10 * d10 + f.w11 * d11;
      float gx = aw * sx * ((1.f - f.fy) * (d01 - d00) + f.fy * (d11 - d10));
      float gy = aw * sy * ((1.f - f.fx) * (d10 - d00) + f.fx * (d11 - d01));
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
      }
      if (active && sub == 0) {
        const int64_t pi = item * (int64_t)L * P + l * P + p;
        grad_attn[pi] = ga;
        *reinterpret_cast<float2*>(grad_loc + pi * 2) = make_float2(gx, gy);
      }
    }
  }
}

template <typename TV, typename TL, int MODE, int D>
static int launch_fwd(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs,
                      void* out, int Lq, int Lv, int M, int L, int P, const LevelTable& lv, int64_t n_items,
                      cudaStream_t st) {
  constexpr int G = D / Vec16<TV>::N;
  const int64_t threads = n_items * G;
  const int64_t blocks = (threads + 255) / 256;
  msda_gather_fwd_kernel<TV, TL, MODE, D><<<(unsigned)blocks, 256, 0, st>>>(
      (const TV*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, (TV*)out, Lq, Lv, M, L, P, lv, n_items);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

template <typename TV, typename TL, int MODE, int D>
static int launch_bwd(const void* go, const void* value, const void* loc, const void* attn, const float* ref,
                      int64_t ref_bs, float* gv, float* gl, float* ga, int Lq, int Lv, int M, int L, int P,
                      const LevelTable& lv, int64_t n_items, cudaStream_t st) {
  constexpr int G = D / Vec16<TV>::N;
  const int64_t threads = n_items * G;
  const int64_t blocks = (threads + 255) / 256;
  msda_gather_bwd_kernel<TV, TL, MODE, D><<<(unsigned)blocks, 256, 0, st>>>(
      (const TV*)go, (const TV*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, gv, gl, ga, Lq, Lv, M, L, P,
      lv, n_items);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

#define EMRT_DISPATCH_D(TV, TL, MODE, CALL, ...)                                               \
  switch (D) {                                                                                 \
    case 16: return CALL<TV, TL, MODE, 16>(__VA_ARGS__);                                       \
    case 32: return CALL<TV, TL, MODE, 32>(__VA_ARGS__);                                       \
    case 64: return CALL<TV, TL, MODE, 64>(__VA_ARGS__);                                       \
    default: return set_error(EMRT_ERR_UNSUPPORTED, "head dim D=%d not in {16,32,64}", D);     \
  }

}  // namespace emrt

using namespace emrt;

static int check_common(const void* value, const void* loc, const void* attn, const float* ref, int B, int Lq,
                        int Lv, int M, int D, int L, int P, int value_dtype, int loc_dtype, int mode) {
  EMRT_REQUIRE(value && loc && attn, "NULL tensor pointer");
  EMRT_REQUIRE(B > 0 && Lq > 0 && Lv > 0 && M > 0 && D > 0 && P > 0, "non-positive dimension");
  EMRT_REQUIRE(value_dtype == EMRT_F32 || value_dtype == EMRT_BF16, "value_dtype must be F32 or BF16");
  EMRT_REQUIRE(loc_dtype == EMRT_F32 || loc_dtype == EMRT_F16 || loc_dtype == EMRT_BF16, "bad loc_dtype");
  EMRT_REQUIRE((mode & ~(EMRT_LOC_PIXEL_OFFSET | EMRT_VALUE_HEAD_MAJOR | EMRT_QUERY_PIXEL_GRID)) == 0, "bad mode flags");
  EMRT_REQUIRE(!(mode & EMRT_LOC_PIXEL_OFFSET) || ref != nullptr, "PIXEL_OFFSET mode needs reference points");
  (void)L;
  return EMRT_OK;
}

#define EMRT_GATHER_DISPATCH(LAUNCH, ...)                                                                      \
  do {                                                                                                         \
    if (mode & EMRT_VALUE_HEAD_MAJOR)                                                                          \
      return set_error(EMRT_ERR_UNSUPPORTED, "head-major value layout needs bf16, D=32, L=3, P=6 (forward)");   \
    const bool px = (mode & EMRT_LOC_PIXEL_OFFSET) != 0;                                                       \
    if (value_dtype == EMRT_F32) {                                                                             \
      if (loc_dtype != EMRT_F32) return set_error(EMRT_ERR_UNSUPPORTED, "F32 value needs F32 loc/attn");       \
      if (px) { EMRT_DISPATCH_D(float, float, 1, LAUNCH, __VA_ARGS__) }                                        \
      else    { EMRT_DISPATCH_D(float, float, 0, LAUNCH, __VA_ARGS__) }                                        \
    } else if (loc_dtype == EMRT_F32) {                                                                        \
      if (px) { EMRT_DISPATCH_D(__nv_bfloat16, float, 1, LAUNCH, __VA_ARGS__) }                                \
      else    { EMRT_DISPATCH_D(__nv_bfloat16, float, 0, LAUNCH, __VA_ARGS__) }                                \
    } else if (loc_dtype == EMRT_F16) {                                                                        \
      if (px) { EMRT_DISPATCH_D(__nv_bfloat16, __half, 1, LAUNCH, __VA_ARGS__) }                               \
      else    { EMRT_DISPATCH_D(__nv_bfloat16, __half, 0, LAUNCH, __VA_ARGS__) }                               \
    } else {                                                                                                   \
      if (px) { EMRT_DISPATCH_D(__nv_bfloat16, __nv_bfloat16, 1, LAUNCH, __VA_ARGS__) }                        \
      else    { EMRT_DISPATCH_D(__nv_bfloat16, __nv_bfloat16, 0, LAUNCH, __VA_ARGS__) }                        \
    }                                                                                                          \
  } while (0)

extern "C" int emrt_msda_gather_fwd_hint(const void* value, const void* loc, const void* attn, const float* ref,
                                         int64_t ref_batch_stride, void* out, int B, int Lq, int Lv, int M, int D,
                                         int L, int P, const int32_t* shapes_hw_host,
                                         const int32_t* level_start_host, int value_dtype, int loc_dtype, int mode,
                                         const int32_t* window_center_host, void* stream) {
  if (int e = check_common(value, loc, attn, ref, B, Lq, Lv, M, D, L, P, value_dtype, loc_dtype, mode)) return e;
  EMRT_REQUIRE(out != nullptr, "out is NULL");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, level_start_host, Lv)) return e;
  const int64_t n_items = (int64_t)B * Lq * M;
  cudaStream_t st = as_stream(stream);
  if (value_dtype == EMRT_BF16 && (mode & EMRT_QUERY_PIXEL_GRID) && !getenv("EMRT_GATHER_NO_WIN")) {
    const int e = gather_fwd_win(value, loc, attn, ref, ref_batch_stride, out, B, Lq, Lv, M, D, L, P, lv, loc_dtype, mode,
                                 window_center_host, st);
    if (e != EMRT_ERR_UNSUPPORTED) return e;
  }
  mode &= ~EMRT_QUERY_PIXEL_GRID;
  if (value_dtype == EMRT_BF16 && !getenv("EMRT_GATHER_V0")) {
    const int e = gather_fwd_v1(value, loc, attn, ref, ref_batch_stride, out, B, Lq, Lv, M, D, L, P, lv, loc_dtype, mode, st);
    if (e != EMRT_ERR_UNSUPPORTED) return e;
  }
  EMRT_GATHER_DISPATCH(launch_fwd, value, loc, attn, ref, ref_batch_stride, out, Lq, Lv, M, L, P, lv, n_items, st);
}

extern "C" int emrt_msda_gather_fwd(const void* value, const void* loc, const void* attn, const float* ref,
                                    int64_t ref_batch_stride, void* out, int B, int Lq, int Lv, int M, int D,
                                    int L, int P, const int32_t* shapes_hw_host, const int32_t* level_start_host,
                                    int value_dtype, int loc_dtype, int mode, void* stream) {
  return emrt_msda_gather_fwd_hint(value, loc, attn, ref, ref_batch_stride, out, B, Lq, Lv, M, D, L, P, shapes_hw_host,
                                   level_start_host, value_dtype, loc_dtype, mode, nullptr, stream);
}

extern "C" int emrt_msda_gather_bwd_hint(const void* grad_out, const void* value, const void* loc, const void* attn,
                                         const float* ref, int64_t ref_batch_stride, float* grad_value,
                                         float* grad_loc, float* grad_attn, int B, int Lq, int Lv, int M, int D, int L,
                                         int P, const int32_t* shapes_hw_host, const int32_t* level_start_host,
                                         int value_dtype, int loc_dtype, int mode, const int32_t* window_center_host,
                                         void* stream) {
  if (int e = check_common(value, loc, attn, ref, B, Lq, Lv, M, D, L, P, value_dtype, loc_dtype, mode)) return e;
  EMRT_REQUIRE(grad_out && grad_value && grad_loc && grad_attn, "NULL gradient pointer");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, level_start_host, Lv)) return e;
  const int64_t n_items = (int64_t)B * Lq * M;
  cudaStream_t st = as_stream(stream);
  if (value_dtype == EMRT_BF16 && (mode & EMRT_QUERY_PIXEL_GRID) && !(mode & EMRT_VALUE_HEAD_MAJOR) &&
      !getenv("EMRT_GATHER_NO_WIN")) {
    const int e = gather_bwd_win(grad_out, value, loc, attn, ref, ref_batch_stride, grad_value, grad_loc, grad_attn, B, Lq,
                                 Lv, M, D, L, P, lv, loc_dtype, mode, window_center_host, st);
    if (e != EMRT_ERR_UNSUPPORTED) return e;
  }
  mode &= ~EMRT_QUERY_PIXEL_GRID;
  EMRT_GATHER_DISPATCH(launch_bwd, grad_out, value, loc, attn, ref, ref_batch_stride, grad_value, grad_loc,
                       grad_attn, Lq, Lv, M, L, P, lv, n_items, st);
}

extern "C" int emrt_msda_gather_bwd(const void* grad_out, const void* value, const void* loc, const void* attn,
                                    const float* ref, int64_t ref_batch_stride, float* grad_value, float* grad_loc,
                                    float* grad_attn, int B, int Lq, int Lv, int M, int D, int L, int P,
                                    const int32_t* shapes_hw_host, const int32_t* level_start_host,
                                    int value_dtype, int loc_dtype, int mode, void* stream) {
  return emrt_msda_gather_bwd_hint(grad_out, value, loc, attn, ref, ref_batch_stride, grad_value, grad_loc, grad_attn, B,
                                   Lq, Lv, M, D, L, P, shapes_hw_host, level_start_host, value_dtype, loc_dtype, mode,
                                   nullptr, stream);
}
