// Pieces shared by the gather kernels: 16-bit pair loads and the bilinear footprint
// (grid_sample(bilinear, zeros, align_corners=False) semantics, src/models/EMRT_utils/utils.py:87-88).
#pragma once
#include "common.cuh"

namespace emrt {

template <typename TL> struct Pair;
template <> struct Pair<float> {
  __device__ static __forceinline__ float2 load(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
};
template <> struct Pair<__half> {
  __device__ static __forceinline__ float2 load(const __half* p) {
    unsigned int r = __ldg(reinterpret_cast<const unsigned int*>(p));
    return __half22float2(*reinterpret_cast<__half2*>(&r));
  }
};
template <> struct Pair<__nv_bfloat16> {
  __device__ static __forceinline__ float2 load(const __nv_bfloat16* p) {
    unsigned int r = __ldg(reinterpret_cast<const unsigned int*>(p));
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
};
template <typename TL> __device__ __forceinline__ float load1(const TL* p) { return to_float(__ldg(p)); }
template <> __device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((unsigned int)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
template <> __device__ __forceinline__ float load1<__half>(const __half* p) {
  unsigned short r = __ldg(reinterpret_cast<const unsigned short*>(p));
  return __half2float(*reinterpret_cast<__half*>(&r));
}

// One bilinear footprint: 4 clamped pixel indices (relative to the level start) + 4 weights (0 when outside).
struct Footprint {
  int i00, i01, i10, i11;
  float w00, w01, w10, w11;
  float fx, fy;
  bool v00, v01, v10, v11;
};

__device__ __forceinline__ Footprint make_footprint(float x, float y, int H, int W) {
  Footprint f;
  // corners (floor, floor+1), each dropped if outside the map.  NaN / far-away samples are rejected before the
  // float->int conversion (all four corners outside anyway).
  const bool any = (x > -1.f) && (y > -1.f) && (x < (float)W) && (y < (float)H);
  const float xs = any ? x : 0.f, ys = any ? y : 0.f;
  const float x0f = floorf(xs), y0f = floorf(ys);
  const int x0 = (int)x0f, y0 = (int)y0f;
  f.fx = xs - x0f;
  f.fy = ys - y0f;
  const bool xl = any && x0 >= 0, xh = any && (x0 + 1) < W;
  const bool yl = any && y0 >= 0, yh = any && (y0 + 1) < H;
  f.v00 = xl && yl; f.v01 = xh && yl; f.v10 = xl && yh; f.v11 = xh && yh;
  const int xa = max(x0, 0), xb = min(x0 + 1, W - 1);
  const int ya = max(y0, 0), yb = min(y0 + 1, H - 1);
  f.i00 = ya * W + xa; f.i01 = ya * W + xb; f.i10 = yb * W + xa; f.i11 = yb * W + xb;
  const float gx = 1.f - f.fx, gy = 1.f - f.fy;
  f.w00 = f.v00 ? gx * gy : 0.f;
  f.w01 = f.v01 ? f.fx * gy : 0.f;
  f.w10 = f.v10 ? gx * f.fy : 0.f;
  f.w11 = f.v11 ? f.fx * f.fy : 0.f;
  return f;
}

// Specialised forward for bf16 values, head dim 32, 3 levels x 6 points (EMRT's configuration); returns
// EMRT_ERR_UNSUPPORTED (without touching the error text) when the shape is not covered.
int gather_fwd_v1(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out,
                  int B, int Lq, int Lv, int M, int D, int L, int P, const LevelTable& lv, int loc_dtype, int mode,
                  cudaStream_t st);

// Window-staged forward (msda_gather_win.cu): TMA copies the value windows a region of queries can reach into
// shared memory; same contract as gather_fwd_v1, additionally needs Lq == Lv on a regular 3-level pyramid.
int gather_fwd_win(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out,
                   int B, int Lq, int Lv, int M, int D, int L, int P, const LevelTable& lv, int loc_dtype, int mode,
                   const int32_t* win_center_host, cudaStream_t st);

// Windowed backward (msda_gather_bwd_win.cu): grad_value accumulated in fixed point in shared-memory windows (integer
// shared-memory reductions), one float reduction per window pixel to L2; same contract as the generic backward,
// pixel-major bf16 value, Lq == Lv on a regular 3-level pyramid.
int gather_bwd_win(const void* go, const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs,
                   float* gv, float* gl, float* ga, int B, int Lq, int Lv, int M, int D, int L, int P,
                   const LevelTable& lv, int loc_dtype, int mode, const int32_t* win_center_host, cudaStream_t st);

}  // namespace emrt
