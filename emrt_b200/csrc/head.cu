// Segmentation-head tail: UpHead's last x2 bilinear upsample (paddle_EMRT.py:178-180), slide_inference's
// window accumulation / divide (src/api/infer.py:69-79), ss_inference's resize + softmax + argmax
// (infer.py:150-154, predict.py:162-166) and calculate_area (src/utils/metrics.py:20-69).
//
// All of it is HBM-bound byte/float work.  The fused kernel reads only the half-resolution window logits and
// writes only the label map: per pixel it upsamples every covering window on the fly and sums them in window
// order (deterministic, no atomics, no full-resolution fp32 canvas round trip).
#include <cstdlib>

#include <cstring>

#include "common.cuh"

namespace emrt {

// half-pixel source coordinate for align_corners=False: src = (dst + 0.5) * scale - 0.5, clamped at 0
struct Tap { int i0, i1; float f; };
__device__ __forceinline__ Tap make_tap(int dst, int n_in, float scale) {
  float s = ((float)dst + 0.5f) * scale - 0.5f;
  s = s < 0.f ? 0.f : s;
  Tap t;
  t.i0 = min((int)s, n_in - 1);
  t.i1 = min(t.i0 + 1, n_in - 1);
  t.f = s - (float)t.i0;
  return t;
}

template <typename T>
__device__ __forceinline__ float bilerp(const T* __restrict__ plane, int w, const Tap& ty, const Tap& tx) {
  const float a = to_float(plane[ty.i0 * w + tx.i0]);
  const float b = to_float(plane[ty.i0 * w + tx.i1]);
  const float c = to_float(plane[ty.i1 * w + tx.i0]);
  const float d = to_float(plane[ty.i1 * w + tx.i1]);
  const float top = a + tx.f * (b - a);
  const float bot = c + tx.f * (d - c);
  return top + ty.f * (bot - top);
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t planes, int h, int w) {
  const int H = 2 * h, W = 2 * w;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= planes * H * W) return;
  const int ox = (int)(idx % W);
  const int oy = (int)((idx / W) % H);
  const int64_t pl = idx / ((int64_t)W * H);
  out[idx] = bilerp<T>(in + pl * h * w, w, make_tap(oy, h, 0.5f), make_tap(ox, w, 0.5f));
}

// ---- ordered list of the windows that touch a CTA's pixel tile ---------------------------------------------
constexpr int TILE_X = 32, TILE_Y = 8, MAX_LIST = 64;

struct WinList {
  int n;            // number of entries, or -1 when more than MAX_LIST windows touch the tile (scan globally)
  int idx[MAX_LIST];
};

__device__ __forceinline__ void build_window_list(WinList& wl, int* warp_cnt, int n_win, int img, int ty0, int tx0,
                                                  int tile_h, int tile_w, int hc, int wc,
                                                  const int32_t* __restrict__ win_img,
                                                  const int32_t* __restrict__ win_y0,
                                                  const int32_t* __restrict__ win_x0) {
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  if (t == 0) wl.n = 0;
  __syncthreads();
  for (int base = 0; base < n_win; base += 256) {
    const int w = base + t;
    bool hit = false;
    if (w < n_win && __ldg(win_img + w) == img) {
      const int y0 = __ldg(win_y0 + w), x0 = __ldg(win_x0 + w);
      hit = (y0 < ty0 + tile_h) && (y0 + hc > ty0) && (x0 < tx0 + tile_w) && (x0 + wc > tx0);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    int off = wl.n, total = 0;
    for (int i = 0; i < 8; ++i) { if (i < warp) off += warp_cnt[i]; total += warp_cnt[i]; }
    off += __popc(ballot & ((1u << lane) - 1u));
    if (hit && off >= 0 && off < MAX_LIST) wl.idx[off] = w;
    __syncthreads();
    if (t == 0) wl.n = (wl.n < 0 || wl.n + total > MAX_LIST) ? -1 : wl.n + total;
    __syncthreads();
  }
}

// softmax(axis=1) -> argmax(axis=1) with first-max tie rule (infer.py:152-153).
template <int NC>
__device__ __forceinline__ int softmax_argmax(const float (&l)[NC], int nc, float (&prob)[NC]) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NC; ++c) if (c < nc) mx = fmaxf(mx, l[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) if (c < nc) { prob[c] = expf(l[c] - mx); s += prob[c]; }
  const float inv = 1.f / s;
  int best = 0;
  float bv = -INFINITY;
#pragma unroll
  for (int c = 0; c < NC; ++c) if (c < nc) {
    prob[c] *= inv;
    if (prob[c] > bv) { bv = prob[c]; best = c; }
  }
  return best;
}

// argmax over the logits themselves: argmax(softmax(l)) == argmax(l) (exp is monotonic; equal logits give equal
// probabilities, so the first-max tie rule picks the same class) — used whenever probabilities are not requested.
template <int NC>
__device__ __forceinline__ int argmax_first(const float (&l)[NC], int nc) {
  int best = 0;
  float bv = l[0];
#pragma unroll
  for (int c = 1; c < NC; ++c) if (c < nc && l[c] > bv) { bv = l[c]; best = c; }
  return best;
}

__device__ __forceinline__ void store_label(void* labels, int label_dtype, int64_t i, int v) {
  if (label_dtype == EMRT_U8) reinterpret_cast<uint8_t*>(labels)[i] = (uint8_t)v;
  else reinterpret_cast<int32_t*>(labels)[i] = v;
}

// a6 accumulate on a full-resolution canvas: canvas += sum of window logits (window order), count += cover.
__global__ void __launch_bounds__(256)
window_accumulate_kernel(const float* __restrict__ win_logits, float* __restrict__ canvas, float* __restrict__ count,
                         int n_win, int nc, int hc, int wc, int H, int W, const int32_t* __restrict__ win_img,
                         const int32_t* __restrict__ win_y0, const int32_t* __restrict__ win_x0) {
  __shared__ WinList wl;
  __shared__ int warp_cnt[8];
  const int img = blockIdx.z;
  const int tx0 = blockIdx.x * TILE_X, ty0 = blockIdx.y * TILE_Y;
  build_window_list(wl, warp_cnt, n_win, img, ty0, tx0, TILE_Y, TILE_X, hc, wc, win_img, win_y0, win_x0);
  const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
  if (x >= W || y >= H) return;
  const int n = wl.n < 0 ? n_win : wl.n;
  float cnt = 0.f;
  const int64_t plane = (int64_t)H * W;
  float* cv = canvas + (int64_t)img * nc * plane + (int64_t)y * W + x;
  for (int i = 0; i < n; ++i) {
    const int w = wl.n < 0 ? i : wl.idx[i];
    if (wl.n < 0 && __ldg(win_img + w) != img) continue;
    const int ly = y - __ldg(win_y0 + w), lx = x - __ldg(win_x0 + w);
    if (ly < 0 || ly >= hc || lx < 0 || lx >= wc) continue;
    cnt += 1.f;
    const float* src = win_logits + ((int64_t)w * nc * hc + ly) * wc + lx;
    for (int c = 0; c < nc; ++c) cv[c * plane] += __ldg(src + (int64_t)c * hc * wc);
  }
  if (count) count[(int64_t)img * plane + (int64_t)y * W + x] += cnt;
}

template <int NC>
__global__ void __launch_bounds__(256)
finalize_argmax_kernel(const float* __restrict__ canvas, const float* __restrict__ count, void* __restrict__ labels,
                       int label_dtype, float* __restrict__ probs_out, float* __restrict__ logits_out, int n_img,
                       int nc, int H, int W, int Ho, int Wo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t oplane = (int64_t)Ho * Wo, plane = (int64_t)H * W;
  if (idx >= n_img * oplane) return;
  const int img = (int)(idx / oplane);
  const int oy = (int)((idx % oplane) / Wo), ox = (int)(idx % Wo);
  const float* cv = canvas + (int64_t)img * nc * plane;
  const float* ct = count ? count + (int64_t)img * plane : nullptr;
  float l[NC], prob[NC];
  if (Ho == H && Wo == W) {
    const int64_t o = (int64_t)oy * W + ox;
    const float c0 = ct ? ct[o] : 1.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) if (c < nc) l[c] = cv[c * plane + o] / c0;   // logit / count (infer.py:79)
    if (logits_out) {
#pragma unroll
      for (int c = 0; c < NC; ++c) if (c < nc) logits_out[((int64_t)img * nc + c) * plane + o] = l[c];
    }
  } else {
    // F.interpolate(logit, ori_shape, 'bilinear', align_corners=False) on the averaged logits (infer.py:151)
    const Tap ty = make_tap(oy, H, (float)H / (float)Ho), tx = make_tap(ox, W, (float)W / (float)Wo);
    const int64_t o00 = (int64_t)ty.i0 * W + tx.i0, o01 = (int64_t)ty.i0 * W + tx.i1;
    const int64_t o10 = (int64_t)ty.i1 * W + tx.i0, o11 = (int64_t)ty.i1 * W + tx.i1;
    const float c00 = ct ? ct[o00] : 1.f, c01 = ct ? ct[o01] : 1.f, c10 = ct ? ct[o10] : 1.f, c11 = ct ? ct[o11] : 1.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) if (c < nc) {
      const float* p = cv + c * plane;
      const float a = p[o00] / c00, b = p[o01] / c01, cc = p[o10] / c10, d = p[o11] / c11;
      const float top = a + tx.f * (b - a), bot = cc + tx.f * (d - cc);
      l[c] = top + ty.f * (bot - top);
    }
  }
  const int best = probs_out ? softmax_argmax<NC>(l, nc, prob) : argmax_first<NC>(l, nc);
  store_label(labels, label_dtype, idx, best);
  if (probs_out) {
#pragma unroll
    for (int c = 0; c < NC; ++c) if (c < nc) probs_out[((int64_t)img * nc + c) * oplane + (int64_t)oy * Wo + ox] = prob[c];
  }
}

// a5 + a6 + a7 in one pass over the label map.
template <typename T, int NC>
__global__ void __launch_bounds__(256)
stitch_argmax_fused_kernel(const T* __restrict__ half_logits, void* __restrict__ labels, int label_dtype,
                           float* __restrict__ logits_out, int n_win, int nc, int hc, int wc, int H, int W,
                           const int32_t* __restrict__ win_img, const int32_t* __restrict__ win_y0,
                           const int32_t* __restrict__ win_x0) {
  __shared__ WinList wl;
  __shared__ int warp_cnt[8];
  const int img = blockIdx.z;
  const int tx0 = blockIdx.x * TILE_X, ty0 = blockIdx.y * TILE_Y;
  build_window_list(wl, warp_cnt, n_win, img, ty0, tx0, TILE_Y, TILE_X, hc, wc, win_img, win_y0, win_x0);
  const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
  if (x >= W || y >= H) return;
  const int hh = hc / 2, hw = wc / 2;
  const int n = wl.n < 0 ? n_win : wl.n;
  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
  float cnt = 0.f;
  for (int i = 0; i < n; ++i) {
    const int w = wl.n < 0 ? i : wl.idx[i];
    if (wl.n < 0 && __ldg(win_img + w) != img) continue;
    const int ly = y - __ldg(win_y0 + w), lx = x - __ldg(win_x0 + w);
    if (ly < 0 || ly >= hc || lx < 0 || lx >= wc) continue;
    cnt += 1.f;
    const Tap ty = make_tap(ly, hh, 0.5f), tx = make_tap(lx, hw, 0.5f);
    const T* src = half_logits + (int64_t)w * nc * hh * hw;
#pragma unroll
    for (int c = 0; c < NC; ++c) if (c < nc) acc[c] += bilerp<T>(src + (int64_t)c * hh * hw, hw, ty, tx);
  }
  const int64_t plane = (int64_t)H * W, o = (int64_t)y * W + x;
  if (logits_out) {
#pragma unroll
    for (int c = 0; c < NC; ++c) if (c < nc) {
      acc[c] = acc[c] / cnt;                                   // logit / count (infer.py:79)
      logits_out[((int64_t)img * nc + c) * plane + o] = acc[c];
    }
  }
  store_label(labels, label_dtype, (int64_t)img * plane + o, argmax_first<NC>(acc, nc));
}

// a5 + a6 + a7, one 2x2 quad of label pixels per thread (H, W even).  For x2 upsampling the four pixels of a quad read
// at most a 3x3 patch of the half-resolution logits, so a window costs 9 loads per class for 4 outputs instead of
// 16, and the tap / address arithmetic is shared.  PY / PX = parity of the quad's first row / column in the
// window's own coordinates (uniform per CTA and window); the per-pixel arithmetic is make_tap + bilerp unchanged,
// so results equal the one-pixel kernel bit for bit.
template <typename T, int NC, int PY, int PX>
__device__ __forceinline__ void quad_accumulate(const T* __restrict__ src, int nc, int hh, int hw, int ly, int lx,
                                                float (&acc)[4][NC]) {
  // rows / columns of the patch: local coordinate 2k -> taps (k-1, k); 2k+1 -> taps (k, k+1); clamped loads reproduce
  // make_tap's edge rule exactly (a + f * (a - a) == a)
  constexpr int NR = PY ? 2 : 3, NCOL = PX ? 2 : 3;
  const int ky = ly >> 1, kx = lx >> 1;
  int ry[NR], rx[NCOL];
#pragma unroll
  for (int i = 0; i < NR; ++i) ry[i] = min(max(ky - (PY ? 0 : 1) + i, 0), hh - 1) * hw;
#pragma unroll
  for (int i = 0; i < NCOL; ++i) rx[i] = min(max(kx - (PX ? 0 : 1) + i, 0), hw - 1);
  // output row j (0,1) uses patch rows (a, a+1) with fraction f: even local -> (0,1) f = 0.75 ; odd local -> (k,k+1) f = 0.25
  // PY == 0: row 0 is even -> rows (0,1) f .75 ; row 1 odd -> rows (1,2) f .25.   PY == 1: row 0 odd -> (0,1) f .25 ; row 1 even -> (0,1) f .75
  constexpr int ra[2] = {0, PY ? 0 : 1}, ca[2] = {0, PX ? 0 : 1};
  float fy[2], fx[2];
  fy[0] = make_tap(ly, hh, 0.5f).f; fy[1] = make_tap(ly + 1, hh, 0.5f).f;
  fx[0] = make_tap(lx, hw, 0.5f).f; fx[1] = make_tap(lx + 1, hw, 0.5f).f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < nc) {
      const T* pl = src + (int64_t)c * hh * hw;
      float v[NR][NCOL];
#pragma unroll
      for (int i = 0; i < NR; ++i)
#pragma unroll
        for (int j = 0; j < NCOL; ++j) v[i][j] = to_float(pl[ry[i] + rx[j]]);
#pragma unroll
      for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
          const float a = v[ra[oy]][ca[ox]], b = v[ra[oy]][ca[ox] + 1];
          const float cc = v[ra[oy] + 1][ca[ox]], d = v[ra[oy] + 1][ca[ox] + 1];
          const float top = a + fx[ox] * (b - a);
          const float bot = cc + fx[ox] * (d - cc);
          acc[oy * 2 + ox][c] += top + fy[oy] * (bot - top);
        }
    }
  }
}

constexpr int QTILE = 32;   // label pixels per CTA side: 16 x 16 threads, one quad each

// Evaluation extras of the fused kernel: the label map goes straight into
//   * metrics.calculate_area's per-class areas (src/utils/metrics.py:20-69, ignore-index contract) per image, and
//   * predict.py:171-174's palette image (uint8 [n_img, H, W, 3], colour = palette[class]),
// so neither the one-hot tensors nor a second pass over the labels exist.
struct StitchEval {
  const void* gt;                // ground-truth labels [n_img, H, W] (I32 | U8) or NULL
  int gt_dtype;
  int ignore_index;
  unsigned long long* areas;     // [n_img, 3, nc]: intersect, pred, label (accumulated)
  const uint8_t* palette;        // [nc, 3] or NULL
  uint8_t* color;                // [n_img, H, W, 3] or NULL
};

// a5 + a6 + a7, one 16 x 2 strip of label pixels per thread (bf16 logits, label map only).  The quad kernel above is bound by
// instruction issue: 9 two-byte loads per class and window for four labels.  Here a thread owns 16 consecutive label columns
// of a row pair, so for a window whose origin is a multiple of 16 (x) and 2 (y) — every origin of the reference's 512 / 384
// plans — one class of one window costs three 16-byte loads (half-resolution rows k-1, k, k+1, columns j..j+7) plus the
// two neighbour columns, and 6 flops per label.  Classes run in the OUTER loop (one accumulator strip + running first-max),
// windows in list order inside: per pixel the same sums in the same order, with the same expressions
// (a + f * (b - a), make_tap's fractions 0.75 / 0.25, clamped edge taps), as the one-pixel kernel — labels are bit-equal.
// Windows at other origins take the per-pixel path inside the same loop.
constexpr int STRIP_W = 16, STRIP_TILE_X = 32 * STRIP_W, STRIP_TILE_Y = 16;

__device__ __forceinline__ float lerp_tap(float a, float b, float f) { return a + f * (b - a); }

template <int NC>
__global__ void __launch_bounds__(256)
stitch_argmax_strip_kernel(const __nv_bfloat16* __restrict__ half_logits, void* __restrict__ labels, int label_dtype,
                           int n_win, int nc, int hc, int wc, int H, int W, const int32_t* __restrict__ win_img,
                           const int32_t* __restrict__ win_y0, const int32_t* __restrict__ win_x0) {
  __shared__ WinList wl;
  __shared__ int warp_cnt[8];
  __shared__ int s_y0[MAX_LIST], s_x0[MAX_LIST];
  const int img = blockIdx.z;
  const int tx0 = blockIdx.x * STRIP_TILE_X, ty0 = blockIdx.y * STRIP_TILE_Y;
  build_window_list(wl, warp_cnt, n_win, img, ty0, tx0, STRIP_TILE_Y, STRIP_TILE_X, hc, wc, win_img, win_y0, win_x0);
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  if (wl.n >= 0 && t < wl.n) { s_y0[t] = __ldg(win_y0 + wl.idx[t]); s_x0[t] = __ldg(win_x0 + wl.idx[t]); }
  __syncthreads();
  const int X = tx0 + threadIdx.x * STRIP_W, y = ty0 + threadIdx.y * 2;
  if (X >= W || y >= H) return;
  const int hh = hc / 2, hw = wc / 2;
  const int n = wl.n < 0 ? n_win : wl.n;
  float best[2][STRIP_W];
  uint32_t bidx[2][2] = {{0u, 0u}, {0u, 0u}};          // 4 bits per pixel
  for (int c = 0; c < nc; ++c) {
    float acc[2][STRIP_W];
#pragma unroll
    for (int x = 0; x < STRIP_W; ++x) acc[0][x] = acc[1][x] = 0.f;
    for (int i = 0; i < n; ++i) {
      int w, wy0, wx0;
      if (wl.n < 0) {
        w = i;
        if (__ldg(win_img + w) != img) continue;
        wy0 = __ldg(win_y0 + w); wx0 = __ldg(win_x0 + w);
      } else {
        w = wl.idx[i]; wy0 = s_y0[i]; wx0 = s_x0[i];
      }
      const int ly = y - wy0, lx = X - wx0;
      if (ly + 1 < 0 || ly >= hc || lx + STRIP_W - 1 < 0 || lx >= wc) continue;
      const __nv_bfloat16* pl = half_logits + ((int64_t)w * nc + c) * hh * hw;
      if (((wy0 & 1) | (wx0 & (STRIP_W - 1))) == 0) {
        // aligned window: the strip is wholly inside it
        const int k = ly >> 1, j0 = lx >> 1;
        const int jl = max(j0 - 1, 0), jr = min(j0 + 8, hw - 1);
        const int rows[3] = {max(k - 1, 0) * hw, k * hw, min(k + 1, hh - 1) * hw};
        float h[3][10];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(pl + rows[r] + j0));
          h[r][0] = __bfloat162float(pl[rows[r] + jl]);
          h[r][9] = __bfloat162float(pl[rows[r] + jr]);
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            h[r][1 + 2 * q] = __uint_as_float(u[q] << 16);
            h[r][2 + 2 * q] = __uint_as_float(u[q] & 0xffff0000u);
          }
        }
#pragma unroll
        for (int x = 0; x < STRIP_W; ++x) {
          // even local column 2j: taps (j-1, j), f = 0.75; odd 2j+1: taps (j, j+1), f = 0.25 (make_tap, scale 0.5)
          const int a = (x >> 1) + (x & 1);
          const float f = (x & 1) ? 0.25f : 0.75f;
          const float t0 = lerp_tap(h[0][a], h[0][a + 1], f);
          const float t1 = lerp_tap(h[1][a], h[1][a + 1], f);
          const float t2 = lerp_tap(h[2][a], h[2][a + 1], f);
          acc[0][x] += lerp_tap(t0, t1, 0.75f);
          acc[1][x] += lerp_tap(t1, t2, 0.25f);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int py = ly + r;
          if (py < 0 || py >= hc) continue;
          const Tap ty = make_tap(py, hh, 0.5f);
#pragma unroll
          for (int x = 0; x < STRIP_W; ++x) {
            const int px = lx + x;
            if (px < 0 || px >= wc) continue;
            acc[r][x] += bilerp<__nv_bfloat16>(pl, hw, ty, make_tap(px, hw, 0.5f));
          }
        }
      }
    }
    // first-max rule (infer.py:152-153): strictly greater replaces, classes in increasing order
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int x = 0; x < STRIP_W; ++x) {
        const bool up = c == 0 || acc[r][x] > best[r][x];
        best[r][x] = up ? acc[r][x] : best[r][x];
        const uint32_t sh = 4 * (x & 7);
        bidx[r][x >> 3] = up ? ((bidx[r][x >> 3] & ~(0xFu << sh)) | ((uint32_t)c << sh)) : bidx[r][x >> 3];
      }
  }
  const int64_t plane = (int64_t)H * W;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int64_t o = (int64_t)img * plane + (int64_t)(y + r) * W + X;
    if (label_dtype == EMRT_U8) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t nib = bidx[r][q >> 1] >> (16 * (q & 1));
        pk[q] = (nib & 0xFu) | ((nib & 0xF0u) << 4) | ((nib & 0xF00u) << 8) | ((nib & 0xF000u) << 12);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(labels) + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t nib = bidx[r][q >> 1] >> (16 * (q & 1));
        *reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(labels) + o + 4 * q) =
            make_int4(nib & 0xF, (nib >> 4) & 0xF, (nib >> 8) & 0xF, (nib >> 12) & 0xF);
      }
    }
  }
}

template <typename T, int NC, bool EVAL>
__global__ void __launch_bounds__(256)
stitch_argmax_quad_kernel(const T* __restrict__ half_logits, void* __restrict__ labels, int label_dtype,
                          float* __restrict__ logits_out, int n_win, int nc, int hc, int wc, int H, int W,
                          const int32_t* __restrict__ win_img, const int32_t* __restrict__ win_y0,
                          const int32_t* __restrict__ win_x0, const __grid_constant__ StitchEval ev) {
  __shared__ WinList wl;
  __shared__ int warp_cnt[8];
  __shared__ unsigned int hist[3 * NC];
  const int img = blockIdx.z;
  const int tx0 = blockIdx.x * QTILE, ty0 = blockIdx.y * QTILE;
  if (EVAL) {
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < 3 * NC) hist[t] = 0u;          // ordered before its use by the barriers inside build_window_list
  }
  build_window_list(wl, warp_cnt, n_win, img, ty0, tx0, QTILE, QTILE, hc, wc, win_img, win_y0, win_x0);
  const int x = tx0 + 2 * threadIdx.x, y = ty0 + 2 * threadIdx.y;
  const bool inside = x < W && y < H;
  if (!EVAL && !inside) return;
  int gt_lab[4] = {0, 0, 0, 0}, pred_lab[4] = {-1, -1, -1, -1};
  if (inside) {
  const int hh = hc / 2, hw = wc / 2;
  const int n = wl.n < 0 ? n_win : wl.n;
  float acc[4][NC];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[k][c] = 0.f;
  float cnt[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < n; ++i) {
    const int w = wl.n < 0 ? i : wl.idx[i];
    if (wl.n < 0 && __ldg(win_img + w) != img) continue;
    const int ly = y - __ldg(win_y0 + w), lx = x - __ldg(win_x0 + w);
    const T* src = half_logits + (int64_t)w * nc * hh * hw;
    if (ly >= 0 && ly + 1 < hc && lx >= 0 && lx + 1 < wc) {
      // whole quad inside the window (always the case for even origins and sizes)
#pragma unroll
      for (int k = 0; k < 4; ++k) cnt[k] += 1.f;
      switch ((ly & 1) * 2 + (lx & 1)) {
        case 0: quad_accumulate<T, NC, 0, 0>(src, nc, hh, hw, ly, lx, acc); break;
        case 1: quad_accumulate<T, NC, 0, 1>(src, nc, hh, hw, ly, lx, acc); break;
        case 2: quad_accumulate<T, NC, 1, 0>(src, nc, hh, hw, ly, lx, acc); break;
        default: quad_accumulate<T, NC, 1, 1>(src, nc, hh, hw, ly, lx, acc); break;
      }
    } else {
      // quad straddles the window border (odd origin): per-pixel path
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int py = ly + (k >> 1), px = lx + (k & 1);
        if (py < 0 || py >= hc || px < 0 || px >= wc) continue;
        cnt[k] += 1.f;
        const Tap ty = make_tap(py, hh, 0.5f), tx = make_tap(px, hw, 0.5f);
#pragma unroll
        for (int c = 0; c < NC; ++c) if (c < nc) acc[k][c] += bilerp<T>(src + (int64_t)c * hh * hw, hw, ty, tx);
      }
    }
  }
  const int64_t plane = (int64_t)H * W;
  int best[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (logits_out) {
#pragma unroll
      for (int c = 0; c < NC; ++c) if (c < nc) acc[k][c] = acc[k][c] / cnt[k];     // logit / count (infer.py:79)
    }
    // argmax(softmax(sum / count)) == argmax(sum): exp and the division by a positive count are monotonic, and equal
    // sums give equal probabilities, so the first-max tie rule (infer.py:152-153) picks the same class
    best[k] = argmax_first<NC>(acc[k], nc);
  }
#pragma unroll
  for (int oy = 0; oy < 2; ++oy) {
    const int64_t o = (int64_t)(y + oy) * W + x;
    if (logits_out) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < nc)
          *reinterpret_cast<float2*>(logits_out + ((int64_t)img * nc + c) * plane + o) =
              make_float2(acc[oy * 2][c], acc[oy * 2 + 1][c]);
    }
    if (label_dtype == EMRT_U8) {
      *reinterpret_cast<uchar2*>(reinterpret_cast<uint8_t*>(labels) + (int64_t)img * plane + o) =
          make_uchar2((unsigned char)best[oy * 2], (unsigned char)best[oy * 2 + 1]);
    } else {
      *reinterpret_cast<int2*>(reinterpret_cast<int32_t*>(labels) + (int64_t)img * plane + o) =
          make_int2(best[oy * 2], best[oy * 2 + 1]);
    }
    if (EVAL && ev.color) {
      uint8_t* dst = ev.color + ((int64_t)img * plane + o) * 3;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dst[k * 3 + ch] = __ldg(ev.palette + best[oy * 2 + k] * 3 + ch);
    }
    if (EVAL && ev.gt) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int64_t gi = (int64_t)img * plane + o + k;
        gt_lab[oy * 2 + k] = ev.gt_dtype == EMRT_U8 ? (int)__ldg(reinterpret_cast<const uint8_t*>(ev.gt) + gi)
                                                    : __ldg(reinterpret_cast<const int32_t*>(ev.gt) + gi);
        pred_lab[oy * 2 + k] = best[oy * 2 + k];
      }
    }
  }
  }   // inside
  if (EVAL && ev.gt) {
    // metrics.calculate_area: mask = label != ignore_index; areas of pred / label / their intersection per class.
    // One ballot per (pixel slot, class, area kind): a warp adds its counts with 3 * nc shared-memory atomics.
    const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c >= nc) break;
      unsigned n_i = 0, n_p = 0, n_l = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool keep = inside && gt_lab[k] != ev.ignore_index;
        const bool ip = keep && pred_lab[k] == c, il = keep && gt_lab[k] == c;
        n_p += __popc(__ballot_sync(0xffffffffu, ip));
        n_l += __popc(__ballot_sync(0xffffffffu, il));
        n_i += __popc(__ballot_sync(0xffffffffu, ip && il));
      }
      if (lane == 0) {
        if (n_i) atomicAdd(&hist[c], n_i);
        if (n_p) atomicAdd(&hist[NC + c], n_p);
        if (n_l) atomicAdd(&hist[2 * NC + c], n_l);
      }
    }
    __syncthreads();
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < 3 * NC && (t % NC) < nc && hist[t])
      atomicAdd(ev.areas + ((int64_t)img * 3 + t / NC) * nc + (t % NC), (unsigned long long)hist[t]);
  }
}

__global__ void __launch_bounds__(256)
calculate_area_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ label, int64_t n, int nc,
                      int ignore_index, unsigned long long* __restrict__ areas) {
  extern __shared__ unsigned int hist[];   // 3*nc
  for (int i = threadIdx.x; i < 3 * nc; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = label[i], p = pred[i];
    if (l == ignore_index) continue;            // mask = label != ignore_index (metrics.py:44)
    if (p >= 0 && p < nc) atomicAdd(&hist[nc + p], 1u);
    if (l >= 0 && l < nc) atomicAdd(&hist[2 * nc + l], 1u);
    if (p == l && p >= 0 && p < nc) atomicAdd(&hist[p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * nc; i += blockDim.x)
    if (hist[i]) atomicAdd(areas + i, (unsigned long long)hist[i]);
}

}  // namespace emrt

using namespace emrt;

extern "C" int emrt_upsample2x(const void* in, float* out, int n, int nc, int h, int w, int in_dtype, void* stream) {
  EMRT_REQUIRE(in && out && n > 0 && nc > 0 && h > 0 && w > 0, "bad upsample2x arguments");
  const int64_t total = (int64_t)n * nc * 4 * h * w;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (in_dtype == EMRT_F32) upsample2x_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, out, (int64_t)n * nc, h, w);
  else if (in_dtype == EMRT_BF16) upsample2x_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, out, (int64_t)n * nc, h, w);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad in_dtype %d", in_dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_window_accumulate(const float* win_logits, float* canvas, float* count, int n_win, int n_img,
                                      int nc, int hc, int wc, int H, int W, const int32_t* win_img,
                                      const int32_t* win_y0, const int32_t* win_x0, void* stream) {
  EMRT_REQUIRE(win_logits && canvas && win_img && win_y0 && win_x0, "NULL pointer");
  EMRT_REQUIRE(n_win > 0 && n_img > 0 && nc > 0 && hc > 0 && wc > 0 && H > 0 && W > 0, "non-positive dimension");
  dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y, n_img), block(TILE_X, TILE_Y);
  window_accumulate_kernel<<<grid, block, 0, as_stream(stream)>>>(win_logits, canvas, count, n_win, nc, hc, wc, H, W,
                                                                   win_img, win_y0, win_x0);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_finalize_argmax(const float* canvas, const float* count, void* labels, int label_dtype,
                                    float* probs_out, float* logits_out, int n_img, int nc, int H, int W, int Ho,
                                    int Wo, void* stream) {
  EMRT_REQUIRE(canvas && labels, "NULL pointer");
  EMRT_REQUIRE(n_img > 0 && nc > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "non-positive dimension");
  EMRT_REQUIRE(label_dtype == EMRT_I32 || label_dtype == EMRT_U8, "label_dtype must be I32 or U8");
  EMRT_REQUIRE(!logits_out || (Ho == H && Wo == W), "logits_out needs Ho==H and Wo==W");
  if (nc > 32) return set_error(EMRT_ERR_UNSUPPORTED, "nc=%d > 32", nc);
  const int64_t total = (int64_t)n_img * Ho * Wo;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (nc <= 8)
    finalize_argmax_kernel<8><<<blocks, 256, 0, st>>>(canvas, count, labels, label_dtype, probs_out, logits_out, n_img, nc, H, W, Ho, Wo);
  else
    finalize_argmax_kernel<32><<<blocks, 256, 0, st>>>(canvas, count, labels, label_dtype, probs_out, logits_out, n_img, nc, H, W, Ho, Wo);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

static int stitch_argmax_launch(const void* half_logits, int in_dtype, void* labels, int label_dtype,
                                float* logits_out, int n_win, int n_img, int nc, int hc, int wc, int H, int W,
                                const int32_t* win_img, const int32_t* win_y0, const int32_t* win_x0,
                                const StitchEval* eval, void* stream);

extern "C" int emrt_stitch_argmax_fused(const void* half_logits, int in_dtype, void* labels, int label_dtype,
                                        float* logits_out, int n_win, int n_img, int nc, int hc, int wc, int H, int W,
                                        const int32_t* win_img, const int32_t* win_y0, const int32_t* win_x0,
                                        void* stream) {
  return stitch_argmax_launch(half_logits, in_dtype, labels, label_dtype, logits_out, n_win, n_img, nc, hc, wc, H, W, win_img,
                              win_y0, win_x0, nullptr, stream);
}

extern "C" int emrt_stitch_argmax_eval(const void* half_logits, int in_dtype, void* labels, int label_dtype, int n_win,
                                       int n_img, int nc, int hc, int wc, int H, int W, const int32_t* win_img,
                                       const int32_t* win_y0, const int32_t* win_x0, const void* gt, int gt_dtype,
                                       int ignore_index, long long* areas, const uint8_t* palette, uint8_t* color,
                                       void* stream) {
  EMRT_REQUIRE((gt != nullptr) == (areas != nullptr), "gt and areas go together");
  EMRT_REQUIRE((palette != nullptr) == (color != nullptr), "palette and color go together");
  EMRT_REQUIRE(gt || color, "nothing to evaluate: use emrt_stitch_argmax_fused");
  EMRT_REQUIRE(!gt || gt_dtype == EMRT_I32 || gt_dtype == EMRT_U8, "gt_dtype must be I32 or U8");
  if (!(H % 2 == 0 && W % 2 == 0 && nc <= 8))
    return set_error(EMRT_ERR_UNSUPPORTED, "the fused evaluation needs even H, W and nc <= 8 (got %d x %d, nc %d)", H, W, nc);
  StitchEval ev;
  ev.gt = gt; ev.gt_dtype = gt_dtype; ev.ignore_index = ignore_index;
  ev.areas = reinterpret_cast<unsigned long long*>(areas); ev.palette = palette; ev.color = color;
  return stitch_argmax_launch(half_logits, in_dtype, labels, label_dtype, nullptr, n_win, n_img, nc, hc, wc, H, W, win_img,
                              win_y0, win_x0, &ev, stream);
}

static int stitch_argmax_launch(const void* half_logits, int in_dtype, void* labels, int label_dtype,
                                float* logits_out, int n_win, int n_img, int nc, int hc, int wc, int H, int W,
                                const int32_t* win_img, const int32_t* win_y0, const int32_t* win_x0,
                                const StitchEval* eval, void* stream) {
  EMRT_REQUIRE(half_logits && labels && win_img && win_y0 && win_x0, "NULL pointer");
  EMRT_REQUIRE(n_win > 0 && n_img > 0 && nc > 0 && hc > 0 && wc > 0 && H > 0 && W > 0, "non-positive dimension");
  EMRT_REQUIRE(hc % 2 == 0 && wc % 2 == 0, "window size must be even (x2 upsample of the half-resolution logits)");
  EMRT_REQUIRE(label_dtype == EMRT_I32 || label_dtype == EMRT_U8, "label_dtype must be I32 or U8");
  if (nc > 32) return set_error(EMRT_ERR_UNSUPPORTED, "nc=%d > 32", nc);
  cudaStream_t st = as_stream(stream);
  // label map only, bf16 logits, 16-pixel column alignment: the strip kernel (windows at unaligned origins are handled
  // inside it, per pixel).  EMRT_STITCH_QUAD=1 keeps the quad kernel (the bit-equality test compares the two).
  if (!eval && !logits_out && in_dtype == EMRT_BF16 && H % 2 == 0 && W % STRIP_W == 0 && wc % STRIP_W == 0 && nc <= 8 &&
      (reinterpret_cast<uintptr_t>(half_logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(labels) & 15) == 0 &&
      !getenv("EMRT_STITCH_QUAD") && !getenv("EMRT_STITCH_PIXEL")) {
    dim3 sgrid((W + STRIP_TILE_X - 1) / STRIP_TILE_X, (H + STRIP_TILE_Y - 1) / STRIP_TILE_Y, n_img), sblock(32, 8);
    stitch_argmax_strip_kernel<8><<<sgrid, sblock, 0, st>>>((const __nv_bfloat16*)half_logits, labels, label_dtype, n_win, nc,
                                                            hc, wc, H, W, win_img, win_y0, win_x0);
    EMRT_LAUNCH_CHECK();
    return EMRT_OK;
  }
  if (H % 2 == 0 && W % 2 == 0 && nc <= 8 && (eval || !getenv("EMRT_STITCH_PIXEL"))) {
    dim3 qgrid((W + QTILE - 1) / QTILE, (H + QTILE - 1) / QTILE, n_img), qblock(16, 16);
    StitchEval ev;
    memset(&ev, 0, sizeof(ev));
    if (eval) ev = *eval;
#define EMRT_QUAD(T)                                                                                                   \
    if (eval) stitch_argmax_quad_kernel<T, 8, true><<<qgrid, qblock, 0, st>>>((const T*)half_logits, labels, label_dtype, \
                  logits_out, n_win, nc, hc, wc, H, W, win_img, win_y0, win_x0, ev);                                       \
    else stitch_argmax_quad_kernel<T, 8, false><<<qgrid, qblock, 0, st>>>((const T*)half_logits, labels, label_dtype,       \
                  logits_out, n_win, nc, hc, wc, H, W, win_img, win_y0, win_x0, ev)
    if (in_dtype == EMRT_F32) { EMRT_QUAD(float); }
    else if (in_dtype == EMRT_BF16) { EMRT_QUAD(__nv_bfloat16); }
    else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad in_dtype %d", in_dtype);
#undef EMRT_QUAD
    EMRT_LAUNCH_CHECK();
    return EMRT_OK;
  }
  if (eval) return set_error(EMRT_ERR_UNSUPPORTED, "the fused evaluation needs even H, W and nc <= 8");
  dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y, n_img), block(TILE_X, TILE_Y);
#define EMRT_ST(T, NC)                                                                                           \
  stitch_argmax_fused_kernel<T, NC><<<grid, block, 0, st>>>((const T*)half_logits, labels, label_dtype, logits_out, \
                                                            n_win, nc, hc, wc, H, W, win_img, win_y0, win_x0)
  if (in_dtype == EMRT_F32) { if (nc <= 8) EMRT_ST(float, 8); else EMRT_ST(float, 32); }
  else if (in_dtype == EMRT_BF16) { if (nc <= 8) EMRT_ST(__nv_bfloat16, 8); else EMRT_ST(__nv_bfloat16, 32); }
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad in_dtype %d", in_dtype);
#undef EMRT_ST
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_calculate_area(const int32_t* pred, const int32_t* label, int64_t n, int nc, int ignore_index,
                                   long long* areas, void* stream) {
  EMRT_REQUIRE(pred && label && areas && n > 0 && nc > 0 && nc <= 1024, "bad calculate_area arguments");
  const int64_t want = (n + 255) / 256;
  const unsigned blocks = (unsigned)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  calculate_area_kernel<<<blocks, 256, 3 * nc * sizeof(unsigned int), as_stream(stream)>>>(
      pred, label, n, nc, ignore_index, reinterpret_cast<unsigned long long*>(areas));
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
