// Backward sampling-gather for encoder self-attention (EMRT_QUERY_PIXEL_GRID), grad_value accumulated WITHOUT float
// atomics in shared memory.
//
// The generic backward (msda_gather.cu) sends every bilinear corner of every sample to L2 as a red.global.add.v4.f32:
// B*Lq*M*L*P*4 corners x 32 channels — 396 M vector reductions per call at the training batch (16 tiles of 512^2),
// 1.5 ms, bound by the L2 atomic rate.  Here one CTA owns a region of queries (8 x 16 level-0 pixels + the co-located
// pixels of levels 1, 2, as in the forward window kernel) of one (batch, head); the gradient of the value windows that
// region reaches (region +- R pixels) is accumulated in shared memory and sent to L2 once per window pixel:
//   * shared memory has no native fp32 atomic add (atomicAdd(float) is an ATOMS.CAS loop), but it has a native integer
//     one.  Contributions are therefore accumulated in FIXED POINT: the CTA scales its grad_out rows by a power of two
//     so that max|grad_out| < 2^20 (exact), converts each contribution w * grad_out to an integer with the 1.5 * 2^23
//     magic-number add (round to nearest, no F2I), and adds it with red.shared.add.s32.  At most TH*TW*(1+1/4+1/16)*P
//     = 1008 contributions of weight <= 1 can meet in one window element, so the sum stays below 2^30; the rounding
//     error per contribution is 2^-21 of the CTA's max |grad_out|.  Integer addition is associative, so a CTA's sums do
//     not depend on warp scheduling (the cross-CTA reduction in L2 still does).
//   * 8 lanes per (query, head): lane (side, sub) owns the left or right pixel of the footprint and channels
//     [8 sub, 8 sub + 8).  At step k it handles channel 8 sub + ((k + 2 g + side) & 7), g = the query's slot in the warp:
//     window pixels are 128-byte aligned, so this rotation is what makes the 32 lanes of one red.shared hit 32 distinct
//     banks (one wavefront per instruction) instead of the same four.
//   * the windows are processed in two passes (level 0; levels 1 + 2) through one 93 KB buffer, flushed in between with
//     red.global.add.v4.f32 — one vector reduction per (window pixel, 4 channels), 8x fewer than the generic kernel,
//     skipping elements that stayed zero.
//   * grad_loc / grad_attn: bf16 x bf16 + fp32 dot products (fma.rn.f32.bf16) of the grad_out row with the value
//     corners (read through L1 from the pixel-major value tensor the training path keeps), reduced over the 8 lanes by
//     shuffles and written once — no atomics, as in the generic kernel.
//   * a sample that leaves the window but not the map falls back to float reductions in global memory, so results do
//     not depend on R or on the window-centre hint.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "msda_common.cuh"
#include "tc_common.cuh"

namespace emrt {

constexpr int BW_L = 3, BW_P = 6, BW_LP = BW_L * BW_P, BW_D = 32, BW_QPB = 4, BW_MAX_M = 16;
constexpr float BW_MAGIC = 12582912.f;            // 1.5 * 2^23: (x + MAGIC) has x rounded to nearest in its low mantissa bits
constexpr int BW_MAGIC_BITS = 0x4B400000;

struct BwdWinParams {
  int32_t WW[BW_L], WH[BW_L];     // window size in pixels
  uint32_t win_off[BW_L];         // byte offset of the level's int32 window (levels 0 and 1 both start at 0: two passes)
  uint32_t cnt_off[BW_L];         // byte offset of the level's per-pixel contribution counters (int32)
  uint32_t pass_bytes[2];         // bytes to clear before each pass (windows + counters)
  uint32_t go_off;                // byte offset of the CTA's grad_out rows (bf16, 64 bytes per query)
  int32_t R, TH, TW, tw_shift, regions_x, regions_y;
  int32_t Lq, Lv, M;
  LevelTable lv;
  int8_t cshift[BW_MAX_M][BW_L][2];
};

__device__ __forceinline__ uint4 bw_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bw_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// PLAIN = true replaces the reduction by a plain store: WRONG results, used only by the timing experiment that
// measures how much of the kernel is the shared-memory atomic unit (EMRT_BWD_WIN_TIMING_PLAIN_STORES=1, never in tests)
template <bool PLAIN>
__device__ __forceinline__ void bw_red_s32(uint32_t addr, int v) {
  if (PLAIN) asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
  else asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void bw_red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// acc += bf16(half HI of a) * bf16(half HI of b), fp32 accumulate
template <int HI>
__device__ __forceinline__ void bw_fhfma(float& acc, uint32_t a, uint32_t b) {
  unsigned short a_lo, a_hi, b_lo, b_hi;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(a_lo), "=h"(a_hi) : "r"(a));
  asm("mov.b32 {%0,%1}, %2;" : "=h"(b_lo), "=h"(b_hi) : "r"(b));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(HI ? a_hi : a_lo), "h"(HI ? b_hi : b_lo));
}
__device__ __forceinline__ float bw_dot8(const uint4& a, const uint4& b) {
  float s0 = 0.f, s1 = 0.f;
  bw_fhfma<0>(s0, a.x, b.x); bw_fhfma<1>(s1, a.x, b.x);
  bw_fhfma<0>(s0, a.y, b.y); bw_fhfma<1>(s1, a.y, b.y);
  bw_fhfma<0>(s0, a.z, b.z); bw_fhfma<1>(s1, a.z, b.z);
  bw_fhfma<0>(s0, a.w, b.w); bw_fhfma<1>(s1, a.w, b.w);
  return s0 + s1;
}
__device__ __forceinline__ float bw_absmax8(float m, const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m = fmaxf(m, __uint_as_float((w[i] << 16) & 0x7fffffffu));
    m = fmaxf(m, __uint_as_float(w[i] & 0x7fff0000u));
  }
  return m;
}

// first global query index of a warp batch (same numbering as the forward window kernel: level 0 row-major, then the
// co-located pixels of level 1, then level 2; TW a power of two >= 16, so a batch is 4 consecutive pixels of one row)
__device__ __forceinline__ int bw_batch_query_base(const BwdWinParams& p, const int (&qb)[BW_L], int n0, int n01, int sh0,
                                                   int batch) {
  const bool l1 = batch >= n0, l2 = batch >= n01;
  const int k = batch - (l2 ? n01 : (l1 ? n0 : 0));
  const int sh = sh0 - (l2 ? 2 : (l1 ? 1 : 0));
  const int y = k >> sh, x = k - (y << sh);
  const int W = l2 ? p.lv.W[2] : (l1 ? p.lv.W[1] : p.lv.W[0]);
  return (l2 ? qb[2] : (l1 ? qb[1] : qb[0])) + y * W + (x << 2);
}

// BW_WARPS warps per CTA, 2 CTAs per SM (8 warps: 128 registers, no spills; 10 warps: 96 registers); batches are claimed
// dynamically from a shared counter, reset for each pass
template <typename TL, int MODE, int BW_WARPS, bool PLAIN>
__global__ void __launch_bounds__(BW_WARPS * 32, 2)
msda_gather_bwd_win_kernel(const __nv_bfloat16* __restrict__ grad_out, const __nv_bfloat16* __restrict__ value,
                           const TL* __restrict__ loc, const TL* __restrict__ attn, const float* __restrict__ ref,
                           int64_t ref_bs, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                           float* __restrict__ grad_attn, const __grid_constant__ BwdWinParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_max[BW_WARPS];
  __shared__ int s_next;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x, region = blockIdx.y, b = blockIdx.z;
  const int ry = region / p.regions_x, rx = region - ry * p.regions_x;
  const uint32_t smem_base = smem_u32(smem);

  int ox[BW_L], oy[BW_L], qb[BW_L];
#pragma unroll
  for (int l = 0; l < BW_L; ++l) {
    const int x0 = (rx * p.TW) >> l, y0 = (ry * p.TH) >> l;
    const int cx = m < BW_MAX_M ? p.cshift[m][l][0] : 0, cy = m < BW_MAX_M ? p.cshift[m][l][1] : 0;
    ox[l] = min(max(x0 - p.R + cx, -1), p.lv.W[l] + 1 - p.WW[l]);
    oy[l] = min(max(y0 - p.R + cy, -1), p.lv.H[l] + 1 - p.WH[l]);
    qb[l] = p.lv.start[l] + y0 * p.lv.W[l] + x0;
  }
  const int n0 = (p.TH * p.TW) >> 2, n01 = n0 + (n0 >> 2), n_batches = n01 + (n0 >> 4), sh0 = p.tw_shift - 2;

  // ---- the CTA's grad_out rows (head m of its queries) into shared memory, and their largest magnitude ---------------
  float mx = 0.f;
  for (int idx = threadIdx.x; idx < n_batches * (BW_QPB * 4); idx += BW_WARPS * 32) {
    const int k = idx >> 2, part = idx & 3;
    const int q = bw_batch_query_base(p, qb, n0, n01, sh0, k >> 2) + (k & 3);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(grad_out + (((int64_t)b * p.Lq + q) * p.M + m) * BW_D + part * 8));
    bw_sts128(smem_base + p.go_off + (uint32_t)idx * 16, v);
    mx = bw_absmax8(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < BW_WARPS; ++w) mx = fmaxf(mx, s_max[w]);
  // max in [2^(e-127), 2^(e-126)) -> scale 2^(146-e): max * scale in [2^19, 2^20).  Tiny / zero / non-finite max: scale 1.
  const uint32_t ebits = (__float_as_uint(mx) >> 23) & 0xffu;
  const bool scaled = ebits >= 20u && ebits < 255u;
  const float scale = scaled ? __uint_as_float((273u - ebits) << 23) : 1.f;
  const float inv_scale = scaled ? __uint_as_float((ebits - 19u) << 23) : 1.f;

  const int g = lane >> 3, s = lane & 7, side = s >> 2, sub = s & 3;
  const int rot = 2 * g + side;
  const int pix_stride = p.M * BW_D;      // elements between neighbouring pixels of one head (the host checks 32-bit offsets fit)
  const float sgn = side ? 1.f : -1.f;

  for (int pass = 0; pass < 2; ++pass) {
    // ---- zero this pass's windows ------------------------------------------------------------------------------------
    const uint32_t pass_bytes = p.pass_bytes[pass];
    for (uint32_t o = threadIdx.x * 16; o < pass_bytes; o += BW_WARPS * 32 * 16) bw_sts128(smem_base + o, make_uint4(0, 0, 0, 0));
    if (threadIdx.x == 0) s_next = BW_WARPS;
    __syncthreads();

    for (int batch = warp; batch < n_batches;) {
      const int q = bw_batch_query_base(p, qb, n0, n01, sh0, batch) + g;
      const int64_t item = ((int64_t)b * p.Lq + q) * p.M + m;
      // this lane's eight grad_out channels: packed bf16 in natural order for the dot products, fp32 in the rotated
      // order of the scatter steps
      const uint32_t grow = smem_base + p.go_off + (uint32_t)(batch * BW_QPB + g) * (BW_D * 2);
      const uint4 gnat = bw_lds128(grow + sub * 16);
      float gr[8];
      uint32_t koff[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = (k + rot) & 7;
        unsigned short h;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(grow + (uint32_t)(sub * 8 + c) * 2));
        gr[k] = __uint_as_float((uint32_t)h << 16);
        koff[k] = (uint32_t)c * 4;
      }
      // the 18 sampling offsets / attention weights of this (query, head): three per lane, one memory latency, then
      // handed round the query's 8 lanes by shuffles
      const TL* lp = loc + item * (BW_LP * 2);
      const TL* ap = attn + item * BW_LP;
      float px_[3], py_[3], pa_[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int j = min(r * 8 + s, BW_LP - 1);
        const float2 xy = Pair<TL>::load(lp + j * 2);
        px_[r] = xy.x; py_[r] = xy.y; pa_[r] = load1<TL>(ap + j);
      }
      const float* rp = (MODE == EMRT_LOC_PIXEL_OFFSET) ? ref + b * ref_bs + (int64_t)q * (BW_L * 2) : nullptr;
      float* gl = grad_loc + item * (BW_LP * 2);
      float* ga_out = grad_attn + item * BW_LP;

#pragma unroll
      for (int l = 0; l < BW_L; ++l) {
        if ((l == 0) != (pass == 0)) continue;
        const int H = p.lv.H[l], W = p.lv.W[l], WWl = p.WW[l], WHl = p.WH[l];
        const float fW = (float)W, fH = (float)H;
        const uint32_t wbase = smem_base + p.win_off[l] + (uint32_t)sub * 32;
        const uint32_t cbase = smem_base + p.cnt_off[l];
        const int64_t lvl_off = (((int64_t)b * p.Lv + p.lv.start[l]) * p.M + m) * BW_D + sub * 8;
        const __nv_bfloat16* vptr = value + lvl_off;
        float* gvptr = grad_value + lvl_off;
        float rxp = 0.f, ryp = 0.f;
        if (MODE == EMRT_LOC_PIXEL_OFFSET) {
          const float2 r = __ldg(reinterpret_cast<const float2*>(rp + 2 * l));
          rxp = r.x * fW - 0.5f;
          ryp = r.y * fH - 0.5f;
        }
        const float sx = (MODE == EMRT_LOC_PIXEL_OFFSET) ? 1.f : fW;
        const float sy = (MODE == EMRT_LOC_PIXEL_OFFSET) ? 1.f : fH;
#pragma unroll
        for (int h = 0; h < BW_P; h += 3) {
          // three points at a time: their six value loads are in flight together
          float aw_[3], fx_[3], fy_[3];
          int xi_[3], yi_[3];
          bool vt_[3], vb_[3];
          uint4 tv_[3], bv_[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int j = l * BW_P + h + i, src = (lane & 24) + (j & 7);
            const float ox_ = __shfl_sync(0xffffffffu, px_[j >> 3], src), oy_ = __shfl_sync(0xffffffffu, py_[j >> 3], src);
            aw_[i] = __shfl_sync(0xffffffffu, pa_[j >> 3], src);
            float x, y;
            if (MODE == EMRT_LOC_PIXEL_OFFSET) { x = rxp + ox_; y = ryp + oy_; }
            else { x = ox_ * fW - 0.5f; y = oy_ * fH - 0.5f; }
            // a sample contributes iff -1 < x < W and -1 < y < H (make_footprint's rule; also rejects NaN); NaN / -inf go to
            // -2, huge values stay convertible
            const bool live = (x > -1.f) && (y > -1.f);
            const float xs = fminf(fmaxf(x, -2.f), 16777216.f), ys = fminf(fmaxf(y, -2.f), 16777216.f);
            xi_[i] = __float2int_rd(xs); yi_[i] = __float2int_rd(ys);
            fx_[i] = xs - (float)xi_[i]; fy_[i] = ys - (float)yi_[i];
            const int cx = xi_[i] + side;
            const bool vx = live && (unsigned)cx < (unsigned)W;
            vt_[i] = vx && (unsigned)yi_[i] < (unsigned)H;
            vb_[i] = vx && (unsigned)(yi_[i] + 1) < (unsigned)H;
            const int pix_t = (yi_[i] * W + cx) * pix_stride;
            tv_[i] = make_uint4(0, 0, 0, 0); bv_[i] = make_uint4(0, 0, 0, 0);
            if (vt_[i]) tv_[i] = __ldg(reinterpret_cast<const uint4*>(vptr + (uint32_t)pix_t));
            if (vb_[i]) bv_[i] = __ldg(reinterpret_cast<const uint4*>(vptr + (uint32_t)(pix_t + W * pix_stride)));
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int pt = h + i;
            const float aw = aw_[i], fx = fx_[i], fy = fy_[i];
            const int xi = xi_[i], yi = yi_[i];
            const float d_t = bw_dot8(gnat, tv_[i]), d_b = bw_dot8(gnat, bv_[i]);
            const float gy1 = 1.f - fy, wxs = side ? fx : 1.f - fx;
            const float along = gy1 * d_t + fy * d_b;
            float ga = wxs * along, gxp = sgn * along, gyp = wxs * (d_b - d_t);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              ga += __shfl_xor_sync(0xffffffffu, ga, o);
              gxp += __shfl_xor_sync(0xffffffffu, gxp, o);
              gyp += __shfl_xor_sync(0xffffffffu, gyp, o);
            }
            if (s == 0) {
              ga_out[l * BW_P + pt] = ga;
              *reinterpret_cast<float2*>(gl + (l * BW_P + pt) * 2) = make_float2(aw * sx * gxp, aw * sy * gyp);
            }
            // ---- grad_value: w * grad_out into the window (fixed point) or, outside it, into global memory ------------
            const int wx = xi - ox[l], wy = yi - oy[l];
            const float wa = aw * wxs;
            if ((unsigned)wx < (unsigned)(WWl - 1) && (unsigned)wy < (unsigned)(WHl - 1)) {
              const uint32_t wpix = (uint32_t)(wy * WWl + wx + side);
              const uint32_t a_t = wbase + wpix * (BW_D * 4);
              const uint32_t a_b = a_t + (uint32_t)WWl * (BW_D * 4);
              const float wst = wa * gy1 * scale, wsb = wa * fy * scale;
              // the raw bits of (x + MAGIC) are MAGIC_BITS + round(x): the constant is taken out at the flush, from the
              // pixel's contribution count (one red.shared per (query, side, row) instead of one subtraction per element)
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                bw_red_s32<PLAIN>(a_t + koff[k], __float_as_int(fmaf(gr[k], wst, BW_MAGIC)));
                bw_red_s32<PLAIN>(a_b + koff[k], __float_as_int(fmaf(gr[k], wsb, BW_MAGIC)));
              }
              if (sub == 0) {
                const uint32_t c_t = cbase + wpix * 4;
                bw_red_s32<PLAIN>(c_t, 1);
                bw_red_s32<PLAIN>(c_t + (uint32_t)WWl * 4, 1);
              }
            } else {
              const int pix_t = (yi * W + xi + side) * pix_stride;
              if (vt_[i]) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(gvptr + pix_t + (koff[k] >> 2), gr[k] * (wa * gy1));
              }
              if (vb_[i]) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(gvptr + pix_t + W * pix_stride + (koff[k] >> 2), gr[k] * (wa * fy));
              }
            }
          }
        }
      }
      int next = 0;
      if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(next) : "r"(smem_u32(&s_next)) : "memory");
      batch = __shfl_sync(0xffffffffu, next, 0);
    }
    __syncthreads();

    // ---- flush: one vector reduction per (in-map window pixel, 4 channels) that received anything ---------------------
#pragma unroll
    for (int l = 0; l < BW_L; ++l) {
      if ((l == 0) != (pass == 0)) continue;
      const int WWl = p.WW[l], npx = WWl * p.WH[l], W = p.lv.W[l], H = p.lv.H[l];
      float* gbase = grad_value + (((int64_t)b * p.Lv + p.lv.start[l]) * p.M + m) * BW_D;
      for (int idx = threadIdx.x; idx < npx * 8; idx += BW_WARPS * 32) {
        const int pix = idx >> 3, c4 = idx & 7;
        const int wyp = pix / WWl, wxp = pix - wyp * WWl;
        const int X = ox[l] + wxp, Y = oy[l] + wyp;
        if ((unsigned)X < (unsigned)W && (unsigned)Y < (unsigned)H) {
          uint32_t cnt;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cnt) : "r"(smem_base + p.cnt_off[l] + (uint32_t)pix * 4));
          if (cnt) {
            const uint4 v = bw_lds128(smem_base + p.win_off[l] + (uint32_t)pix * (BW_D * 4) + c4 * 16);
            const uint32_t off = cnt * (uint32_t)BW_MAGIC_BITS;     // modulo 2^32, like the sums
            bw_red_add_v4(gbase + (int64_t)(Y * W + X) * pix_stride + c4 * 4, (float)(int)(v.x - off) * inv_scale,
                          (float)(int)(v.y - off) * inv_scale, (float)(int)(v.z - off) * inv_scale,
                          (float)(int)(v.w - off) * inv_scale);
          }
        }
      }
    }
    __syncthreads();
  }
}

// =====================================================================================================================
// v2: the forward window kernel's structure applied to the backward.  One CTA per SM (16 warps) owns (head, region, batch):
//   * the region's three value windows are staged once by TMA — a 5-D tensor map {32 ch, M, W, H, B} cuts one head's
//     window out of the pixel-major value tensor the training path keeps; out-of-map pixels are zero-filled, which is
//     grid_sample's zeros padding, so the dot products need no validity logic;
//   * stage A: one lane per (query, point) computes the footprint once into a 16-byte record {window pixel | slow / dead
//     code, fx, fy, attention weight}; stage B: the query's 8 lanes read the record (one broadcast LDS.128), their
//     16-byte slices of the top / bottom pixel from the value window, form the dot products for grad_loc / grad_attn and
//     scatter w * grad_out into the fixed-point window exactly as v1 does.
// Shared memory: fixed-point window + counters of one pass (96 KB, passes: level 0; levels 1 + 2), the three bf16 value
// windows (91 KB), the CTA's grad_out rows (10.5 KB), records (1.1 KB per warp): 216 KB.
constexpr uint32_t BW2_SLOW = 0x80000000u, BW2_DEAD = 0xffffffffu;

struct Bwd2Params {
  CUtensorMap tmap[BW_L];         // level l: {32, M, W_l, H_l, B} bf16 over the pixel-major value tensor, box {32,1,WW,WH,1}
  int32_t WW[BW_L], WH[BW_L];
  uint32_t val_off[BW_L];         // bf16 value windows
  uint32_t win_off[BW_L], cnt_off[BW_L], pass_bytes[2];
  uint32_t go_off, rec_off;
  int32_t R, TH, TW, tw_shift, regions_x, regions_y;
  int32_t Lq, Lv, M;
  LevelTable lv;
  int8_t cshift[BW_MAX_M][BW_L][2];
};

__device__ __forceinline__ void bw_tma_load_5d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

template <typename TL, int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
msda_gather_bwd_win2_kernel(const __nv_bfloat16* __restrict__ grad_out, const __nv_bfloat16* __restrict__ value,
                            const TL* __restrict__ loc, const TL* __restrict__ attn, const float* __restrict__ ref,
                            int64_t ref_bs, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                            float* __restrict__ grad_attn, const __grid_constant__ Bwd2Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_bar;
  __shared__ float s_max[NW];
  __shared__ int s_next;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x, region = blockIdx.y, b = blockIdx.z;
  const int ry = region / p.regions_x, rx = region - ry * p.regions_x;
  const uint32_t smem_base = smem_u32(smem);

  int ox[BW_L], oy[BW_L], qb[BW_L];
#pragma unroll
  for (int l = 0; l < BW_L; ++l) {
    const int x0 = (rx * p.TW) >> l, y0 = (ry * p.TH) >> l;
    const int cx = m < BW_MAX_M ? p.cshift[m][l][0] : 0, cy = m < BW_MAX_M ? p.cshift[m][l][1] : 0;
    ox[l] = min(max(x0 - p.R + cx, -1), p.lv.W[l] + 1 - p.WW[l]);
    oy[l] = min(max(y0 - p.R + cy, -1), p.lv.H[l] + 1 - p.WH[l]);
    qb[l] = p.lv.start[l] + y0 * p.lv.W[l] + x0;
  }
  const int n0 = (p.TH * p.TW) >> 2, n01 = n0 + (n0 >> 2), n_batches = n01 + (n0 >> 4), sh0 = p.tw_shift - 2;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
#pragma unroll
    for (int l = 0; l < BW_L; ++l) bytes += (uint32_t)(p.WW[l] * p.WH[l]) * (BW_D * 2);
    mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
    for (int l = 0; l < BW_L; ++l) bw_tma_load_5d(smem_base + p.val_off[l], &p.tmap[l], &s_bar, 0, m, ox[l], oy[l], b);
  }

  // ---- the CTA's grad_out rows (head m of its queries) into shared memory, and their largest magnitude ---------------
  float mx = 0.f;
  for (int idx = threadIdx.x; idx < n_batches * (BW_QPB * 4); idx += NW * 32) {
    const int k = idx >> 2, part = idx & 3;
    const bool l1 = (k >> 2) >= n0, l2 = (k >> 2) >= n01;
    const int kk = (k >> 2) - (l2 ? n01 : (l1 ? n0 : 0));
    const int sh = sh0 - (l2 ? 2 : (l1 ? 1 : 0));
    const int yy = kk >> sh, xx = kk - (yy << sh);
    const int Wl = l2 ? p.lv.W[2] : (l1 ? p.lv.W[1] : p.lv.W[0]);
    const int q = (l2 ? qb[2] : (l1 ? qb[1] : qb[0])) + yy * Wl + (xx << 2) + (k & 3);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(grad_out + (((int64_t)b * p.Lq + q) * p.M + m) * BW_D + part * 8));
    bw_sts128(smem_base + p.go_off + (uint32_t)idx * 16, v);
    mx = bw_absmax8(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < NW; ++w) mx = fmaxf(mx, s_max[w]);
  const uint32_t ebits = (__float_as_uint(mx) >> 23) & 0xffu;
  const bool scaled = ebits >= 20u && ebits < 255u;
  const float scale = scaled ? __uint_as_float((273u - ebits) << 23) : 1.f;
  const float inv_scale = scaled ? __uint_as_float((ebits - 19u) << 23) : 1.f;

  const int g = lane >> 3, s = lane & 7, side = s >> 2, sub = s & 3;
  const int rot = 2 * g + side;
  const int pix_stride = p.M * BW_D;
  const float sgn = side ? 1.f : -1.f;
  const uint32_t rec_base = smem_base + p.rec_off + (uint32_t)warp * (BW_QPB * BW_LP * 16);
  // stage-A identity: lane -> (query a_qi = lane % 4, point a_pp = lane / 4), lanes 24-31 idle
  const bool a_on = lane < BW_QPB * BW_P;
  const int a_qi = a_on ? (lane & 3) : 0, a_pp = a_on ? (lane >> 2) : 0;
  uint32_t koff[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) koff[k] = (uint32_t)((k + rot) & 7) * 4;
  bool windows_ready = false;

  for (int pass = 0; pass < 2; ++pass) {
    const uint32_t pass_bytes = p.pass_bytes[pass];
    for (uint32_t o = threadIdx.x * 16; o < pass_bytes; o += NW * 32 * 16) bw_sts128(smem_base + o, make_uint4(0, 0, 0, 0));
    if (threadIdx.x == 0) s_next = NW;
    __syncthreads();
    if (!windows_ready) {
      mbar_wait(&s_bar, 0);
      windows_ready = true;
    }

    for (int batch = warp; batch < n_batches;) {
      const bool l1 = batch >= n0, l2 = batch >= n01;
      const int kk = batch - (l2 ? n01 : (l1 ? n0 : 0));
      const int sh = sh0 - (l2 ? 2 : (l1 ? 1 : 0));
      const int yy = kk >> sh, xx = kk - (yy << sh);
      const int Wq = l2 ? p.lv.W[2] : (l1 ? p.lv.W[1] : p.lv.W[0]);
      const int q_base = (l2 ? qb[2] : (l1 ? qb[1] : qb[0])) + yy * Wq + (xx << 2);

      // ---- stage A: footprint records of this pass's levels, one per lane ---------------------------------------------
      if (a_on) {
        const int64_t item = ((int64_t)b * p.Lq + (q_base + a_qi)) * p.M + m;
        const TL* lp = loc + item * (BW_LP * 2) + a_pp * 2;
        const TL* ap = attn + item * BW_LP + a_pp;
        const float* rp = (MODE == EMRT_LOC_PIXEL_OFFSET) ? ref + b * ref_bs + (int64_t)(q_base + a_qi) * (BW_L * 2) : nullptr;
#pragma unroll
        for (int l = 0; l < BW_L; ++l) {
          if ((l == 0) != (pass == 0)) continue;
          const float2 xy = Pair<TL>::load(lp + l * (BW_P * 2));
          const float aw = load1<TL>(ap + l * BW_P);
          const float fW = (float)p.lv.W[l], fH = (float)p.lv.H[l];
          float x, y;
          if (MODE == EMRT_LOC_PIXEL_OFFSET) {
            const float2 r = __ldg(reinterpret_cast<const float2*>(rp + 2 * l));
            x = r.x * fW - 0.5f + xy.x;
            y = r.y * fH - 0.5f + xy.y;
          } else {
            x = xy.x * fW - 0.5f;
            y = xy.y * fH - 0.5f;
          }
          // live iff -1 < x < W and -1 < y < H (make_footprint's rule; rejects NaN)
          const bool live = (x > -1.f) && (y > -1.f) && (x < fW) && (y < fH);
          const float xs = fmaxf(x, -2.f), ys = fmaxf(y, -2.f);
          const int xi = live ? __float2int_rd(xs) : 0, yi = live ? __float2int_rd(ys) : 0;
          const float fx = live ? xs - (float)xi : 0.f, fy = live ? ys - (float)yi : 0.f;
          const int wx = xi - ox[l], wy = yi - oy[l];
          const bool inwin = live && (unsigned)wx < (unsigned)(p.WW[l] - 1) && (unsigned)wy < (unsigned)(p.WH[l] - 1);
          const uint32_t w0 = inwin ? (uint32_t)(wy * p.WW[l] + wx)
                                    : (live ? (BW2_SLOW | ((uint32_t)(yi + 2) << 16) | (uint32_t)(xi + 2)) : BW2_DEAD);
          const uint32_t dst = rec_base + (uint32_t)(a_qi * BW_LP + l * BW_P + a_pp) * 16;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(w0), "r"(__float_as_uint(fx)),
                       "r"(__float_as_uint(fy)), "r"(__float_as_uint(live ? aw : 0.f)) : "memory");
        }
      }
      __syncwarp();

      // ---- stage B ---------------------------------------------------------------------------------------------------
      const int q = q_base + g;
      const int64_t item = ((int64_t)b * p.Lq + q) * p.M + m;
      const uint32_t grow = smem_base + p.go_off + (uint32_t)(batch * BW_QPB + g) * (BW_D * 2);
      const uint4 gnat = bw_lds128(grow + sub * 16);
      float gr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        unsigned short h;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(grow + (uint32_t)sub * 16 + (koff[k] >> 1)));
        gr[k] = __uint_as_float((uint32_t)h << 16);
      }
      float* gl = grad_loc + item * (BW_LP * 2);
      float* ga_out = grad_attn + item * BW_LP;
      const uint32_t my_rec = rec_base + (uint32_t)(g * BW_LP) * 16;

#pragma unroll
      for (int l = 0; l < BW_L; ++l) {
        if ((l == 0) != (pass == 0)) continue;
        const int W = p.lv.W[l], H = p.lv.H[l], WWl = p.WW[l];
        const uint32_t vbase = smem_base + p.val_off[l] + (uint32_t)side * (BW_D * 2) + (uint32_t)sub * 16;
        const uint32_t wbase = smem_base + p.win_off[l] + (uint32_t)side * (BW_D * 4) + (uint32_t)sub * 32;
        const uint32_t cbase = smem_base + p.cnt_off[l] + (uint32_t)side * 4;
        const int64_t lvl_off = (((int64_t)b * p.Lv + p.lv.start[l]) * p.M + m) * BW_D + sub * 8;
        const float sx = (MODE == EMRT_LOC_PIXEL_OFFSET) ? 1.f : (float)W;
        const float sy = (MODE == EMRT_LOC_PIXEL_OFFSET) ? 1.f : (float)H;
#pragma unroll
        for (int h = 0; h < BW_P; h += 3) {
          uint4 rec[3], tv[3], bv[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) rec[i] = bw_lds128(my_rec + (uint32_t)(l * BW_P + h + i) * 16);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            tv[i] = make_uint4(0, 0, 0, 0); bv[i] = make_uint4(0, 0, 0, 0);
            if (rec[i].x < BW2_SLOW) {
              const uint32_t a = vbase + rec[i].x * (BW_D * 2);
              tv[i] = bw_lds128(a);
              bv[i] = bw_lds128(a + (uint32_t)WWl * (BW_D * 2));
            } else if (rec[i].x != BW2_DEAD) {
              // left the window but not the map: this lane's corners from global memory, with their validity
              const int xi = (int)(rec[i].x & 0xffffu) - 2, yi = (int)((rec[i].x >> 16) & 0x7fffu) - 2;
              const int cx = xi + side;
              const bool vx = (unsigned)cx < (unsigned)W;
              const __nv_bfloat16* vptr = value + lvl_off;
              const int pix_t = (yi * W + cx) * pix_stride;
              if (vx && (unsigned)yi < (unsigned)H) tv[i] = __ldg(reinterpret_cast<const uint4*>(vptr + pix_t));
              if (vx && (unsigned)(yi + 1) < (unsigned)H) bv[i] = __ldg(reinterpret_cast<const uint4*>(vptr + pix_t + W * pix_stride));
            }
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int pt = h + i;
            const float fx = __uint_as_float(rec[i].y), fy = __uint_as_float(rec[i].z), aw = __uint_as_float(rec[i].w);
            const float d_t = bw_dot8(gnat, tv[i]), d_b = bw_dot8(gnat, bv[i]);
            const float gy1 = 1.f - fy, wxs = side ? fx : 1.f - fx;
            const float along = gy1 * d_t + fy * d_b;
            float ga = wxs * along, gxp = sgn * along, gyp = wxs * (d_b - d_t);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              ga += __shfl_xor_sync(0xffffffffu, ga, o);
              gxp += __shfl_xor_sync(0xffffffffu, gxp, o);
              gyp += __shfl_xor_sync(0xffffffffu, gyp, o);
            }
            if (s == 0) {
              ga_out[l * BW_P + pt] = ga;
              *reinterpret_cast<float2*>(gl + (l * BW_P + pt) * 2) = make_float2(aw * sx * gxp, aw * sy * gyp);
            }
            const float wa = aw * wxs;
            if (rec[i].x < BW2_SLOW) {
              const uint32_t a_t = wbase + rec[i].x * (BW_D * 4);
              const uint32_t a_b = a_t + (uint32_t)WWl * (BW_D * 4);
              const float wst = wa * gy1 * scale, wsb = wa * fy * scale;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                bw_red_s32<false>(a_t + koff[k], __float_as_int(fmaf(gr[k], wst, BW_MAGIC)));
                bw_red_s32<false>(a_b + koff[k], __float_as_int(fmaf(gr[k], wsb, BW_MAGIC)));
              }
              if (sub == 0) {
                const uint32_t c_t = cbase + rec[i].x * 4;
                bw_red_s32<false>(c_t, 1);
                bw_red_s32<false>(c_t + (uint32_t)WWl * 4, 1);
              }
            } else if (rec[i].x != BW2_DEAD) {
              const int xi = (int)(rec[i].x & 0xffffu) - 2, yi = (int)((rec[i].x >> 16) & 0x7fffu) - 2;
              const int cx = xi + side;
              const bool vx = (unsigned)cx < (unsigned)W;
              float* gvptr = grad_value + lvl_off;
              const int pix_t = (yi * W + cx) * pix_stride;
              if (vx && (unsigned)yi < (unsigned)H) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(gvptr + pix_t + (koff[k] >> 2), gr[k] * (wa * gy1));
              }
              if (vx && (unsigned)(yi + 1) < (unsigned)H) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(gvptr + pix_t + W * pix_stride + (koff[k] >> 2), gr[k] * (wa * fy));
              }
            }
          }
        }
      }
      int next = 0;
      if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(next) : "r"(smem_u32(&s_next)) : "memory");
      batch = __shfl_sync(0xffffffffu, next, 0);
      __syncwarp();                                  // this batch's record reads before the next batch's record writes
    }
    __syncthreads();

    // ---- flush ---------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int l = 0; l < BW_L; ++l) {
      if ((l == 0) != (pass == 0)) continue;
      const int WWl = p.WW[l], npx = WWl * p.WH[l], W = p.lv.W[l], H = p.lv.H[l];
      float* gbase = grad_value + (((int64_t)b * p.Lv + p.lv.start[l]) * p.M + m) * BW_D;
      for (int idx = threadIdx.x; idx < npx * 8; idx += NW * 32) {
        const int pix = idx >> 3, c4 = idx & 7;
        const int wyp = pix / WWl, wxp = pix - wyp * WWl;
        const int X = ox[l] + wxp, Y = oy[l] + wyp;
        if ((unsigned)X < (unsigned)W && (unsigned)Y < (unsigned)H) {
          uint32_t cnt;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cnt) : "r"(smem_base + p.cnt_off[l] + (uint32_t)pix * 4));
          if (cnt) {
            const uint4 v = bw_lds128(smem_base + p.win_off[l] + (uint32_t)pix * (BW_D * 4) + c4 * 16);
            const uint32_t off = cnt * (uint32_t)BW_MAGIC_BITS;
            bw_red_add_v4(gbase + (int64_t)(Y * W + X) * pix_stride + c4 * 4, (float)(int)(v.x - off) * inv_scale,
                          (float)(int)(v.y - off) * inv_scale, (float)(int)(v.z - off) * inv_scale,
                          (float)(int)(v.w - off) * inv_scale);
          }
        }
      }
    }
    __syncthreads();
  }
}

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

static int bw_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <typename TL, int MODE, int BW_WARPS, bool PLAIN = false>
static int launch_bwd_win(const void* go, const void* value, const void* loc, const void* attn, const float* ref,
                          int64_t ref_bs, float* gv, float* gl, float* ga, int B, const BwdWinParams& p,
                          size_t smem_bytes, cudaStream_t st) {
  auto kern = msda_gather_bwd_win_kernel<TL, MODE, BW_WARPS, PLAIN>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));   // per context
  const int n_regions = p.regions_x * p.regions_y;
  if (B > 65535 || n_regions > 65535) return EMRT_ERR_UNSUPPORTED;
  kern<<<dim3((unsigned)p.M, (unsigned)n_regions, (unsigned)B), BW_WARPS * 32, smem_bytes, st>>>(
      (const __nv_bfloat16*)go, (const __nv_bfloat16*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, gv, gl, ga, p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

template <typename TL, int MODE, int NW>
static int launch_bwd_win2(const void* go, const void* value, const void* loc, const void* attn, const float* ref,
                           int64_t ref_bs, float* gv, float* gl, float* ga, int B, const Bwd2Params& p, size_t smem_bytes,
                           cudaStream_t st) {
  auto kern = msda_gather_bwd_win2_kernel<TL, MODE, NW>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));   // per context
  const int n_regions = p.regions_x * p.regions_y;
  if (B > 65535 || n_regions > 65535) return EMRT_ERR_UNSUPPORTED;
  kern<<<dim3((unsigned)p.M, (unsigned)n_regions, (unsigned)B), NW * 32, smem_bytes, st>>>(
      (const __nv_bfloat16*)go, (const __nv_bfloat16*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, gv, gl, ga, p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

// v2 launcher: EMRT_ERR_UNSUPPORTED when the windows do not fit one CTA's shared memory (the caller then tries v1).
static int gather_bwd_win2(const void* go, const void* value, const void* loc, const void* attn, const float* ref,
                           int64_t ref_bs, float* gv, float* gl, float* ga, int B, int Lq, int Lv, int M,
                           const LevelTable& lv, int loc_dtype, int mode, const int32_t* win_center_host, cudaStream_t st) {
  constexpr int L = BW_L;
  Bwd2Params p;
  memset(&p, 0, sizeof(p));
  p.R = bw_env_int("EMRT_BWD_WIN_R", 7);
  p.TH = bw_env_int("EMRT_BWD_WIN_TH", 8);
  p.TW = bw_env_int("EMRT_BWD_WIN_TW", 16);
  if (p.R < 1 || p.TH < 4 || p.TW < 16 || (p.TH & 3) || (p.TW & (p.TW - 1)) || lv.H[0] % p.TH || lv.W[0] % p.TW) return EMRT_ERR_UNSUPPORTED;
  for (p.tw_shift = 0; (1 << p.tw_shift) < p.TW; ++p.tw_shift) {}
  if ((int64_t)Lv * M * BW_D >= (1LL << 30) || lv.W[0] > 32000 || lv.H[0] > 32000) return EMRT_ERR_UNSUPPORTED;
  p.regions_x = lv.W[0] / p.TW;
  p.regions_y = lv.H[0] / p.TH;
  p.Lq = Lq; p.Lv = Lv; p.M = M; p.lv = lv;
  if (win_center_host && M <= BW_MAX_M && !getenv("EMRT_WIN_NO_HINT"))
    for (int mm = 0; mm < M; ++mm)
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < 2; ++k)
          p.cshift[mm][l][k] = (int8_t)std::min(std::max(win_center_host[(mm * L + l) * 2 + k], -100), 100);
  uint32_t px[BW_L];
  for (int l = 0; l < L; ++l) {
    p.WW[l] = std::min((p.TW >> l) + 2 * p.R + 1, lv.W[l] + 2);
    p.WH[l] = std::min((p.TH >> l) + 2 * p.R + 1, lv.H[l] + 2);
    if (p.WW[l] > 256 || p.WH[l] > 256) return EMRT_ERR_UNSUPPORTED;
    px[l] = (uint32_t)(p.WW[l] * p.WH[l]);
  }
  // region A (re-used by the two passes): pass 0 [window 0][counters 0]; pass 1 [window 1][window 2][counters 1][counters 2]
  p.win_off[0] = 0;
  p.cnt_off[0] = px[0] * (BW_D * 4);
  p.pass_bytes[0] = (p.cnt_off[0] + px[0] * 4 + 15u) & ~15u;
  p.win_off[1] = 0;
  p.win_off[2] = px[1] * (BW_D * 4);
  p.cnt_off[1] = (px[1] + px[2]) * (BW_D * 4);
  p.cnt_off[2] = p.cnt_off[1] + px[1] * 4;
  p.pass_bytes[1] = (p.cnt_off[2] + px[2] * 4 + 15u) & ~15u;
  uint32_t off = (std::max(p.pass_bytes[0], p.pass_bytes[1]) + 127u) & ~127u;
  for (int l = 0; l < L; ++l) {
    p.val_off[l] = off;
    off += (px[l] * (BW_D * 2) + 127u) & ~127u;
    // one head's plane of level l inside the pixel-major [B, Lv, M, 32] tensor
    const uint64_t dims[5] = {(uint64_t)BW_D, (uint64_t)M, (uint64_t)lv.W[l], (uint64_t)lv.H[l], (uint64_t)B};
    const uint64_t strides[4] = {(uint64_t)BW_D * 2, (uint64_t)M * BW_D * 2, (uint64_t)lv.W[l] * M * BW_D * 2,
                                 (uint64_t)Lv * M * BW_D * 2};
    const uint32_t box[5] = {(uint32_t)BW_D, 1u, (uint32_t)p.WW[l], (uint32_t)p.WH[l], 1u};
    const __nv_bfloat16* base = (const __nv_bfloat16*)value + (int64_t)lv.start[l] * M * BW_D;
    if (int e = make_tensor_map(&p.tmap[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box,
                                CU_TENSOR_MAP_SWIZZLE_NONE))
      return e;
  }
  p.go_off = off;
  const int n_queries = p.TH * p.TW + ((p.TH * p.TW) >> 2) + ((p.TH * p.TW) >> 4);
  off += (uint32_t)n_queries * (BW_D * 2);
  p.rec_off = (off + 15u) & ~15u;
  const int warps = bw_env_int("EMRT_BWD_WIN_WARPS", 16) == 12 ? 12 : 16;
  const size_t smem_bytes = (size_t)p.rec_off + (size_t)warps * BW_QPB * BW_LP * 16;
  if (smem_bytes > 227 * 1024) return EMRT_ERR_UNSUPPORTED;
  const bool pxm = (mode & EMRT_LOC_PIXEL_OFFSET) != 0;
#define EMRT_BWIN2(TL)                                                                                             \
  if (warps == 12)                                                                                                 \
    return pxm ? launch_bwd_win2<TL, 1, 12>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st)    \
               : launch_bwd_win2<TL, 0, 12>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st);   \
  return pxm ? launch_bwd_win2<TL, 1, 16>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st)      \
             : launch_bwd_win2<TL, 0, 16>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st)
  switch (loc_dtype) {
    case EMRT_F32: EMRT_BWIN2(float);
    case EMRT_F16: EMRT_BWIN2(__half);
    case EMRT_BF16: EMRT_BWIN2(__nv_bfloat16);
    default: return EMRT_ERR_UNSUPPORTED;
  }
#undef EMRT_BWIN2
}

// Returns EMRT_ERR_UNSUPPORTED (error text untouched) when the shape is not the regular 3-level pyramid this kernel
// tiles; the caller then runs the generic backward.
int gather_bwd_win(const void* go, const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs,
                   float* gv, float* gl, float* ga, int B, int Lq, int Lv, int M, int D, int L, int P,
                   const LevelTable& lv, int loc_dtype, int mode, const int32_t* win_center_host, cudaStream_t st) {
  if (D != BW_D || L != BW_L || P != BW_P || Lq != Lv || (mode & EMRT_VALUE_HEAD_MAJOR)) return EMRT_ERR_UNSUPPORTED;
  for (int l = 1; l < L; ++l)
    if (lv.H[l] != (lv.H[0] >> l) || lv.W[l] != (lv.W[0] >> l) || (lv.H[l] << l) != lv.H[0] || (lv.W[l] << l) != lv.W[0])
      return EMRT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(value) & 15) != 0 || (reinterpret_cast<uintptr_t>(go) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(gv) & 15) != 0)
    return EMRT_ERR_UNSUPPORTED;
  if (!getenv("EMRT_BWD_WIN_V1")) {
    const int e = gather_bwd_win2(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, Lq, Lv, M, lv, loc_dtype, mode,
                                  win_center_host, st);
    if (e != EMRT_ERR_UNSUPPORTED) return e;
  }
  BwdWinParams p;
  memset(&p, 0, sizeof(p));
  p.R = bw_env_int("EMRT_BWD_WIN_R", 7);
  p.TH = bw_env_int("EMRT_BWD_WIN_TH", 8);
  p.TW = bw_env_int("EMRT_BWD_WIN_TW", 16);
  if (p.R < 1 || p.TH < 4 || p.TW < 16 || (p.TH & 3) || (p.TW & (p.TW - 1)) || lv.H[0] % p.TH || lv.W[0] % p.TW) return EMRT_ERR_UNSUPPORTED;
  for (p.tw_shift = 0; (1 << p.tw_shift) < p.TW; ++p.tw_shift) {}
  if ((int64_t)B * Lv * M * D >= (1LL << 40)) return EMRT_ERR_UNSUPPORTED;
  p.regions_x = lv.W[0] / p.TW;
  p.regions_y = lv.H[0] / p.TH;
  p.Lq = Lq; p.Lv = Lv; p.M = M; p.lv = lv;
  if (win_center_host && M <= BW_MAX_M && !getenv("EMRT_WIN_NO_HINT"))
    for (int mm = 0; mm < M; ++mm)
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < 2; ++k)
          p.cshift[mm][l][k] = (int8_t)std::min(std::max(win_center_host[(mm * L + l) * 2 + k], -100), 100);
  uint32_t bytes[BW_L];
  for (int l = 0; l < L; ++l) {
    p.WW[l] = std::min((p.TW >> l) + 2 * p.R + 1, lv.W[l] + 2);
    p.WH[l] = std::min((p.TH >> l) + 2 * p.R + 1, lv.H[l] + 2);
    bytes[l] = (uint32_t)(p.WW[l] * p.WH[l]) * (BW_D * 4);
  }
  // pass 0: [window 0][counters 0]; pass 1: [window 1][window 2][counters 1][counters 2]; both start at 0
  p.win_off[0] = 0;
  p.cnt_off[0] = bytes[0];
  p.pass_bytes[0] = (bytes[0] + bytes[0] / BW_D + 15u) & ~15u;
  p.win_off[1] = 0;
  p.win_off[2] = bytes[1];
  p.cnt_off[1] = bytes[1] + bytes[2];
  p.cnt_off[2] = p.cnt_off[1] + bytes[1] / BW_D;
  p.pass_bytes[1] = (p.cnt_off[2] + bytes[2] / BW_D + 15u) & ~15u;
  p.go_off = (std::max(p.pass_bytes[0], p.pass_bytes[1]) + 127u) & ~127u;
  if ((int64_t)Lv * M * D >= (1LL << 30)) return EMRT_ERR_UNSUPPORTED;   // 32-bit element offsets inside one level plane
  const int n_queries = p.TH * p.TW + ((p.TH * p.TW) >> 2) + ((p.TH * p.TW) >> 4);
  const size_t smem_bytes = (size_t)p.go_off + (size_t)n_queries * (BW_D * 2);
  if (smem_bytes > 227 * 1024) return EMRT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(value) & 15) != 0 || (reinterpret_cast<uintptr_t>(go) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(gv) & 15) != 0)
    return EMRT_ERR_UNSUPPORTED;
  const bool px = (mode & EMRT_LOC_PIXEL_OFFSET) != 0;
  const int warps = bw_env_int("EMRT_BWD_WIN_WARPS", 8) == 10 ? 10 : 8;
  if (loc_dtype == EMRT_F16 && px && getenv("EMRT_BWD_WIN_TIMING_PLAIN_STORES"))
    return launch_bwd_win<__half, 1, 8, true>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st);
#define EMRT_BWIN(TL)                                                                                              \
  if (warps == 10)                                                                                                 \
    return px ? launch_bwd_win<TL, 1, 10>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st)      \
              : launch_bwd_win<TL, 0, 10>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st);     \
  return px ? launch_bwd_win<TL, 1, 8>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st)         \
            : launch_bwd_win<TL, 0, 8>(go, value, loc, attn, ref, ref_bs, gv, gl, ga, B, p, smem_bytes, st)
  switch (loc_dtype) {
    case EMRT_F32: EMRT_BWIN(float);
    case EMRT_F16: EMRT_BWIN(__half);
    case EMRT_BF16: EMRT_BWIN(__nv_bfloat16);
    default: return EMRT_ERR_UNSUPPORTED;
  }
#undef EMRT_BWIN
}

}  // namespace emrt
