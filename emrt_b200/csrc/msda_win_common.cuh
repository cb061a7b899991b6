// Pieces shared by the window-staged forward gathers (msda_gather_win.cu: run-time geometry; msda_gather_win7.cu: the
// default geometry as compile-time constants).
#pragma once
#include "msda_common.cuh"
#include "tc_common.cuh"

namespace emrt {

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

constexpr int WIN_L = 3, WIN_P = 6, WIN_LP = WIN_L * WIN_P, WIN_D = 32;
constexpr int WIN_MAX_M = 16;                // heads a window-centre hint can describe
constexpr int WIN_QPB = 4;                   // queries per warp batch (one LDS.128 serves 4 queries x 128 bytes)
constexpr uint32_t WIN_SLOW = 0x80000000u;   // record flag (sign bit of the bottom weight, weights are >= 0):
                                             // take the global-memory path for this point

struct WinParams {
  CUtensorMap tmap[WIN_L];       // level l: {32 ch, W_l, H_l, B*M} bf16, box {32, WW_l, WH_l, 1}, zero OOB fill
  int32_t WW[WIN_L], WH[WIN_L];  // window size in pixels
  uint32_t win_off[WIN_L];       // byte offset of window l in dynamic shared memory (offset 0 holds 128 zero bytes)
  uint32_t rec_off;              // byte offset of the footprint records
  int32_t R, TH, TW, tw_shift, regions_x, regions_y;
  int8_t cshift[WIN_MAX_M][WIN_L][2];   // per (head, level): window centre shift (x, y) in level pixels — a locality hint
  int32_t Lq, Lv, M;
  int32_t pixel_major;           // value is [B,Lv,M,D] (5-D tensor maps, head = coordinate 1) instead of head-major [B,M,Lv,D]
  LevelTable lv;
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// acc += bf16(half of word) * bf16(half of wpair), fp32 accumulate (FHFMA.BF16)
template <int HI, int WHI>
__device__ __forceinline__ void fhfma(float& acc, uint32_t word, uint32_t wpair) {
  unsigned short a_lo, a_hi, w_lo, w_hi;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(a_lo), "=h"(a_hi) : "r"(word));
  asm("mov.b32 {%0,%1}, %2;" : "=h"(w_lo), "=h"(w_hi) : "r"(wpair));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(HI ? a_hi : a_lo), "h"(WHI ? w_hi : w_lo));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 8 channels of one pixel row pair: acc[2i] += lo(d[i]) * w, acc[2i+1] += hi(d[i]) * w, with w = half WHI of wpair
template <int WHI>
__device__ __forceinline__ void fma_row(float (&acc)[8], const uint4& d, uint32_t wpair) {
  fhfma<0, WHI>(acc[0], d.x, wpair); fhfma<1, WHI>(acc[1], d.x, wpair);
  fhfma<0, WHI>(acc[2], d.y, wpair); fhfma<1, WHI>(acc[3], d.y, wpair);
  fhfma<0, WHI>(acc[4], d.z, wpair); fhfma<1, WHI>(acc[5], d.z, wpair);
  fhfma<0, WHI>(acc[6], d.w, wpair); fhfma<1, WHI>(acc[7], d.w, wpair);
}

// First global query index of a warp batch.  A region's queries are numbered level 0 first (TH x TW pixels, row-major),
// then the co-located (TH/2 x TW/2) pixels of level 1, then level 2; TW is a power of two >= 16 (host-checked), so the
// WIN_QPB = 4 queries of a batch are consecutive pixels of one row and the batch count is exact (no padding).
// Warp-uniform selects and shifts only: no per-CTA lookup table to build.  qb[l] = index of the region's first query
// of level l, n0 / n01 = batch counts of level 0 / levels 0 + 1, sh0 = log2(TW / 4).
__device__ __forceinline__ int batch_query_base(const WinParams& p, const int (&qb)[WIN_L], int n0, int n01, int sh0, int batch) {
  const bool l1 = batch >= n0, l2 = batch >= n01;
  const int k = batch - (l2 ? n01 : (l1 ? n0 : 0));
  const int sh = sh0 - (l2 ? 2 : (l1 ? 1 : 0));
  const int y = k >> sh, x = k - (y << sh);
  const int W = l2 ? p.lv.W[2] : (l1 ? p.lv.W[1] : p.lv.W[0]);
  return (l2 ? qb[2] : (l1 ? qb[1] : qb[0])) + y * W + (x << 2);
}

// Slow path of one point: it left the staged window (|offset| > R) but not the map.  Global loads with the explicit
// zero-padding weights of make_footprint, from the sample position stage A kept; out-of-line so the unrolled fast path
// stays small.
static __device__ __noinline__ void slow_point(const WinParams& p, const __nv_bfloat16* __restrict__ value, int b, int m, int l,
                                        float x, float y, float aw, int s, uint4* d0, uint4* d1, uint32_t* wp) {
  const int side = s >> 2;
  const Footprint f = make_footprint(x, y, p.lv.H[l], p.lv.W[l]);
  // first pixel of (b, level l, head m) and the distance between neighbouring pixels of one head, in elements
  const int64_t first = p.pixel_major ? (((int64_t)b * p.Lv + p.lv.start[l]) * p.M + m) * WIN_D
                                      : (((int64_t)b * p.M + m) * p.Lv + p.lv.start[l]) * WIN_D;
  const int64_t pix_bytes = (p.pixel_major ? p.M : 1) * (WIN_D * 2);
  const char* base = reinterpret_cast<const char*>(value + first) + (s & 3) * 16;
  *d0 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(side ? f.i01 : f.i00) * pix_bytes));
  *d1 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(side ? f.i11 : f.i10) * pix_bytes));
  *wp = side ? pack_bf16(f.w01 * aw, f.w11 * aw) : pack_bf16(f.w00 * aw, f.w10 * aw);
}

// Raw inputs of one (query, point) as they sit in global memory; fetched one batch ahead of their use.
template <typename TL> struct RawLoc;
template <> struct RawLoc<float> { float2 xy; float aw; };
template <> struct RawLoc<__half> { unsigned int xy; unsigned short aw; };
template <> struct RawLoc<__nv_bfloat16> { unsigned int xy; unsigned short aw; };
__device__ __forceinline__ void raw_fetch(RawLoc<float>& r, const float* lp, const float* ap) {
  r.xy = __ldg(reinterpret_cast<const float2*>(lp)); r.aw = __ldg(ap);
}
template <typename TL> __device__ __forceinline__ void raw_fetch(RawLoc<TL>& r, const TL* lp, const TL* ap) {
  r.xy = __ldg(reinterpret_cast<const unsigned int*>(lp));
  r.aw = __ldg(reinterpret_cast<const unsigned short*>(ap));
}
__device__ __forceinline__ void raw_decode(const RawLoc<float>& r, float& x, float& y, float& aw) { x = r.xy.x; y = r.xy.y; aw = r.aw; }
__device__ __forceinline__ void raw_decode(const RawLoc<__half>& r, float& x, float& y, float& aw) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r.xy));
  x = f.x; y = f.y; aw = __half2float(*reinterpret_cast<const __half*>(&r.aw));
}
__device__ __forceinline__ void raw_decode(const RawLoc<__nv_bfloat16>& r, float& x, float& y, float& aw) {
  x = __uint_as_float(r.xy << 16); y = __uint_as_float(r.xy & 0xffff0000u); aw = __uint_as_float((unsigned int)r.aw << 16);
}


}  // namespace emrt
