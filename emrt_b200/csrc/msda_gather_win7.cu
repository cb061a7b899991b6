// Window-staged forward gather with the DEFAULT geometry compiled in (region 8 x 16 level-0 pixels, halo R = 7, 3 levels x 6
// points, D = 32, 10 warps x 2 CTAs per SM) — the kernel the encoder self-attention of a 512^2 / 256^2 tile runs.
//
// Same algorithm, data flow and arithmetic as msda_gather_fwd_win_kernel (msda_gather_win.cu; read its header first).
// What changed is only how many instructions a batch of 4 queries costs.  The ncu source view of the run-time-geometry
// kernel (profiles/r2b_gather_win_lines.txt) attributes 762 warp-instructions to a batch: 288 FHFMA + 45 LDS that ARE the
// gather, and ~430 around them.  A pure LDS.128 + FHFMA microbenchmark of the same access pattern
// (scripts/microbench/lds_gather_floor.cu, profiles/r2a_lds_floor.txt) runs in 0.41 ms against the kernel's 0.77 ms, so the
// instructions around the gather, not the shared-memory pipe, set the pace.  Removed here:
//   * geometry as constants: window sizes / pitches / offsets, the batch -> query map and the record layout fold into
//     immediates (no LDCU / R2UR / UIADD3 traffic, `a + row_bytes` becomes an LDS immediate offset);
//   * windows are NOT clamped into the map any more: the tensor map zero-fills whatever lies outside, which is exactly
//     zeros padding, so the origin is plain arithmetic and the window sizes never depend on the image size;
//   * the liveness test of a sample runs only for the rare sample that left its window;
//   * the shared batch counter is claimed by an elect.sync-predicated atom (ptxas expands a uniform atom.add / atom.inc
//     into a 17-instruction warp-aggregation sequence, VOTEU / FLO / POPC / SHFL, even under a one-lane branch);
//   * per-CTA setup shrinks to what really depends on blockIdx (it was ~70 instructions per batch, amortised over only
//     4.2 batches per warp).
// A sample outside its staged window but inside the map still takes the global-memory slow path: results never depend on
// the geometry or on the hint.
#include <cstdlib>
#include <cstring>

#include "msda_win_common.cuh"

namespace emrt {

namespace w7 {
constexpr int R = 7, TH = 8, TW = 16, WARPS = 10;
constexpr int WW0 = TW + 2 * R + 1, WH0 = TH + 2 * R + 1;                  // 31 x 23
constexpr int WW1 = TW / 2 + 2 * R + 1, WH1 = TH / 2 + 2 * R + 1;          // 23 x 19
constexpr int WW2 = TW / 4 + 2 * R + 1, WH2 = TH / 4 + 2 * R + 1;          // 19 x 17
constexpr uint32_t align128(uint32_t x) { return (x + 127u) & ~127u; }
constexpr uint32_t OFF0 = 128;                                             // [0, 128): zero block
constexpr uint32_t OFF1 = OFF0 + align128(WW0 * WH0 * WIN_D * 2);
constexpr uint32_t OFF2 = OFF1 + align128(WW1 * WH1 * WIN_D * 2);
constexpr uint32_t REC_OFF = OFF2 + align128(WW2 * WH2 * WIN_D * 2);
constexpr uint32_t REC_BYTES = WIN_QPB * WIN_LP * 16;                      // per warp: {address, weight pair} x 2 sides
constexpr uint32_t XY_OFF = REC_OFF + WARPS * REC_BYTES;
constexpr uint32_t XY_BYTES = WIN_QPB * WIN_LP * 8;                        // per warp: slow-point positions
constexpr uint32_t SMEM_BYTES = XY_OFF + WARPS * XY_BYTES;
constexpr uint32_t WIN_TX_BYTES = (WW0 * WH0 + WW1 * WH1 + WW2 * WH2) * WIN_D * 2;
constexpr int N0 = TH * TW / 4, N1 = N0 / 4, N2 = N0 / 16, N_BATCHES = N0 + N1 + N2;   // 32 + 8 + 2 = 42
template <int L> struct Lv;
template <> struct Lv<0> { static constexpr int WW = WW0, WH = WH0; static constexpr uint32_t OFF = OFF0; };
template <> struct Lv<1> { static constexpr int WW = WW1, WH = WH1; static constexpr uint32_t OFF = OFF1; };
template <> struct Lv<2> { static constexpr int WW = WW2, WH = WH2; static constexpr uint32_t OFF = OFF2; };
}  // namespace w7

struct Win7Params {
  CUtensorMap tmap[WIN_L];       // level l: {32 ch, W_l, H_l, B*M} bf16 (head-major) or {32, M, W_l, H_l, B} (pixel-major)
  int8_t cshift[WIN_MAX_M][WIN_L][2];
  int32_t regions_x;
  int32_t Lq, Lv, M;
  int32_t pixel_major;
  LevelTable lv;
};

// One level of stage A for this lane's (query, point): footprint -> two 8-byte records {smem address, weight pair}.
template <int L, int MODE, typename TL>
__device__ __forceinline__ void stage_a_level(const RawLoc<TL>& raw, const float2& rref, float fW, float fH, int ox, int oy,
                                              uint32_t smem_base, uint32_t rdst, uint32_t xydst, unsigned& slow_lv) {
  using G = w7::Lv<L>;
  float x, y, aw;
  raw_decode(raw, x, y, aw);
  if (MODE == EMRT_LOC_PIXEL_OFFSET) {
    x = rref.x * fW - 0.5f + x;
    y = rref.y * fH - 0.5f + y;
  } else {
    x = x * fW - 0.5f;
    y = y * fH - 0.5f;
  }
  // NaN and -inf go to -1e4 (far below every window origin, still an exact int); +inf converts to INT_MAX: both fail the
  // unsigned window test below
  const float xs = fmaxf(x, -1.0e4f), ys = fmaxf(y, -1.0e4f);
  const int xi = __float2int_rd(xs), yi = __float2int_rd(ys);
  const float fx = xs - (float)xi, fy = ys - (float)yi;
  const int wx = xi - ox, wy = yi - oy;
  // both pixel pairs inside the window: everything the window holds outside the map is zero (tensor-map fill = zeros
  // padding), so no liveness logic is needed on this path
  const bool fast = (unsigned)wx < (unsigned)(G::WW - 1) && (unsigned)wy < (unsigned)(G::WH - 1);
  const float gx = 1.f - fx, gy = 1.f - fy;
  const float gxa = gx * aw, fxa = fx * aw;
  uint32_t addr = smem_base + G::OFF + (uint32_t)(wy * G::WW + wx) * (WIN_D * 2);
  uint32_t wl = pack_bf16(gxa * gy, gxa * fy);   // left pixel: top, bottom
  uint32_t wr = pack_bf16(fxa * gy, fxa * fy);   // right pixel: top, bottom
  if (!fast) {
    // outside the window.  Dead (all four corners outside the map; also NaN / inf): weight 0 on the zero block.  Live: the
    // zero block again, the SLOW flag in the sign bit of the bottom weight, the attention weight in the low half and the
    // sample position in the side buffer for the fix-up.
    const bool live = (x > -1.f) && (y > -1.f) && (x < fW) && (y < fH);
    addr = smem_base;
    wl = wr = live ? (WIN_SLOW | (pack_bf16(aw, 0.f) & 0xffffu)) : 0u;
    if (live) {
      asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(xydst + L * (WIN_P * 8)), "f"(x), "f"(y) : "memory");
      slow_lv |= 1u << L;
    }
  }
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(rdst + L * (WIN_P * 8)), "r"(addr), "r"(wl) : "memory");
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(rdst + L * (WIN_P * 8) + WIN_LP * 8), "r"(addr), "r"(wr) : "memory");
}

// One level of stage B: six records, twelve LDS.128 (two halves of three points), 96 FHFMA.
template <int L>
__device__ __forceinline__ void stage_b_level(float (&acc)[8], uint32_t my_rec, uint32_t lane_off, uint32_t (&wpair)[WIN_P]) {
  using G = w7::Lv<L>;
  constexpr uint32_t ROW = G::WW * (WIN_D * 2);
  uint32_t addr[WIN_P];
#pragma unroll
  for (int pp = 0; pp < WIN_P; pp += 2)
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(addr[pp]), "=r"(wpair[pp]), "=r"(addr[pp + 1]), "=r"(wpair[pp + 1])
                 : "r"(my_rec + (L * WIN_P + pp) * 8));
#pragma unroll
  for (int h = 0; h < WIN_P; h += 3) {
    uint4 d0[3], d1[3];
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      const uint32_t a = addr[h + pp] + lane_off;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(d0[pp].x), "=r"(d0[pp].y), "=r"(d0[pp].z), "=r"(d0[pp].w) : "r"(a));
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(d1[pp].x), "=r"(d1[pp].y), "=r"(d1[pp].z), "=r"(d1[pp].w)
                   : "r"(a), "n"(ROW));
    }
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      fma_row<0>(acc, d0[pp], wpair[h + pp]);
      fma_row<1>(acc, d1[pp], wpair[h + pp]);
    }
  }
}

// slow_point of msda_win_common.cuh takes a WinParams; it only reads these fields
__device__ __forceinline__ void slow_fix(const Win7Params& p, const __nv_bfloat16* __restrict__ value, int b, int m, int l, float x,
                                         float y, float aw, int s, float (&acc)[8]) {
  const int side = s >> 2;
  const Footprint f = make_footprint(x, y, p.lv.H[l], p.lv.W[l]);
  const int64_t first = p.pixel_major ? (((int64_t)b * p.Lv + p.lv.start[l]) * p.M + m) * WIN_D
                                      : (((int64_t)b * p.M + m) * p.Lv + p.lv.start[l]) * WIN_D;
  const int64_t pix_bytes = (p.pixel_major ? p.M : 1) * (WIN_D * 2);
  const char* base = reinterpret_cast<const char*>(value + first) + (s & 3) * 16;
  const uint4 e0 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(side ? f.i01 : f.i00) * pix_bytes));
  const uint4 e1 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(side ? f.i11 : f.i10) * pix_bytes));
  const uint32_t w = side ? pack_bf16(f.w01 * aw, f.w11 * aw) : pack_bf16(f.w00 * aw, f.w10 * aw);
  fma_row<0>(acc, e0, w);
  fma_row<1>(acc, e1, w);
}

template <typename TL, int MODE, bool STATIC>
__global__ void __launch_bounds__(w7::WARPS * 32, 2)
msda_gather_fwd_win7_kernel(const __nv_bfloat16* __restrict__ value, const TL* __restrict__ loc, const TL* __restrict__ attn,
                            const float* __restrict__ ref, int64_t ref_bs, __nv_bfloat16* __restrict__ out,
                            const __grid_constant__ Win7Params p) {
  using namespace w7;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_bar;
  __shared__ int s_next;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x;                 // grid (M, regions, B): the heads of a region run back to back
  const int region = blockIdx.y;
  const int b = blockIdx.z;
  const int ry = region / p.regions_x, rx = region - ry * p.regions_x;
  const uint32_t smem_base = smem_u32(smem);

  // window origins in level coordinates: region origin - R + the head's centre hint; whatever falls outside the map is
  // zero-filled by the tensor map
  int ox[WIN_L], oy[WIN_L];
#pragma unroll
  for (int l = 0; l < WIN_L; ++l) {
    ox[l] = ((rx * TW) >> l) - R + p.cshift[m & (WIN_MAX_M - 1)][l][0];
    oy[l] = ((ry * TH) >> l) - R + p.cshift[m & (WIN_MAX_M - 1)][l][1];
  }
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    s_next = WARPS;                           // batches 0..WARPS-1 are taken statically, one per warp
    fence_barrier_init();
  }
  if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(smem)[threadIdx.x] = 0u;   // what weight-0 records point at
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&s_bar, WIN_TX_BYTES);
    if (p.pixel_major) {
      tma_load_5d(smem_base + OFF2, &p.tmap[2], &s_bar, 0, m, ox[2], oy[2], b);
      tma_load_5d(smem_base + OFF1, &p.tmap[1], &s_bar, 0, m, ox[1], oy[1], b);
      tma_load_5d(smem_base + OFF0, &p.tmap[0], &s_bar, 0, m, ox[0], oy[0], b);
    } else {
      tma_load_4d(smem_base + OFF2, &p.tmap[2], &s_bar, 0, ox[2], oy[2], b * p.M + m);
      tma_load_4d(smem_base + OFF1, &p.tmap[1], &s_bar, 0, ox[1], oy[1], b * p.M + m);
      tma_load_4d(smem_base + OFF0, &p.tmap[0], &s_bar, 0, ox[0], oy[0], b * p.M + m);
    }
  }

  const uint32_t rec_base = smem_base + REC_OFF + (uint32_t)warp * REC_BYTES;
  const uint32_t xy_base = smem_base + XY_OFF + (uint32_t)warp * XY_BYTES;
  const int g = lane >> 3, s = lane & 7, side = s >> 2;

  // stage A lane map: query a_qi = lane % 4, point a_pp = lane / 4 (lanes 24-31 idle), the same in every level's round
  const bool a_on = lane < WIN_QPB * WIN_P;
  const int a_qi = a_on ? (lane & (WIN_QPB - 1)) : 0;
  const int a_pp = a_on ? (lane >> 2) : 0;
  const uint32_t a_rdst = rec_base + ((uint32_t)(a_qi * 2 * WIN_LP) + (uint32_t)a_pp) * 8;
  const uint32_t a_xydst = xy_base + (uint32_t)(a_qi * WIN_LP + a_pp) * 8;
  float fW[WIN_L], fH[WIN_L];
#pragma unroll
  for (int l = 0; l < WIN_L; ++l) { fW[l] = (float)p.lv.W[l]; fH[l] = (float)p.lv.H[l]; }

  RawLoc<TL> raw[WIN_L];
  float2 rref[WIN_L];
  const uint32_t item_stride = (uint32_t)p.M * WIN_LP;
  const TL* loc_lane = loc + (((int64_t)b * p.Lq) * p.M + m) * (WIN_LP * 2) + a_pp * 2;
  const TL* attn_lane = attn + (((int64_t)b * p.Lq) * p.M + m) * WIN_LP + a_pp;
  const float* ref_b = ref + (MODE == EMRT_LOC_PIXEL_OFFSET ? b * ref_bs : 0);
  __nv_bfloat16* out_lane = out + (((int64_t)b * p.Lq) * p.M + m) * WIN_D + g * (p.M * WIN_D) + (s & 3) * 8 + side * 4;
  // first query of the region at every level
  const int qb0 = p.lv.start[0] + (ry * TH) * p.lv.W[0] + rx * TW;
  const int qb1 = p.lv.start[1] + (ry * (TH / 2)) * p.lv.W[1] + rx * (TW / 2);
  const int qb2 = p.lv.start[2] + (ry * (TH / 4)) * p.lv.W[2] + rx * (TW / 4);
  const int W0 = p.lv.W[0], W1 = p.lv.W[1], W2 = p.lv.W[2];
  // batch -> first query: level 0 rows hold 4 batches, level 1 rows 2, level 2 rows 1
  auto query_of = [&](int batch) {
    if (batch < N0) return qb0 + (batch >> 2) * W0 + ((batch & 3) << 2);
    if (batch < N0 + N1) return qb1 + ((batch - N0) >> 1) * W1 + (((batch - N0) & 1) << 2);
    return qb2 + (batch - N0 - N1) * W2;
  };
  int q_next = 0;
  auto fetch = [&](int batch) {
    q_next = query_of(batch);
    const uint32_t qq = (uint32_t)(q_next + a_qi);
    const TL* lp = loc_lane + (size_t)(qq * item_stride) * 2;
    const TL* ap = attn_lane + (size_t)(qq * item_stride);
    const float* rp = ref_b + (size_t)(qq * (WIN_L * 2u));
#pragma unroll
    for (int l = 0; l < WIN_L; ++l) {
      raw_fetch(raw[l], lp + l * (WIN_P * 2), ap + l * WIN_P);
      if (MODE == EMRT_LOC_PIXEL_OFFSET) rref[l] = __ldg(reinterpret_cast<const float2*>(rp + 2 * l));
    }
  };

  int batch = warp;
  fetch(batch);                              // WARPS <= N_BATCHES: every warp owns at least one batch
  bool windows_ready = false;
  const uint32_t my_rec = rec_base + (uint32_t)(g * 2 + side) * (WIN_LP * 8);
  const uint32_t lane_off = (uint32_t)s * 16;

  while (batch < N_BATCHES) {
    const int cur_q = q_next;
    int next = batch + WARPS;
    uint32_t leader = 0;
    if (!STATIC)                             // one elected lane claims the next batch (see the header on atom + elect.sync)
      asm volatile("{\n .reg .pred p;\n elect.sync %1|p, 0xffffffff;\n @p atom.shared.add.u32 %0, [%2], 1;\n}\n"
                   : "+r"(next), "=r"(leader) : "r"(smem_u32(&s_next)) : "memory");
    // ---- stage A: one footprint record per (query, point), level by level, from the prefetched inputs --------------
    unsigned slow_lv = 0u;
    if (a_on) {
      stage_a_level<0, MODE>(raw[0], rref[0], fW[0], fH[0], ox[0], oy[0], smem_base, a_rdst, a_xydst, slow_lv);
      stage_a_level<1, MODE>(raw[1], rref[1], fW[1], fH[1], ox[1], oy[1], smem_base, a_rdst, a_xydst, slow_lv);
      stage_a_level<2, MODE>(raw[2], rref[2], fW[2], fH[2], ox[2], oy[2], smem_base, a_rdst, a_xydst, slow_lv);
    }
    batch = STATIC ? next : __shfl_sync(0xffffffffu, next, leader);
    const unsigned slow_levels = __reduce_or_sync(0xffffffffu, slow_lv);
    if (batch < N_BATCHES) fetch(batch);     // the next batch's loads land while this one gathers
    __syncwarp();
    if (!windows_ready) {
      mbar_wait(&s_bar, 0);
      windows_ready = true;
    }

    // ---- stage B: 8 lanes per query; lane s reads bytes [16 s, 16 s + 16) of the 128-byte pixel pair ----------------
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // (fix-ups, per level, under a warp-uniform and rare branch: points that left the staged window but not the map read
    // global memory)
    auto fix_level = [&](int l, const uint32_t (&wpl)[WIN_P]) {
#pragma unroll 1
      for (int pp = 0; pp < WIN_P; ++pp) {
        uint32_t w = wpl[0];
#pragma unroll
        for (int k = 1; k < WIN_P; ++k) { if (pp == k) w = wpl[k]; }
        if (w & WIN_SLOW) {
          float sx, sy;
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(sx), "=f"(sy) : "r"(xy_base + (uint32_t)(g * WIN_LP + l * WIN_P + pp) * 8));
          slow_fix(p, value, b, m, l, sx, sy, __uint_as_float(w << 16), s, acc);
        }
      }
    };
    {
      uint32_t wp[WIN_P];
      stage_b_level<0>(acc, my_rec, lane_off, wp);
      if (slow_levels & 1u) fix_level(0, wp);
    }
    {
      uint32_t wp[WIN_P];
      stage_b_level<1>(acc, my_rec, lane_off, wp);
      if (slow_levels & 2u) fix_level(1, wp);
    }
    {
      uint32_t wp[WIN_P];
      stage_b_level<2>(acc, my_rec, lane_off, wp);
      if (slow_levels & 4u) fix_level(2, wp);
    }
    // left + right pixel halves: lane s keeps channels [8 (s&3) + 4 side, +4)
    float keep[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float mine = side ? acc[4 + i] : acc[i];
      const float send = side ? acc[i] : acc[4 + i];
      keep[i] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    uint2 o;
    o.x = pack_bf16(keep[0], keep[1]);
    o.y = pack_bf16(keep[2], keep[3]);
    *reinterpret_cast<uint2*>(out_lane + (uint32_t)cur_q * (uint32_t)(p.M * WIN_D)) = o;
    __syncwarp();   // records are rewritten by the next batch
  }
  if (!windows_ready) mbar_wait(&s_bar, 0);   // never leave with a TMA still writing this CTA's shared memory
}

template <typename TL, int MODE, bool STATIC>
static int launch_win7(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out, int B,
                       int n_regions, const Win7Params& p, cudaStream_t st) {
  auto kern = msda_gather_fwd_win7_kernel<TL, MODE, STATIC>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w7::SMEM_BYTES));
  kern<<<dim3((unsigned)p.M, (unsigned)n_regions, (unsigned)B), w7::WARPS * 32, w7::SMEM_BYTES, st>>>(
      (const __nv_bfloat16*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, (__nv_bfloat16*)out, p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

// EMRT_ERR_UNSUPPORTED (error text untouched) when the shape / environment asks for something else: the caller then runs
// the run-time-geometry kernel.
int gather_fwd_win7(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out, int B,
                    int Lq, int Lv, int M, int D, int L, int P, const LevelTable& lv, int loc_dtype, int mode,
                    const int32_t* win_center_host, cudaStream_t st) {
  using namespace w7;
  if (D != WIN_D || L != WIN_L || P != WIN_P || Lq != Lv || M > WIN_MAX_M) return EMRT_ERR_UNSUPPORTED;
  if (getenv("EMRT_WIN_R") || getenv("EMRT_WIN_TH") || getenv("EMRT_WIN_TW") || getenv("EMRT_WIN_WARPS") || getenv("EMRT_WIN_GENERIC"))
    return EMRT_ERR_UNSUPPORTED;             // the sweeps of profiles/r1s_gather_sweep.txt run the run-time-geometry kernel
  for (int l = 1; l < L; ++l)
    if (lv.H[l] != (lv.H[0] >> l) || lv.W[l] != (lv.W[0] >> l) || (lv.H[l] << l) != lv.H[0] || (lv.W[l] << l) != lv.W[0])
      return EMRT_ERR_UNSUPPORTED;
  if (lv.H[0] % TH || lv.W[0] % TW) return EMRT_ERR_UNSUPPORTED;
  if ((int64_t)Lq * M * WIN_LP * 2 >= (1LL << 31)) return EMRT_ERR_UNSUPPORTED;   // 32-bit per-batch-element offsets
  if ((reinterpret_cast<uintptr_t>(value) & 15) != 0) return EMRT_ERR_UNSUPPORTED;
  Win7Params p;
  memset(&p, 0, sizeof(p));
  p.regions_x = lv.W[0] / TW;
  const int n_regions = p.regions_x * (lv.H[0] / TH);
  if (B > 65535 || n_regions > 65535) return EMRT_ERR_UNSUPPORTED;
  p.Lq = Lq; p.Lv = Lv; p.M = M; p.lv = lv;
  p.pixel_major = (mode & EMRT_VALUE_HEAD_MAJOR) ? 0 : 1;
  if (win_center_host && !getenv("EMRT_WIN_NO_HINT"))
    for (int mm = 0; mm < M; ++mm)
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < 2; ++k)
          p.cshift[mm][l][k] = (int8_t)std::min(std::max(win_center_host[(mm * L + l) * 2 + k], -100), 100);
  const int WWs[3] = {WW0, WW1, WW2}, WHs[3] = {WH0, WH1, WH2};
  for (int l = 0; l < L; ++l) {
    if (p.pixel_major) {
      const uint64_t dims[5] = {(uint64_t)WIN_D, (uint64_t)M, (uint64_t)lv.W[l], (uint64_t)lv.H[l], (uint64_t)B};
      const uint64_t strides[4] = {(uint64_t)WIN_D * 2, (uint64_t)M * WIN_D * 2, (uint64_t)lv.W[l] * M * WIN_D * 2,
                                   (uint64_t)Lv * M * WIN_D * 2};
      const uint32_t box[5] = {(uint32_t)WIN_D, 1u, (uint32_t)WWs[l], (uint32_t)WHs[l], 1u};
      const __nv_bfloat16* base = (const __nv_bfloat16*)value + (int64_t)lv.start[l] * M * WIN_D;
      if (int e = make_tensor_map(&p.tmap[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
        return e;
    } else {
      const uint64_t dims[4] = {(uint64_t)WIN_D, (uint64_t)lv.W[l], (uint64_t)lv.H[l], (uint64_t)B * M};
      const uint64_t strides[3] = {(uint64_t)WIN_D * 2, (uint64_t)lv.W[l] * WIN_D * 2, (uint64_t)Lv * WIN_D * 2};
      const uint32_t box[4] = {(uint32_t)WIN_D, (uint32_t)WWs[l], (uint32_t)WHs[l], 1u};
      const __nv_bfloat16* base = (const __nv_bfloat16*)value + (int64_t)lv.start[l] * WIN_D;
      if (int e = make_tensor_map(&p.tmap[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
        return e;
    }
  }
  const bool stat = getenv("EMRT_WIN_STATIC") != nullptr && atoi(getenv("EMRT_WIN_STATIC")) != 0;
  const bool px = (mode & EMRT_LOC_PIXEL_OFFSET) != 0;
#define EMRT_WIN7(TL)                                                                                                   \
  if (stat) return px ? launch_win7<TL, 1, true>(value, loc, attn, ref, ref_bs, out, B, n_regions, p, st)               \
                      : launch_win7<TL, 0, true>(value, loc, attn, ref, ref_bs, out, B, n_regions, p, st);              \
  return px ? launch_win7<TL, 1, false>(value, loc, attn, ref, ref_bs, out, B, n_regions, p, st)                        \
            : launch_win7<TL, 0, false>(value, loc, attn, ref, ref_bs, out, B, n_regions, p, st)
  switch (loc_dtype) {
    case EMRT_F32: EMRT_WIN7(float);
    case EMRT_F16: EMRT_WIN7(__half);
    case EMRT_BF16: EMRT_WIN7(__nv_bfloat16);
    default: return EMRT_ERR_UNSUPPORTED;
  }
#undef EMRT_WIN7
}

}  // namespace emrt
