// Forward sampling-gather with the value windows staged in shared memory by TMA (encoder self-attention).
//
// ncu on the L1-path kernel (profiles/r1b_gather_ncu.txt) shows it bound by the SM's L1 data pipe
// (l1tex__data_pipe_lsu_wavefronts 84 % of peak): every bilinear corner is a 64-byte segment of another 128-byte
// line, so one LDG.128 costs ~6 wavefronts for 512 bytes.  HBM traffic is already minimal (168 MB for 183 MB
// algorithmic).  The remaining lever is wavefronts per byte, so this kernel
//   1. lets one CTA own a REGION of the image (TH x TW level-0 pixels and the co-located pixels of the coarser
//      levels — all of them queries) for one (batch, head), and has TMA copy the three value windows that region
//      can reach (region +- R pixels) into shared memory, out-of-map pixels zero-filled by the tensor map — which
//      IS grid_sample's zeros padding, so the fast path has no validity logic at all;
//   2. computes each bilinear footprint once (one lane per (query, point)) into a 16-byte record
//      {smem address of the top pixel pair, of the bottom pair, bf16 weights left (top,bottom), right (top,bottom)};
//   3. gathers with 8 lanes per query: lanes 0-3 read the left pixel's 64 bytes, lanes 4-7 the right pixel's — one
//      contiguous 128-byte span per query and row, i.e. a conflict-free LDS.128 at the full 128 B/clk;
//   4. accumulates with the sm_100 mixed-precision FMA (fma.rn.f32.bf16, SASS FHFMA.BF16): bf16 x bf16 + fp32 with the
//      half-word selected inside the instruction, so there is no unpack work.
// A sample that lands outside the staged window but inside the map (|offset| > R) takes a per-point slow path
// that reads global memory with explicit validity weights, so results do not depend on R.
// The query -> region mapping is only a locality promise (EMRT_QUERY_PIXEL_GRID): any reference points give
// correct results, far-away ones just run the slow path.  The same holds for the optional window-centre hint (per head
// and level, the mean sampling offset taken from the layer's bias): it only moves the windows to where that head samples.
// Value layouts: head-major [B,M,Lv,32] (4-D tensor maps) or the reference's pixel-major [B,Lv,M,32] (5-D maps, the
// head is a coordinate) — the training path keeps the latter.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "msda_win_common.cuh"

namespace emrt {

static_assert(WIN_QPB * WIN_P <= 32, "stage A: one level's records of a batch fit one warp round");

// STATIC: batches are dealt to the warps round-robin (no shared counter) — right when the batch count of a region is a
// multiple of the warp count (168 queries = 42 batches = 7 warps x 6); otherwise the warps claim batches dynamically, so
// none idles while the CTA's shared-memory windows stay resident.
template <typename TL, int MODE, int WIN_WARPS, bool STATIC>
__global__ void __launch_bounds__(WIN_WARPS * 32, WIN_WARPS > 12 ? 1 : 2)
msda_gather_fwd_win_kernel(const __nv_bfloat16* __restrict__ value, const TL* __restrict__ loc,
                           const TL* __restrict__ attn, const float* __restrict__ ref, int64_t ref_bs,
                           __nv_bfloat16* __restrict__ out, const __grid_constant__ WinParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_bar;
  __shared__ int s_next;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid (M, regions, B): the M heads of a region run back to back (they share the region's loc / attn rows)
  const int m = blockIdx.x;
  const int region = blockIdx.y;
  const int b = blockIdx.z;
  const int ry = region / p.regions_x, rx = region - ry * p.regions_x;
  const uint32_t smem_base = smem_u32(smem);

  // window origins (level coordinates; may start at -1: the zero halo)
  int ox[WIN_L], oy[WIN_L];
#pragma unroll
  for (int l = 0; l < WIN_L; ++l) {
    const int x0 = (rx * p.TW) >> l, y0 = (ry * p.TH) >> l;
    const int cx = m < WIN_MAX_M ? p.cshift[m][l][0] : 0, cy = m < WIN_MAX_M ? p.cshift[m][l][1] : 0;
    ox[l] = min(max(x0 - p.R + cx, -1), p.lv.W[l] + 1 - p.WW[l]);
    oy[l] = min(max(y0 - p.R + cy, -1), p.lv.H[l] + 1 - p.WH[l]);
  }
  // queries of this region: TH*TW at level 0, a quarter of that at each coarser level
  const int n_batches = ((p.TH * p.TW) >> 2) + ((p.TH * p.TW) >> 4) + ((p.TH * p.TW) >> 6);

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    s_next = WIN_WARPS;                      // batches 0..WIN_WARPS-1 are taken statically, one per warp
    fence_barrier_init();
  }
  if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(smem)[threadIdx.x] = 0u;   // what weight-0 records point at
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
#pragma unroll
    for (int l = 0; l < WIN_L; ++l) bytes += (uint32_t)(p.WW[l] * p.WH[l]) * (WIN_D * 2);
    mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
    for (int l = WIN_L - 1; l >= 0; --l) {
      if (p.pixel_major) tma_load_5d(smem_base + p.win_off[l], &p.tmap[l], &s_bar, 0, m, ox[l], oy[l], b);
      else tma_load_4d(smem_base + p.win_off[l], &p.tmap[l], &s_bar, 0, ox[l], oy[l], b * p.M + m);
    }
  }

  const uint32_t rec_base = smem_base + p.rec_off + (uint32_t)warp * (WIN_QPB * WIN_LP * 16);
  const uint32_t xy_base = smem_base + p.rec_off + (uint32_t)WIN_WARPS * (WIN_QPB * WIN_LP * 16) + (uint32_t)warp * (WIN_QPB * WIN_LP * 8);
  const int g = lane >> 3, s = lane & 7, side = s >> 2;

  // ---- stage A works level by level: round l handles the WIN_QPB x WIN_P records of level l, one per lane
  // (lane -> query a_qi = lane % 4, point a_pp = lane / 4, the same in every round; lanes 24-31 idle), so everything
  // that depends on the level (map size, window origin / pitch / base) is warp-uniform and compile-time indexed.
  // With this lane order the 8-byte record stores of a half-warp hit 32 distinct banks (query stride 72 words = 8 banks).
  const bool a_on = lane < WIN_QPB * WIN_P;
  const int a_qi = a_on ? (lane & (WIN_QPB - 1)) : 0;
  const int a_pp = a_on ? (lane >> 2) : 0;
  // records are kept per (query, side): 8-byte entries {address, that side's weight pair}, the 18 points of a
  // (query, side) contiguous, so stage B fetches two points per LDS.128 (9 loads per lane instead of 18 LDS.64)
  const uint32_t a_rdst = rec_base + ((uint32_t)(a_qi * 2 * WIN_LP) + (uint32_t)a_pp) * 8;
  const uint32_t a_xydst = xy_base + (uint32_t)(a_qi * WIN_LP + a_pp) * 8;
  float fW[WIN_L], fH[WIN_L];
#pragma unroll
  for (int l = 0; l < WIN_L; ++l) { fW[l] = (float)p.lv.W[l]; fH[l] = (float)p.lv.H[l]; }

  RawLoc<TL> raw[WIN_L];
  float2 rref[WIN_L];
  // issue the global loads of one batch (no dependent branches: one memory latency per batch): one 64-bit address per
  // tensor and batch, the three levels at immediate offsets from it
  const TL* loc_lane = loc + (((int64_t)b * p.Lq) * p.M + m) * (WIN_LP * 2) + a_pp * 2;
  const TL* attn_lane = attn + (((int64_t)b * p.Lq) * p.M + m) * WIN_LP + a_pp;
  const float* ref_b = ref + (MODE == EMRT_LOC_PIXEL_OFFSET ? b * ref_bs : 0);
  __nv_bfloat16* out_lane = out + (((int64_t)b * p.Lq) * p.M + m) * WIN_D + g * (p.M * WIN_D) + (s & 3) * 8 + side * 4;
  const uint32_t item_stride = (uint32_t)p.M * WIN_LP;
  int q_next = 0;     // first query of the prefetched batch (warp-uniform)
  int qb[WIN_L];
#pragma unroll
  for (int l = 0; l < WIN_L; ++l) qb[l] = p.lv.start[l] + ((ry * p.TH) >> l) * p.lv.W[l] + ((rx * p.TW) >> l);
  const int n0 = (p.TH * p.TW) >> 2, n01 = n0 + (n0 >> 2), sh0 = p.tw_shift - 2;
  auto fetch = [&](int batch) {
    q_next = batch_query_base(p, qb, n0, n01, sh0, batch);
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
  if (batch < n_batches) fetch(batch);
  bool windows_ready = false;

  while (batch < n_batches) {
    // claim the next batch now: the shared-memory atomic's latency hides under stage A
    const int cur_q = q_next;
    int next = batch + WIN_WARPS;
    if (!STATIC && lane == 0)
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(next) : "r"(smem_u32(&s_next)) : "memory");
    // ---- stage A: one footprint record per (query, point), from the prefetched inputs -----------------------------
    unsigned slow_lv = 0u;
    if (a_on) {
#pragma unroll
      for (int l = 0; l < WIN_L; ++l) {
        float x, y, aw;
        raw_decode(raw[l], x, y, aw);
        if (MODE == EMRT_LOC_PIXEL_OFFSET) {
          x = rref[l].x * fW[l] - 0.5f + x;
          y = rref[l].y * fH[l] - 0.5f + y;
        } else {
          x = x * fW[l] - 0.5f;
          y = y * fH[l] - 0.5f;
        }
        // The window lies inside [-1, W] x [-1, H], so "both pixel pairs inside the window" implies the sample is live;
        // the fmaxf sends NaN and -inf to -2, below every window origin (>= -1); +inf converts to INT_MAX, whose
        // window coordinate fails the unsigned compare as well.
        const float xs = fmaxf(x, -2.f), ys = fmaxf(y, -2.f);
        const int xi = __float2int_rd(xs), yi = __float2int_rd(ys);
        const float fx = xs - (float)xi, fy = ys - (float)yi;
        const int wx = xi - ox[l], wy = yi - oy[l];
        const bool fast = (unsigned)wx < (unsigned)(p.WW[l] - 1) && (unsigned)wy < (unsigned)(p.WH[l] - 1);
        const float gx = 1.f - fx, gy = 1.f - fy;
        const float gxa = gx * aw, fxa = fx * aw;
        uint32_t addr = smem_base + p.win_off[l] + (uint32_t)(wy * p.WW[l] + wx) * (WIN_D * 2);
        uint32_t wl = pack_bf16(gxa * gy, gxa * fy);   // left pixel: top, bottom
        uint32_t wr = pack_bf16(fxa * gy, fxa * fy);   // right pixel: top, bottom
        if (!fast) {
          // a sample whose four corners are all outside the map contributes exactly zero (also rejects NaN / inf):
          // weight 0 on the zero block.  Live but outside the window: the zero block again (so the fast path adds
          // nothing), the SLOW flag in the sign bit of the bottom weight, the attention weight in the low half and the
          // sample position in the side buffer for the fix-up
          const bool live = (x > -1.f) && (y > -1.f) && (x < fW[l]) && (y < fH[l]);
          addr = smem_base;
          wl = wr = live ? (WIN_SLOW | (pack_bf16(aw, 0.f) & 0xffffu)) : 0u;
          if (live) {
            asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a_xydst + l * (WIN_P * 8)), "f"(x), "f"(y) : "memory");
            slow_lv |= 1u << l;
          }
        }
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a_rdst + l * (WIN_P * 8)), "r"(addr), "r"(wl) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a_rdst + l * (WIN_P * 8) + WIN_LP * 8), "r"(addr), "r"(wr) : "memory");
      }
    }
    // next batch: start its loads now, they land while this batch gathers
    batch = STATIC ? next : __shfl_sync(0xffffffffu, next, 0);
    const unsigned slow_levels = __reduce_or_sync(0xffffffffu, slow_lv);   // levels of this batch with a fix-up point
    if (batch < n_batches) fetch(batch);
    __syncwarp();
    if (!windows_ready) {
      mbar_wait(&s_bar, 0);
      windows_ready = true;
    }

    // ---- stage B: 8 lanes per query; lane s reads bytes [16 s, 16 s + 16) of the 128-byte pixel pair ----------------
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const uint32_t my_rec = rec_base + (uint32_t)(g * 2 + side) * (WIN_LP * 8);
#pragma unroll
    for (int l = 0; l < WIN_L; ++l) {
      const uint32_t row_bytes = (uint32_t)p.WW[l] * (WIN_D * 2);
      // the level's six records first: {address, this side's weight pair}, 8 bytes per lane
      uint32_t addr[WIN_P], wpair[WIN_P];
#pragma unroll
      for (int pp = 0; pp < WIN_P; pp += 2)
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(addr[pp]), "=r"(wpair[pp]), "=r"(addr[pp + 1]), "=r"(wpair[pp + 1])
                     : "r"(my_rec + (l * WIN_P + pp) * 8));
      // branch-free fast path in two halves of three points: six independent LDS.128 in flight, then 48 FHFMA
      // (flagged records carry weight -0.0 on the zero block)
#pragma unroll
      for (int h = 0; h < WIN_P; h += 3) {
        uint4 d0[3], d1[3];
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) {
          const uint32_t a = addr[h + pp] + s * 16;
          d0[pp] = lds128(a);
          d1[pp] = lds128(a + row_bytes);
        }
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) {
          fma_row<0>(acc, d0[pp], wpair[h + pp]);
          fma_row<1>(acc, d1[pp], wpair[h + pp]);
        }
      }
      if (slow_levels & (1u << l)) {
        // fix-ups (warp-uniform branch, rare): points that left the staged window but not the map read global memory
#pragma unroll 1
        for (int pp = 0; pp < WIN_P; ++pp) {
          uint32_t w = wpair[0];
#pragma unroll
          for (int k = 1; k < WIN_P; ++k) { if (pp == k) w = wpair[k]; }
          if (w & WIN_SLOW) {
            float sx, sy;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(sx), "=f"(sy) : "r"(xy_base + (uint32_t)(g * WIN_LP + l * WIN_P + pp) * 8));
            uint4 e0, e1;
            slow_point(p, value, b, m, l, sx, sy, __uint_as_float(w << 16), s, &e0, &e1, &w);
            fma_row<0>(acc, e0, w);
            fma_row<1>(acc, e1, w);
          }
        }
      }
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

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <typename TL, int MODE, int NW, bool STATIC>
static int launch_win(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out,
                      int B, const WinParams& p, size_t smem_bytes, cudaStream_t st) {
  auto kern = msda_gather_fwd_win_kernel<TL, MODE, NW, STATIC>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));   // per context
  const int n_regions = p.regions_x * p.regions_y;
  if (B > 65535 || n_regions > 65535) return EMRT_ERR_UNSUPPORTED;
  kern<<<dim3((unsigned)p.M, (unsigned)n_regions, (unsigned)B), NW * 32, smem_bytes, st>>>(
      (const __nv_bfloat16*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, (__nv_bfloat16*)out, p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

// Returns EMRT_ERR_UNSUPPORTED (error text untouched) when the shape is not a regular 3-level pyramid this kernel
// tiles; the caller then falls back to the L1-path kernel.
int gather_fwd_win7(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out, int B,
                    int Lq, int Lv, int M, int D, int L, int P, const LevelTable& lv, int loc_dtype, int mode,
                    const int32_t* win_center_host, cudaStream_t st);

int gather_fwd_win(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out,
                   int B, int Lq, int Lv, int M, int D, int L, int P, const LevelTable& lv, int loc_dtype, int mode,
                   const int32_t* win_center_host, cudaStream_t st) {
  {  // the default geometry compiled in (msda_gather_win7.cu); anything else runs the run-time-geometry kernel below
    const int e = gather_fwd_win7(value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, D, L, P, lv, loc_dtype, mode, win_center_host, st);
    if (e != EMRT_ERR_UNSUPPORTED) return e;
  }
  if (D != WIN_D || L != WIN_L || P != WIN_P || Lq != Lv) return EMRT_ERR_UNSUPPORTED;
  for (int l = 1; l < L; ++l)
    if (lv.H[l] != (lv.H[0] >> l) || lv.W[l] != (lv.W[0] >> l) || (lv.H[l] << l) != lv.H[0] || (lv.W[l] << l) != lv.W[0])
      return EMRT_ERR_UNSUPPORTED;
  WinParams p;
  memset(&p, 0, sizeof(p));
  p.R = env_int("EMRT_WIN_R", 7);
  p.TH = env_int("EMRT_WIN_TH", 8);
  p.TW = env_int("EMRT_WIN_TW", 16);
  // TW a power of two >= 16 and TH a multiple of 4: every level's region rows hold whole batches of 4 queries
  if (p.R < 1 || p.TH < 4 || p.TW < 16 || (p.TH & 3) || (p.TW & (p.TW - 1)) || lv.H[0] % p.TH || lv.W[0] % p.TW) return EMRT_ERR_UNSUPPORTED;
  for (p.tw_shift = 0; (1 << p.tw_shift) < p.TW; ++p.tw_shift) {}
  if ((int64_t)Lq * M * WIN_LP * 2 >= (1LL << 31)) return EMRT_ERR_UNSUPPORTED;   // 32-bit per-batch-element offsets
  p.regions_x = lv.W[0] / p.TW;
  p.regions_y = lv.H[0] / p.TH;
  p.Lq = Lq; p.Lv = Lv; p.M = M; p.lv = lv;
  p.pixel_major = (mode & EMRT_VALUE_HEAD_MAJOR) ? 0 : 1;
  // window-centre hint [M, L, 2] (x, y) in pixels of level l: where head m's samples of level l lie relative to the
  // reference point on average (the sampling_offsets bias).  Only moves the staged windows; any sample outside them
  // still takes the global-memory path, so results never depend on it.
  if (win_center_host && M <= WIN_MAX_M && !getenv("EMRT_WIN_NO_HINT"))
    for (int mm = 0; mm < M; ++mm)
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < 2; ++k)
          p.cshift[mm][l][k] = (int8_t)std::min(std::max(win_center_host[(mm * L + l) * 2 + k], -100), 100);
  uint32_t off = 128;   // [0,128): zero block
  for (int l = 0; l < L; ++l) {
    p.WW[l] = std::min((p.TW >> l) + 2 * p.R + 1, lv.W[l] + 2);
    p.WH[l] = std::min((p.TH >> l) + 2 * p.R + 1, lv.H[l] + 2);
    if (p.WW[l] > 256 || p.WH[l] > 256) return EMRT_ERR_UNSUPPORTED;
    p.win_off[l] = off;
    off += ((uint32_t)(p.WW[l] * p.WH[l]) * (WIN_D * 2) + 127u) & ~127u;
    if (p.pixel_major) {
      // one head's plane of level l inside the reference's own [B, Lv, M, 32] layout: the head is coordinate 1
      const uint64_t dims[5] = {(uint64_t)WIN_D, (uint64_t)M, (uint64_t)lv.W[l], (uint64_t)lv.H[l], (uint64_t)B};
      const uint64_t strides[4] = {(uint64_t)WIN_D * 2, (uint64_t)M * WIN_D * 2, (uint64_t)lv.W[l] * M * WIN_D * 2,
                                   (uint64_t)Lv * M * WIN_D * 2};
      const uint32_t box[5] = {(uint32_t)WIN_D, 1u, (uint32_t)p.WW[l], (uint32_t)p.WH[l], 1u};
      const __nv_bfloat16* base = (const __nv_bfloat16*)value + (int64_t)lv.start[l] * M * WIN_D;
      if (int e = make_tensor_map(&p.tmap[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_NONE))
        return e;
    } else {
      const uint64_t dims[4] = {(uint64_t)WIN_D, (uint64_t)lv.W[l], (uint64_t)lv.H[l], (uint64_t)B * M};
      const uint64_t strides[3] = {(uint64_t)WIN_D * 2, (uint64_t)lv.W[l] * WIN_D * 2, (uint64_t)Lv * WIN_D * 2};
      const uint32_t box[4] = {(uint32_t)WIN_D, (uint32_t)p.WW[l], (uint32_t)p.WH[l], 1u};
      const __nv_bfloat16* base = (const __nv_bfloat16*)value + (int64_t)lv.start[l] * WIN_D;
      if (int e = make_tensor_map(&p.tmap[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_NONE))
        return e;
    }
  }
  p.rec_off = off;
  // 7 warps x 6 batches covers the default 8 x 16 region (42 batches of 4 queries) exactly: static dealing, no counter
  const int n_batches = (p.TH * p.TW + (p.TH >> 1) * (p.TW >> 1) + (p.TH >> 2) * (p.TW >> 2) + WIN_QPB - 1) / WIN_QPB;
  // 10 warps x 2 CTAs per SM: 20 resident warps at 84 registers and 111 KB of windows + records per CTA (measured best of
  // 8 / 10 / 12, profiles/r1s_gather_sweep.txt)
  int warps = env_int("EMRT_WIN_WARPS", 10);
  if (warps != 7 && warps != 8 && warps != 12 && warps != 16 && warps != 24) warps = 10;
  const bool stat = env_int("EMRT_WIN_STATIC", -1) >= 0 ? env_int("EMRT_WIN_STATIC", 0) != 0 : (n_batches % warps == 0);
  const size_t smem_bytes = (size_t)off + (size_t)warps * WIN_QPB * WIN_LP * (16 + 8);   // records + slow-point positions
  if (smem_bytes > 227 * 1024) return EMRT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(value) & 15) != 0) return EMRT_ERR_UNSUPPORTED;
  const bool px = (mode & EMRT_LOC_PIXEL_OFFSET) != 0;
#define EMRT_WIN_NW(TL, NW)                                                                                       \
  if (warps == NW) {                                                                                               \
    if (stat) return px ? launch_win<TL, 1, NW, true>(value, loc, attn, ref, ref_bs, out, B, p, smem_bytes, st)    \
                        : launch_win<TL, 0, NW, true>(value, loc, attn, ref, ref_bs, out, B, p, smem_bytes, st);   \
    return px ? launch_win<TL, 1, NW, false>(value, loc, attn, ref, ref_bs, out, B, p, smem_bytes, st)             \
              : launch_win<TL, 0, NW, false>(value, loc, attn, ref, ref_bs, out, B, p, smem_bytes, st);            \
  }
#define EMRT_WIN(TL) EMRT_WIN_NW(TL, 7) EMRT_WIN_NW(TL, 10) EMRT_WIN_NW(TL, 12) EMRT_WIN_NW(TL, 16) EMRT_WIN_NW(TL, 24) EMRT_WIN_NW(TL, 8) return EMRT_ERR_UNSUPPORTED
  switch (loc_dtype) {
    case EMRT_F32: EMRT_WIN(float);
    case EMRT_F16: EMRT_WIN(__half);
    case EMRT_BF16: EMRT_WIN(__nv_bfloat16);
    default: return EMRT_ERR_UNSUPPORTED;
  }
#undef EMRT_WIN
#undef EMRT_WIN_NW
}

}  // namespace emrt
