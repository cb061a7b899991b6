// 3x3 convolution of the encoder layer's conv branch as an implicit GEMM on tcgen05
// (conv{l}: Conv2D(256, 256, 3, padding 1, no bias), transformer_encoder_decoder.py:125-144,187-189).
//
// The tokens [B, Lv, 256] are NHWC per level, so the nine taps are nine SHIFTED TMA loads of the same tensor: a 4-D box
// {64 ch, W_l, rows, images} whose (x, y) start is the tile origin + (kx-1, ky-1); out-of-map pixels come back as zeros,
// which is the convolution's zero padding.  Every tap contributes a K = 256 GEMM against its [Cout, Cin] weight slab,
// all 36 k-blocks accumulating into one TMEM tile: D[128 pixels x 256] = sum_tap A_tap[128 x 256] W_tap^T.
// Same warp-specialised skeleton as linear_tcgen05.cu (TMA producer / MMA issuer / 8 epilogue warps, two TMEM
// accumulators, staged TMA-store epilogue); one launch covers the three levels (different weights per level).
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace emrt {

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

constexpr int CV_BM = 128, CV_BK = 64, CV_N = 256;
// CG = 2 (CTA pair, tcgen05.mma.cta_group::2): each CTA brings its own 128-pixel A tile and HALF of the weight slab, so a stage
// is 32 KB instead of 48 KB and the ring is six stages deep in the same 192 KB.  Why it matters: with one CTA per tile the
// kernel is bound by the SM's shared-memory bandwidth — per k-block the tensor pipe reads 48 KB of operands and TMA writes 48 KB
// of new ones, 96 KB at 128 B/clk = 750 clocks for 512 clocks of MMAs (measured 814).  Halving B makes it 64 KB = 500 clocks.
template <int CG> struct ConvCfg { static constexpr int STAGES = CG == 2 ? 6 : 4; static constexpr int B_ROWS = CV_N / CG; };
constexpr int CV_EPI_WARP0 = 2, CV_EPI_WARPS = 8, CV_THREADS = 32 * (CV_EPI_WARP0 + CV_EPI_WARPS);
constexpr int CV_MAX_L = 4;

struct ConvParams {
  CUtensorMap tma_x[CV_MAX_L];   // level l: {256 ch, W, H, B} bf16, box {64, W, rows, imgs}, SWIZZLE_128B
  CUtensorMap tma_y[CV_MAX_L];   // level l: {256 ch, H*W, B} bf16, box {32, 32, 1}, SWIZZLE_64B
  CUtensorMap tma_w;             // [L*9*256, 256] bf16 (row = (l*9 + tap)*256 + cout), box {64, 256}, SWIZZLE_128B
  int32_t tile_start[CV_MAX_L + 1];
  int32_t tiles_per_img[CV_MAX_L];   // >= 1 when H*W >= 128, else 0
  int32_t imgs_per_tile[CV_MAX_L];
  int32_t rows_per_tile[CV_MAX_L];
  int32_t W[CV_MAX_L], HW[CV_MAX_L];
  int32_t L;
  float* gn_partial;   // STATS: [tile][lane quarter][32 groups][sum, sum of squares] of the stored (bf16-rounded) output
};

template <int CG>
struct ConvSmem {
  __nv_bfloat16 a[ConvCfg<CG>::STAGES][CV_BM * CV_BK];
  __nv_bfloat16 b[ConvCfg<CG>::STAGES][ConvCfg<CG>::B_ROWS * CV_BK];
  uint8_t stage[CV_EPI_WARPS * 32 * 64];
  uint64_t full[ConvCfg<CG>::STAGES];
  uint64_t empty[ConvCfg<CG>::STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

struct ConvTile { int l, b0, y0; };
__device__ __forceinline__ ConvTile conv_tile(const ConvParams& p, int t) {
  ConvTile c;
  c.l = 0;
  while (c.l + 1 < p.L && t >= p.tile_start[c.l + 1]) ++c.l;
  const int tl = t - p.tile_start[c.l];
  if (p.tiles_per_img[c.l] > 0) {
    c.b0 = tl / p.tiles_per_img[c.l];
    c.y0 = (tl - c.b0 * p.tiles_per_img[c.l]) * p.rows_per_tile[c.l];
  } else {
    c.b0 = tl * p.imgs_per_tile[c.l];
    c.y0 = 0;
  }
  return c;
}

template <int CG>
__device__ __forceinline__ void tma_load_4d_cv(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                               int c2, int c3) {
  if (CG == 2)
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// STATS: the epilogue also leaves the GroupNorm(32) statistics of what it stores — per (tile, 32-pixel slab, group) partial
// sums, combined per (image, level) in a fixed order by conv_gn_finalize_kernel — so that the statistics pass over the conv
// output (one more read of the tensor, 43 us per layer at 72 windows) does not exist.
template <int CG, bool STATS>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_tokens_tc_kernel(const __grid_constant__ ConvParams p) {
  constexpr int CV_STAGES = ConvCfg<CG>::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  ConvSmem<CG>& s = *reinterpret_cast<ConvSmem<CG>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tile_start[p.L];
  constexpr uint32_t STAGE_BYTES = (CV_BM * CV_BK + ConvCfg<CG>::B_ROWS * CV_BK) * 2;
  constexpr int KB = CV_N / CV_BK;          // k-blocks per tap (Cin = 256)
  // CTA pair: tiles 2 k and 2 k + 1 (same level: the host only picks this form when every level has an even tile count)
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int first = (int)blockIdx.x / CG * CG + (int)rank;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < p.L; ++l) { tma_prefetch_desc(&p.tma_x[l]); tma_prefetch_desc(&p.tma_y[l]); }
    tma_prefetch_desc(&p.tma_w);
#pragma unroll
    for (int i = 0; i < CV_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { mbar_init(&s.tmem_full[i], 1); mbar_init(&s.tmem_empty[i], CG * CV_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // (both CTAs of a pair: own A tile, own half of the weight rows; the bytes of both count on the leader's barrier)
      for (int t = first; t < num_tiles; t += gridDim.x) {
        const ConvTile c = conv_tile(p, t);
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&s.empty[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&s.full[stage], CG * STAGE_BYTES);
            const uint32_t lb = leader_addr<CG>(&s.full[stage]);
            tma_load_4d_cv<CG>(s.a[stage], &p.tma_x[c.l], lb, kb * CV_BK, kx - 1, c.y0 + ky - 1, c.b0);
            tma_load_2d_lead<CG>(s.b[stage], &p.tma_w, lb, kb * CV_BK, (c.l * 9 + tap) * CV_N + (int)rank * ConvCfg<CG>::B_ROWS);
            if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (tc_common.cuh, elect_one); in a CTA pair only the leader's warp
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(CV_BM * CG, CV_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t a_base = smem_u32(s.a[0]), b_base = smem_u32(s.b[0]);
      const uint64_t desc_hi = make_smem_desc(0);
      for (int t = first; t < num_tiles; t += gridDim.x) {
        wait_lead<CG>(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CV_N);
        for (int ks = 0; ks < 9 * KB; ++ks) {
          wait_lead<CG>(&s.full[stage], phase);
          tc_fence_after();
          const uint64_t da = desc_hi | (uint64_t)(((a_base + (uint32_t)stage * (uint32_t)sizeof(s.a[0])) & 0x3FFFFu) >> 4);
          const uint64_t db = desc_hi | (uint64_t)(((b_base + (uint32_t)stage * (uint32_t)sizeof(s.b[0])) & 0x3FFFFu) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CV_BK / 16; ++k)
              umma_cg<CG>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ks | k) != 0 ? 1u : 0u);
            commit_cg<CG>(&s.empty[stage]);
            if (ks == 9 * KB - 1) commit_cg<CG>(&s.tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - CV_EPI_WARP0) >> 2;
    constexpr int HALF_N = CV_N / 2;
    const uint32_t stage_addr = smem_u32(s.stage) + (uint32_t)(warp - CV_EPI_WARP0) * (32 * 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t L_tmem_empty[2] = {leader_addr<CG>(&s.tmem_empty[0]), leader_addr<CG>(&s.tmem_empty[1])};
    for (int t = first; t < num_tiles; t += gridDim.x) {
      const ConvTile c = conv_tile(p, t);
      mbar_wait(&s.tmem_full[acc], acc_phase);
      tc_fence_after();
      // this warp's 32 pixels: tile pixel q*32.. -> (image, pixel inside the level map)
      const int tp = q * 32;
      const int img = c.b0 + tp / p.HW[c.l];
      const int pix = c.y0 * p.W[c.l] + tp % p.HW[c.l];
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CV_N + half * HALF_N);
#pragma unroll 1
      for (int cc = 0; cc < HALF_N; cc += 32) {
        uint32_t r[32];
        TMEM_LD_X32(t_row + cc, r);
        TMEM_WAIT_X32(r);
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]), EMRT_BF16);
        if (STATS) {
          // four groups of eight channels in this chunk: (sum, sum of squares) of the ROUNDED values (what GroupNorm reads
          // back), eight numbers per lane.  Reduced over the slab's 32 pixels by recursive halving — exchange half of the
          // numbers with lane ^ 16, half of the rest with ^ 8, then ^ 4, then two plain butterfly steps: 9 shuffles instead
          // of 40, a fixed tree (deterministic); lane 4 k ends up holding number k.
          float v8[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            f32x2_t sm2 = pk2(0.f), sq2 = pk2(0.f);            // (even channels, odd channels): FADD2 / FFMA2
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const f32x2_t d = pk2(__uint_as_float(o[g * 4 + i] << 16), __uint_as_float(o[g * 4 + i] & 0xffff0000u));
              sm2 = add2(sm2, d);
              sq2 = fma2(d, d, sq2);
            }
            float a0, a1, b0, b1;
            upk2(sm2, a0, a1);
            upk2(sq2, b0, b1);
            v8[2 * g] = a0 + a1;
            v8[2 * g + 1] = b0 + b1;
          }
          float v4[4], v2[2];
          const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float mine = h16 ? v8[4 + i] : v8[i], other = h16 ? v8[i] : v8[4 + i];
            v4[i] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float mine = h8 ? v4[2 + i] : v4[i], other = h8 ? v4[i] : v4[2 + i];
            v2[i] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
          }
          float v1 = (h4 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? v2[0] : v2[1], 4);
          v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
          v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
          if ((lane & 3) == 0)      // number k = lane / 4: group k / 2 of the chunk, component k % 2
            p.gn_partial[(((size_t)t * 4 + q) * 32 + (half * HALF_N + cc) / 8) * 2 + (lane >> 2)] = v1;
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 4; ++h)
          sts128(stage_addr + lane * 64 + ((h ^ ((lane >> 1) & 3)) << 4), o[h * 4], o[h * 4 + 1], o[h * 4 + 2], o[h * 4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&p.tma_y[c.l], stage_addr, half * HALF_N + cc, pix, img);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader<CG>(L_tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// stats[b][l][g] = (sum, sum of squares) over the (image, level)'s slabs.  One CTA per (image, level): thread (g, part) adds
// slabs part, part + 8, ... in order, the eight parts are then combined in order: a fixed tree, so the sums do not depend on
// scheduling or on the batch (a serial walk over the up to 128 slabs by one thread per group took 20 us of load latency).
__global__ void __launch_bounds__(256) conv_gn_finalize_kernel(const __grid_constant__ ConvParams p, float* __restrict__ stats) {
  const int bl = blockIdx.x, l = bl % p.L, b = bl / p.L, g = threadIdx.x & 31, part = threadIdx.x >> 5;
  int slab0, nslab;                                   // slabs are numbered tile * 4 + lane quarter
  if (p.tiles_per_img[l] > 0) {
    slab0 = (p.tile_start[l] + b * p.tiles_per_img[l]) * 4; nslab = p.tiles_per_img[l] * 4;
  } else {
    const int slabs = p.HW[l] / 32;                   // 32-pixel slabs per image (HW < 128: 1 or 2)
    slab0 = (p.tile_start[l] + b / p.imgs_per_tile[l]) * 4 + (b % p.imgs_per_tile[l]) * slabs; nslab = slabs;
  }
  float sm = 0.f, sq = 0.f;
  for (int i = part; i < nslab; i += 8) {
    const float2 v = *reinterpret_cast<const float2*>(p.gn_partial + ((size_t)(slab0 + i) * 32 + g) * 2);
    sm += v.x;
    sq += v.y;
  }
  __shared__ float2 sh[8][32];
  sh[part][g] = make_float2(sm, sq);
  __syncthreads();
  if (part == 0) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { sm += sh[k][g].x; sq += sh[k][g].y; }
    *reinterpret_cast<float2*>(stats + ((size_t)(b * p.L + l) * 32 + g) * 2) = make_float2(sm, sq);
  }
}

template <int CG, bool STATS>
static int launch_conv(const ConvParams& p, int tiles, cudaStream_t st, int max_ctas) {
  const int smem_bytes = (int)sizeof(ConvSmem<CG>) + 1024;
  auto kern = conv3x3_tokens_tc_kernel<CG, STATS>;
  // per launch: function attributes are per context (a second GPU in the same process needs its own opt-in) and this is cheap
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int grid = tiles < num_sms() ? tiles : num_sms();
  // a caller that runs this kernel beside another one (second stream) leaves the other SMs to it: the tile loop is persistent,
  // any grid size walks all tiles (profiles/r3t_conv_gather_overlap.txt)
  if (max_ctas >= CG && max_ctas < grid) grid = max_ctas;
  grid = grid / CG * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EMRT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

// Returns EMRT_ERR_UNSUPPORTED (error text untouched) for shapes this kernel does not tile.
// stats_ws (optional): F32 [2 * B * L * 32 sums | tiles * 256 partials] -> the GroupNorm(32) sums of y per (image, level, group)
int64_t conv3x3_stats_workspace_floats(int B, int Lv, int L) { return 2LL * B * L * 32 + ((int64_t)B * Lv / CV_BM + (int64_t)B * L + L) * 256; }

int conv3x3_tokens_tc(const void* x, const void* w_packed, void* y, int B, int Lv, int C, int L, const LevelTable& lv,
                      cudaStream_t st, float* stats_ws, int max_ctas) {
  if (C != CV_N || L > CV_MAX_L) return EMRT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(y)) & 15) return EMRT_ERR_UNSUPPORTED;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.L = L;
  int tiles = 0;
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l], HW = H * W;
    if (W > CV_BM || CV_BM % W != 0 || HW % 32 != 0) return EMRT_ERR_UNSUPPORTED;
    p.W[l] = W; p.HW[l] = HW;
    p.tile_start[l] = tiles;
    uint32_t box_rows, box_imgs;
    if (HW >= CV_BM) {
      if (HW % CV_BM != 0) return EMRT_ERR_UNSUPPORTED;
      p.tiles_per_img[l] = HW / CV_BM; p.imgs_per_tile[l] = 1; p.rows_per_tile[l] = CV_BM / W;
      box_rows = CV_BM / W; box_imgs = 1;
      tiles += B * p.tiles_per_img[l];
    } else {
      if (CV_BM % HW != 0) return EMRT_ERR_UNSUPPORTED;
      p.tiles_per_img[l] = 0; p.imgs_per_tile[l] = CV_BM / HW; p.rows_per_tile[l] = H;
      box_rows = H; box_imgs = CV_BM / HW;
      tiles += (B + p.imgs_per_tile[l] - 1) / p.imgs_per_tile[l];
    }
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x + (int64_t)lv.start[l] * C;
    const uint64_t dx[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t sx[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)Lv * C * 2};
    const uint32_t bx[4] = {(uint32_t)CV_BK, (uint32_t)W, box_rows, box_imgs};
    if (int e = make_tensor_map(&p.tma_x[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, xb, dx, sx, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    __nv_bfloat16* yb = (__nv_bfloat16*)y + (int64_t)lv.start[l] * C;
    const uint64_t dy[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)B};
    const uint64_t sy[2] = {(uint64_t)C * 2, (uint64_t)Lv * C * 2};
    const uint32_t by[3] = {32u, 32u, 1u};
    if (int e = make_tensor_map(&p.tma_y[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, yb, dy, sy, by, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
  }
  p.tile_start[L] = tiles;
  const uint64_t dw[2] = {(uint64_t)C, (uint64_t)L * 9 * C};
  const uint64_t sw[1] = {(uint64_t)C * 2};
  // CTA pairs need tile 2 k and 2 k + 1 on the same level (same weights): every level's tile count even
  bool pair = getenv("EMRT_CONV_1CTA") == nullptr && tiles >= 2;
  for (int l = 0; l <= L; ++l) pair = pair && (p.tile_start[l] % 2 == 0);
  const uint32_t bw[2] = {(uint32_t)CV_BK, (uint32_t)(pair ? CV_N / 2 : CV_N)};
  if (int e = make_tensor_map(&p.tma_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, dw, sw, bw, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  if (!stats_ws) return pair ? launch_conv<2, false>(p, tiles, st, max_ctas) : launch_conv<1, false>(p, tiles, st, max_ctas);
  if ((int64_t)tiles * 256 + 2LL * B * L * 32 > conv3x3_stats_workspace_floats(B, Lv, L)) return EMRT_ERR_UNSUPPORTED;
  p.gn_partial = stats_ws + 2LL * B * L * 32;
  if (int e = pair ? launch_conv<2, true>(p, tiles, st, max_ctas) : launch_conv<1, true>(p, tiles, st, max_ctas)) return e;
  conv_gn_finalize_kernel<<<B * L, 256, 0, st>>>(p, stats_ws);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

}  // namespace emrt
