// y = LayerNorm(x[rows,K] @ Wt[256,K]^T + bias + residual) * gamma + beta  — EMRT_EPI_RESIDUAL_LN on tcgen05.
// (output_proj + norm1, transformer_encoder_decoder.py:106,199-200; linear2 + norm2 / norm3, :157-160,282-295.)
//
// The un-fused form writes the projection (rows x 256 x 2 B), then a LayerNorm pass reads it back with the residual and
// writes again: 5 token streams through HBM.  Here N = 256 is ONE accumulator tile row, so every epilogue thread owns a
// whole token (one TMEM lane) and the LayerNorm runs on the fp32 accumulator before anything is rounded: 3 streams
// (x in, residual in, y out), one launch, and one bf16 rounding fewer on the residual stream.
//
// Same skeleton as linear_tcgen05.cu (persistent, warp-specialised, 320 threads, 2-deep TMEM ring = 512 columns):
//   warp 0  TMA producer (A ring; B resident when K <= 256, streamed with A otherwise)
//   warp 1  tcgen05.mma issuer, M = 128, N = 256
//   warps 2..9 epilogue: two warps per TMEM lane quarter, each owning 128 of the 256 columns of its 32 rows.
// Epilogue of a tile, per warp (32 rows x 128 columns, in four 32-column chunks):
//   pass 1  x = acc + bias + residual -> written BACK to TMEM (tcgen05.st), row sums;  residual chunks arrive by the
//           warp's own TMA loads (box 32 rows x 32 columns, SWIZZLE_64B: one 64-byte row per lane, conflict-free) into two
//           2 KB buffers, the next chunk in flight while this one is consumed
//   pass 2  sum (x - mean)^2 from TMEM          (the two warps of a row pair exchange partial sums through shared memory)
//   pass 3  normalise, gamma / beta, round, stage (the same two buffers), TMA store.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace emrt {

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

namespace {

constexpr int BM = 128, BK = 64, BN = 256, UMMA_K = 16;
constexpr int EPI_WARP0 = 2, NUM_EPI_WARPS = 8, NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int CHUNK = 32;                    // columns per epilogue step
constexpr int HALF_N = BN / 2;               // columns per epilogue warp
constexpr int CHUNKS = HALF_N / CHUNK;       // 4
constexpr uint32_t BUF_BYTES = 32 * CHUNK * 2;   // one staged chunk: 32 rows x 64 bytes

struct LnGemmParams {
  CUtensorMap tma_a;     // x        [rows, K]   bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tma_b;     // Wt       [256, K]    bf16, box {64, 256}, SWIZZLE_128B
  CUtensorMap tma_res;   // residual [rows, 256] bf16, box {32, 32},  SWIZZLE_64B
  CUtensorMap tma_y;     // y        [rows, 256] bf16, box {32, 32},  SWIZZLE_64B
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  int32_t K, tiles_m;
  // GN variant: + GELU(GroupNorm_l(conv)) + skip behind the LayerNorm (the encoder layer's conv branch, t_e_d.py:187-189,203)
  CUtensorMap tma_conv;  // conv [rows, 256] bf16, box {32, 32}, SWIZZLE_64B
  CUtensorMap tma_skip;  // skip [rows, 256] bf16, same box
  const float* gn_stats; // [B, L, 32, 2] (sum, sum of squares) of conv per (image, level, group)
  const float* gn_gamma; // [L, 256]
  const float* gn_beta;
  float gn_eps;
  int32_t L, Lv, B;
  LevelTable lv;
};
constexpr int GN_MAX_L = 4, GN_GROUPS = 32;

template <int STAGES, bool B_RES, bool GN = false>
struct LnSmem {
  __nv_bfloat16 a[STAGES][BM * BK];
  __nv_bfloat16 b[B_RES ? 4 : STAGES][BN * BK];
  uint8_t buf[NUM_EPI_WARPS][2][BUF_BYTES];
  uint8_t buf2[GN ? NUM_EPI_WARPS : 1][2][GN ? BUF_BYTES : 16];   // GN: the skip chunks (the conv chunks reuse `buf`)
  float gn_gamma[GN ? GN_MAX_L : 1][GN ? BN : 4], gn_beta[GN ? GN_MAX_L : 1][GN ? BN : 4];
  uint64_t gn_full[NUM_EPI_WARPS][2];
  float bias[BN], gamma[BN], beta[BN];
  float xch[2][2][2][BM];                  // [tile parity][pass][column half][row]
  uint64_t full[STAGES], empty[STAGES];
  uint64_t b_full;
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t res_full[NUM_EPI_WARPS][2];
  uint32_t tmem_base;
};

#define TMEM_ST_X32(taddr, r)                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"                                \
               ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),          \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),    \
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),  \
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),  \
                 "r"(taddr)                                                                                       \
               : "memory")

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int STAGES, bool B_RES, bool GN = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_ln_tcgen05_kernel(const __grid_constant__ LnGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using Smem = LnSmem<STAGES, B_RES, GN>;
  Smem& s = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.K / BK;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = B_RES ? A_BYTES : A_BYTES + B_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    tma_prefetch_desc(&p.tma_res);
    tma_prefetch_desc(&p.tma_y);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.b_full, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) { mbar_init(&s.tmem_full[i], 1); mbar_init(&s.tmem_empty[i], NUM_EPI_WARPS); }
    for (int w = 0; w < NUM_EPI_WARPS; ++w) {
      mbar_init(&s.res_full[w][0], 1); mbar_init(&s.res_full[w][1], 1);
      mbar_init(&s.gn_full[w][0], 1); mbar_init(&s.gn_full[w][1], 1);
    }
    if (GN) { tma_prefetch_desc(&p.tma_conv); tma_prefetch_desc(&p.tma_skip); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < BN; i += NUM_THREADS) {
    s.bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    s.gamma[i] = __ldg(p.gamma + i);
    s.beta[i] = __ldg(p.beta + i);
    if (GN)
      for (int l = 0; l < p.L; ++l) { s.gn_gamma[l][i] = __ldg(p.gn_gamma + l * BN + i); s.gn_beta[l][i] = __ldg(p.gn_beta + l * BN + i); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (B_RES && (int)blockIdx.x < p.tiles_m) {
        mbar_arrive_expect_tx(&s.b_full, B_BYTES * (uint32_t)num_kb);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(s.b[kb], &p.tma_b, &s.b_full, kb * BK, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&s.full[stage], STAGE_BYTES);
          tma_load_2d(s.a[stage], &p.tma_a, &s.full[stage], kb * BK, m * BM);
          if (!B_RES) tma_load_2d(s.b[stage], &p.tma_b, &s.full[stage], kb * BK, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      if (B_RES && (int)blockIdx.x < p.tiles_m) { mbar_wait(&s.b_full, 0); tc_fence_after(); }
      for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
        mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(s.a[stage]));
          const uint64_t db = make_smem_desc(smem_u32(s.b[B_RES ? kb : stage]));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&s.empty[stage]);
          if (kb == num_kb - 1) umma_commit(&s.tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;               // which 128 columns
    const int col0 = half * HALF_N;
    const uint32_t buf0 = smem_u32(s.buf[ew][0]);
    uint64_t* rbar = s.res_full[ew];
    uint32_t rphase = 0u;               // bit b: phase parity of residual buffer b
    int acc = 0, par = 0;
    uint32_t acc_phase = 0;
    // this lane's 64-byte row of a staged chunk: 16-byte piece h sits at (h ^ ((lane >> 1) & 3)) (SWIZZLE_64B)
    const uint32_t my_row = (uint32_t)lane * 64u;
    const uint32_t swz = (uint32_t)((lane >> 1) & 3);

    auto load_res = [&](int m, int c) {     // lane 0 only
      mbar_arrive_expect_tx(&rbar[c & 1], BUF_BYTES);
      tma_load_2d(s.buf[ew][c & 1], &p.tma_res, &rbar[c & 1], col0 + c * CHUNK, m * BM + q * 32);
    };
    if (lane == 0 && (int)blockIdx.x < p.tiles_m) { load_res(blockIdx.x, 0); load_res(blockIdx.x, 1); }
    // GN: conv chunk c -> buf[c & 1] (free once pass 1 has consumed the residual; the output is staged over it in place),
    //     skip chunk c -> buf2[c & 1]; one barrier per buffer pair
    uint64_t* gbar = s.gn_full[ew];
    uint32_t gphase = 0u;
    const uint32_t buf2_0 = smem_u32(s.buf2[GN ? ew : 0][0]);
    auto load_gn = [&](int m, int c) {      // lane 0 only
      mbar_arrive_expect_tx(&gbar[c & 1], 2 * BUF_BYTES);
      tma_load_2d(s.buf[ew][c & 1], &p.tma_conv, &gbar[c & 1], col0 + c * CHUNK, m * BM + q * 32);
      tma_load_2d(s.buf2[GN ? ew : 0][c & 1], &p.tma_skip, &gbar[c & 1], col0 + c * CHUNK, m * BM + q * 32);
    };

    for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
      const int row = q * 32 + lane;        // row inside the tile
      const int row0 = m * BM + q * 32;
      mbar_wait(&s.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0);

      // ---- pass 1: x = acc + bias + residual, back into TMEM; row sum ---------------------------------------
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[32];
        TMEM_LD_X32(t_row + c * CHUNK, r);
        mbar_wait(&rbar[c & 1], (rphase >> (c & 1)) & 1u);
        rphase ^= 1u << (c & 1);
        const uint32_t rb = buf0 + (uint32_t)(c & 1) * BUF_BYTES + my_row;
        uint4 rv[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rv[h].x), "=r"(rv[h].y), "=r"(rv[h].z), "=r"(rv[h].w)
                       : "r"(rb + ((((uint32_t)h) ^ swz) << 4)));
        TMEM_WAIT_X32(r);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint32_t w[4] = {rv[h].x, rv[h].y, rv[h].z, rv[h].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = h * 8 + 2 * i;
            const float x0 = __uint_as_float(r[j]) + s.bias[col0 + c * CHUNK + j] + __uint_as_float(w[i] << 16);
            const float x1 = __uint_as_float(r[j + 1]) + s.bias[col0 + c * CHUNK + j + 1] + __uint_as_float(w[i] & 0xffff0000u);
            sum += x0 + x1;
            r[j] = __float_as_uint(x0);
            r[j + 1] = __float_as_uint(x1);
          }
        }
        TMEM_ST_X32(t_row + c * CHUNK, r);
        __syncwarp();                       // every lane has read this buffer: refill it with the chunk after next
        if (lane == 0 && c + 2 < CHUNKS) load_res(m, c + 2);
      }
      if (GN && lane == 0) { load_gn(m, 0); load_gn(m, 1); }      // both buffers are free: their latency hides behind pass 2
      s.xch[par][0][half][row] = sum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float mean = (s.xch[par][0][0][row] + s.xch[par][0][1][row]) * (1.f / BN);
      tmem_wait_st();

      // ---- pass 2: centred second moment ------------------------------------------------------------------
      float sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[32];
        TMEM_LD_X32(t_row + c * CHUNK, r);
        TMEM_WAIT_X32(r);
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(r[j]) - mean; sq = fmaf(d, d, sq); }
      }
      s.xch[par][1][half][row] = sq;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float rstd = rsqrtf((s.xch[par][1][0][row] + s.xch[par][1][1][row]) * (1.f / BN) + p.eps);

      // ---- pass 3: normalise (+ conv branch), round, stage, TMA store --------------------------------------------
      // GN: this lane's row -> (image, level) -> the group statistics of its 8-column pieces (one group per 16-byte piece)
      int gl = 0;
      const float* gst = nullptr;
      float ginv = 0.f;
      if (GN) {
        const int64_t rg = (int64_t)m * BM + row;
        int b = (int)(rg / p.Lv);
        b = b < p.B ? b : p.B - 1;
        const int t = (int)(rg - (int64_t)b * p.Lv);
        while (gl + 1 < p.L && t >= p.lv.start[gl + 1]) ++gl;
        gst = p.gn_stats + ((int64_t)(b * p.L + gl) * GN_GROUPS) * 2;
        ginv = 1.f / (float)(p.lv.H[gl] * p.lv.W[gl] * (BN / GN_GROUPS));
      }
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[32];
        TMEM_LD_X32(t_row + c * CHUNK, r);
        if (!GN && c >= 2) {                // the store of chunk c - 2 has drained this buffer
          if (lane == 0) tma_store_wait_read_1();
          __syncwarp();
        }
        const uint32_t sb = buf0 + (uint32_t)(c & 1) * BUF_BYTES + my_row;
        uint4 cv[4] = {}, sk[4] = {};
        if (GN) {
          mbar_wait(&gbar[c & 1], (gphase >> (c & 1)) & 1u);
          gphase ^= 1u << (c & 1);
          const uint32_t kb2 = buf2_0 + (uint32_t)(c & 1) * BUF_BYTES + my_row;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            cv[h] = lds128(sb + ((((uint32_t)h) ^ swz) << 4));
            sk[h] = lds128(kb2 + ((((uint32_t)h) ^ swz) << 4));
          }
        }
        TMEM_WAIT_X32(r);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t o[4];
          float gmean = 0.f, grstd = 0.f;
          if (GN) {
            const float2 st2 = __ldg(reinterpret_cast<const float2*>(gst) + ((col0 + c * CHUNK) >> 3) + h);
            gmean = st2.x * ginv;
            grstd = rsqrtf(fmaxf(st2.y * ginv - gmean * gmean, 0.f) + p.gn_eps);
          }
          const uint32_t cw[4] = {cv[h].x, cv[h].y, cv[h].z, cv[h].w}, sw[4] = {sk[h].x, sk[h].y, sk[h].z, sk[h].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = h * 8 + 2 * i, col = col0 + c * CHUNK + j;
            float y0 = (__uint_as_float(r[j]) - mean) * rstd * s.gamma[col] + s.beta[col];
            float y1 = (__uint_as_float(r[j + 1]) - mean) * rstd * s.gamma[col + 1] + s.beta[col + 1];
            if (GN) {
              const float c0 = __uint_as_float(cw[i] << 16), c1 = __uint_as_float(cw[i] & 0xffff0000u);
              y0 += gelu_erf(fmaf(c0 - gmean, grstd * s.gn_gamma[gl][col], s.gn_beta[gl][col])) + __uint_as_float(sw[i] << 16);
              y1 += gelu_erf(fmaf(c1 - gmean, grstd * s.gn_gamma[gl][col + 1], s.gn_beta[gl][col + 1])) + __uint_as_float(sw[i] & 0xffff0000u);
            }
            o[i] = pack2(y0, y1, EMRT_BF16);
          }
          sts128(sb + ((((uint32_t)h) ^ swz) << 4), o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tma_y, buf0 + (uint32_t)(c & 1) * BUF_BYTES, col0 + c * CHUNK, row0);
          tma_store_commit();
          if (GN && c + 2 < CHUNKS) {       // chunk c + 2 goes into the buffers of chunk c: wait until the store has read them
            tma_store_wait_read();
            load_gn(m, c + 2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s.tmem_empty[acc]);
        const int mn = m + gridDim.x;       // the next tile's first two residual chunks, as soon as the stores have drained
        if (mn < p.tiles_m) {
          tma_store_wait_read_1();
          load_res(mn, 0);
          tma_store_wait_read();
          load_res(mn, 1);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      par ^= 1;
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <int STAGES, bool B_RES, bool GN = false>
int launch_ln(LnGemmParams& p, cudaStream_t st) {
  using Smem = LnSmem<STAGES, B_RES, GN>;
  constexpr int smem_bytes = (int)sizeof(Smem) + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB shared-memory limit of one CTA");
  auto kern = linear_ln_tcgen05_kernel<STAGES, B_RES, GN>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int grid = p.tiles_m < num_sms() ? p.tiles_m : num_sms();
  kern<<<grid, NUM_THREADS, smem_bytes, st>>>(p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

}  // namespace

int linear_ln_tcgen05(const emrt_linear_args* a, cudaStream_t st) {
  if (a->N != BN) return set_error(EMRT_ERR_UNSUPPORTED, "RESIDUAL_LN epilogue is built for N = 256 (one accumulator row), got N=%d", a->N);
  if (a->K % BK != 0) return set_error(EMRT_ERR_UNSUPPORTED, "RESIDUAL_LN epilogue needs K %% 64 == 0, got K=%d", a->K);
  if (a->x_dtype != EMRT_BF16 || a->w_dtype != EMRT_BF16 || !a->w_transposed || a->y_dtype != EMRT_BF16)
    return set_error(EMRT_ERR_UNSUPPORTED, "RESIDUAL_LN epilogue needs bf16 x / y / residual and packed bf16 [N,K] weights");
  if (!a->residual || !a->ln_gamma || !a->ln_beta)
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "RESIDUAL_LN needs residual, ln_gamma and ln_beta");
  if (a->rows >= (1LL << 31)) return set_error(EMRT_ERR_UNSUPPORTED, "RESIDUAL_LN: rows must be < 2^31");
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w) | reinterpret_cast<uintptr_t>(a->y) |
       reinterpret_cast<uintptr_t>(a->residual)) & 15)
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "RESIDUAL_LN needs 16-byte aligned x, w, y, residual");
  LnGemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = a->bias; p.gamma = a->ln_gamma; p.beta = a->ln_beta; p.eps = a->ln_eps;
  p.K = a->K;
  p.tiles_m = (int)((a->rows + BM - 1) / BM);
  {
    const uint64_t d[2] = {(uint64_t)a->K, (uint64_t)a->rows}, sb[1] = {(uint64_t)a->K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (int e = make_tensor_map(&p.tma_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->x, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t d[2] = {(uint64_t)a->K, (uint64_t)BN}, sb[1] = {(uint64_t)a->K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    if (int e = make_tensor_map(&p.tma_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->w, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t d[2] = {(uint64_t)BN, (uint64_t)a->rows}, sb[1] = {(uint64_t)BN * 2};
    const uint32_t box[2] = {(uint32_t)CHUNK, 32u};
    if (int e = make_tensor_map(&p.tma_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->residual, d, sb, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (int e = make_tensor_map(&p.tma_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->y, d, sb, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
  }
  if (a->gn) {
    const emrt_gn_branch* g = a->gn;
    if (!g->conv || !g->skip || !g->stats || !g->gamma || !g->beta)
      return set_error(EMRT_ERR_INVALID_ARGUMENT, "RESIDUAL_LN + conv branch: NULL pointer in emrt_gn_branch");
    if (g->groups != GN_GROUPS || g->L < 1 || g->L > GN_MAX_L || g->Lv <= 0 || a->rows % g->Lv != 0)
      return set_error(EMRT_ERR_UNSUPPORTED, "RESIDUAL_LN + conv branch: needs 32 groups, 1..4 levels and rows %% Lv == 0");
    if ((reinterpret_cast<uintptr_t>(g->conv) | reinterpret_cast<uintptr_t>(g->skip)) & 15)
      return set_error(EMRT_ERR_INVALID_ARGUMENT, "RESIDUAL_LN + conv branch: conv / skip must be 16-byte aligned");
    if (int e = fill_levels(p.lv, g->L, g->shapes_hw, nullptr, g->Lv)) return e;
    p.gn_stats = g->stats; p.gn_gamma = g->gamma; p.gn_beta = g->beta; p.gn_eps = g->eps;
    p.L = g->L; p.Lv = g->Lv; p.B = (int)(a->rows / g->Lv);
    const uint64_t d[2] = {(uint64_t)BN, (uint64_t)a->rows}, sb[1] = {(uint64_t)BN * 2};
    const uint32_t box[2] = {(uint32_t)CHUNK, 32u};
    if (int e = make_tensor_map(&p.tma_conv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->conv, d, sb, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (int e = make_tensor_map(&p.tma_skip, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->skip, d, sb, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (a->K <= 4 * BK) return set_error(EMRT_ERR_UNSUPPORTED, "RESIDUAL_LN + conv branch is built for the streamed-B form (K > 256: linear2 of the FFN)");
    return launch_ln<3, false, true>(p, st);
  }
  if (a->K <= 4 * BK) return launch_ln<3, true>(p, st);
  return launch_ln<3, false>(p, st);
}

}  // namespace emrt
