// The encoder layer's FFN as ONE kernel:  y = LayerNorm(x + linear2(relu(linear1(x)))) * gamma + beta  (+ conv branch)
// (forward_ffn + the layer's final add, transformer_encoder_decoder.py:157-160,187-189,203).
//
// Why: as two GEMMs the hidden tensor [rows, 1024] crosses HBM twice — 793 MB written by linear1 (which runs AT this GPU's
// pure-write ceiling, 3.9 TB/s: profiles/r2o_hbm_write_read.txt) and 793 MB read back by linear2: 1.6 GB of the layer's
// 2.6 GB of FFN traffic and 0.5 ms of its 1.9 ms.  Here the hidden activations never leave the SM: a CTA owns a 128-token
// row tile and walks the hidden dimension in chunks of 64 units:
//     GEMM1(c):  hacc[c&1] (TMEM, 64 cols)  = x[128 x 256] @ W1[c*64 .. +64, :]^T          (4 k-blocks, UMMA N = 64)
//     convert:   h_c = bf16(relu(hacc + b1))  -> shared memory, canonical K-major SWIZZLE_128B tile (the A operand of GEMM2)
//     GEMM2(c):  y (TMEM, 256 cols)        += h_c[128 x 64] @ W2[:, c*64 .. +64]^T          (1 k-block, UMMA N = 256)
// and the weights (1 MB per row tile) stream from L2 through a 3 x 32 KB TMA ring, the way the 3x3 conv streams its taps.
// The LayerNorm + conv-branch epilogue of linear_ln_tcgen05.cu follows on the same tile.
//
// Warp roles (448 threads, one CTA per SM, persistent over row tiles):
//   warp 0        TMA producer: x tile (64 KB, once per tile), then W1(0), W1(1), W2(0), W1(2), W2(1), ... in MMA order
//   warp 1        tcgen05.mma issuer: G1(0), [G1(c+1), G2(c)] ... — G1(c+1) runs while chunk c is being converted
//   warps 2..9    H warps: TMEM -> +b1 -> ReLU -> bf16 -> swizzled shared memory (two warps per TMEM lane quarter)
//   warps 10..13  LN warps (one per lane quarter, a thread owns a token): pass 1 adds bias + residual to the y accumulator,
//                 rounds to bf16 and parks the row in 128 spare TMEM columns (packed pairs) — that frees y for the next tile's
//                 GEMM2 after ~1 us instead of after the whole epilogue; passes 2 / 3 (centred variance; normalise + GELU(
//                 GroupNorm(conv)) + skip, TMA store) then run from the parked copy while the tensor pipe works on the next tile.
// TMEM: y [0,256) | hacc0 [256,320) | hacc1 [320,384) | parked pre-LayerNorm row, bf16 pairs [384,512).
// Numerics: linear1's output is rounded to bf16 (as the two-kernel form stores it); linear2 + bias + residual is rounded to
// bf16 once before the LayerNorm (the two-kernel form rounded linear2's output in its first version; oracle: `ffn_fused`).
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace emrt {

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

namespace {

constexpr int BM = 128, BK = 64, DM = 256, CH = 64, UMMA_K = 16;
constexpr int H_WARP0 = 2, NUM_H_WARPS = 8, LN_WARP0 = 10, NUM_LN_WARPS = 4;
constexpr int NUM_THREADS = 32 * (LN_WARP0 + NUM_LN_WARPS);
constexpr int TMEM_COLS = 512, Y_COL = 0, HACC_COL = 256, XP_COL = 384;
constexpr int W_STAGES = 3;
constexpr uint32_t W_STAGE_BYTES = 32768;          // W1 chunk: 4 k-blocks x [64 x 64] bf16; W2 chunk: one [256 x 64] k-block
constexpr int CHUNK = 32;                          // LayerNorm epilogue: columns per step
constexpr int CHUNKS = DM / CHUNK;                 // 8
constexpr uint32_t BUF_BYTES = 32 * CHUNK * 2;     // one staged chunk: 32 rows x 64 bytes
constexpr int GN_MAX_L = 4, GN_GROUPS = 32;

struct FfnParams {
  CUtensorMap tma_x;     // x  [rows, 256] bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tma_w1;    // W1 [F, 256]    bf16, box {64, 64},  SWIZZLE_128B
  CUtensorMap tma_w2;    // W2 [256, F]    bf16, box {64, 256}, SWIZZLE_128B
  CUtensorMap tma_res;   // x  [rows, 256] bf16, box {32, 32},  SWIZZLE_64B (the residual, read back from L2)
  CUtensorMap tma_y;     // y  [rows, 256] bf16, same box
  CUtensorMap tma_conv;  // conv / skip [rows, 256] bf16, same box
  CUtensorMap tma_skip;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  const float* gn_stats; // [B, L, 32, 2]
  const float* gn_gamma; // [L, 256]
  const float* gn_beta;
  float gn_eps;
  int32_t L, Lv, B;
  LevelTable lv;
  int32_t tiles_m, num_chunks;
};

struct FfnSmem {
  __nv_bfloat16 x[DM / BK][BM * BK];               // 64 KB: the row tile, A operand of every GEMM1
  __nv_bfloat16 hs[2][BM * CH];                    // 2 x 16 KB: relu(linear1) chunk, A operand of GEMM2
  uint8_t w[W_STAGES][W_STAGE_BYTES];              // 96 KB weight ring
  uint8_t buf[NUM_LN_WARPS][2][BUF_BYTES];         // residual chunks (pass 1), conv chunks / output staging (pass 3)
  uint8_t buf2[NUM_LN_WARPS][2][BUF_BYTES];        // skip chunks
  uint64_t x_full, x_empty;
  uint64_t w_full[W_STAGES], w_empty[W_STAGES];
  uint64_t hacc_full[2], hacc_empty[2];
  uint64_t hs_full[2], hs_empty[2];
  uint64_t y_full, y_empty;
  uint64_t res_full[NUM_LN_WARPS][2], gn_full[NUM_LN_WARPS][2];
  uint32_t tmem_base;
};

#define TMEM_ST_X16(taddr, r)                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" \
               ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),          \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),    \
                 "r"(taddr)                                                                                       \
               : "memory")

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

template <bool GN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_fused_tcgen05_kernel(const __grid_constant__ FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FfnSmem& s = *reinterpret_cast<FfnSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = p.num_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_x);
    tma_prefetch_desc(&p.tma_w1);
    tma_prefetch_desc(&p.tma_w2);
    tma_prefetch_desc(&p.tma_res);
    tma_prefetch_desc(&p.tma_y);
    if (GN) { tma_prefetch_desc(&p.tma_conv); tma_prefetch_desc(&p.tma_skip); }
    mbar_init(&s.x_full, 1);
    mbar_init(&s.x_empty, 1);
#pragma unroll
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.hacc_full[i], 1); mbar_init(&s.hacc_empty[i], NUM_H_WARPS);
      mbar_init(&s.hs_full[i], NUM_H_WARPS); mbar_init(&s.hs_empty[i], 1);
    }
    mbar_init(&s.y_full, 1);
    mbar_init(&s.y_empty, NUM_LN_WARPS);
    for (int w = 0; w < NUM_LN_WARPS; ++w)
      for (int i = 0; i < 2; ++i) { mbar_init(&s.res_full[w][i], 1); mbar_init(&s.gn_full[w][i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0, xph = 0;
      auto load_w1 = [&](int c) {
        mbar_wait(&s.w_empty[ws], wph ^ 1);
        mbar_arrive_expect_tx(&s.w_full[ws], W_STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < DM / BK; ++kb) tma_load_2d(s.w[ws] + kb * (CH * BK * 2), &p.tma_w1, &s.w_full[ws], kb * BK, c * CH);
        if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
      };
      auto load_w2 = [&](int c) {
        mbar_wait(&s.w_empty[ws], wph ^ 1);
        mbar_arrive_expect_tx(&s.w_full[ws], W_STAGE_BYTES);
        tma_load_2d(s.w[ws], &p.tma_w2, &s.w_full[ws], c * CH, 0);
        if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
      };
      for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
        mbar_wait(&s.x_empty, xph ^ 1);      // the previous tile's last GEMM1 has read x
        xph ^= 1;
        mbar_arrive_expect_tx(&s.x_full, (uint32_t)(BM * DM * 2));
#pragma unroll
        for (int kb = 0; kb < DM / BK; ++kb) tma_load_2d(s.x[kb], &p.tma_x, &s.x_full, kb * BK, m * BM);
        load_w1(0);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) load_w1(c + 1);
          load_w2(c);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(BM, CH), idesc2 = make_idesc(BM, DM);
      int ws = 0;
      uint32_t wph = 0, xph = 0, yph = 0;
      uint32_t hacc_e = 0, hs_f = 0;          // bit b: phase parity of buffer b
      for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
        mbar_wait(&s.x_full, xph);
        xph ^= 1;
        tc_fence_after();
        auto g1 = [&](int c) {
          const int b = c & 1;
          mbar_wait(&s.hacc_empty[b], ((hacc_e >> b) & 1u) ^ 1u);
          hacc_e ^= 1u << b;
          mbar_wait(&s.w_full[ws], wph);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(HACC_COL + b * CH);
#pragma unroll
          for (int kb = 0; kb < DM / BK; ++kb) {
            const uint64_t da = make_smem_desc(smem_u32(s.x[kb]));
            const uint64_t db = make_smem_desc(smem_u32(s.w[ws] + kb * (CH * BK * 2)));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&s.w_empty[ws]);
          umma_commit(&s.hacc_full[b]);
          if (c == NC - 1) umma_commit(&s.x_empty);
          if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
        };
        auto g2 = [&](int c) {
          const int b = c & 1;
          mbar_wait(&s.hs_full[b], (hs_f >> b) & 1u);
          hs_f ^= 1u << b;
          if (c == 0) {                        // the LN warps have taken the previous tile's row sums out of y
            mbar_wait(&s.y_empty, yph ^ 1);
            yph ^= 1;
          }
          mbar_wait(&s.w_full[ws], wph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(s.hs[b]));
          const uint64_t db = make_smem_desc(smem_u32(s.w[ws]));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(tmem_base + Y_COL, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (c | k) != 0 ? 1u : 0u);
          umma_commit(&s.w_empty[ws]);
          umma_commit(&s.hs_empty[b]);
          if (c == NC - 1) umma_commit(&s.y_full);
          if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
        };
        g1(0);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) g1(c + 1);
          g2(c);
        }
      }
    }
  } else if (warp < LN_WARP0) {
    // ===================== H warps: relu(linear1) chunk -> bf16 A operand in shared memory =====================
    const int q = warp & 3;                                 // TMEM lane quarter
    const int hh = (warp - H_WARP0) >> 2;                   // which 32 of the chunk's 64 columns
    const int row = q * 32 + lane;
    uint32_t hacc_f = 0, hs_e = 0;
    const uint32_t my_row = (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
      for (int c = 0; c < NC; ++c) {
        const int b = c & 1;
        mbar_wait(&s.hacc_full[b], (hacc_f >> b) & 1u);
        hacc_f ^= 1u << b;
        tc_fence_after();
        uint32_t r[32];
        TMEM_LD_X32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(HACC_COL + b * CH + hh * 32), r);
        TMEM_WAIT_X32(r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.hacc_empty[b]);
        const float4* b1v = reinterpret_cast<const float4*>(p.b1 + c * CH + hh * 32);
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b1v + j);
          o[2 * j] = pack_relu_bf16x2(__uint_as_float(r[4 * j]) + bb.x, __uint_as_float(r[4 * j + 1]) + bb.y);
          o[2 * j + 1] = pack_relu_bf16x2(__uint_as_float(r[4 * j + 2]) + bb.z, __uint_as_float(r[4 * j + 3]) + bb.w);
        }
        mbar_wait(&s.hs_empty[b], ((hs_e >> b) & 1u) ^ 1u);  // GEMM2(c - 2) has read this buffer
        hs_e ^= 1u << b;
        const uint32_t dst = smem_u32(s.hs[b]) + my_row;
#pragma unroll
        for (int j = 0; j < 4; ++j)            // 16-byte piece hh*4 + j of the row's 128 bytes, SWIZZLE_128B position
          sts128(dst + ((((uint32_t)(hh * 4 + j)) ^ swz) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.hs_full[b]);
      }
    }
  } else {
    // ===================== LN warps: bias + residual + LayerNorm (+ conv branch) =====================
    const int lw = warp - LN_WARP0;
    const int q = warp & 3;
    const uint32_t buf0 = smem_u32(s.buf[lw][0]);
    const uint32_t buf2_0 = smem_u32(s.buf2[lw][0]);
    uint64_t* rbar = s.res_full[lw];
    uint64_t* gbar = s.gn_full[lw];
    uint32_t rphase = 0u, gphase = 0u, yph = 0u;
    const uint32_t my_row = (uint32_t)lane * 64u;
    const uint32_t swz = (uint32_t)((lane >> 1) & 3);
    auto load_res = [&](int m, int c) {     // lane 0 only
      mbar_arrive_expect_tx(&rbar[c & 1], BUF_BYTES);
      tma_load_2d(s.buf[lw][c & 1], &p.tma_res, &rbar[c & 1], c * CHUNK, m * BM + q * 32);
    };
    auto load_gn = [&](int m, int c) {      // lane 0 only
      mbar_arrive_expect_tx(&gbar[c & 1], 2 * BUF_BYTES);
      tma_load_2d(s.buf[lw][c & 1], &p.tma_conv, &gbar[c & 1], c * CHUNK, m * BM + q * 32);
      tma_load_2d(s.buf2[lw][c & 1], &p.tma_skip, &gbar[c & 1], c * CHUNK, m * BM + q * 32);
    };
    if (lane == 0 && (int)blockIdx.x < p.tiles_m) { load_res(blockIdx.x, 0); load_res(blockIdx.x, 1); }

    for (int m = blockIdx.x; m < p.tiles_m; m += gridDim.x) {
      const int row = q * 32 + lane;
      const int row0 = m * BM + q * 32;
      mbar_wait(&s.y_full, yph);
      yph ^= 1;
      tc_fence_after();
      const uint32_t t_y = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)Y_COL;
      const uint32_t t_xp = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)XP_COL;

      // ---- pass 1: x = bf16(acc + bias + residual) parked in TMEM as packed pairs; row sum of the rounded values --------
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[32];
        TMEM_LD_X32(t_y + c * CHUNK, r);
        mbar_wait(&rbar[c & 1], (rphase >> (c & 1)) & 1u);
        rphase ^= 1u << (c & 1);
        const uint32_t rb = buf0 + (uint32_t)(c & 1) * BUF_BYTES + my_row;
        uint4 rv[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) rv[h] = lds128(rb + ((((uint32_t)h) ^ swz) << 4));
        TMEM_WAIT_X32(r);
        uint32_t o[16];
        const float4* b2v = reinterpret_cast<const float4*>(p.b2 + c * CHUNK);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint32_t w[4] = {rv[h].x, rv[h].y, rv[h].z, rv[h].w};
          const float4 ba = __ldg(b2v + 2 * h), bb = __ldg(b2v + 2 * h + 1);
          const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = h * 8 + 2 * i;
            const float x0 = __uint_as_float(r[j]) + bias[2 * i] + bf16_lo(w[i]);
            const float x1 = __uint_as_float(r[j + 1]) + bias[2 * i + 1] + bf16_hi(w[i]);
            const uint32_t pk = pack_bf16x2(x0, x1);
            o[h * 4 + i] = pk;
            sum += bf16_lo(pk) + bf16_hi(pk);
          }
        }
        TMEM_ST_X16(t_xp + c * (CHUNK / 2), o);
        __syncwarp();                       // every lane has read this buffer: refill it with the chunk after next
        if (lane == 0 && c + 2 < CHUNKS) load_res(m, c + 2);
      }
      // y has been read completely: the next tile's GEMM2 may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.y_empty);
      if (GN && lane == 0) { load_gn(m, 0); load_gn(m, 1); }
      const float mean = sum * (1.f / DM);
      tmem_wait_st();

      // ---- pass 2: centred second moment of the parked row --------------------------------------------------
      float sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; c += 2) {
        uint32_t r[32];
        TMEM_LD_X32(t_xp + c * (CHUNK / 2), r);
        TMEM_WAIT_X32(r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d0 = bf16_lo(r[j]) - mean, d1 = bf16_hi(r[j]) - mean;
          sq = fmaf(d0, d0, sq);
          sq = fmaf(d1, d1, sq);
        }
      }
      const float rstd = rsqrtf(sq * (1.f / DM) + p.eps);

      // ---- pass 3: normalise (+ conv branch), round, stage, TMA store ------------------------------------------
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
        ginv = 1.f / (float)(p.lv.H[gl] * p.lv.W[gl] * (DM / GN_GROUPS));
      }
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[16];
        TMEM_LD_X16(t_xp + c * (CHUNK / 2), r);
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
        TMEM_WAIT_X16(r);
        const float4* gav = reinterpret_cast<const float4*>(p.gamma + c * CHUNK);
        const float4* bev = reinterpret_cast<const float4*>(p.beta + c * CHUNK);
        const float4* ggv = reinterpret_cast<const float4*>(p.gn_gamma + (GN ? gl * DM : 0) + c * CHUNK);
        const float4* gbv = reinterpret_cast<const float4*>(p.gn_beta + (GN ? gl * DM : 0) + c * CHUNK);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t o[4];
          float gmean = 0.f, grstd = 0.f;
          if (GN) {
            const float2 st2 = __ldg(reinterpret_cast<const float2*>(gst) + ((c * CHUNK) >> 3) + h);
            gmean = st2.x * ginv;
            grstd = rsqrtf(fmaxf(st2.y * ginv - gmean * gmean, 0.f) + p.gn_eps);
          }
          const float4 g0 = __ldg(gav + 2 * h), g1 = __ldg(gav + 2 * h + 1), e0 = __ldg(bev + 2 * h), e1 = __ldg(bev + 2 * h + 1);
          const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float be[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
          float gg[8] = {}, gb[8] = {};
          if (GN) {
            const float4 a0 = __ldg(ggv + 2 * h), a1 = __ldg(ggv + 2 * h + 1), c0 = __ldg(gbv + 2 * h), c1 = __ldg(gbv + 2 * h + 1);
            gg[0] = a0.x; gg[1] = a0.y; gg[2] = a0.z; gg[3] = a0.w; gg[4] = a1.x; gg[5] = a1.y; gg[6] = a1.z; gg[7] = a1.w;
            gb[0] = c0.x; gb[1] = c0.y; gb[2] = c0.z; gb[3] = c0.w; gb[4] = c1.x; gb[5] = c1.y; gb[6] = c1.z; gb[7] = c1.w;
          }
          const uint32_t cw[4] = {cv[h].x, cv[h].y, cv[h].z, cv[h].w}, sw[4] = {sk[h].x, sk[h].y, sk[h].z, sk[h].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t xw = r[h * 4 + i];
            float y0 = (bf16_lo(xw) - mean) * rstd * ga[2 * i] + be[2 * i];
            float y1 = (bf16_hi(xw) - mean) * rstd * ga[2 * i + 1] + be[2 * i + 1];
            if (GN) {
              y0 += gelu_erf(fmaf(bf16_lo(cw[i]) - gmean, grstd * gg[2 * i], gb[2 * i])) + bf16_lo(sw[i]);
              y1 += gelu_erf(fmaf(bf16_hi(cw[i]) - gmean, grstd * gg[2 * i + 1], gb[2 * i + 1])) + bf16_hi(sw[i]);
            }
            o[i] = pack_bf16x2(y0, y1);
          }
          sts128(sb + ((((uint32_t)h) ^ swz) << 4), o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tma_y, buf0 + (uint32_t)(c & 1) * BUF_BYTES, c * CHUNK, row0);
          tma_store_commit();
          if (GN && c + 2 < CHUNKS) {       // chunk c + 2 goes into the buffers of chunk c: wait until the store has read them
            tma_store_wait_read();
            load_gn(m, c + 2);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        const int mn = m + gridDim.x;       // the next tile's first two residual chunks, as soon as the stores have drained
        if (mn < p.tiles_m) {
          tma_store_wait_read_1();
          load_res(mn, 0);
          tma_store_wait_read();
          load_res(mn, 1);
        }
      }
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

template <bool GN>
int launch_ffn(FfnParams& p, cudaStream_t st) {
  constexpr int smem_bytes = (int)sizeof(FfnSmem) + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB shared-memory limit of one CTA");
  auto kern = ffn_fused_tcgen05_kernel<GN>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int grid = p.tiles_m < num_sms() ? p.tiles_m : num_sms();
  kern<<<grid, NUM_THREADS, smem_bytes, st>>>(p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

}  // namespace
}  // namespace emrt

using namespace emrt;

extern "C" int emrt_ffn_fused_fwd(const emrt_ffn_args* a, void* stream) {
  EMRT_REQUIRE(a != nullptr, "args is NULL");
  EMRT_REQUIRE(a->x && a->w1 && a->w2 && a->y && a->b1 && a->b2 && a->ln_gamma && a->ln_beta, "NULL pointer in emrt_ffn_args");
  EMRT_REQUIRE(a->rows > 0 && a->rows < (1LL << 31), "rows out of range");
  if (a->d_model != DM) return set_error(EMRT_ERR_UNSUPPORTED, "fused FFN is built for d_model = 256 (one accumulator row), got %d", a->d_model);
  if (a->d_ff < CH || a->d_ff % CH != 0) return set_error(EMRT_ERR_UNSUPPORTED, "fused FFN needs d_ff %% 64 == 0, got %d", a->d_ff);
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w1) | reinterpret_cast<uintptr_t>(a->w2) |
       reinterpret_cast<uintptr_t>(a->y) | reinterpret_cast<uintptr_t>(a->b1) | reinterpret_cast<uintptr_t>(a->b2) |
       reinterpret_cast<uintptr_t>(a->ln_gamma) | reinterpret_cast<uintptr_t>(a->ln_beta)) & 15)
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "fused FFN needs 16-byte aligned tensors and parameter vectors");
  if (int e = emrt_device_check()) return e;
  cudaStream_t st = as_stream(stream);
  FfnParams p;
  memset(&p, 0, sizeof(p));
  p.b1 = a->b1; p.b2 = a->b2; p.gamma = a->ln_gamma; p.beta = a->ln_beta; p.eps = a->ln_eps;
  p.tiles_m = (int)((a->rows + BM - 1) / BM);
  p.num_chunks = a->d_ff / CH;
  {
    const uint64_t d[2] = {(uint64_t)DM, (uint64_t)a->rows}, sb[1] = {(uint64_t)DM * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (int e = make_tensor_map(&p.tma_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->x, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    const uint32_t box2[2] = {(uint32_t)CHUNK, 32u};
    if (int e = make_tensor_map(&p.tma_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->x, d, sb, box2, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (int e = make_tensor_map(&p.tma_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->y, d, sb, box2, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
  }
  {
    const uint64_t d[2] = {(uint64_t)DM, (uint64_t)a->d_ff}, sb[1] = {(uint64_t)DM * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)CH};
    if (int e = make_tensor_map(&p.tma_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->w1, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t d[2] = {(uint64_t)a->d_ff, (uint64_t)DM}, sb[1] = {(uint64_t)a->d_ff * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)DM};
    if (int e = make_tensor_map(&p.tma_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->w2, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  if (a->gn) {
    const emrt_gn_branch* g = a->gn;
    if (!g->conv || !g->skip || !g->stats || !g->gamma || !g->beta)
      return set_error(EMRT_ERR_INVALID_ARGUMENT, "fused FFN + conv branch: NULL pointer in emrt_gn_branch");
    if (g->groups != GN_GROUPS || g->L < 1 || g->L > GN_MAX_L || g->Lv <= 0 || a->rows % g->Lv != 0)
      return set_error(EMRT_ERR_UNSUPPORTED, "fused FFN + conv branch: needs 32 groups, 1..4 levels and rows %% Lv == 0");
    if ((reinterpret_cast<uintptr_t>(g->conv) | reinterpret_cast<uintptr_t>(g->skip) | reinterpret_cast<uintptr_t>(g->gamma) |
         reinterpret_cast<uintptr_t>(g->beta)) & 15)
      return set_error(EMRT_ERR_INVALID_ARGUMENT, "fused FFN + conv branch: conv / skip / gamma / beta must be 16-byte aligned");
    if (int e = fill_levels(p.lv, g->L, g->shapes_hw, nullptr, g->Lv)) return e;
    p.gn_stats = g->stats; p.gn_gamma = g->gamma; p.gn_beta = g->beta; p.gn_eps = g->eps;
    p.L = g->L; p.Lv = g->Lv; p.B = (int)(a->rows / g->Lv);
    const uint64_t d[2] = {(uint64_t)DM, (uint64_t)a->rows}, sb[1] = {(uint64_t)DM * 2};
    const uint32_t box[2] = {(uint32_t)CHUNK, 32u};
    if (int e = make_tensor_map(&p.tma_conv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->conv, d, sb, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (int e = make_tensor_map(&p.tma_skip, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->skip, d, sb, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    return launch_ffn<true>(p, st);
  }
  p.gn_gamma = a->ln_gamma; p.gn_beta = a->ln_beta;     // never read; keeps the pointer arithmetic defined
  return launch_ffn<false>(p, st);
}
