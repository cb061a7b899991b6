// The encoder layer's FFN as ONE kernel:  y = LayerNorm(x + linear2(relu(linear1(x)))) * gamma + beta  (+ conv branch)
// (forward_ffn + the layer's final add, transformer_encoder_decoder.py:157-160,187-189,203).
//
// Why: as two GEMMs the hidden tensor [rows, 1024] crosses HBM twice — 793 MB written by linear1 (which runs AT this GPU's
// pure-write ceiling, 3.9 TB/s: profiles/r2o_hbm_write_read.txt) and 793 MB read back by linear2: 1.6 GB of the layer's
// 2.6 GB of FFN traffic and 0.5 ms of its 1.9 ms.  Here the hidden activations never leave the SM: a CTA owns a 128-token
// row tile and walks the hidden dimension in chunks of 128 units:
//     GEMM1(c):  hacc (TMEM, 128 cols)  = x[128 x 256] @ W1[128 c .. +128, :]^T              (4 k-blocks, UMMA N = 128)
//     convert:   h_c = bf16(relu(hacc + b1)) -> shared memory, two canonical K-major SWIZZLE_128B [128 x 64] tiles
//     GEMM2(c):  y (TMEM, 256 cols)    += h_c[128 x 128] @ W2[:, 128 c .. +128]^T            (2 k-blocks, UMMA N = 256)
// and the weights (1 MB per row tile) stream from L2 through a ring of six 16 KB TMA units, the way the 3x3 conv streams its
// taps.  The LayerNorm + conv-branch epilogue of linear_ln_tcgen05.cu follows on the same tile.
//
// What paces it (measured, EMRT_FFN_PROF): the tensor pipe's OPERAND FETCH.  A cta_group::1 tcgen05.mma reads A and B from
// shared memory at 64 B/clk, so one M = 128, K = 16 instruction costs (4096 + 32 N) / 64 = 64 + N/2 clocks, not the N/2 of
// the arithmetic: N = 256 runs at 67 % of the nominal rate (which is where cuBLAS's measured 1.6 PFLOP/s sits against the
// nominal 2.4), N = 128 at 50 %, N = 64 at 33 %.  The first version of this kernel (64-unit chunks: N = 64 for GEMM1) and the
// second (N = 128 everywhere) both ran at exactly the time this model gives.  Hence: GEMM2 with N = 256, GEMM1 with the
// largest N the 512 TMEM columns leave room for (128).
//
// Warp roles (448 threads, one CTA per SM, persistent over row tiles):
//   warp 0        TMA producer: x tile (64 KB, once per tile), then the weight units in MMA order
//   warp 1        tcgen05.mma issuer: G1(0), then [G1(c+1), G2(c)] — GEMM2(c-1) and GEMM1(c+1) run while chunk c is converted
//   warps 2..5    H warps (one per TMEM lane quarter): TMEM -> +b1 -> ReLU -> bf16 -> swizzled shared memory
//   warps 6..13   LN warps (two per lane quarter, 128 columns each; a thread owns half a token): pass 1 adds bias + residual
//                 to the y accumulator, rounds to bf16 and parks the row in 128 spare TMEM columns (packed pairs) — that frees
//                 y for the next tile's GEMM2 after ~1 us instead of after the whole epilogue; passes 2 / 3 (centred variance;
//                 normalise + GELU(GroupNorm(conv)) + skip, TMA store) run from the parked copy under the next tile's MMAs.
// TMEM: y [0,256) | hacc [256,384) | parked pre-LayerNorm row, bf16 pairs [384,512).
// Numerics: linear1's output is rounded to bf16 (as the two-kernel form stores it); linear2 + bias + residual is rounded to
// bf16 once before the LayerNorm (oracle: kernel_storage_rounding models both).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

namespace emrt {

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

namespace {

constexpr int BM = 128, BK = 64, DM = 256, CH = 128, UMMA_K = 16;
constexpr int NUM_H_WARPS = 4, LN_WARP0 = 6, NUM_LN_WARPS = 8;      // warps 0 / 1: TMA producer / MMA issuer, 2..5: H, 6..13: LN
constexpr int NUM_THREADS = 32 * (LN_WARP0 + NUM_LN_WARPS);
constexpr int TMEM_COLS = 512, Y_COL = 0, HACC_COL = 256, XP_COL = 384;
// weight ring (96 KB): units of one [128 / CG rows x 64 k] bf16 box — a W1 k-block of this CTA's share of the chunk; a W2
// k-block (this CTA's 256 / CG output rows) takes two adjacent units.  CG = 2: the CTA pair splits every B operand in half.
template <int CG> struct Ring {
  static constexpr int UNITS = 6 * CG;
  static constexpr uint32_t UNIT_BYTES = (128 / CG) * BK * 2;
};
constexpr int HS_SLOTS = 2;                        // hidden k-block tiles [128 x 64] between the H warps and GEMM2
constexpr int LCH = 16;                            // LayerNorm epilogue: columns per step
constexpr int LN_COLS = DM / 2;                    // columns per LN warp
constexpr int LCHUNKS = LN_COLS / LCH;             // 8
constexpr int LN_BUFS = 4;                         // chunk buffers per LN warp
constexpr uint32_t LBUF_BYTES = 32 * LCH * 2;      // one staged chunk: 32 rows x 32 bytes
constexpr int GN_MAX_L = 4, GN_GROUPS = 32;

struct FfnParams {
  CUtensorMap tma_x;     // x  [rows, 256] bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tma_w1;    // W1 [F, 256]    bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tma_w2;    // W2 [256, F]    bf16, box {64, 256}, SWIZZLE_128B
  CUtensorMap tma_res;   // x  [rows, 256] bf16, box {16, 32},  SWIZZLE_32B (the residual, read back from L2)
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
  int32_t debug;         // timing experiments (wrong results): 1 = LN warps stop after pass 1, 4 = no conv / skip loads and no stores
  long long* prof;       // EMRT_FFN_PROF: per CTA, cycles the MMA thread spent waiting on each barrier kind
  long long* prof_w;     // ... and the weight waits split: first W1 unit of a chunk / other W1 units / W2 units
  long long* prof_ln;    // ... and LN warp 0 of each CTA: waiting for y_full / pass 1 / merge of the halves / pass 3
};

template <int CG>
struct FfnSmem {
  __nv_bfloat16 x[DM / BK][BM * BK];               // 64 KB: the row tile, A operand of every GEMM1
  __nv_bfloat16 hs[HS_SLOTS][BM * BK];             // 2 x 16 KB: relu(linear1) k-blocks, A operand of GEMM2
  uint8_t w[Ring<CG>::UNITS][Ring<CG>::UNIT_BYTES];   // 96 KB weight ring
  uint8_t lbuf[NUM_LN_WARPS][LN_BUFS][LBUF_BYTES]; // per LN warp: ring of 1 KB chunk buffers (residual | conv, skip / output)
  float xch[2][2][BM];                             // [pass][column half][row]: partial row sums of the two LN warps of a row
  uint64_t x_full, x_empty;
  uint64_t w_full[Ring<CG>::UNITS], w_empty[Ring<CG>::UNITS];
  uint64_t hacc_full[2], hacc_empty[2];             // per 64-column half of the GEMM1 accumulator
  uint64_t hs_full[HS_SLOTS], hs_empty[HS_SLOTS];
  uint64_t y_full, y_empty;
  uint64_t l_full[NUM_LN_WARPS][LN_BUFS];
  uint32_t tmem_base;
};

#define TMEM_ST_X8(taddr, r)                                                                                      \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"                            \
               ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(taddr) : "memory")

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
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
__device__ __forceinline__ uint32_t pack_relu_pair(f32x2_t v) {
  float a, b;
  upk2(v, a, b);
  return pack_relu_bf16x2(a, b);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

#ifdef EMRT_FFN_WATCHDOG
// debug build: every wait of this file reports what it is stuck on instead of hanging the GPU
__device__ __noinline__ void wd_wait(uint64_t* bar, uint32_t parity, int line) {
  for (long long it = 0;; ++it) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (it > 4000000) {
      printf("ffn_fused watchdog: block %d warp %d lane %d stuck at line %d (parity %u)\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
             (int)(threadIdx.x & 31), line, parity);
      __trap();
    }
  }
}
#define mbar_wait(bar, parity) wd_wait(bar, parity, __LINE__)
#endif

#ifdef EMRT_FFN_WATCHDOG
#define wait_lead_line(CGV, bar, parity) wd_wait(bar, parity, __LINE__)
#else
#define wait_lead_line(CGV, bar, parity) wait_lead<CGV>(bar, parity)
#endif

template <bool GN, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_fused_tcgen05_kernel(const __grid_constant__ FfnParams p) {
  using R = Ring<CG>;
  constexpr int W_UNITS = R::UNITS;
  constexpr uint32_t W_UNIT_BYTES = R::UNIT_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t align_off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  FfnSmem<CG>& s = *reinterpret_cast<FfnSmem<CG>*>(smem_raw + align_off);
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int cluster_id = (int)blockIdx.x / CG, num_clusters = (int)gridDim.x / CG;
  const int npairs = (p.tiles_m + CG - 1) / CG;       // row tiles are dealt to the pair two at a time: tile CG * mp + rank
  {
    // the launch asks for sizeof(FfnSmem) plus whatever slack the 227 KB limit leaves: fail loudly if the window's base is
    // not aligned well enough for the struct to fit behind the SWIZZLE_128B alignment
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (align_off + (uint32_t)sizeof(FfnSmem<CG>) > dyn) {
      if (threadIdx.x == 0 && blockIdx.x == 0) printf("ffn_fused: shared window misaligned by %u bytes, %u available\n", align_off, dyn);
      __trap();
    }
  }

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
    for (int i = 0; i < W_UNITS; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.hacc_full[i], 1);
      mbar_init(&s.hacc_empty[i], CG * NUM_H_WARPS);   // (the leader's copy collects the arrivals of both CTAs)
    }
#pragma unroll
    for (int i = 0; i < HS_SLOTS; ++i) { mbar_init(&s.hs_full[i], CG * NUM_H_WARPS); mbar_init(&s.hs_empty[i], 1); }
    mbar_init(&s.y_full, 1);
    mbar_init(&s.y_empty, CG * NUM_LN_WARPS);
    for (int w = 0; w < NUM_LN_WARPS; ++w)
      for (int i = 0; i < LN_BUFS; ++i) mbar_init(&s.l_full[w][i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // both CTAs' barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  const uint32_t L_hacc_empty[2] = {leader_addr<CG>(&s.hacc_empty[0]), leader_addr<CG>(&s.hacc_empty[1])};
  const uint32_t L_y_empty = leader_addr<CG>(&s.y_empty);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (both CTAs of a pair run it: each loads its own x tile and its own half of every weight box; the bytes of both are
    // counted on the leader's barrier, which only the leader arms)
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0, xph = 0;
      const uint32_t L_x_full = leader_addr<CG>(&s.x_full);
      auto advance = [&]() { if (++ws == W_UNITS) { ws = 0; wph ^= 1; } };
      long long t_pw = 0;
      const long long t_p0 = clock64();
      // four units: this CTA's [128 / CG hidden x 64 k] boxes of chunk c.  ONE full barrier for the chunk (its first unit's,
      // armed with the bytes of all four): every barrier the issuing warp has to poll costs it ~80 clocks even when the
      // data is long there, and twelve polls per chunk were a fifth of its time (EMRT_FFN_PROF)
      auto load_w1 = [&](int c) {
        const int ws0 = ws;
        const uint32_t wph0 = wph;
        for (int kb = 0; kb < DM / BK; ++kb) {
          if (p.prof) { const long long t0_ = clock64(); mbar_wait(&s.w_empty[ws], wph ^ 1); t_pw += clock64() - t0_; }
          else mbar_wait(&s.w_empty[ws], wph ^ 1);
          advance();
        }
        ws = ws0;
        wph = wph0;
        const uint32_t lead = leader_addr<CG>(&s.w_full[ws0]);
        if (rank == 0) mbar_arrive_expect_tx(&s.w_full[ws0], CG * (DM / BK) * W_UNIT_BYTES);
        for (int kb = 0; kb < DM / BK; ++kb) {
          if (rank == 0 && kb > 0) mbar_arrive(&s.w_full[ws]);        // keeps the unit's barrier phase in step
          // the unit holds this CTA's B rows of the chunk's two 64-unit halves, half a first: the pair's CTAs split each half
          // (rows [32 r, +32) of it each), so that a half's 64 hidden units stay contiguous across the pair
          if (CG == 2) {
            tma_load_2d_lead<CG>(s.w[ws], &p.tma_w1, lead, kb * BK, c * CH + (int)rank * 32);
            tma_load_2d_lead<CG>(s.w[ws] + W_UNIT_BYTES / 2, &p.tma_w1, lead, kb * BK, c * CH + 64 + (int)rank * 32);
          } else {
            tma_load_2d_lead<CG>(s.w[ws], &p.tma_w1, lead, kb * BK, c * CH);
          }
          advance();
        }
      };
      auto load_w2 = [&](int j) {            // two adjacent units (ws is even here): this CTA's [256 / CG out x 64 k] box of k-block j
        if (p.prof) { const long long t0_ = clock64(); mbar_wait(&s.w_empty[ws], wph ^ 1); mbar_wait(&s.w_empty[ws + 1], wph ^ 1); t_pw += clock64() - t0_; }
        else { mbar_wait(&s.w_empty[ws], wph ^ 1); mbar_wait(&s.w_empty[ws + 1], wph ^ 1); }
        if (rank == 0) {
          mbar_arrive_expect_tx(&s.w_full[ws], CG * 2 * W_UNIT_BYTES);
          mbar_arrive(&s.w_full[ws + 1]);    // the second unit's barrier only keeps its phase in step
        }
        tma_load_2d_lead<CG>(s.w[ws], &p.tma_w2, leader_addr<CG>(&s.w_full[ws]), j * BK, (int)rank * (DM / CG));
        advance();
        advance();
      };
      auto load_x = [&](int mp) {
        const int m = mp * CG + (int)rank;
        mbar_wait(&s.x_empty, xph ^ 1);      // the previous tile's last GEMM1 has read x
        xph ^= 1;
        if (rank == 0) mbar_arrive_expect_tx(&s.x_full, (uint32_t)(CG * BM * DM * 2));
#pragma unroll
        for (int kb = 0; kb < DM / BK; ++kb) tma_load_2d_lead<CG>(s.x[kb], &p.tma_x, L_x_full, kb * BK, m * BM);
      };
      // The weight units form one stream across tiles (ring order = MMA order).  The NEXT tile's x is requested inside this
      // tile's last chunk — right where its buffer frees (the last GEMM1) — so that the wait for it does not hold back the
      // next tile's first weight units: with the x wait at the top of the tile the issuing thread idled ~5 k clocks per tile
      // pair on weights that could have been in flight (EMRT_FFN_PROF).
      if (cluster_id < npairs) load_x(cluster_id);
      for (int mp = cluster_id; mp < npairs; mp += num_clusters) {
        load_w1(0);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) load_w1(c + 1);
          if (c == NC - 1 && mp + num_clusters < npairs) load_x(mp + num_clusters);
          load_w2(2 * c);
          load_w2(2 * c + 1);
        }
      }
      if (p.prof && rank == 0) { p.prof[(size_t)blockIdx.x * 8 + 6] = t_pw; p.prof[(size_t)blockIdx.x * 8 + 7] = clock64() - t_p0; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop (warp-uniform control flow, barrier waits by every lane) and one elected lane issues the
    // tcgen05 instructions: the descriptors then live in uniform registers.  Under `if (lane == 0)` the compiler wrapped
    // every tcgen05.mma in a per-active-lane election loop and moved each operand through R2UR — ~65 instructions per four
    // MMAs on one thread's dependency chain, which (not the tensor pipe, 40 % active) was what set the pace.
    if (rank == 0) {
      constexpr uint32_t idesc1 = make_idesc(BM * CG, CH / 2), idesc2 = make_idesc(BM * CG, DM);
      int ws = 0;
      uint32_t wph = 0, xph = 0, yph = 0, hacc_e = 0;
      uint32_t jg = 0;                        // hidden k-blocks consumed so far (runs across tiles): hs slot = jg & 1
      long long t_x = 0, t_w = 0, t_hacc = 0, t_hs = 0, t_y = 0, t_w1a = 0, t_w2 = 0;
      const long long t_begin = clock64();
#define PROF_WAIT(acc, ...) do { if (p.prof) { const long long t0_ = clock64(); __VA_ARGS__; acc += clock64() - t0_; } else { __VA_ARGS__; } } while (0)
      auto advance = [&]() { if (++ws == W_UNITS) { ws = 0; wph ^= 1; } };
      const uint32_t x_base = smem_u32(s.x[0]), w_base = smem_u32(s.w[0]), hs_base = smem_u32(s.hs[0]);
      const uint64_t desc_hi = make_smem_desc(0);           // everything but the start address
      auto desc_of = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      for (int mp = cluster_id; mp < npairs; mp += num_clusters) {
        PROF_WAIT(t_x, wait_lead_line(CG, &s.x_full, xph));
        xph ^= 1;
        tc_fence_after();
        // GEMM1 of chunk c in two N = 64 halves, each into its own 64 accumulator columns: the H warps read half a while the
        // tensor pipe fills half b, and the next chunk's half a only needs THAT half read — the accumulator is double-buffered
        // inside the 128 columns the LayerNorm's parked row leaves free.  Both halves read the chunk's four W1 units (half b
        // releases them).
        auto g1 = [&](int c) {
          constexpr uint32_t HALF_B_BYTES = W_UNIT_BYTES / 2;
          const int ws0 = ws;
          const uint32_t wph0 = wph;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            PROF_WAIT(t_hacc, wait_lead_line(CG, &s.hacc_empty[hf], hacc_e ^ 1u));   // the H warps have read this half of the previous chunk
            if (hf == 0) PROF_WAIT(t_w1a, wait_lead_line(CG, &s.w_full[ws0], wph0));      // one barrier for the chunk's four units
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)(HACC_COL + hf * 64);
            // one elected block for the half's sixteen MMAs (nothing to wait for between the chunk's four units any more):
            // at N = 64 an MMA is 32 tensor clocks, so every instruction between two of them counts
            if (elect_one()) {
              int u = ws0;
#pragma unroll
              for (int kb = 0; kb < DM / BK; ++kb) {
                const uint64_t da = desc_of(x_base + (uint32_t)kb * (BM * BK * 2));
                const uint64_t db = desc_of(w_base + (uint32_t)u * W_UNIT_BYTES + (uint32_t)hf * HALF_B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  umma_cg<CG>(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
                if (hf == 1) commit_cg<CG>(&s.w_empty[u]);
                if (++u == W_UNITS) u = 0;
              }
              commit_cg<CG>(&s.hacc_full[hf]);
              if (hf == 1 && c == NC - 1) commit_cg<CG>(&s.x_empty);
            }
            __syncwarp();
          }
          ws = ws0;
          wph = wph0;
#pragma unroll
          for (int kb = 0; kb < DM / BK; ++kb) advance();
          hacc_e ^= 1u;
        };
        auto g2 = [&](int j) {                 // hidden k-block j of this tile: y += h[:, 64 j .. +64] @ W2[:, 64 j .. +64]^T
          const uint32_t slot = jg & 1u;
          PROF_WAIT(t_hs, wait_lead_line(CG, &s.hs_full[slot], (jg >> 1) & 1u));
          ++jg;
          if (j == 0) {                        // the LN warps have taken the previous tile's row out of y
            PROF_WAIT(t_y, wait_lead_line(CG, &s.y_empty, yph ^ 1));
            yph ^= 1;
          }
          PROF_WAIT(t_w2, wait_lead_line(CG, &s.w_full[ws], wph));       // (the pair's second barrier only keeps its phase)
          tc_fence_after();
          const uint64_t da = desc_of(hs_base + slot * (uint32_t)(BM * BK * 2));
          const uint64_t db = desc_of(w_base + (uint32_t)ws * W_UNIT_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_cg<CG>(tmem_base + (uint32_t)Y_COL, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (j | k) != 0 ? 1u : 0u);
            commit_cg<CG>(&s.w_empty[ws]);
            commit_cg<CG>(&s.w_empty[ws + 1]);
            commit_cg<CG>(&s.hs_empty[slot]);
            if (j == 2 * NC - 1) commit_cg<CG>(&s.y_full);
          }
          __syncwarp();
          advance();
          advance();
        };
        // G1(c+1) is queued before G2(c): while chunk c is converted (TMEM -> registers -> shared memory) the tensor pipe
        // has G2(c-1) and G1(c+1) to run
        g1(0);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) g1(c + 1);
          g2(2 * c);
          g2(2 * c + 1);
        }
      }
      if (p.prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 8;
        o[0] = clock64() - t_begin; o[1] = t_x; o[2] = t_w + t_w1a + t_w2; o[3] = t_hacc; o[4] = t_hs; o[5] = t_y;
        if (p.prof_w) { p.prof_w[(size_t)blockIdx.x * 4] = t_w1a; p.prof_w[(size_t)blockIdx.x * 4 + 1] = t_w; p.prof_w[(size_t)blockIdx.x * 4 + 2] = t_w2; }
      }
#undef PROF_WAIT
    }
  } else if (warp < LN_WARP0) {
    // ===================== H warps: relu(linear1) chunk -> bf16 A operand in shared memory =====================
    // warp q: rows [32 q, +32) of the chunk's 128 accumulator columns = two k-blocks; per k-block one full 128-byte swizzled
    // row per lane of the [128 x 64] tile
    const int q = warp & 3;                                 // TMEM lane quarter
    const int row = q * 32 + lane;
    uint32_t hacc_f = 0;
    uint32_t cg = 0;                                        // chunks converted so far (runs across tiles)
    const uint32_t my_row = (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t t_h = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)HACC_COL;
    const uint32_t L_hs_full[2] = {leader_addr<CG>(&s.hs_full[0]), leader_addr<CG>(&s.hs_full[1])};
    for (int mp = cluster_id; mp < npairs; mp += num_clusters) {
      for (int c = 0; c < NC; ++c, ++cg) {
        // the chunk's linear1 bias (the same for every lane: L1 broadcast); the first quarter before the wait, so that its
        // latency is not part of the conversion's critical path
        const float4* b1v = reinterpret_cast<const float4*>(p.b1 + c * CH);
        float4 bias[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bias[j] = __ldg(b1v + j);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {       // hidden k-block kk of the chunk: accumulator columns [64 kk, +64), its own barrier pair
          mbar_wait(&s.hacc_full[kk], hacc_f);
          tc_fence_after();
          uint32_t ra[32], rb[32];
          TMEM_LD_X32(t_h + kk * 64, ra);
          TMEM_LD_X32(t_h + kk * 64 + 32, rb);
          TMEM_WAIT_X32(ra);
          TMEM_WAIT_X32(rb);
          // this half of hacc has been read: GEMM1 of the next chunk may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader<CG>(L_hacc_empty[kk]);
          uint32_t o[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = kk == 0 ? bias[j] : __ldg(b1v + 16 + j);
            o[2 * j] = pack_relu_pair(add2(pk2(__uint_as_float(ra[4 * j]), __uint_as_float(ra[4 * j + 1])), pk2(bb.x, bb.y)));
            o[2 * j + 1] = pack_relu_pair(add2(pk2(__uint_as_float(ra[4 * j + 2]), __uint_as_float(ra[4 * j + 3])), pk2(bb.z, bb.w)));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b1v + kk * 16 + 8 + j);
            o[16 + 2 * j] = pack_relu_pair(add2(pk2(__uint_as_float(rb[4 * j]), __uint_as_float(rb[4 * j + 1])), pk2(bb.x, bb.y)));
            o[16 + 2 * j + 1] = pack_relu_pair(add2(pk2(__uint_as_float(rb[4 * j + 2]), __uint_as_float(rb[4 * j + 3])), pk2(bb.z, bb.w)));
          }
          // slot kk has been used cg times before: the GEMM2 of its previous use (k-block kk of chunk cg - 1) is done
          mbar_wait(&s.hs_empty[kk], (cg & 1u) ^ 1u);
          const uint32_t dst = smem_u32(s.hs[kk]) + my_row;
#pragma unroll
          for (int j = 0; j < 8; ++j)          // 16-byte piece j of the row's 128 bytes, at its SWIZZLE_128B position
            sts128(dst + ((((uint32_t)j) ^ swz) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) arrive_leader<CG>(L_hs_full[kk]);
        }
        hacc_f ^= 1u;
      }
    }
  } else {
    // ===================== LN warps: bias + residual + LayerNorm (+ conv branch) =====================
    // Two warps per TMEM lane quarter, 128 columns each; the two partial row sums meet in shared memory (named barrier per
    // quarter).  Streams (residual in pass 1; conv + skip in and the output out in pass 3) move in 16-column chunks — boxes
    // of 32 rows x 32 bytes, SWIZZLE_32B, one row per lane — through the warp's ring of four 1 KB buffers.
    const int lw = warp - LN_WARP0;
    const int q = warp & 3;
    const int half = lw >> 2;
    const int col0 = half * LN_COLS;
    const uint32_t lb0 = smem_u32(s.lbuf[lw][0]);
    uint64_t* lbar = s.l_full[lw];
    uint32_t lphase = 0u;                    // bit i: phase parity of buffer i
    uint32_t yph = 0u;
    const uint32_t my_row = (uint32_t)lane * 32u;
    const uint32_t swz = (uint32_t)((lane >> 2) & 1);
    auto load_res = [&](int m, int c) {     // lane 0 only: residual chunk c -> buffer c % 4
      const int i = c & (LN_BUFS - 1);
      mbar_arrive_expect_tx(&lbar[i], LBUF_BYTES);
      tma_load_2d(s.lbuf[lw][i], &p.tma_res, &lbar[i], col0 + c * LCH, m * BM + q * 32);
    };
    auto load_gn = [&](int m, int c) {      // lane 0 only: conv chunk c -> buffer 2 (c % 2), skip chunk c -> the next one
      const int i = 2 * (c & 1);
      mbar_arrive_expect_tx(&lbar[i], 2 * LBUF_BYTES);
      tma_load_2d(s.lbuf[lw][i], &p.tma_conv, &lbar[i], col0 + c * LCH, m * BM + q * 32);
      tma_load_2d(s.lbuf[lw][i + 1], &p.tma_skip, &lbar[i], col0 + c * LCH, m * BM + q * 32);
    };
    auto wait_buf = [&](int i) {
      mbar_wait(&lbar[i], (lphase >> i) & 1u);
      lphase ^= 1u << i;
    };
    if (lane == 0 && cluster_id < npairs)
      for (int c = 0; c < LN_BUFS; ++c) load_res(cluster_id * CG + (int)rank, c);

    for (int mp = cluster_id; mp < npairs; mp += num_clusters) {
      const int m = mp * CG + (int)rank;
      const int row = q * 32 + lane;
      const int row0 = m * BM + q * 32;
      const long long tl0 = p.prof_ln ? clock64() : 0;
      mbar_wait(&s.y_full, yph);
      yph ^= 1;
      const long long tl1 = p.prof_ln ? clock64() : 0;
      tc_fence_after();
      const uint32_t t_y = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Y_COL + col0);
      const uint32_t t_xp = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(XP_COL + col0 / 2);

      // ---- pass 1: x = bf16(acc + bias + residual) parked in TMEM as packed pairs; shifted sums of the rounded values ----
      // (sum and sum of squares of x - shift, shift = the thread's first element: one pass gives the mean and the centred
      // second moment without the cancellation of E[x^2] - mean^2; the two column halves are merged with Chan's formula.
      // The earlier separate variance pass re-read the parked row: 64 KB of TMEM reads per tile on a port the H warps need.)
      float shift = 0.f;
      f32x2_t sum_p = pk2(0.f), sum2_p = pk2(0.f);      // (even columns, odd columns): combined after the loop
#pragma unroll 1
      for (int c = 0; c < LCHUNKS; ++c) {
        uint32_t r[16];
        TMEM_LD_X16(t_y + c * LCH, r);
        const int bi = c & (LN_BUFS - 1);
        wait_buf(bi);
        const uint32_t rb = lb0 + (uint32_t)bi * LBUF_BYTES + my_row;
        uint4 rv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) rv[h] = lds128(rb + ((((uint32_t)h) ^ swz) << 4));
        TMEM_WAIT_X16(r);
        uint32_t o[8];
        const float4* b2v = reinterpret_cast<const float4*>(p.b2 + col0 + c * LCH);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t w[4] = {rv[h].x, rv[h].y, rv[h].z, rv[h].w};
          const float4 ba = __ldg(b2v + 2 * h), bb = __ldg(b2v + 2 * h + 1);
          const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = h * 8 + 2 * i;
            float x0, x1;
            upk2(add2(add2(pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pk2(bias[2 * i], bias[2 * i + 1])),
                      pk2(bf16_lo(w[i]), bf16_hi(w[i]))), x0, x1);
            const uint32_t pk = pack_bf16x2(x0, x1);
            o[h * 4 + i] = pk;
            if (c == 0 && h == 0 && i == 0) shift = bf16_lo(pk);
            const f32x2_t d = sub2(pk2(bf16_lo(pk), bf16_hi(pk)), pk2(shift));
            sum_p = add2(sum_p, d);
            sum2_p = fma2(d, d, sum2_p);
          }
        }
        TMEM_ST_X8(t_xp + c * (LCH / 2), o);
        __syncwarp();                       // every lane has read this buffer: refill it with the chunk 4 further on
        if (lane == 0 && c + LN_BUFS < LCHUNKS) load_res(m, c + LN_BUFS);
      }
      // this warp's half of y has been read completely: when all eight have, the next tile's GEMM2 may overwrite it
      tc_fence_before();
      __syncwarp();
      const long long tl2 = p.prof_ln ? clock64() : 0;
      if (lane == 0) {
        arrive_leader<CG>(L_y_empty);
        if (GN && !(p.debug & 5)) { load_gn(m, 0); load_gn(m, 1); }     // every buffer is free: two (conv, skip) pairs
      }
      if (p.debug & 1) {
        __syncwarp();
        if (lane == 0 && mp + num_clusters < npairs) for (int c = 0; c < LN_BUFS; ++c) load_res((mp + num_clusters) * CG + (int)rank, c);
        continue;
      }
      // this half: mean_h and M2_h = sum (x - mean_h)^2 over its 128 columns; merge the two halves
      float sum, sum2;
      {
        float a0, a1, b0, b1;
        upk2(sum_p, a0, a1);
        upk2(sum2_p, b0, b1);
        sum = a0 + a1;
        sum2 = b0 + b1;
      }
      {
        const float mh = shift + sum * (1.f / LN_COLS);
        s.xch[0][half][row] = mh;
        s.xch[1][half][row] = sum2 - sum * sum * (1.f / LN_COLS);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float ma = s.xch[0][0][row], mb = s.xch[0][1][row];
      const float mean = 0.5f * (ma + mb);
      const float m2 = s.xch[1][0][row] + s.xch[1][1][row] + (ma - mb) * (ma - mb) * (0.5f * LN_COLS);
      const float rstd = rsqrtf(fmaxf(m2, 0.f) * (1.f / DM) + p.eps);
      tmem_wait_st();
      // (the next tile's pass 1 may not overwrite xch before the partner has read it: both warps of the pair pass this
      // point again — the barrier below — only after their reads)
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");

      const long long tl3 = p.prof_ln ? clock64() : 0;
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
      for (int c = 0; c < LCHUNKS; ++c) {
        uint32_t r[8];
        TMEM_LD_X8(t_xp + c * (LCH / 2), r);
        // GN: conv in buffer 2 (c % 2), skip in the next one, output staged over conv.  Plain: output staged in buffer c % 4.
        const int bi = GN ? 2 * (c & 1) : (c & (LN_BUFS - 1));
        // the other pair held chunk c - 1, whose store was issued at the end of the previous step: as soon as that store has
        // read its buffer, chunk c + 1 is requested there — a whole step ahead of its use
        if (GN && lane == 0 && c >= 1 && c + 1 < LCHUNKS) {
          tma_store_wait_read();
          if (!(p.debug & 4)) load_gn(m, c + 1);
        }
        if (!GN && c >= LN_BUFS) {          // the store of chunk c - 4 has drained this buffer
          if (lane == 0) tma_store_wait_read_n<LN_BUFS - 1>();
          __syncwarp();
        }
        const uint32_t sb = lb0 + (uint32_t)bi * LBUF_BYTES + my_row;
        uint4 cv[2] = {}, sk[2] = {};
        if (GN) {
          if (!(p.debug & 4)) wait_buf(bi);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            cv[h] = lds128(sb + ((((uint32_t)h) ^ swz) << 4));
            sk[h] = lds128(sb + LBUF_BYTES + ((((uint32_t)h) ^ swz) << 4));
          }
        }
        TMEM_WAIT_X8(r);
        const int cc = col0 + c * LCH;
        const float4* gav = reinterpret_cast<const float4*>(p.gamma + cc);
        const float4* bev = reinterpret_cast<const float4*>(p.beta + cc);
        const float4* ggv = reinterpret_cast<const float4*>(p.gn_gamma + (GN ? gl * DM : 0) + cc);
        const float4* gbv = reinterpret_cast<const float4*>(p.gn_beta + (GN ? gl * DM : 0) + cc);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[4];
          float gmean = 0.f, grstd = 0.f;
          if (GN) {
            const float2 st2 = __ldg(reinterpret_cast<const float2*>(gst) + (cc >> 3) + h);
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
          // two columns per instruction (FFMA2 / FMUL2 / FADD2): the LN warps' instruction count is what they cost the
          // H warps and the issuing warp, so it is worth halving
          const f32x2_t mean2 = pk2(mean), rstd2 = pk2(rstd), gmean2 = pk2(gmean), grstd2 = pk2(grstd);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t xw = r[h * 4 + i];
            f32x2_t y = fma2(mul2(sub2(pk2(bf16_lo(xw), bf16_hi(xw)), mean2), rstd2), pk2(ga[2 * i], ga[2 * i + 1]), pk2(be[2 * i], be[2 * i + 1]));
            if (GN) {
              const f32x2_t u = fma2(sub2(pk2(bf16_lo(cw[i]), bf16_hi(cw[i])), gmean2), mul2(grstd2, pk2(gg[2 * i], gg[2 * i + 1])),
                                     pk2(gb[2 * i], gb[2 * i + 1]));
              y = add2(y, add2(gelu_erf2(u), pk2(bf16_lo(sw[i]), bf16_hi(sw[i]))));
            }
            float y0, y1;
            upk2(y, y0, y1);
            o[i] = pack_bf16x2(y0, y1);
          }
          sts128(sb + ((((uint32_t)h) ^ swz) << 4), o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (!(p.debug & 4)) {
            tma_store_2d(&p.tma_y, lb0 + (uint32_t)bi * LBUF_BYTES, cc, row0);
            tma_store_commit();
          }
        }
      }
      __syncwarp();
      if (p.prof_ln && lw == 0 && lane == 0) {
        long long* o = p.prof_ln + (size_t)blockIdx.x * 4;
        o[0] += tl1 - tl0; o[1] += tl2 - tl1; o[2] += tl3 - tl2; o[3] += clock64() - tl3;
      }
      if (lane == 0) {
        // the next tile's first four residual chunks, as soon as the stores have drained
        if (mp + num_clusters < npairs) {
          tma_store_wait_read();
          for (int c = 0; c < LN_BUFS; ++c) load_res((mp + num_clusters) * CG + (int)rank, c);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // neither CTA leaves (or frees TMEM) while the pair's MMAs / arrivals can touch it
  if (warp == 1) {
    __syncwarp();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <bool GN, int CG>
int launch_ffn(FfnParams& p, cudaStream_t st) {
  // the struct plus whatever alignment slack the 227 KB limit leaves (the kernel checks that it fits behind its 1024-byte
  // alignment: the dynamic window of a kernel without static shared memory starts aligned)
  constexpr int smem_bytes = (int)sizeof(FfnSmem<CG>) + 1024 <= 232448 ? (int)sizeof(FfnSmem<CG>) + 1024 : 232448;
  static_assert(sizeof(FfnSmem<CG>) <= 232448, "exceeds the 227 KB shared-memory limit of one CTA");
  auto kern = ffn_fused_tcgen05_kernel<GN, CG>;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int npairs = (p.tiles_m + CG - 1) / CG;
  const int max_clusters = num_sms() / CG;
  const int grid = CG * (npairs < max_clusters ? npairs : max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (getenv("EMRT_FFN_PROF")) {          // diagnosis only: where the MMA thread waits (synchronous, prints to stderr)
    long long* d = nullptr;
    EMRT_CUDA_CHECK(cudaMalloc(&d, (size_t)grid * 8 * sizeof(long long)));
    EMRT_CUDA_CHECK(cudaMemset(d, 0, (size_t)grid * 8 * sizeof(long long)));
    p.prof = d;
    long long* dw = nullptr;
    EMRT_CUDA_CHECK(cudaMalloc(&dw, (size_t)grid * 4 * sizeof(long long)));
    EMRT_CUDA_CHECK(cudaMemset(dw, 0, (size_t)grid * 4 * sizeof(long long)));
    p.prof_w = dw;
    long long* dl = nullptr;
    EMRT_CUDA_CHECK(cudaMalloc(&dl, (size_t)grid * 4 * sizeof(long long)));
    EMRT_CUDA_CHECK(cudaMemset(dl, 0, (size_t)grid * 4 * sizeof(long long)));
    p.prof_ln = dl;
    EMRT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    EMRT_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<long long> h((size_t)grid * 8);
    EMRT_CUDA_CHECK(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<long long> hw((size_t)grid * 4);
    EMRT_CUDA_CHECK(cudaMemcpy(hw.data(), dw, hw.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dw);
    {
      std::vector<long long> hl((size_t)grid * 4);
      EMRT_CUDA_CHECK(cudaMemcpy(hl.data(), dl, hl.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      cudaFree(dl);
      double l4[4] = {0, 0, 0, 0};
      for (int i = 0; i < grid; ++i) for (int k = 0; k < 4; ++k) l4[k] += (double)hl[(size_t)i * 4 + k] / grid;
      fprintf(stderr, "ffn_fused LN warp 0, cycles per CTA: waiting for y_full %.0f, pass 1 %.0f, merge %.0f, pass 3 %.0f\n", l4[0], l4[1], l4[2], l4[3]);
    }
    double w3[3] = {0, 0, 0};
    for (int i = 0; i < grid; i += CG) for (int k = 0; k < 3; ++k) w3[k] += (double)hw[(size_t)i * 4 + k] / (grid / CG);
    fprintf(stderr, "ffn_fused weight waits: first W1 unit of a chunk %.0f, other W1 units %.0f, W2 units %.0f\n", w3[0], w3[1], w3[2]);
    cudaFree(d);
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < grid; i += CG) for (int k = 0; k < 8; ++k) a[k] += (double)h[(size_t)i * 8 + k] / (grid / CG);
    fprintf(stderr, "ffn_fused MMA thread, cycles per CTA (avg of %d): total %.0f | wait x %.0f, weights %.0f, hacc_empty %.0f, hs_full %.0f, y_empty %.0f | tiles/CTA %.1f | producer: total %.0f, waiting for a free ring unit %.0f\n",
            grid, a[0], a[1], a[2], a[3], a[4], a[5], (double)p.tiles_m / grid, a[7], a[6]);
    count_launch();
    return EMRT_OK;
  }
  EMRT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
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
  if (a->d_ff < CH || a->d_ff % CH != 0) return set_error(EMRT_ERR_UNSUPPORTED, "fused FFN needs d_ff %% 128 == 0, got %d", a->d_ff);
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w1) | reinterpret_cast<uintptr_t>(a->w2) |
       reinterpret_cast<uintptr_t>(a->y) | reinterpret_cast<uintptr_t>(a->b1) | reinterpret_cast<uintptr_t>(a->b2) |
       reinterpret_cast<uintptr_t>(a->ln_gamma) | reinterpret_cast<uintptr_t>(a->ln_beta)) & 15)
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "fused FFN needs 16-byte aligned tensors and parameter vectors");
  if (int e = emrt_device_check()) return e;
  cudaStream_t st = as_stream(stream);
  FfnParams p;
  memset(&p, 0, sizeof(p));
  const bool one_cta = getenv("EMRT_FFN_1CTA") != nullptr;      // the cta_group::1 form (one CTA per tile, full B operand per SM)
  const uint32_t cg = one_cta ? 1u : 2u;                          // every weight box holds this CTA's share of the B rows
  p.b1 = a->b1; p.b2 = a->b2; p.gamma = a->ln_gamma; p.beta = a->ln_beta; p.eps = a->ln_eps;
  p.tiles_m = (int)((a->rows + BM - 1) / BM);
  p.num_chunks = a->d_ff / CH;
  { const char* e = getenv("EMRT_FFN_DEBUG"); p.debug = e ? atoi(e) : 0; }
  {
    const uint64_t d[2] = {(uint64_t)DM, (uint64_t)a->rows}, sb[1] = {(uint64_t)DM * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (int e = make_tensor_map(&p.tma_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->x, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    const uint32_t box2[2] = {(uint32_t)LCH, 32u};
    if (int e = make_tensor_map(&p.tma_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->x, d, sb, box2, CU_TENSOR_MAP_SWIZZLE_32B)) return e;
    if (int e = make_tensor_map(&p.tma_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->y, d, sb, box2, CU_TENSOR_MAP_SWIZZLE_32B)) return e;
  }
  {
    const uint64_t d[2] = {(uint64_t)DM, (uint64_t)a->d_ff}, sb[1] = {(uint64_t)DM * 2};
    const uint32_t box[2] = {(uint32_t)BK, cg == 2 ? 32u : (uint32_t)CH};    // CTA pair: a quarter of the chunk per load (two per unit)
    if (int e = make_tensor_map(&p.tma_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->w1, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t d[2] = {(uint64_t)a->d_ff, (uint64_t)DM}, sb[1] = {(uint64_t)a->d_ff * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)DM / cg};
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
    const uint32_t box[2] = {(uint32_t)LCH, 32u};
    if (int e = make_tensor_map(&p.tma_conv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->conv, d, sb, box, CU_TENSOR_MAP_SWIZZLE_32B)) return e;
    if (int e = make_tensor_map(&p.tma_skip, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->skip, d, sb, box, CU_TENSOR_MAP_SWIZZLE_32B)) return e;
    return one_cta ? launch_ffn<true, 1>(p, st) : launch_ffn<true, 2>(p, st);
  }
  p.gn_gamma = a->ln_gamma; p.gn_beta = a->ln_beta;     // never read; keeps the pointer arithmetic defined
  return one_cta ? launch_ffn<false, 1>(p, st) : launch_ffn<false, 2>(p, st);
}
