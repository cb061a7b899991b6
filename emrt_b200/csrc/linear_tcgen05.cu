// bf16 nn.Linear on 5th-generation tensor cores: y = epilogue(x[rows,K] @ Wt[N,K]^T + bias)
// (value_proj / sampling_offsets+attention_weights / output_proj, transformer_encoder_decoder.py:83,89-96,106).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B) of A [128 x 64] and B [BLOCK_N x 64] tiles
//   warp 1      MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16, fp32 accumulators in TMEM
//   warps 2..5  epilogue       tcgen05.ld (one accumulator row per thread) -> bias / mask / ReLU / softmax -> global
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) between TMA and MMA, and a 2-deep TMEM ring
// (tmem_full/tmem_empty) between MMA and epilogue so tile i+1's MMAs overlap tile i's epilogue.
// With K = 256 these GEMMs are HBM-bound (AI = 128 FLOP/B), so the design goal is to keep
// >= 64 KB of TMA loads in flight per SM and to fuse every row-wise consumer into the epilogue.
#include <cuda.h>

#include <cstring>

#include "common.cuh"

namespace emrt {

constexpr int BM = 128;       // rows per tile (UMMA M)
constexpr int BK = 64;        // bf16 elements per k-block = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr int TMEM_COLS = 512;

enum { EPI_KIND_GENERIC = 0, EPI_KIND_QPROJ = 1 };

struct GemmParams {
  CUtensorMap tma_a;   // x  [rows, K] bf16, box {64, 128}
  CUtensorMap tma_b;   // Wt [N, K]    bf16, box {64, BLOCK_N}
  const float* bias;
  const float* row_scale;
  void* y;
  void* y2;
  int64_t rows;
  int32_t K, N;
  int32_t y_dtype;
  int32_t flags;
  int32_t group;       // QPROJ: softmax group (L*P)
  int32_t hm_rows, hm_D;   // HEAD_MAJOR: rows per batch element, head dim
  int32_t tiles_m, tiles_n;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row core-matrix groups),
//   [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TMEM_LD_X16(taddr, r)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))
// tcgen05.wait::ld with the loaded registers as in/out operands, so no use of them can be hoisted above the wait
#define TMEM_WAIT_X16(r)                                                                                          \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                   \
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),   \
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) \
               :: "memory")

// ---- epilogue stores: 8 consecutive columns of one row ---------------------------------------------------------
__device__ __forceinline__ void store8(void* y, int y_dtype, int64_t elem_off, const float (&v)[8]) {
  if (y_dtype == EMRT_F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + elem_off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    if (y_dtype == EMRT_BF16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + elem_off) = r;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(y) + elem_off) = r;
    }
  }
}
__device__ __forceinline__ void store2(void* y, int y_dtype, int64_t elem_off, float a, float b) {
  if (y_dtype == EMRT_F32) {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(y) + elem_off) = make_float2(a, b);
  } else if (y_dtype == EMRT_BF16) {
    *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(y) + elem_off) = __floats2bfloat162_rn(a, b);
  } else {
    *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(y) + elem_off) = __floats2half2_rn(a, b);
  }
}

template <int BLOCK_N, int STAGES>
struct alignas(1024) GemmSmem {
  __nv_bfloat16 a[STAGES][BM * BK];
  __nv_bfloat16 b[STAGES][BLOCK_N * BK];
  float bias[1024];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <int BLOCK_N, int STAGES, int EPI, int GROUP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N for M=128");
  static_assert((BLOCK_N * BK * 2) % 1024 == 0, "B stage must keep 1024-byte alignment");
  extern __shared__ uint8_t smem_raw[];
  using Smem = GemmSmem<BLOCK_N, STAGES>;
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_tiles = p.tiles_m * p.tiles_n;
  constexpr uint32_t STAGE_BYTES = (BM * BK + BLOCK_N * BK) * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { mbar_init(&s.tmem_full[i], 1); mbar_init(&s.tmem_empty[i], 4 * 32); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < p.N && i < 1024; i += NUM_THREADS) s.bias[i] = p.bias ? p.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&s.full[stage], STAGE_BYTES);
          tma_load_2d(s.a[stage], &p.tma_a, &s.full[stage], kb * BK, m_blk * BM);
          tma_load_2d(s.b[stage], &p.tma_b, &s.full[stage], kb * BK, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(s.a[stage]));
          const uint64_t db = make_smem_desc(smem_u32(s.b[stage]));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 bytes (= 2 x 16 B) per K=16 step inside the 128-byte swizzle row
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&s.empty[stage]);                     // frees the smem slot when these MMAs retire
          if (kb == num_kb - 1) umma_commit(&s.tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
      mbar_wait(&s.tmem_full[acc], acc_phase);
      tc_fence_after();
      const int64_t row = (int64_t)m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.rows;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      const int n0 = n_blk * BLOCK_N;

      if (EPI == EPI_KIND_GENERIC) {
        const float rs = (p.flags & EMRT_EPI_ROW_MASK) ? (row_ok ? __ldg(p.row_scale + row) : 0.f) : 1.f;
        const bool relu = p.flags & EMRT_EPI_RELU;
        // HEAD_MAJOR: element (row = b*hm_rows + pix, col = m*hm_D + d) goes to [b][m][pix][d]
        const bool hm = p.flags & EMRT_EPI_HEAD_MAJOR;
        int64_t hm_b = 0, hm_pix = 0;
        if (hm && row_ok) { hm_b = row / p.hm_rows; hm_pix = row - hm_b * p.hm_rows; }
        const int hm_heads = hm ? p.N / p.hm_D : 1;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t r[2][16];
          TMEM_LD_X16(t_row + c, r[0]);
          if (c + 16 < BLOCK_N) TMEM_LD_X16(t_row + c + 16, r[1]);
          TMEM_WAIT_X16(r[0]);
          if (c + 16 < BLOCK_N) TMEM_WAIT_X16(r[1]);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int col = n0 + c + h * 8;
            if (c + h * 8 < BLOCK_N && row_ok && col < p.N) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x = (__uint_as_float(r[h >> 1][(h & 1) * 8 + i]) + s.bias[col + i]) * rs;
                v[i] = relu ? fmaxf(x, 0.f) : x;
              }
              const int64_t dst = hm ? ((hm_b * hm_heads + col / p.hm_D) * p.hm_rows + hm_pix) * p.hm_D + col % p.hm_D
                                     : row * p.N + col;
              store8(p.y, p.y_dtype, dst, v);
            }
          }
        }
      } else {
        // MSDA query projection: N = 3*tp columns = [2*tp pixel offsets | tp attention logits], BLOCK_N == tp.
        constexpr int tp = BLOCK_N;
        if (n_blk < 2) {
#pragma unroll 1
          for (int c = 0; c < BLOCK_N; c += 16) {
            uint32_t r[16];
            TMEM_LD_X16(t_row + c, r);
            TMEM_WAIT_X16(r);
            if (row_ok) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[h * 8 + i]) + s.bias[n0 + c + h * 8 + i];
                store8(p.y, p.y_dtype, row * (2 * tp) + n0 + c + h * 8, v);
              }
            }
          }
        } else {
          // softmax over each group of GROUP consecutive logits (t_e_d.py:92-96).  The whole 144-column row lives in
          // registers (static indexing): 9 aligned x16 TMEM loads, then 8 groups of 18.
          uint32_t r[BLOCK_N / 16][16];
#pragma unroll
          for (int c = 0; c < BLOCK_N / 16; ++c) TMEM_LD_X16(t_row + c * 16, r[c]);
#pragma unroll
          for (int c = 0; c < BLOCK_N / 16; ++c) TMEM_WAIT_X16(r[c]);
#pragma unroll
          for (int g = 0; g < BLOCK_N / GROUP; ++g) {
            float v[GROUP];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
              const int col = g * GROUP + i;
              v[i] = __uint_as_float(r[col / 16][col % 16]) + s.bias[n0 + col];
              mx = fmaxf(mx, v[i]);
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < GROUP; ++i) { v[i] = __expf(v[i] - mx); sum += v[i]; }
            const float inv = 1.f / sum;
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < GROUP; i += 2)
                store2(p.y2, p.y_dtype, row * tp + g * GROUP + i, v[i] * inv, v[i + 1] * inv);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&s.tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D bf16 row-major [outer, inner] tensor, box {BK, box_outer}, 128-byte swizzle, zero fill out of bounds.
static int make_tma_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(EMRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EMRT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return EMRT_OK;
}

template <int BLOCK_N, int STAGES, int EPI, int GROUP>
static int launch_tc(GemmParams& p, const emrt_linear_args* a, cudaStream_t st) {
  using Smem = GemmSmem<BLOCK_N, STAGES>;
  const int smem_bytes = (int)sizeof(Smem) + 1024;
  if (int e = make_tma_2d(&p.tma_a, a->x, (uint64_t)a->K, (uint64_t)a->rows, BM)) return e;
  if (int e = make_tma_2d(&p.tma_b, a->w, (uint64_t)a->K, (uint64_t)a->N, BLOCK_N)) return e;
  p.tiles_m = (int)((a->rows + BM - 1) / BM);
  p.tiles_n = (a->N + BLOCK_N - 1) / BLOCK_N;
  auto kern = linear_tcgen05_kernel<BLOCK_N, STAGES, EPI, GROUP>;
  static bool attr_set = false;
  if (!attr_set) {
    EMRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, smem_bytes, st>>>(p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

int linear_tcgen05(const emrt_linear_args* a, cudaStream_t st) {
  if (a->x_dtype != EMRT_BF16 || a->w_dtype != EMRT_BF16 || !a->w_transposed)
    return set_error(EMRT_ERR_UNSUPPORTED, "tcgen05 linear needs bf16 x and bf16 pre-packed [N,K] weights (emrt_pack_weight)");
  if (a->K % 8 != 0 || a->N % 8 != 0 || a->N > 1024)
    return set_error(EMRT_ERR_UNSUPPORTED, "tcgen05 linear needs K %% 8 == 0, N %% 8 == 0, N <= 1024 (K=%d N=%d)", a->K, a->N);
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w) | reinterpret_cast<uintptr_t>(a->y)) & 15)
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "tcgen05 linear needs 16-byte aligned x, w, y");
  if (a->y_dtype != EMRT_F32 && a->y_dtype != EMRT_BF16 && a->y_dtype != EMRT_F16)
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad y_dtype %d", a->y_dtype);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = a->bias; p.row_scale = a->row_scale; p.y = a->y; p.y2 = a->y2;
  p.rows = a->rows; p.K = a->K; p.N = a->N; p.y_dtype = a->y_dtype; p.flags = a->epilogue; p.group = a->qproj_group;
  p.hm_rows = a->hm_rows; p.hm_D = a->hm_D;
  if (a->epilogue & EMRT_EPI_HEAD_MAJOR) {
    if (a->hm_rows <= 0 || a->hm_D <= 0 || a->hm_D % 8 != 0 || a->N % a->hm_D != 0 || a->rows % a->hm_rows != 0)
      return set_error(EMRT_ERR_INVALID_ARGUMENT, "HEAD_MAJOR needs hm_D %% 8 == 0, N %% hm_D == 0, rows %% hm_rows == 0");
  }
  if (a->epilogue & EMRT_EPI_MSDA_QPROJ) {
    if (a->epilogue != EMRT_EPI_MSDA_QPROJ) return set_error(EMRT_ERR_UNSUPPORTED, "MSDA_QPROJ cannot be combined");
    if (a->N != 3 * 144) return set_error(EMRT_ERR_UNSUPPORTED, "MSDA_QPROJ epilogue is built for M*L*P = 144 (N = 432), got N=%d", a->N);
    if (!a->y2 || a->qproj_group != 18)
      return set_error(EMRT_ERR_UNSUPPORTED, "MSDA_QPROJ needs y2 and softmax group L*P = 18 (EMRT: 3 levels x 6 points), got %d", a->qproj_group);
    return launch_tc<144, 6, EPI_KIND_QPROJ, 18>(p, a, st);
  }
  if (a->epilogue & ~(EMRT_EPI_ROW_MASK | EMRT_EPI_RELU | EMRT_EPI_HEAD_MAJOR))
    return set_error(EMRT_ERR_UNSUPPORTED, "tcgen05 linear: unsupported epilogue flags %d", a->epilogue);
  if (a->N <= 64) return launch_tc<64, 8, EPI_KIND_GENERIC, 2>(p, a, st);
  if (a->N <= 128) return launch_tc<128, 6, EPI_KIND_GENERIC, 2>(p, a, st);
  return launch_tc<256, 4, EPI_KIND_GENERIC, 2>(p, a, st);
}

}  // namespace emrt
