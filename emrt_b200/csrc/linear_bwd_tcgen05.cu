// Weight gradient of nn.Linear on tcgen05: dW[K,N] += x[rows,K]^T dy[rows,N] (bf16 operands, fp32 accumulate).
//
// GEMM view: D[M = k][N = n] = sum_r A[m][r] B[r][n] with the REDUCTION running over the rows.  Both operands are
// row-major [rows, *] in HBM, i.e. MN-major for this product, so the smem descriptors use the MN-major SWIZZLE_128B
// canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: a TMA box {64 elements, 64 rows} lands as 64 rows of
// 128 swizzled bytes = eight 1024-byte atoms (SBO = 1024 B between 8-row groups); a 128-wide operand is two such boxes
// (LBO = 8192 B).  The instruction descriptor sets the a_major / b_major (transpose) bits.
// Split-K: every output tile (128 k x 128 n) is shared by gridDim.x / tiles CTAs, each reducing a contiguous range of
// 64-row chunks into one TMEM accumulator and adding its partial tile to dW with red.global.add.v4.f32.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace emrt {

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

constexpr int DW_T = 128;        // tile: 128 (k) x 128 (n)
constexpr int DW_R = 64;         // rows per pipeline stage
constexpr int DW_STAGES = 6;
constexpr int DW_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct DwParams {
  CUtensorMap tma_x;    // x  [rows, K] bf16, box {64, 64}
  CUtensorMap tma_dy;   // dy [rows, N] bf16, box {64, 64}
  float* dw;
  int32_t K, N;
  int32_t tiles_k, tiles_n, splits, chunks;
};

struct DwSmem {
  __nv_bfloat16 a[DW_STAGES][2][DW_R * 64];   // [stage][64-wide k block][64 rows x 64 elements]
  __nv_bfloat16 b[DW_STAGES][2][DW_R * 64];
  uint64_t full[DW_STAGES];
  uint64_t empty[DW_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// kind::f16 instruction descriptor, D = f32, A = B = bf16, both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(DW_THREADS, 1)
linear_bwd_weight_tc_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DwSmem& s = *reinterpret_cast<DwSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int tk = tile / p.tiles_n, tn = tile % p.tiles_n;
  const int per = (p.chunks + p.splits - 1) / p.splits;
  const int c_begin = split * per, c_end = min(c_begin + per, p.chunks);
  constexpr uint32_t STAGE_BYTES = 4 * DW_R * 64 * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_x);
    tma_prefetch_desc(&p.tma_dy);
#pragma unroll
    for (int i = 0; i < DW_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&s.empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&s.full[stage], STAGE_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tma_load_2d(s.a[stage][h], &p.tma_x, &s.full[stage], tk * DW_T + h * 64, c * DW_R);
          tma_load_2d(s.b[stage][h], &p.tma_dy, &s.full[stage], tn * DW_T + h * 64, c * DW_R);
        }
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_mn(DW_T, DW_T);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&s.full[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_mn(smem_u32(s.a[stage][0]));
        const uint64_t db = make_smem_desc_mn(smem_u32(s.b[stage][0]));
#pragma unroll
        for (int k = 0; k < DW_R / 16; ++k)   // 16 reduction rows = two 1024-byte atoms = +2048 bytes
          umma_bf16(tmem_base, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (c > c_begin || k > 0) ? 1u : 0u);
        umma_commit(&s.empty[stage]);
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&s.acc_full);
    }
  } else if (c_end > c_begin) {
    const int q = warp & 3;
    mbar_wait(&s.acc_full, 0);
    tc_fence_after();
    const int k = tk * DW_T + q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < DW_T; c += 32) {
      uint32_t r[32];
      TMEM_LD_X32(t_row + c, r);
      TMEM_WAIT_X32(r);
      if (k < p.K) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = tn * DW_T + c + j;
          if (n < p.N)     // N % 4 == 0
            red_add_v4(p.dw + (int64_t)k * p.N + n, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

// column sums of dy (the bias gradient): grid-stride over row blocks, fp32 atomics per CTA
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ dy, float* __restrict__ db, int64_t rows, int N, int rows_per_cta) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc += to_float(dy[r * N + n]);
    atomicAdd(db + n, acc);
  }
}

// bf16, N a multiple of 8 and <= 2048: each thread owns 8 adjacent columns (one 16-byte load per row) and every
// (256 / (N/8))-th row of the CTA's slab; the row lanes are combined in shared memory, one fp32 atomic per column and CTA.
__global__ void __launch_bounds__(256)
colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db, int64_t rows, int N, int rows_per_cta) {
  __shared__ float s_part[2048];
  const int G = N / 8, RL = 256 / G;
  const int cg = threadIdx.x % G, rl = threadIdx.x / G;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  if (rl < RL) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* src = dy + cg * 8;
    int64_t r = r0 + rl;
    for (; r + 3 * RL < r1; r += 4 * RL) {              // four independent loads in flight
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(src + (r + u * RL) * N));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += __uint_as_float(w[i] << 16);
          acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
      }
    }
    for (; r < r1; r += RL) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + r * N));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[rl * N + cg * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += 256) {
    float t = 0.f;
    for (int k = 0; k < RL; ++k) t += s_part[k * N + n];
    atomicAdd(db + n, t);
  }
}

int colsum(const void* dy, float* db, int64_t rows, int N, int dtype, cudaStream_t st) {
  if (dtype == EMRT_BF16 && N % 8 == 0 && N >= 8 && N <= 2048 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
      !getenv("EMRT_COLSUM_SLOW")) {
    const int rows_per_cta = 128;
    const unsigned blocks = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
    colsum_bf16x8_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, db, rows, N, rows_per_cta);
    EMRT_LAUNCH_CHECK();
    return EMRT_OK;
  }
  const int rows_per_cta = 256;
  const unsigned blocks = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
  if (dtype == EMRT_BF16) colsum_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, db, rows, N, rows_per_cta);
  else if (dtype == EMRT_F32) colsum_kernel<float><<<blocks, 256, 0, st>>>((const float*)dy, db, rows, N, rows_per_cta);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "colsum: bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

// Returns EMRT_ERR_UNSUPPORTED (error text untouched) when the tensor-core path does not apply.
int linear_bwd_weight_tc(const void* x, const void* dy, float* dw, int64_t rows, int K, int N, cudaStream_t st) {
  if (K % 8 != 0 || N % 8 != 0 || rows >= (1LL << 31)) return EMRT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw)) & 15) return EMRT_ERR_UNSUPPORTED;
  DwParams p;
  memset(&p, 0, sizeof(p));
  const uint32_t box[2] = {64u, (uint32_t)DW_R};
  const uint64_t dx[2] = {(uint64_t)K, (uint64_t)rows}, sx[1] = {(uint64_t)K * 2};
  if (int e = make_tensor_map(&p.tma_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dx, sx, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  const uint64_t dd[2] = {(uint64_t)N, (uint64_t)rows}, sd[1] = {(uint64_t)N * 2};
  if (int e = make_tensor_map(&p.tma_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, dd, sd, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  p.dw = dw; p.K = K; p.N = N;
  p.tiles_k = (K + DW_T - 1) / DW_T;
  p.tiles_n = (N + DW_T - 1) / DW_T;
  p.chunks = (int)((rows + DW_R - 1) / DW_R);
  const int tiles = p.tiles_k * p.tiles_n;
  p.splits = num_sms() / tiles;
  if (p.splits < 1) p.splits = 1;
  if (p.splits > p.chunks) p.splits = p.chunks;
  const int smem_bytes = (int)sizeof(DwSmem) + 1024;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(linear_bwd_weight_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));   // per context
  linear_bwd_weight_tc_kernel<<<tiles * p.splits, DW_THREADS, smem_bytes, st>>>(p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}


// ---- 3x3 conv weight gradient on the same skeleton (conv{l}.0.weight, transformer_encoder_decoder.py:125-144) ----------------
// ws[l][tap][ci][co] += sum over 64-pixel chunks of  x_shifted[chunk, ci]^T dy[chunk, co]:  the A operand of a chunk is the
// conv's own shifted 4-D TMA box {64 ch, W_l, 64 / W_l rows, 1 image} at (kx - 1, y0 + ky - 1) — out-of-map pixels arrive as
// zeros (the padding) — and lands in shared memory exactly like the 2-D {64, 64} box of the Linear kernel above (64 rows of
// 128 swizzled bytes, pixel-major), so descriptors, MMA loop and epilogue are unchanged.  One CTA = one (level, tap,
// 128 x 128 tile of [Cin, Cout], split of the level's chunks); splits per level are proportional to its pixel count.
constexpr int CW_MAX_L = 4;
struct ConvDwParams {
  CUtensorMap tma_x[CW_MAX_L];   // level l: {C, W, H, B} bf16, box {64, W, 64 / W, 1}, SWIZZLE_128B
  CUtensorMap tma_dy;            // [B * Lv, C] bf16, box {64, 64}
  float* ws;
  int32_t C, L, Lv;
  int32_t cta_start[CW_MAX_L + 1];
  int32_t splits[CW_MAX_L], chunks_per_img[CW_MAX_L], rows_per_chunk[CW_MAX_L], start[CW_MAX_L], chunks[CW_MAX_L];
};

__device__ __forceinline__ void tma_load_4d_dw(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(DW_THREADS, 1)
conv3x3_bwd_weight_tc_kernel(const __grid_constant__ ConvDwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DwSmem& s = *reinterpret_cast<DwSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int l = 0;
  while (l + 1 < p.L && (int)blockIdx.x >= p.cta_start[l + 1]) ++l;
  const int tiles_c = p.C / DW_T;                                   // tiles along Cin and along Cout
  int j = blockIdx.x - p.cta_start[l];
  const int split = j % p.splits[l]; j /= p.splits[l];
  const int tn = j % tiles_c; j /= tiles_c;
  const int tk = j % tiles_c; j /= tiles_c;
  const int tap = j;
  const int ky = tap / 3, kx = tap - ky * 3;
  const int per = (p.chunks[l] + p.splits[l] - 1) / p.splits[l];
  const int c_begin = split * per, c_end = min(c_begin + per, p.chunks[l]);
  constexpr uint32_t STAGE_BYTES = 4 * DW_R * 64 * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_x[l]);
    tma_prefetch_desc(&p.tma_dy);
#pragma unroll
    for (int i = 0; i < DW_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int b = c / p.chunks_per_img[l], ci = c - b * p.chunks_per_img[l];
        const int y0 = ci * p.rows_per_chunk[l];
        const int row0 = b * p.Lv + p.start[l] + ci * DW_R;
        mbar_wait(&s.empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&s.full[stage], STAGE_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tma_load_4d_dw(s.a[stage][h], &p.tma_x[l], &s.full[stage], tk * DW_T + h * 64, kx - 1, y0 + ky - 1, b);
          tma_load_2d(s.b[stage][h], &p.tma_dy, &s.full[stage], tn * DW_T + h * 64, row0);
        }
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_mn(DW_T, DW_T);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&s.full[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_mn(smem_u32(s.a[stage][0]));
        const uint64_t db = make_smem_desc_mn(smem_u32(s.b[stage][0]));
#pragma unroll
        for (int k = 0; k < DW_R / 16; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (c > c_begin || k > 0) ? 1u : 0u);
        umma_commit(&s.empty[stage]);
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&s.acc_full);
    }
  } else if (c_end > c_begin) {
    const int q = warp & 3;
    mbar_wait(&s.acc_full, 0);
    tc_fence_after();
    const int k = tk * DW_T + q * 32 + lane;                       // cin
    float* dst = p.ws + (((int64_t)l * 9 + tap) * p.C + k) * p.C + tn * DW_T;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < DW_T; c += 32) {
      uint32_t r[32];
      TMEM_LD_X32(t_row + c, r);
      TMEM_WAIT_X32(r);
#pragma unroll
      for (int jj = 0; jj < 32; jj += 4)
        red_add_v4(dst + c + jj, __uint_as_float(r[jj]), __uint_as_float(r[jj + 1]), __uint_as_float(r[jj + 2]), __uint_as_float(r[jj + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

// ws [L, 9, C, C] fp32 is zeroed here and receives the gradient.  EMRT_ERR_UNSUPPORTED (error text untouched) for shapes
// this kernel does not tile (C % 128, level width not dividing 64, level size not a multiple of 64).
int conv3x3_bwd_weight_tc(const void* x, const void* dy, float* ws, int B, int Lv, int C, int L, const LevelTable& lv,
                          cudaStream_t st) {
  if (C % DW_T != 0 || L > CW_MAX_L || (int64_t)B * Lv >= (1LL << 31)) return EMRT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(ws)) & 15) return EMRT_ERR_UNSUPPORTED;
  for (int l = 0; l < L; ++l)
    if (lv.W[l] > DW_R || DW_R % lv.W[l] != 0 || (lv.H[l] * lv.W[l]) % DW_R != 0) return EMRT_ERR_UNSUPPORTED;
  ConvDwParams p;
  memset(&p, 0, sizeof(p));
  p.ws = ws; p.C = C; p.L = L; p.Lv = Lv;
  int64_t total_chunks = 0;
  for (int l = 0; l < L; ++l) total_chunks += (int64_t)B * lv.H[l] * lv.W[l] / DW_R;
  const int tiles = (C / DW_T) * (C / DW_T) * 9;
  // ~3 waves of CTAs over the three levels, each level's splits proportional to its share of the chunks
  const double per_cta = (double)total_chunks * tiles / (3.0 * num_sms());
  int ctas = 0;
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    p.chunks_per_img[l] = H * W / DW_R;
    p.rows_per_chunk[l] = DW_R / W;
    p.start[l] = lv.start[l];
    p.chunks[l] = B * p.chunks_per_img[l];
    int sp = (int)(p.chunks[l] / (per_cta > 1.0 ? per_cta : 1.0) + 0.5);
    if (sp < 1) sp = 1;
    if (sp > p.chunks[l]) sp = p.chunks[l];
    p.splits[l] = sp;
    p.cta_start[l] = ctas;
    ctas += tiles * sp;
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x + (int64_t)lv.start[l] * C;
    const uint64_t dx[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t sx[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)Lv * C * 2};
    const uint32_t bx[4] = {64u, (uint32_t)W, (uint32_t)(DW_R / W), 1u};
    if (int e = make_tensor_map(&p.tma_x[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, xb, dx, sx, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  p.cta_start[L] = ctas;
  const uint32_t box[2] = {64u, (uint32_t)DW_R};
  const uint64_t dd[2] = {(uint64_t)C, (uint64_t)B * Lv}, sd[1] = {(uint64_t)C * 2};
  if (int e = make_tensor_map(&p.tma_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, dd, sd, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  EMRT_CUDA_CHECK(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)L * 9 * C * C, st));
  const int smem_bytes = (int)sizeof(DwSmem) + 1024;
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_bwd_weight_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  conv3x3_bwd_weight_tc_kernel<<<ctas, DW_THREADS, smem_bytes, st>>>(p);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

}  // namespace emrt
