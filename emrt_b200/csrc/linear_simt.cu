// fp32-accumulate SIMT GEMM: the parity-mode path of nn.Linear (y = x @ W + b, Paddle layout W [in,out];
// transformer_encoder_decoder.py:36-42,83,89,92,106) and the comparison kernel for the tcgen05 path.
// Not the performance path: bf16 projections run in linear_tcgen05.cu.
#include "common.cuh"

namespace emrt {

constexpr int SBM = 128, SBN = 64, SBK = 16;

template <typename T> __device__ __forceinline__ float ld_elem(const T* p) { return to_float(__ldg(p)); }
template <> __device__ __forceinline__ float ld_elem<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((unsigned int)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

template <typename TX, typename TW, typename TY, bool WT>
__global__ void __launch_bounds__(256)
linear_simt_kernel(const TX* __restrict__ x, const TW* __restrict__ w, const float* __restrict__ bias,
                   TY* __restrict__ y, int64_t rows, int K, int N, int epilogue,
                   const float* __restrict__ row_scale) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * SBM;
  const int col0 = blockIdx.y * SBN;
  const int ty = t / 16, tx = t % 16;   // 16 x 16 threads, each 8 rows x 4 cols
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SBK) {
    {  // A tile: 128 rows x 16 k; thread -> row t/2, k segment (t%2)*8
      const int r = t >> 1, ks = (t & 1) * 8;
      const int64_t gr = row0 + r;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int gk = k0 + ks + i;
        As[ks + i][r] = (gr < rows && gk < K) ? ld_elem<TX>(x + gr * K + gk) : 0.f;
      }
    }
    if (!WT) {  // W [K,N]: thread -> k = t/16, n segment (t%16)*4
      const int kk = t >> 4, ns = (t & 15) * 4;
      const int gk = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gn = col0 + ns + i;
        Bs[kk][ns + i] = (gk < K && gn < N) ? ld_elem<TW>(w + (int64_t)gk * N + gn) : 0.f;
      }
    } else {    // W^T [N,K]: thread -> n = t/4, k segment (t%4)*4
      const int nn = t >> 2, ks = (t & 3) * 4;
      const int gn = col0 + nn;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gk = k0 + ks + i;
        Bs[ks + i][nn] = (gk < K && gn < N) ? ld_elem<TW>(w + (int64_t)gn * K + gk) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gr = row0 + ty * 8 + i;
    if (gr >= rows) continue;
    const float rs = (epilogue & EMRT_EPI_ROW_MASK) ? __ldg(row_scale + gr) : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = col0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
      v *= rs;
      if (epilogue & EMRT_EPI_RELU) v = fmaxf(v, 0.f);
      y[gr * N + gn] = from_float<TY>(v);
    }
  }
}

template <typename TX, typename TW, typename TY>
static int launch_simt(const emrt_linear_args* a, cudaStream_t st) {
  dim3 grid((unsigned)((a->rows + SBM - 1) / SBM), (unsigned)((a->N + SBN - 1) / SBN));
  if (a->w_transposed)
    linear_simt_kernel<TX, TW, TY, true><<<grid, 256, 0, st>>>((const TX*)a->x, (const TW*)a->w, a->bias, (TY*)a->y,
                                                                 a->rows, a->K, a->N, a->epilogue, a->row_scale);
  else
    linear_simt_kernel<TX, TW, TY, false><<<grid, 256, 0, st>>>((const TX*)a->x, (const TW*)a->w, a->bias, (TY*)a->y,
                                                                  a->rows, a->K, a->N, a->epilogue, a->row_scale);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

template <typename TX, typename TW>
static int dispatch_y(const emrt_linear_args* a, cudaStream_t st) {
  switch (a->y_dtype) {
    case EMRT_F32: return launch_simt<TX, TW, float>(a, st);
    case EMRT_BF16: return launch_simt<TX, TW, __nv_bfloat16>(a, st);
    case EMRT_F16: return launch_simt<TX, TW, __half>(a, st);
    default: return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad y_dtype %d", a->y_dtype);
  }
}

int linear_simt(const emrt_linear_args* a, cudaStream_t st) {
  if (a->epilogue & ~(EMRT_EPI_ROW_MASK | EMRT_EPI_RELU))
    return set_error(EMRT_ERR_UNSUPPORTED, "SIMT linear supports only ROW_MASK|RELU epilogues (got %d)", a->epilogue);
  if (a->x_dtype == EMRT_F32 && a->w_dtype == EMRT_F32) return dispatch_y<float, float>(a, st);
  if (a->x_dtype == EMRT_BF16 && a->w_dtype == EMRT_BF16) return dispatch_y<__nv_bfloat16, __nv_bfloat16>(a, st);
  if (a->x_dtype == EMRT_BF16 && a->w_dtype == EMRT_F32) return dispatch_y<__nv_bfloat16, float>(a, st);
  if (a->x_dtype == EMRT_F32 && a->w_dtype == EMRT_BF16) return dispatch_y<float, __nv_bfloat16>(a, st);
  return set_error(EMRT_ERR_UNSUPPORTED, "SIMT linear: unsupported x/w dtypes %d/%d", a->x_dtype, a->w_dtype);
}

// ---- weight packing: dst[dst_row0 + n, k] = bf16(src[k, n]) ------------------------------------------------
template <typename TS>
__global__ void pack_weight_kernel(const TS* __restrict__ src, __nv_bfloat16* __restrict__ dst, int K, int N,
                                   int dst_row0) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? ld_elem<TS>(src + (int64_t)k * N + n) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) dst[(int64_t)(dst_row0 + n) * K + k] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

}  // namespace emrt

using namespace emrt;

namespace emrt { int linear_tcgen05(const emrt_linear_args* a, cudaStream_t st); }

extern "C" int emrt_linear_fwd(const emrt_linear_args* a, void* stream) {
  EMRT_REQUIRE(a != nullptr, "args is NULL");
  EMRT_REQUIRE(a->x && a->w && a->y, "NULL tensor pointer");
  EMRT_REQUIRE(a->rows > 0 && a->K > 0 && a->N > 0, "non-positive dimension");
  EMRT_REQUIRE(!(a->epilogue & EMRT_EPI_ROW_MASK) || a->row_scale, "ROW_MASK needs row_scale");
  cudaStream_t st = as_stream(stream);
  int impl = a->impl;
  if (impl == 0) impl = (a->x_dtype == EMRT_BF16 && a->w_dtype == EMRT_BF16 && a->w_transposed) ? 2 : 1;
  if (impl == 2) return linear_tcgen05(a, st);
  if (a->x_nchw_hw > 0) return set_error(EMRT_ERR_UNSUPPORTED, "channel-major x (x_nchw_hw) is a tcgen05-path operand");
  return linear_simt(a, st);
}

extern "C" int emrt_pack_weight(const void* src, int src_dtype, void* dst_bf16, int K, int N, int dst_row0,
                                void* stream) {
  EMRT_REQUIRE(src && dst_bf16 && K > 0 && N > 0 && dst_row0 >= 0, "bad pack_weight arguments");
  dim3 grid((N + 31) / 32, (K + 31) / 32), block(32, 8);
  cudaStream_t st = as_stream(stream);
  if (src_dtype == EMRT_F32)
    pack_weight_kernel<float><<<grid, block, 0, st>>>((const float*)src, (__nv_bfloat16*)dst_bf16, K, N, dst_row0);
  else if (src_dtype == EMRT_BF16)
    pack_weight_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst_bf16, K, N, dst_row0);
  else
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "pack_weight: src_dtype must be F32 or BF16");
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
