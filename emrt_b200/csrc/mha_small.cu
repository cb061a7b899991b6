// Decoder self-attention core (MultiHeadAttention over the 110 PSP query tokens, layers.py:282-301):
//   out[b, q, m*D + d] = sum_k softmax_k(scale * <Q[b,q,m,:], K[b,k,m,:]>) V[b,k,m,d]
// The sequences are tiny (Lq = Lk = 110, D = 32), so one CTA owns a (batch, head): K and V sit in shared memory as fp32,
// a warp owns a query at a time — lanes over keys for the scores and the softmax reductions, lanes over channels for
// the weighted sum.  Inputs are the projected q / k / v in the token layout [B, L, M*D] (row strides given, so the fused
// [q | k] projection output can be read in place).
#include "common.cuh"

namespace emrt {

constexpr int MHA_MAX_LK = 256;

template <typename T, int D>
__global__ void __launch_bounds__(256)
mha_small_kernel(const T* __restrict__ q, int64_t q_ld, const T* __restrict__ k, int64_t k_ld, const T* __restrict__ v,
                 int64_t v_ld, T* __restrict__ out, int Lq, int Lk, int M, float scale) {
  extern __shared__ float sm[];
  float* ks = sm;                       // [Lk][D + 1]
  float* vs = sm + Lk * (D + 1);        // [Lk][D]
  float* ps = vs + Lk * D;              // [8 warps][MHA_MAX_LK]
  const int m = blockIdx.x % M;
  const int64_t b = blockIdx.x / M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < Lk * D; i += blockDim.x) {
    const int kk = i / D, d = i - kk * D;
    ks[kk * (D + 1) + d] = to_float(k[(b * Lk + kk) * k_ld + m * D + d]);
    vs[kk * D + d] = to_float(v[(b * Lk + kk) * v_ld + m * D + d]);
  }
  __syncthreads();
  float* pw = ps + warp * MHA_MAX_LK;
  for (int qi = warp; qi < Lq; qi += 8) {
    float qv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) qv[d] = to_float(q[(b * Lq + qi) * q_ld + m * D + d]) * scale;   // (q k^T) * D^-0.5
    float mx = -INFINITY;
    for (int kk = lane; kk < Lk; kk += 32) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(qv[d], ks[kk * (D + 1) + d], s);
      pw[kk] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int kk = lane; kk < Lk; kk += 32) {
      const float e = expf(pw[kk] - mx);
      pw[kk] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.f / sum;
    // lanes over channels (D <= 32)
    if (lane < D) {
      float acc = 0.f;
      for (int kk = 0; kk < Lk; ++kk) acc = fmaf(pw[kk], vs[kk * D + lane], acc);
      out[(b * Lq + qi) * (int64_t)(M * D) + m * D + lane] = from_float<T>(acc * inv);
    }
    __syncwarp();
  }
}

}  // namespace emrt

using namespace emrt;

extern "C" int emrt_mha_small(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld,
                              void* out, int B, int Lq, int Lk, int M, int D, float scale, int dtype, void* stream) {
  EMRT_REQUIRE(q && k && v && out && B > 0 && Lq > 0 && Lk > 0 && M > 0, "bad mha_small arguments");
  if (D != 32) return set_error(EMRT_ERR_UNSUPPORTED, "mha_small is built for head dim 32 (got %d)", D);
  if (Lk > MHA_MAX_LK) return set_error(EMRT_ERR_UNSUPPORTED, "mha_small holds K/V of one head in shared memory: Lk <= %d (got %d)", MHA_MAX_LK, Lk);
  const size_t smem = sizeof(float) * ((size_t)Lk * (32 + 1) + (size_t)Lk * 32 + 8 * MHA_MAX_LK);
  cudaStream_t st = as_stream(stream);
  static bool attr = false;
  if (!attr) {
    EMRT_CUDA_CHECK(cudaFuncSetAttribute(mha_small_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    EMRT_CUDA_CHECK(cudaFuncSetAttribute(mha_small_kernel<__nv_bfloat16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  const unsigned grid = (unsigned)(B * M);
  if (dtype == EMRT_F32)
    mha_small_kernel<float, 32><<<grid, 256, smem, st>>>((const float*)q, q_ld, (const float*)k, k_ld, (const float*)v, v_ld, (float*)out, Lq, Lk, M, scale);
  else if (dtype == EMRT_BF16)
    mha_small_kernel<__nv_bfloat16, 32><<<grid, 256, smem, st>>>((const __nv_bfloat16*)q, q_ld, (const __nv_bfloat16*)k, k_ld, (const __nv_bfloat16*)v, v_ld, (__nv_bfloat16*)out, Lq, Lk, M, scale);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
