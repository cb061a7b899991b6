// Decoder self-attention core (MultiHeadAttention over the 110 PSP query tokens, layers.py:282-301):
//   out[b, q, m*D + d] = sum_k softmax_k(scale * <Q[b,q,m,:], K[b,k,m,:]>) V[b,k,m,d]
// The sequences are tiny (Lq = Lk = 110, D = 32), so one CTA owns a (batch, head): K and V sit in shared memory as fp32
// and every LANE owns a query — its 32 q values and 32 accumulators live in registers, the K / V rows are read as
// broadcast LDS.128 (all lanes the same address: one wavefront per 16 bytes for 32 queries).  ONE pass over the keys in
// blocks of eight with a running maximum (the accumulators are rescaled by exp(old max - new max) once per block, branch
// free), and the dot products / weighted V sums on fp32 PAIRS (FFMA2: the K / V rows arrive as aligned register pairs from
// the LDS.128): a third of the instructions of the two-pass scalar form (65 us per call, issue-bound at 16 warps per SM).
// The first version (a warp per query, lanes over keys, scalar shared-memory reads) was bound by the shared-memory pipe:
// 350 wavefronts per query, 117 us per call.  Inputs are the projected q / k / v in the token layout [B, L, M*D] (row strides
// given, so the fused [q | k] projection output can be read in place).
#include "common.cuh"

namespace emrt {

constexpr int MHA_MAX_LK = 256;
constexpr int MHA_WARPS = 4;

constexpr int MHA_KB = 8;        // keys per block of the running-maximum pass

// four independent partial sums (two packed chains): a single fmaf chain of 32 would leave a warp one instruction per 4 clocks
template <int D>
__device__ __forceinline__ float mha_dot(const f32x2_t (&qv)[D / 2], const float* __restrict__ row) {
  f32x2_t s01 = pk2(0.f, 0.f), s23 = pk2(0.f, 0.f);
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    const float4 kv = *reinterpret_cast<const float4*>(row + d);
    s01 = fma2(qv[d / 2], pk2(kv.x, kv.y), s01);
    s23 = fma2(qv[d / 2 + 1], pk2(kv.z, kv.w), s23);
  }
  float s0, s1, s2, s3;
  upk2(s01, s0, s1);
  upk2(s23, s2, s3);
  return (s0 + s1) + (s2 + s3);
}

template <typename T, int D>
__global__ void __launch_bounds__(MHA_WARPS * 32)
mha_small_kernel(const T* __restrict__ q, int64_t q_ld, const T* __restrict__ k, int64_t k_ld, const T* __restrict__ v,
                 int64_t v_ld, T* __restrict__ out, int Lq, int Lk, int M, float scale, bool vec_io) {
  extern __shared__ __align__(16) float sm[];
  const int Lkp = (Lk + MHA_KB - 1) / MHA_KB * MHA_KB;      // rows up to a whole block: K anything finite, V zero
  float* ks = sm;                 // [Lkp][D]
  float* vs = sm + Lkp * D;       // [Lkp][D]
  const int m = blockIdx.x % M;
  const int64_t b = blockIdx.x / M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (vec_io) {
    constexpr int VEC = Vec16<T>::N;
    for (int i = threadIdx.x * VEC; i < Lk * D; i += blockDim.x * VEC) {
      const int kk = i / D, d = i - kk * D;
      float tk[VEC], tv[VEC];
      Vec16<T>::load(k + (b * Lk + kk) * k_ld + m * D + d, tk);
      Vec16<T>::load(v + (b * Lk + kk) * v_ld + m * D + d, tv);
#pragma unroll
      for (int j = 0; j < VEC; ++j) { ks[i + j] = tk[j]; vs[i + j] = tv[j]; }
    }
  } else {
    for (int i = threadIdx.x; i < Lk * D; i += blockDim.x) {
      const int kk = i / D, d = i - kk * D;
      ks[i] = to_float(k[(b * Lk + kk) * k_ld + m * D + d]);
      vs[i] = to_float(v[(b * Lk + kk) * v_ld + m * D + d]);
    }
  }
  for (int i = Lk * D + threadIdx.x; i < Lkp * D; i += blockDim.x) { ks[i] = 0.f; vs[i] = 0.f; }
  __syncthreads();
  for (int q0 = (blockIdx.y * MHA_WARPS + warp) * 32; q0 < Lq; q0 += 32 * MHA_WARPS * gridDim.y) {
    const int qi = q0 + lane;
    const bool valid = qi < Lq;
    const T* qrow = q + (b * Lq + (valid ? qi : Lq - 1)) * q_ld + m * D;
    float qf[D];
    // a lane's query row is 32 contiguous values: 16-byte loads (every lane another row — element-wise loads cost 32
    // wavefronts each, 1024 per warp for the row; the alignment of the row is checked on the host)
    if (vec_io) {
      constexpr int VEC = Vec16<T>::N;
#pragma unroll
      for (int d = 0; d < D; d += VEC) {
        float t[VEC];
        Vec16<T>::load(qrow + d, t);
#pragma unroll
        for (int k = 0; k < VEC; ++k) qf[d + k] = t[k] * scale;               // (q k^T) * D^-0.5
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) qf[d] = to_float(qrow[d]) * scale;
    }
    f32x2_t qv[D / 2];
#pragma unroll
    for (int d = 0; d < D; d += 2) qv[d / 2] = pk2(qf[d], qf[d + 1]);
    float mx = -INFINITY, sum = 0.f;
    f32x2_t acc2[D / 2];
#pragma unroll
    for (int d = 0; d < D / 2; ++d) acc2[d] = pk2(0.f, 0.f);
#pragma unroll 1
    for (int kb = 0; kb < Lk; kb += MHA_KB) {
      float sc[MHA_KB];
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < MHA_KB; ++j) {
        sc[j] = kb + j < Lk ? mha_dot<D>(qv, ks + (kb + j) * D) : -INFINITY;       // (uniform: every lane has the same keys)
        bm = fmaxf(bm, sc[j]);
      }
      const float nm = fmaxf(mx, bm);
      const float corr = expf(mx - nm);              // first block: exp(-inf) = 0 on zero accumulators
      mx = nm;
      sum *= corr;
      const f32x2_t corr2 = pk2(corr, corr);
#pragma unroll
      for (int d = 0; d < D / 2; ++d) acc2[d] = mul2(acc2[d], corr2);
#pragma unroll
      for (int j = 0; j < MHA_KB; ++j) {
        const float e = expf(sc[j] - nm);            // padded keys: exp(-inf) = 0 on zero V rows
        sum += e;
        const f32x2_t e2 = pk2(e, e);
        const float* vr = vs + (kb + j) * D;
#pragma unroll
        for (int d = 0; d < D; d += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(vr + d);
          acc2[d / 2] = fma2(e2, pk2(vv.x, vv.y), acc2[d / 2]);
          acc2[d / 2 + 1] = fma2(e2, pk2(vv.z, vv.w), acc2[d / 2 + 1]);
        }
      }
    }
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; d += 2) upk2(acc2[d / 2], acc[d], acc[d + 1]);
    if (valid) {
      const float inv = 1.f / sum;
      T* orow = out + (b * Lq + qi) * (int64_t)(M * D) + m * D;
      if (vec_io) {
        constexpr int VEC = Vec16<T>::N;
#pragma unroll
        for (int d = 0; d < D; d += VEC) {
          float t[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) t[k] = acc[d + k] * inv;
          Vec16<T>::store(orow + d, t);
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) orow[d] = from_float<T>(acc[d] * inv);
      }
    }
  }
}

}  // namespace emrt

using namespace emrt;

extern "C" int emrt_mha_small(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld,
                              void* out, int B, int Lq, int Lk, int M, int D, float scale, int dtype, void* stream) {
  EMRT_REQUIRE(q && k && v && out && B > 0 && Lq > 0 && Lk > 0 && M > 0, "bad mha_small arguments");
  if (D != 32) return set_error(EMRT_ERR_UNSUPPORTED, "mha_small is built for head dim 32 (got %d)", D);
  if (Lk > MHA_MAX_LK) return set_error(EMRT_ERR_UNSUPPORTED, "mha_small holds K/V of one head in shared memory: Lk <= %d (got %d)", MHA_MAX_LK, Lk);
  const size_t smem = sizeof(float) * ((size_t)((Lk + MHA_KB - 1) / MHA_KB * MHA_KB) * 32 * 2);
  cudaStream_t st = as_stream(stream);
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(mha_small_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(mha_small_kernel<__nv_bfloat16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));

  // one CTA (4 warps x 32 queries) per (batch, head) and 128 queries
  const int splits = (Lq + 32 * MHA_WARPS - 1) / (32 * MHA_WARPS);
  const dim3 grid((unsigned)(B * M), (unsigned)splits);
  // 16-byte accesses need 16-byte aligned rows: base pointers and row strides (elements) multiples of 16 bytes
  const int esz = dtype == EMRT_F32 ? 4 : 2;
  const bool vec_io = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                        reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                      ((q_ld | k_ld | v_ld) * esz) % 16 == 0 && ((int64_t)M * D * esz) % 16 == 0;
  if (dtype == EMRT_F32)
    mha_small_kernel<float, 32><<<grid, MHA_WARPS * 32, smem, st>>>((const float*)q, q_ld, (const float*)k, k_ld, (const float*)v, v_ld, (float*)out, Lq, Lk, M, scale, vec_io);
  else if (dtype == EMRT_BF16)
    mha_small_kernel<__nv_bfloat16, 32><<<grid, MHA_WARPS * 32, smem, st>>>((const __nv_bfloat16*)q, q_ld, (const __nv_bfloat16*)k, k_ld, (const __nv_bfloat16*)v, v_ld, (__nv_bfloat16*)out, Lq, Lk, M, scale, vec_io);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
