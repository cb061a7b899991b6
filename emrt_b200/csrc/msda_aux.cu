// softmax over levels x points + sampling-location arithmetic on raw projections
// (transformer_encoder_decoder.py:92-102) for the un-fused path; the fused path does the same inside the
// tcgen05 GEMM epilogue (linear_tcgen05.cu, EMRT_EPI_MSDA_QPROJ).
#include "common.cuh"

namespace emrt {

// one thread per (row, head)
template <typename TO, int MODE>
__global__ void __launch_bounds__(256)
msda_softmax_loc_kernel(const float* __restrict__ off_raw, int64_t off_ld, const float* __restrict__ logit_raw,
                        int64_t logit_ld, const float* __restrict__ ref, int64_t ref_bs, TO* __restrict__ loc_out,
                        TO* __restrict__ attn_out, int Lq, int M, int L, int P,
                        const __grid_constant__ LevelTable lv, int64_t n_items) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  const int m = (int)(item % M);
  const int64_t row = item / M;
  const int q = (int)(row % Lq);
  const int64_t b = row / Lq;
  const int LP = L * P;
  const float* lg = logit_raw + row * logit_ld + (int64_t)m * LP;
  // F.softmax(attention_weights, -1) over L*P (t_e_d.py:95): max-subtracted, fp32
  float mx = -INFINITY;
  for (int i = 0; i < LP; ++i) mx = fmaxf(mx, __ldg(lg + i));
  float sum = 0.f;
  for (int i = 0; i < LP; ++i) sum += expf(__ldg(lg + i) - mx);
  const float inv = 1.f / sum;
  TO* ao = attn_out + item * LP;
  for (int i = 0; i < LP; ++i) ao[i] = from_float<TO>(expf(__ldg(lg + i) - mx) * inv);

  const float* of = off_raw + row * off_ld + (int64_t)m * LP * 2;
  TO* lo = loc_out + item * LP * 2;
  for (int l = 0; l < L; ++l) {
    float rx = 0.f, ry = 0.f, iw = 1.f, ih = 1.f;
    if (MODE == EMRT_LOC_NORMALIZED) {
      const float* rp = ref + b * ref_bs + ((int64_t)q * L + l) * 2;
      rx = __ldg(rp); ry = __ldg(rp + 1);
      iw = (float)lv.W[l]; ih = (float)lv.H[l];
    }
    for (int p = 0; p < P; ++p) {
      const int i = (l * P + p) * 2;
      float ox = __ldg(of + i), oy = __ldg(of + i + 1);
      if (MODE == EMRT_LOC_NORMALIZED) {   // ref + off / (W_l, H_l)  (t_e_d.py:98-102)
        ox = rx + ox / iw;
        oy = ry + oy / ih;
      }
      lo[i] = from_float<TO>(ox);
      lo[i + 1] = from_float<TO>(oy);
    }
  }
}

// y = LayerNorm(residual + x) * gamma + beta over the last dim N (one warp per row; N <= 1024, N % 32 == 0).
template <typename T>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const T* __restrict__ x, const T* __restrict__ residual, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ y, int64_t rows, int N, float eps) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[32];
  const int per = N / 32;
  float s = 0.f;
  for (int i = 0; i < per; ++i) {
    const int c = i * 32 + lane;
    float a = to_float(x[row * N + c]);
    if (residual) a += to_float(residual[row * N + c]);
    v[i] = a;
    s += a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)N;
  float ss = 0.f;
  for (int i = 0; i < per; ++i) { const float d = v[i] - mean; ss += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)N + eps);
  for (int i = 0; i < per; ++i) {
    const int c = i * 32 + lane;
    y[row * N + c] = from_float<T>((v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
  }
}

// out[i] = a[i] + b[i % b_period]  (with_pos_embed, transformer_encoder_decoder.py:154-155,198), 16 B per thread
template <typename T>
__global__ void __launch_bounds__(256)
add_bcast_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n_vec,
                 int64_t period_vec) {
  constexpr int VEC = Vec16<T>::N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float x[VEC], y[VEC];
    Vec16<T>::load(a + i * VEC, x);
    Vec16<T>::load(b + (i % period_vec) * VEC, y);
#pragma unroll
    for (int k = 0; k < VEC; ++k) x[k] += y[k];
    Vec16<T>::store(out + i * VEC, x);
  }
}

}  // namespace emrt

using namespace emrt;

extern "C" int emrt_add_bcast(const void* a, const void* b, void* out, int64_t n, int64_t b_period, int dtype,
                              void* stream) {
  EMRT_REQUIRE(a && b && out && n > 0 && b_period > 0 && n % b_period == 0, "bad add_bcast arguments");
  const int vec = dtype == EMRT_F32 ? 4 : 8;
  EMRT_REQUIRE(n % vec == 0 && b_period % vec == 0, "sizes must be multiples of the 16-byte vector width");
  const int64_t nv = n / vec, pv = b_period / vec;
  const int64_t want = (nv + 255) / 256;
  const unsigned blocks = (unsigned)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  cudaStream_t st = as_stream(stream);
  if (dtype == EMRT_F32) add_bcast_kernel<float><<<blocks, 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, nv, pv);
  else if (dtype == EMRT_BF16) add_bcast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, nv, pv);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_msda_softmax_loc(const float* off_raw, int64_t off_ld, const float* logit_raw, int64_t logit_ld,
                                     const float* ref, int64_t ref_batch_stride, void* loc_out, void* attn_out,
                                     int B, int Lq, int M, int L, int P, const int32_t* shapes_hw_host,
                                     int out_dtype, int mode, void* stream) {
  EMRT_REQUIRE(off_raw && logit_raw && loc_out && attn_out, "NULL tensor pointer");
  EMRT_REQUIRE(B > 0 && Lq > 0 && M > 0 && P > 0, "non-positive dimension");
  EMRT_REQUIRE(mode == EMRT_LOC_NORMALIZED || mode == EMRT_LOC_PIXEL_OFFSET, "bad loc mode");
  EMRT_REQUIRE(mode != EMRT_LOC_NORMALIZED || ref != nullptr, "NORMALIZED mode needs reference points");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, -1)) return e;
  const int64_t n_items = (int64_t)B * Lq * M;
  const unsigned blocks = (unsigned)((n_items + 255) / 256);
  cudaStream_t st = as_stream(stream);
#define EMRT_SL(TO, MODE)                                                                                       \
  msda_softmax_loc_kernel<TO, MODE><<<blocks, 256, 0, st>>>(off_raw, off_ld, logit_raw, logit_ld, ref,          \
                                                             ref_batch_stride, (TO*)loc_out, (TO*)attn_out, Lq, \
                                                             M, L, P, lv, n_items)
  const bool px = mode == EMRT_LOC_PIXEL_OFFSET;
  if (out_dtype == EMRT_F32) { if (px) EMRT_SL(float, 1); else EMRT_SL(float, 0); }
  else if (out_dtype == EMRT_F16) { if (px) EMRT_SL(__half, 1); else EMRT_SL(__half, 0); }
  else if (out_dtype == EMRT_BF16) { if (px) EMRT_SL(__nv_bfloat16, 1); else EMRT_SL(__nv_bfloat16, 0); }
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad out_dtype %d", out_dtype);
#undef EMRT_SL
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_add_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                                  void* y, int64_t rows, int N, float eps, int dtype, void* stream) {
  EMRT_REQUIRE(x && gamma && beta && y && rows > 0, "bad add_layernorm arguments");
  EMRT_REQUIRE(N % 32 == 0 && N <= 1024, "N must be a multiple of 32 and <= 1024");
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (dtype == EMRT_F32)
    add_layernorm_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (const float*)residual, gamma, beta, (float*)y, rows, N, eps);
  else if (dtype == EMRT_BF16)
    add_layernorm_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, (__nv_bfloat16*)y, rows, N, eps);
  else
    return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
