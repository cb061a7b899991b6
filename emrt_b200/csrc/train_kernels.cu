// Backward-pass kernels of the MSDA module (what Paddle autograd derives for transformer_encoder_decoder.py:83-106):
//   * weight / bias gradients of nn.Linear:   dW[K,N] += x[rows,K]^T dy[rows,N],  db[N] += sum_rows dy
//   * softmax(L*P) + sampling-location backward:  (grad_loc, grad_attn, attn) -> d[offsets | logits]
//   * dV row-mask + cast
// The data gradient of nn.Linear (dx = dy W^T) needs no kernel of its own: Paddle's [in,out] weight layout is exactly
// the pre-packed [N',K'] operand of emrt_linear_fwd for that product.
// dW here is an fp32-accumulate SIMT tile kernel with split-row partial sums (atomics into the fp32 gradient); it is
// the parity implementation — the tcgen05 version (MN-major operands) is the next step.
#include <cstdlib>

#include <type_traits>
#include "common.cuh"

namespace emrt {

constexpr int WG_T = 64;      // output tile: 64 (k) x 64 (n)
constexpr int WG_R = 16;      // rows per smem step

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return to_float(__ldg(p)); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((unsigned int)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
linear_bwd_weight_kernel(const TX* __restrict__ x, const TY* __restrict__ dy, float* __restrict__ dw,
                         float* __restrict__ db, int64_t rows, int K, int N, int rows_per_cta) {
  __shared__ float xs[WG_R][WG_T + 1];
  __shared__ float ys[WG_R][WG_T + 1];
  const int k0 = blockIdx.x * WG_T, n0 = blockIdx.y * WG_T;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < rows ? r_begin + rows_per_cta : rows;
  const int t = threadIdx.x;
  const int tk = (t / 16) * 4, tn = (t % 16) * 4;      // 4 x 4 outputs per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;                                      // column sum of dy (threads 0..63 of the k-tile-0 CTAs)
  const int lr = t / 16, lc = (t % 16) * 4;             // loader: row lr, 4 consecutive columns
  for (int64_t r0 = r_begin; r0 < r_end; r0 += WG_R) {
    const int64_t r = r0 + lr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lc + i, n = n0 + lc + i;
      xs[lr][lc + i] = (r < r_end && k < K) ? ldf<TX>(x + r * K + k) : 0.f;
      ys[lr][lc + i] = (r < r_end && n < N) ? ldf<TY>(dy + r * N + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < WG_R; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = xs[rr][tk + i]; b[i] = ys[rr][tn + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (db != nullptr && blockIdx.x == 0 && t < WG_T) {
#pragma unroll
      for (int rr = 0; rr < WG_R; ++rr) bsum += ys[rr][t];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk + i, n = n0 + tn + j;
      if (k < K && n < N) atomicAdd(dw + (int64_t)k * N + n, acc[i][j]);
    }
  if (db != nullptr && blockIdx.x == 0 && t < WG_T && n0 + t < N) atomicAdd(db + n0 + t, bsum);
}

// One thread per (row, head): d_logit_i = a_i (g_i - sum_j a_j g_j) over the head's L*P weights (softmax backward,
// t_e_d.py:95), d_offset = grad_loc (pixel-offset mode) or grad_loc / (W_l, H_l) (normalised mode, t_e_d.py:98-102).
// dq [rows, 3*M*L*P] = [offset grads (2*M*L*P) | logit grads (M*L*P)], the layout of the fused query projection.
template <typename TA, typename TO, int MODE>
__global__ void __launch_bounds__(256)
msda_qproj_bwd_kernel(const float* __restrict__ grad_loc, const float* __restrict__ grad_attn,
                      const TA* __restrict__ attn, TO* __restrict__ dq, int M, int L, int P,
                      const __grid_constant__ LevelTable lv, int64_t n_items) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  const int m = (int)(item % M);
  const int64_t row = item / M;
  const int LP = L * P, tp = M * LP;
  const float* ga = grad_attn + item * LP;
  const TA* a = attn + item * LP;
  float dot = 0.f;
  for (int i = 0; i < LP; ++i) dot = fmaf(to_float(a[i]), __ldg(ga + i), dot);
  TO* out = dq + row * (3 * tp);
  for (int i = 0; i < LP; ++i)
    out[2 * tp + m * LP + i] = from_float<TO>(to_float(a[i]) * (__ldg(ga + i) - dot));
  const float* gl = grad_loc + item * LP * 2;
  for (int l = 0; l < L; ++l) {
    const float sx = MODE == EMRT_LOC_NORMALIZED ? 1.f / (float)lv.W[l] : 1.f;
    const float sy = MODE == EMRT_LOC_NORMALIZED ? 1.f / (float)lv.H[l] : 1.f;
    for (int p = 0; p < P; ++p) {
      const int i = (l * P + p) * 2;
      out[m * LP * 2 + i] = from_float<TO>(__ldg(gl + i) * sx);
      out[m * LP * 2 + i + 1] = from_float<TO>(__ldg(gl + i + 1) * sy);
    }
  }
}

// Vectorised form of the above for 16-bit attn / bf16 dq, L*P even and 2*M*L*P a multiple of 4 (EMRT: 18 and 288).
// The offset half of a dq row is an element-wise cast of the row's grad_loc (float4 in, 4 x bf16 out, scaled by
// 1 / (W_l, H_l) in normalised mode); the logit half is one thread per (row, head) with float2 / 32-bit accesses.
// Everything is contiguous across items, so the kernel streams at HBM speed instead of issuing 4-byte strided loads.
template <typename TA, int MODE>
__global__ void __launch_bounds__(256)
msda_qproj_bwd_fast_kernel(const float* __restrict__ grad_loc, const float* __restrict__ grad_attn,
                           const TA* __restrict__ attn, __nv_bfloat16* __restrict__ dq, int M, int L, int P,
                           const __grid_constant__ LevelTable lv, int64_t rows) {
  const int LP = L * P, tp = M * LP, off_len = 2 * tp, row_len = 3 * tp;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // ---- offsets: dq[row, 0 : 2 tp] = grad_loc[row] * scale --------------------------------------------------------------
  const int q4 = off_len / 4;
  for (int64_t c = tid; c < rows * q4; c += nth) {
    const int64_t row = c / q4;
    const int j = (int)(c - row * q4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(grad_loc + row * off_len + j));
    float o[4] = {v.x, v.y, v.z, v.w};
    if (MODE == EMRT_LOC_NORMALIZED) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = (j + k) % (2 * LP), l = i / (2 * P);
        o[k] *= 1.f / (float)(((j + k) & 1) ? lv.H[l] : lv.W[l]);
      }
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dq + row * row_len + j) = pk;
  }
  // ---- logits: softmax backward per (row, head) -------------------------------------------------------------------------
  const int64_t n_items = rows * M;
  for (int64_t item = tid; item < n_items; item += nth) {
    const int m = (int)(item % M);
    const int64_t row = item / M;
    const float2* ga = reinterpret_cast<const float2*>(grad_attn + item * LP);
    const uint32_t* a = reinterpret_cast<const uint32_t*>(attn + item * LP);
    uint32_t* out = reinterpret_cast<uint32_t*>(dq + row * row_len + off_len + m * LP);
    float dot = 0.f;
    for (int i = 0; i < LP / 2; ++i) {
      const float2 g2 = __ldg(ga + i);
      const uint32_t aw = __ldg(a + i);
      float a0, a1;
      if (sizeof(TA) == 2 && std::is_same<TA, __half>::value) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw));
        a0 = f.x; a1 = f.y;
      } else {
        a0 = __uint_as_float(aw << 16); a1 = __uint_as_float(aw & 0xffff0000u);
      }
      dot = fmaf(a1, g2.y, fmaf(a0, g2.x, dot));
    }
    for (int i = 0; i < LP / 2; ++i) {
      const float2 g2 = __ldg(ga + i);
      const uint32_t aw = __ldg(a + i);
      float a0, a1;
      if (sizeof(TA) == 2 && std::is_same<TA, __half>::value) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw));
        a0 = f.x; a1 = f.y;
      } else {
        a0 = __uint_as_float(aw << 16); a1 = __uint_as_float(aw & 0xffff0000u);
      }
      __nv_bfloat162 o = __floats2bfloat162_rn(a0 * (g2.x - dot), a1 * (g2.y - dot));
      out[i] = *reinterpret_cast<uint32_t*>(&o);
    }
  }
}

// dst[r, c] = src[r, c] * row_scale[r] (row_scale may be NULL), fp32 -> TO; 4 elements per thread
template <typename TO>
__global__ void __launch_bounds__(256)
scale_rows_cast_kernel(const float* __restrict__ src, const float* __restrict__ row_scale, TO* __restrict__ dst,
                       int64_t rows, int cols) {
  const int64_t n4 = rows * cols / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    const float s = row_scale ? __ldg(row_scale + (i * 4) / cols) : 1.f;
    dst[i * 4 + 0] = from_float<TO>(v.x * s);
    dst[i * 4 + 1] = from_float<TO>(v.y * s);
    dst[i * 4 + 2] = from_float<TO>(v.z * s);
    dst[i * 4 + 3] = from_float<TO>(v.w * s);
  }
}


// Gradient of the reference points (sampling_locations = reference_points + offsets / (W_l, H_l), t_e_d.py:98-102, so
// d ref[b,q,l,:] = sum over heads and points of d loc[b,q,m,l,p,:]; in PIXEL_OFFSET mode grad_loc is per pixel and
// d x / d ref_x = W_l).  One thread per (reference batch, query, level); a batch-shared reference (ref_batches == 1, the
// decoder's sigmoid(Linear(query_pos_embed)), t_e_d.py:466) sums over the batch in a fixed order: no atomics.
// One WARP per (reference batch, query, level): lane j sums items j, j + 32, ... of the (batch, head, point) list, then a
// fixed shuffle tree — deterministic, and 32 loads in flight instead of one serial chain of B * M * P.
__global__ void msda_ref_bwd_kernel(const float* __restrict__ grad_loc, float* __restrict__ grad_ref, int B, int ref_batches,
                                    int Lq, int M, int L, int P, LevelTable lv, int pixel_mode) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= (int64_t)ref_batches * Lq * L) return;
  const int l = (int)(i % L);
  const int q = (int)((i / L) % Lq);
  const int rb = (int)(i / ((int64_t)L * Lq));
  const int nb = (B - rb + ref_batches - 1) / ref_batches, MP = M * P;
  float gx = 0.f, gy = 0.f;
  for (int j = lane; j < nb * MP; j += 32) {
    const int bi = j / MP, r = j - bi * MP, m = r / P, pp = r - m * P;
    const int b = rb + bi * ref_batches;
    const float2 v = __ldg(reinterpret_cast<const float2*>(grad_loc) + ((((int64_t)b * Lq + q) * M + m) * L + l) * P + pp);
    gx += v.x;
    gy += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { gx += __shfl_xor_sync(0xffffffffu, gx, o); gy += __shfl_xor_sync(0xffffffffu, gy, o); }
  if (lane == 0) {
    if (pixel_mode) { gx *= (float)lv.W[l]; gy *= (float)lv.H[l]; }
    reinterpret_cast<float2*>(grad_ref)[i] = make_float2(gx, gy);
  }
}

}  // namespace emrt

using namespace emrt;

namespace emrt {
int linear_bwd_weight_tc(const void* x, const void* dy, float* dw, int64_t rows, int K, int N, cudaStream_t st);
int colsum(const void* dy, float* db, int64_t rows, int N, int dtype, cudaStream_t st);
}

extern "C" int emrt_linear_bwd_weight(const void* x, const void* dy, float* dw, float* db, int64_t rows, int K, int N,
                                      int x_dtype, int dy_dtype, void* stream) {
  EMRT_REQUIRE(x && dy && dw, "NULL tensor pointer");
  EMRT_REQUIRE(rows > 0 && K > 0 && N > 0, "non-positive dimension");
  if (x_dtype == EMRT_BF16 && dy_dtype == EMRT_BF16 && !getenv("EMRT_DW_SIMT")) {
    // tensor-core path (MN-major tcgen05, split-K); the bias gradient is a separate column-sum pass over dy
    const int e = linear_bwd_weight_tc(x, dy, dw, rows, K, N, as_stream(stream));
    if (e == EMRT_OK) return db ? colsum(dy, db, rows, N, dy_dtype, as_stream(stream)) : EMRT_OK;
    if (e != EMRT_ERR_UNSUPPORTED) return e;
  }
  int rows_per_cta = 1024;
  int64_t chunks = (rows + rows_per_cta - 1) / rows_per_cta;
  if (chunks > 65535) { rows_per_cta = (int)((rows + 65534) / 65535); rows_per_cta = (rows_per_cta + WG_R - 1) / WG_R * WG_R; chunks = (rows + rows_per_cta - 1) / rows_per_cta; }
  dim3 grid((K + WG_T - 1) / WG_T, (N + WG_T - 1) / WG_T, (unsigned)chunks);
  cudaStream_t st = as_stream(stream);
#define EMRT_WG(TX, TY) linear_bwd_weight_kernel<TX, TY><<<grid, 256, 0, st>>>((const TX*)x, (const TY*)dy, dw, db, rows, K, N, rows_per_cta)
  if (x_dtype == EMRT_F32 && dy_dtype == EMRT_F32) EMRT_WG(float, float);
  else if (x_dtype == EMRT_BF16 && dy_dtype == EMRT_BF16) EMRT_WG(__nv_bfloat16, __nv_bfloat16);
  else if (x_dtype == EMRT_BF16 && dy_dtype == EMRT_F32) EMRT_WG(__nv_bfloat16, float);
  else if (x_dtype == EMRT_F32 && dy_dtype == EMRT_BF16) EMRT_WG(float, __nv_bfloat16);
  else return set_error(EMRT_ERR_UNSUPPORTED, "linear_bwd_weight: unsupported dtypes %d/%d", x_dtype, dy_dtype);
#undef EMRT_WG
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_msda_qproj_bwd(const float* grad_loc, const float* grad_attn, const void* attn, void* dq,
                                   int64_t rows, int M, int L, int P, const int32_t* shapes_hw_host, int attn_dtype,
                                   int dq_dtype, int mode, void* stream) {
  EMRT_REQUIRE(grad_loc && grad_attn && attn && dq, "NULL tensor pointer");
  EMRT_REQUIRE(rows > 0 && M > 0 && P > 0, "non-positive dimension");
  EMRT_REQUIRE(mode == EMRT_LOC_NORMALIZED || mode == EMRT_LOC_PIXEL_OFFSET, "bad loc mode");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, -1)) return e;
  const int64_t n_items = rows * M;
  const unsigned blocks = (unsigned)((n_items + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (dq_dtype == EMRT_BF16 && (attn_dtype == EMRT_F16 || attn_dtype == EMRT_BF16) && ((L * P) & 1) == 0 &&
      ((2 * M * L * P) & 3) == 0 && (reinterpret_cast<uintptr_t>(grad_loc) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(grad_attn) & 7) == 0 && (reinterpret_cast<uintptr_t>(attn) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(dq) & 7) == 0 && !getenv("EMRT_QPROJ_BWD_SLOW")) {
    const int64_t want = (rows * (2 * M * L * P / 4) + 255) / 256;
    const unsigned fb = (unsigned)(want < (int64_t)num_sms() * 32 ? want : (int64_t)num_sms() * 32);
    const bool pxf = mode == EMRT_LOC_PIXEL_OFFSET;
#define EMRT_QBF(TA, MODE) msda_qproj_bwd_fast_kernel<TA, MODE><<<fb, 256, 0, st>>>(grad_loc, grad_attn, (const TA*)attn, (__nv_bfloat16*)dq, M, L, P, lv, rows)
    if (attn_dtype == EMRT_F16) { if (pxf) EMRT_QBF(__half, 1); else EMRT_QBF(__half, 0); }
    else { if (pxf) EMRT_QBF(__nv_bfloat16, 1); else EMRT_QBF(__nv_bfloat16, 0); }
#undef EMRT_QBF
    EMRT_LAUNCH_CHECK();
    return EMRT_OK;
  }
#define EMRT_QB(TA, TO, MODE) msda_qproj_bwd_kernel<TA, TO, MODE><<<blocks, 256, 0, st>>>(grad_loc, grad_attn, (const TA*)attn, (TO*)dq, M, L, P, lv, n_items)
  const bool px = mode == EMRT_LOC_PIXEL_OFFSET;
  if (attn_dtype == EMRT_F32 && dq_dtype == EMRT_F32) { if (px) EMRT_QB(float, float, 1); else EMRT_QB(float, float, 0); }
  else if (attn_dtype == EMRT_F16 && dq_dtype == EMRT_BF16) { if (px) EMRT_QB(__half, __nv_bfloat16, 1); else EMRT_QB(__half, __nv_bfloat16, 0); }
  else if (attn_dtype == EMRT_BF16 && dq_dtype == EMRT_BF16) { if (px) EMRT_QB(__nv_bfloat16, __nv_bfloat16, 1); else EMRT_QB(__nv_bfloat16, __nv_bfloat16, 0); }
  else if (attn_dtype == EMRT_F16 && dq_dtype == EMRT_F32) { if (px) EMRT_QB(__half, float, 1); else EMRT_QB(__half, float, 0); }
  else return set_error(EMRT_ERR_UNSUPPORTED, "qproj_bwd: unsupported dtypes attn %d / dq %d", attn_dtype, dq_dtype);
#undef EMRT_QB
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_scale_rows_cast(const float* src, const float* row_scale, void* dst, int64_t rows, int cols,
                                    int dst_dtype, void* stream) {
  EMRT_REQUIRE(src && dst && rows > 0 && cols > 0 && cols % 4 == 0, "bad scale_rows_cast arguments");
  const int64_t n4 = rows * cols / 4;
  const int64_t want = (n4 + 255) / 256;
  const unsigned blocks = (unsigned)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  cudaStream_t st = as_stream(stream);
  if (dst_dtype == EMRT_F32) scale_rows_cast_kernel<float><<<blocks, 256, 0, st>>>(src, row_scale, (float*)dst, rows, cols);
  else if (dst_dtype == EMRT_BF16) scale_rows_cast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(src, row_scale, (__nv_bfloat16*)dst, rows, cols);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dst_dtype %d", dst_dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_msda_ref_bwd(const float* grad_loc, float* grad_ref, int B, int ref_batches, int Lq, int M, int L, int P,
                                 const int32_t* shapes_hw_host, int mode, void* stream) {
  EMRT_REQUIRE(grad_loc && grad_ref, "NULL tensor pointer");
  EMRT_REQUIRE(B > 0 && Lq > 0 && M > 0 && P > 0 && (ref_batches == 1 || ref_batches == B), "bad msda_ref_bwd dimensions");
  EMRT_REQUIRE(mode == EMRT_LOC_NORMALIZED || mode == EMRT_LOC_PIXEL_OFFSET, "bad loc mode");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, -1)) return e;
  const int64_t n = (int64_t)ref_batches * Lq * L;
  msda_ref_bwd_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, as_stream(stream)>>>(grad_loc, grad_ref, B, ref_batches, Lq, M, L, P,
                                                                              lv, mode == EMRT_LOC_PIXEL_OFFSET);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
