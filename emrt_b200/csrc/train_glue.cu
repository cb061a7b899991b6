// Backward of the encoder / decoder glue around MSDeformableAttention (SURVEY.md §8e cfg 4: the reference trains the whole
// EncoderDecoder, train.py:146-159 -> transformer_encoder_decoder.py:184-204,282-295, layers.py:236-311):
//   LayerNorm(a + b) backward, GroupNorm (+ GELU) backward on the token layout, ReLU mask, the 110-token self-attention
//   backward, the 3x3 conv weight gradient (SIMT form; the tcgen05 form is conv3x3_bwd_tcgen05.cu), batch / column sums and
//   the sigmoid of the reference-point head.
// Every parameter gradient is reduced in a FIXED order (per-CTA partials in a caller-provided workspace, then one ordered
// reduction): no floating-point atomics, run-to-run reproducible.  Activations T = fp32 (parity path) or bf16; all
// arithmetic fp32.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace emrt {

int colsum(const void* dy, float* db, int64_t rows, int N, int dtype, cudaStream_t st);     // linear_bwd_tcgen05.cu
int conv3x3_bwd_weight_tc(const void* x, const void* dy, float* ws, int B, int Lv, int C, int L, const LevelTable& lv,
                          cudaStream_t st);                                                 // conv3x3_bwd_tcgen05.cu

// out[j] (+)= sum_p parts[p * width + j], p in increasing order (deterministic)
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* __restrict__ parts, float* __restrict__ out, int n_parts, int width, int accumulate) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= width) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += parts[(int64_t)p * width + j];
  out[j] = accumulate ? out[j] + s : s;
}

// ---- LayerNorm backward ------------------------------------------------------------------------------------------------
// y = LN(z) * gamma + beta, z = a + b.  One warp per row, N <= 1024, N % 32 == 0.
//   dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dgamma = sum dy * xhat;  dbeta = sum dy
constexpr int LNB_WARPS = 8, LNB_ROWS_PER_WARP = 8;
template <typename T, int PER>
__global__ void __launch_bounds__(LNB_WARPS * 32)
layernorm_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ gamma, const T* __restrict__ dy,
                     T* __restrict__ dz, float* __restrict__ parts, int64_t rows, float eps) {
  constexpr int N = PER * 32;
  __shared__ float red[LNB_WARPS][2][N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gm[PER], dg[PER], db[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) { gm[i] = __ldg(gamma + lane + 32 * i); dg[i] = 0.f; db[i] = 0.f; }
  // grid-stride over blocks of LNB_WARPS * LNB_ROWS_PER_WARP rows: the number of partials is the (bounded) number of CTAs
  for (int64_t blk = blockIdx.x; blk * (LNB_WARPS * LNB_ROWS_PER_WARP) < rows; blk += gridDim.x)
  for (int rr = 0; rr < LNB_ROWS_PER_WARP; ++rr) {
    const int64_t r = (blk * LNB_WARPS + warp) * LNB_ROWS_PER_WARP + rr;
    if (r >= rows) break;
    float z[PER], d[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int64_t o = r * N + lane + 32 * i;
      z[i] = to_float(a[o]) + (b ? to_float(b[o]) : 0.f);
      d[i] = to_float(dy[o]);
      s += z[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)N;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { z[i] -= mean; v += z[i] * z[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)N + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      z[i] *= rstd;                                   // xhat
      const float g = d[i] * gm[i];
      m1 += g; m2 += g * z[i];
      dg[i] += d[i] * z[i]; db[i] += d[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    m1 /= (float)N; m2 /= (float)N;
#pragma unroll
    for (int i = 0; i < PER; ++i) dz[r * N + lane + 32 * i] = from_float<T>(rstd * (d[i] * gm[i] - m1 - z[i] * m2));
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) { red[warp][0][lane + 32 * i] = dg[i]; red[warp][1][lane + 32 * i] = db[i]; }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * N; j += blockDim.x) {
    const int which = j / N, c = j - which * N;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) t += red[w][which][c];
    parts[(int64_t)blockIdx.x * 2 * N + j] = t;
  }
}

// ---- GroupNorm (+ GELU) backward on tokens -----------------------------------------------------------------------------
// forward: xhat = (x - mean) * rstd per (batch, level, group); h = xhat * gamma_l + beta_l; y = GELU(h) (or h).
constexpr int GNB_CHUNKS = 16;
__device__ __forceinline__ float gelu_grad(float h) {
  // d/dh [h * Phi(h)] = Phi(h) + h * phi(h)
  const float phi = 0.3989422804014327f * __expf(-0.5f * h * h);
  return 0.5f * erfcf(-h * 0.70710678118654752f) + h * phi;
}

// pass A: partial sums.  grid (GNB_CHUNKS, B * L); thread t owns channels t, t + 256, ...
//   group partials  gp[((bl * CHUNKS + chunk) * G + g) * 2] = (sum g, sum g * xhat),  g = dy * act'(h) * gamma
//   channel partials cp[((bl * CHUNKS + chunk) * 2 + {0,1}) * C + c] = (sum dy * act' * xhat, sum dy * act')
template <typename T, bool GELU>
__global__ void __launch_bounds__(256)
groupnorm_bwd_sums_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ stats,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ gp,
                          float* __restrict__ cp, int Lv, int C, int L, int G, float eps, const __grid_constant__ LevelTable lv) {
  extern __shared__ float sh[];                 // [2][C]
  const int bl = blockIdx.y, b = bl / L, l = bl - b * L;
  const int P = lv.H[l] * lv.W[l];
  const int per = (P + GNB_CHUNKS - 1) / GNB_CHUNKS;
  const int p0 = blockIdx.x * per, p1 = min(p0 + per, P);
  const int cpg = C / G;
  const float inv_cnt = 1.f / (float)(P * cpg);
  for (int c = threadIdx.x; c < C; c += 256) {
    const float* st = stats + ((int64_t)bl * G + c / cpg) * 2;
    const float mean = st[0] * inv_cnt;
    const float rstd = rsqrtf(fmaxf(st[1] * inv_cnt - mean * mean, 0.f) + eps);
    const float gm = gamma[l * C + c], bt = beta[l * C + c];
    float s1 = 0.f, s2 = 0.f, sg = 0.f, sb = 0.f;
    const int64_t base = ((int64_t)b * Lv + lv.start[l]) * C + c;
    for (int p = p0; p < p1; ++p) {
      const float xh = (to_float(x[base + (int64_t)p * C]) - mean) * rstd;
      float d = to_float(dy[base + (int64_t)p * C]);
      if (GELU) d *= gelu_grad(fmaf(xh, gm, bt));
      sg += d * xh; sb += d;
      const float g = d * gm;
      s1 += g; s2 += g * xh;
    }
    sh[c] = s1; sh[C + c] = s2;
    const int64_t co = ((int64_t)bl * GNB_CHUNKS + blockIdx.x) * 2 * C;
    cp[co + c] = sg; cp[co + C + c] = sb;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += 256) {
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < cpg; ++k) { t1 += sh[g * cpg + k]; t2 += sh[C + g * cpg + k]; }
    float* dst = gp + (((int64_t)bl * GNB_CHUNKS + blockIdx.x) * G + g) * 2;
    dst[0] = t1; dst[1] = t2;
  }
}

// pass A2: group sums over the chunks (fixed order) -> gs[(bl * G + g) * 2]; channel sums over (batch, chunk) -> dgamma / dbeta
__global__ void __launch_bounds__(256)
groupnorm_bwd_reduce_kernel(const float* __restrict__ gp, const float* __restrict__ cp, float* __restrict__ gs,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int L, int G, int C) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int n_grp = B * L * G * 2;
  if (i < n_grp) {
    const int bl = i / (G * 2), r = i - bl * G * 2;
    float s = 0.f;
    for (int k = 0; k < GNB_CHUNKS; ++k) s += gp[((int64_t)bl * GNB_CHUNKS + k) * G * 2 + r];
    gs[i] = s;
  } else if (i < n_grp + 2 * L * C) {
    const int j = i - n_grp, which = j / (L * C), lc = j - which * L * C, l = lc / C, c = lc - l * C;
    float s = 0.f;
    for (int b = 0; b < B; ++b)
      for (int k = 0; k < GNB_CHUNKS; ++k) s += cp[((((int64_t)b * L + l) * GNB_CHUNKS + k) * 2 + which) * C + c];
    float* dst = which == 0 ? dgamma : dbeta;
    dst[lc] += s;
  }
}

// pass B: dx = rstd * (g - S1 / n - xhat * S2 / n).  One 16-byte vector per thread (VEC channels of one group: VEC <= C / G).
template <typename T, bool GELU>
__global__ void __launch_bounds__(256)
groupnorm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ stats,
                           const float* __restrict__ gs, const float* __restrict__ gamma, const float* __restrict__ beta,
                           T* __restrict__ dx, int Lv, int C, int L, int G, float eps, const __grid_constant__ LevelTable lv) {
  constexpr int VEC = Vec16<T>::N;
  const uint32_t vec_per_tok = (uint32_t)C / VEC;
  const uint32_t vi = blockIdx.x * 256u + threadIdx.x;           // vector index inside this image
  if (vi >= (uint32_t)Lv * vec_per_tok) return;
  const uint32_t t = vi / vec_per_tok;
  const int c0 = (int)(vi - t * vec_per_tok) * VEC;
  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < L && (int)t >= lv.start[l + 1]) ++l;
  const int cpg = C / G, g = c0 / cpg;
  const float inv_cnt = 1.f / (float)(lv.H[l] * lv.W[l] * cpg);
  const float* st = stats + (((int64_t)b * L + l) * G + g) * 2;
  const float mean = st[0] * inv_cnt;
  const float rstd = rsqrtf(fmaxf(st[1] * inv_cnt - mean * mean, 0.f) + eps);
  const float* sg = gs + (((int64_t)b * L + l) * G + g) * 2;
  const float a1 = sg[0] * inv_cnt, a2 = sg[1] * inv_cnt;
  const int64_t o = ((int64_t)b * Lv * vec_per_tok + vi) * VEC;
  float xv[VEC], dv[VEC], out[VEC];
  Vec16<T>::load(x + o, xv);
  Vec16<T>::load(dy + o, dv);
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float gm = __ldg(gamma + l * C + c0 + k);
    const float xh = (xv[k] - mean) * rstd;
    float d = dv[k];
    if (GELU) d *= gelu_grad(fmaf(xh, gm, __ldg(beta + l * C + c0 + k)));
    out[k] = rstd * (d * gm - a1 - xh * a2);
  }
  Vec16<T>::store(dx + o, out);
}

// ---- small elementwise / reduction pieces ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    dx[i] = to_float(y[i]) > 0.f ? dy[i] : from_float<T>(0.f);
}

// out[j] = sum_b x[b * n + j]  (fp32 out), j < n
template <typename T>
__global__ void __launch_bounds__(256) batch_sum_kernel(const T* __restrict__ x, float* __restrict__ out, int B, int64_t n) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += to_float(x[(int64_t)b * n + j]);
  out[j] = s;
}

__global__ void __launch_bounds__(256) sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) y[i] = 1.f / (1.f + expf(-x[i]));
}
__global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

// ---- self-attention backward (110 tokens, head dim 32): one CTA per (batch, head) ----------------------------------------
// P = softmax(scale Q K^T); O = P V.   dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(dP o P));  dQ = scale dS K;  dK = scale dS^T Q
template <typename T, int D>
__global__ void __launch_bounds__(256)
mha_small_bwd_kernel(const T* __restrict__ q, int64_t q_ld, const T* __restrict__ k, int64_t k_ld, const T* __restrict__ v, int64_t v_ld,
                     const T* __restrict__ d_out, T* __restrict__ dq, int64_t dq_ld, T* __restrict__ dk, int64_t dk_ld,
                     T* __restrict__ dv, int64_t dv_ld, int Lq, int Lk, int M, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                      // [Lq][D]
  float* ks = qs + Lq * D;             // [Lk][D]
  float* vs = ks + Lk * D;             // [Lk][D]
  float* os = vs + Lk * D;             // [Lq][D]   dO
  float* ps = os + Lq * D;             // [Lq][Lk]  P, then dS
  const int m = blockIdx.x % M;
  const int64_t b = blockIdx.x / M;
  for (int i = threadIdx.x; i < Lq * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D;
    qs[i] = to_float(q[(b * Lq + r) * q_ld + m * D + d]);
    os[i] = to_float(d_out[(b * Lq + r) * (int64_t)(M * D) + m * D + d]);
  }
  for (int i = threadIdx.x; i < Lk * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D;
    ks[i] = to_float(k[(b * Lk + r) * k_ld + m * D + d]);
    vs[i] = to_float(v[(b * Lk + r) * v_ld + m * D + d]);
  }
  __syncthreads();
  // scores
  for (int i = threadIdx.x; i < Lq * Lk; i += blockDim.x) {
    const int r = i / Lk, c = i - r * Lk;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) s = fmaf(qs[r * D + d], ks[c * D + d], s);
    ps[i] = s * scale;
  }
  __syncthreads();
  // row softmax, then dS in place: one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < Lq; r += nw) {
    float mx = -INFINITY;
    for (int c = lane; c < Lk; c += 32) mx = fmaxf(mx, ps[r * Lk + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < Lk; c += 32) { const float e = expf(ps[r * Lk + c] - mx); ps[r * Lk + c] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int c = lane; c < Lk; c += 32) ps[r * Lk + c] *= inv;
  }
  __syncthreads();
  // dV[c][d] = sum_r P[r][c] dO[r][d]   (before P is overwritten)
  for (int i = threadIdx.x; i < Lk * D; i += blockDim.x) {
    const int c = i / D, d = i - c * D;
    float s = 0.f;
    for (int r = 0; r < Lq; ++r) s = fmaf(ps[r * Lk + c], os[r * D + d], s);
    dv[(b * Lk + c) * dv_ld + m * D + d] = from_float<T>(s);
  }
  __syncthreads();
  for (int r = warp; r < Lq; r += nw) {
    float delta = 0.f;
    for (int c = lane; c < Lk; c += 32) {
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) dp = fmaf(os[r * D + d], vs[c * D + d], dp);
      delta += dp * ps[r * Lk + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
    for (int c = lane; c < Lk; c += 32) {
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) dp = fmaf(os[r * D + d], vs[c * D + d], dp);
      ps[r * Lk + c] = ps[r * Lk + c] * (dp - delta) * scale;       // scale folded in: dQ = dS K, dK = dS^T Q
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Lq * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D;
    float s = 0.f;
    for (int c = 0; c < Lk; ++c) s = fmaf(ps[r * Lk + c], ks[c * D + d], s);
    dq[(b * Lq + r) * dq_ld + m * D + d] = from_float<T>(s);
  }
  for (int i = threadIdx.x; i < Lk * D; i += blockDim.x) {
    const int c = i / D, d = i - c * D;
    float s = 0.f;
    for (int r = 0; r < Lq; ++r) s = fmaf(ps[r * Lk + c], qs[r * D + d], s);
    dk[(b * Lk + c) * dk_ld + m * D + d] = from_float<T>(s);
  }
}

// ---- 3x3 conv weight gradient, SIMT form (fp32 parity path and shapes the tcgen05 kernel does not tile) ------------------
// ws[l][tap][ci][co] = sum_{b, pixel} x[b, pixel + shift(tap), ci] * dy[b, pixel, co]     (zero padding)
// grid (Cin, 9, L), thread = co
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_bwd_weight_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ ws, int B, int Lv, int C,
                               const __grid_constant__ LevelTable lv) {
  const int ci = blockIdx.x, tap = blockIdx.y, l = blockIdx.z;
  const int ky = tap / 3 - 1, kx = tap % 3 - 1;
  const int H = lv.H[l], W = lv.W[l];
  for (int co = threadIdx.x; co < C; co += 256) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const int64_t base = ((int64_t)b * Lv + lv.start[l]) * C;
      for (int y = 0; y < H; ++y) {
        const int yy = y + ky;
        if (yy < 0 || yy >= H) continue;
        for (int xq = 0; xq < W; ++xq) {
          const int xx = xq + kx;
          if (xx < 0 || xx >= W) continue;
          s = fmaf(to_float(x[base + (int64_t)(yy * W + xx) * C + ci]), to_float(dy[base + (int64_t)(y * W + xq) * C + co]), s);
        }
      }
    }
    ws[(((int64_t)l * 9 + tap) * C + ci) * C + co] = s;
  }
}

// dw[l][co][ci][tap] += ws[l][tap][ci][co]   (Paddle Conv2D layout [Cout, Cin, 3, 3] per level)
__global__ void __launch_bounds__(256)
conv3x3_dw_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw, int L, int C) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t n = (int64_t)L * C * C * 9;
  if (i >= n) return;
  const int tap = (int)(i % 9);
  const int ci = (int)((i / 9) % C);
  const int co = (int)((i / (9 * C)) % C);
  const int l = (int)(i / ((int64_t)9 * C * C));
  dw[i] += ws[(((int64_t)l * 9 + tap) * C + ci) * C + co];
}

}  // namespace emrt

using namespace emrt;

static int64_t lnb_ctas(int64_t rows) {
  const int64_t want = (rows + LNB_WARPS * LNB_ROWS_PER_WARP - 1) / (LNB_WARPS * LNB_ROWS_PER_WARP);
  return want < 2 * 148 ? want : 2 * 148;
}
extern "C" int64_t emrt_layernorm_bwd_workspace_floats(int64_t rows, int N) { return lnb_ctas(rows) * 2 * N; }

extern "C" int emrt_layernorm_bwd(const void* a, const void* b, const float* gamma, const void* dy, void* dz, float* dgamma,
                                  float* dbeta, float* workspace, int64_t rows, int N, float eps, int dtype, void* stream) {
  EMRT_REQUIRE(a && gamma && dy && dz && dgamma && dbeta && workspace, "NULL pointer");
  EMRT_REQUIRE(rows > 0 && N > 0, "non-positive dimension");
  if (N != 256 && N != 64 && N != 128 && N != 512)
    return set_error(EMRT_ERR_UNSUPPORTED, "layernorm_bwd: N must be one of 64, 128, 256, 512 (got %d)", N);
  cudaStream_t st = as_stream(stream);
  const unsigned ctas = (unsigned)lnb_ctas(rows);
#define EMRT_LNB(T, PER)                                                                                                   \
  layernorm_bwd_kernel<T, PER><<<ctas, LNB_WARPS * 32, 0, st>>>((const T*)a, (const T*)b, gamma, (const T*)dy, (T*)dz,      \
                                                                workspace, rows, eps)
#define EMRT_LNB_N(T)                                                                                                      \
  switch (N) { case 64: EMRT_LNB(T, 2); break; case 128: EMRT_LNB(T, 4); break; case 256: EMRT_LNB(T, 8); break;           \
               default: EMRT_LNB(T, 16); break; }
  if (dtype == EMRT_F32) { EMRT_LNB_N(float) }
  else if (dtype == EMRT_BF16) { EMRT_LNB_N(__nv_bfloat16) }
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
#undef EMRT_LNB_N
#undef EMRT_LNB
  EMRT_LAUNCH_CHECK();
  // dgamma | dbeta: the per-CTA partials are [cta][2][N]; reduce each half in CTA order
  reduce_parts_kernel<<<(2 * N + 255) / 256, 256, 0, st>>>(workspace, workspace + (int64_t)ctas * 2 * N, (int)ctas, 2 * N, 0);
  EMRT_LAUNCH_CHECK();
  // accumulate into the caller's gradient tensors
  reduce_parts_kernel<<<(N + 255) / 256, 256, 0, st>>>(workspace + (int64_t)ctas * 2 * N, dgamma, 1, N, 1);
  reduce_parts_kernel<<<(N + 255) / 256, 256, 0, st>>>(workspace + (int64_t)ctas * 2 * N + N, dbeta, 1, N, 1);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int64_t emrt_groupnorm_bwd_workspace_floats(int B, int L, int C, int groups) {
  return (int64_t)B * L * GNB_CHUNKS * (groups * 2 + 2 * C) + (int64_t)B * L * groups * 2;
}

extern "C" int emrt_groupnorm_bwd(const void* x, const void* dy, const float* stats, const float* gamma, const float* beta,
                                  void* dx, float* dgamma, float* dbeta, float* workspace, int B, int Lv, int C, int L,
                                  int groups, float eps, const int32_t* shapes_hw_host, int gelu, int dtype, void* stream) {
  EMRT_REQUIRE(x && dy && stats && gamma && beta && dx && dgamma && dbeta && workspace, "NULL pointer");
  EMRT_REQUIRE(B > 0 && B <= 65535 && C > 0 && groups > 0 && C % groups == 0 && C <= 4096, "bad groupnorm_bwd shape");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, Lv)) return e;
  cudaStream_t st = as_stream(stream);
  float* gp = workspace;
  float* cp = gp + (int64_t)B * L * GNB_CHUNKS * groups * 2;
  float* gs = cp + (int64_t)B * L * GNB_CHUNKS * 2 * C;
  const dim3 sgrid(GNB_CHUNKS, (unsigned)(B * L));
  const size_t sh = sizeof(float) * 2 * C;
  const int vec = dtype == EMRT_F32 ? 4 : 8;
  EMRT_REQUIRE((C / groups) % vec == 0 && C % vec == 0, "groupnorm_bwd: channels per group must be a multiple of the 16-byte vector (4 fp32 / 8 bf16)");
  const dim3 agrid((unsigned)(((int64_t)Lv * (C / vec) + 255) / 256), (unsigned)B);
  const int n_red = B * L * groups * 2 + 2 * L * C;
#define EMRT_GNB(T, G_)                                                                                                         \
  do {                                                                                                                          \
    groupnorm_bwd_sums_kernel<T, G_><<<sgrid, 256, sh, st>>>((const T*)x, (const T*)dy, stats, gamma, beta, gp, cp, Lv, C, L,    \
                                                             groups, eps, lv);                                                  \
    count_launch();                                                                                                             \
    groupnorm_bwd_reduce_kernel<<<(n_red + 255) / 256, 256, 0, st>>>(gp, cp, gs, dgamma, dbeta, B, L, groups, C);               \
    count_launch();                                                                                                             \
    groupnorm_bwd_apply_kernel<T, G_><<<agrid, 256, 0, st>>>((const T*)x, (const T*)dy, stats, gs, gamma, beta, (T*)dx, Lv, C,   \
                                                             L, groups, eps, lv);                                               \
  } while (0)
  if (dtype == EMRT_F32) { if (gelu) EMRT_GNB(float, true); else EMRT_GNB(float, false); }
  else if (dtype == EMRT_BF16) { if (gelu) EMRT_GNB(__nv_bfloat16, true); else EMRT_GNB(__nv_bfloat16, false); }
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
#undef EMRT_GNB
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, int dtype, void* stream) {
  EMRT_REQUIRE(dy && y && dx && n > 0, "bad relu_bwd arguments");
  const int64_t want = (n + 255) / 256;
  const unsigned blocks = (unsigned)(want < (int64_t)num_sms() * 32 ? want : (int64_t)num_sms() * 32);
  cudaStream_t st = as_stream(stream);
  if (dtype == EMRT_F32) relu_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)dy, (const float*)y, (float*)dx, n);
  else if (dtype == EMRT_BF16) relu_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dx, n);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_batch_sum(const void* x, float* out, int B, int64_t n, int dtype, void* stream) {
  EMRT_REQUIRE(x && out && B > 0 && n > 0, "bad batch_sum arguments");
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == EMRT_F32) batch_sum_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, out, B, n);
  else if (dtype == EMRT_BF16) batch_sum_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, out, B, n);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_column_sum(const void* x, float* out, int64_t rows, int N, int dtype, void* stream) {
  EMRT_REQUIRE(x && out && rows > 0 && N > 0, "bad column_sum arguments");
  return colsum(x, out, rows, N, dtype, as_stream(stream));        // out[n] += sum_r x[r, n]
}

extern "C" int emrt_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream) {
  EMRT_REQUIRE(x && y && n > 0, "bad sigmoid arguments");
  sigmoid_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, n);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  EMRT_REQUIRE(dy && y && dx && n > 0, "bad sigmoid arguments");
  sigmoid_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dy, y, dx, n);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_mha_small_bwd(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld,
                                  const void* d_out, void* dq, int64_t dq_ld, void* dk, int64_t dk_ld, void* dv, int64_t dv_ld,
                                  int B, int Lq, int Lk, int M, int D, float scale, int dtype, void* stream) {
  EMRT_REQUIRE(q && k && v && d_out && dq && dk && dv && B > 0 && Lq > 0 && Lk > 0 && M > 0, "bad mha_small_bwd arguments");
  if (D != 32) return set_error(EMRT_ERR_UNSUPPORTED, "mha_small_bwd is built for head dim 32 (got %d)", D);
  const size_t smem = sizeof(float) * ((size_t)(2 * Lq + 2 * Lk) * 32 + (size_t)Lq * Lk);
  if (smem > 200 * 1024) return set_error(EMRT_ERR_UNSUPPORTED, "mha_small_bwd keeps one head's Q, K, V, dO and P in shared memory: %zu bytes > 200 KB", smem);
  cudaStream_t st = as_stream(stream);
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(mha_small_bwd_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  EMRT_CUDA_CHECK(cudaFuncSetAttribute(mha_small_bwd_kernel<__nv_bfloat16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const unsigned grid = (unsigned)(B * M);
  if (dtype == EMRT_F32)
    mha_small_bwd_kernel<float, 32><<<grid, 256, smem, st>>>((const float*)q, q_ld, (const float*)k, k_ld, (const float*)v, v_ld,
        (const float*)d_out, (float*)dq, dq_ld, (float*)dk, dk_ld, (float*)dv, dv_ld, Lq, Lk, M, scale);
  else if (dtype == EMRT_BF16)
    mha_small_bwd_kernel<__nv_bfloat16, 32><<<grid, 256, smem, st>>>((const __nv_bfloat16*)q, q_ld, (const __nv_bfloat16*)k, k_ld,
        (const __nv_bfloat16*)v, v_ld, (const __nv_bfloat16*)d_out, (__nv_bfloat16*)dq, dq_ld, (__nv_bfloat16*)dk, dk_ld,
        (__nv_bfloat16*)dv, dv_ld, Lq, Lk, M, scale);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_conv3x3_tokens_bwd_weight(const void* x, const void* dy, float* dw, float* workspace, int B, int Lv, int C,
                                              int L, const int32_t* shapes_hw_host, int dtype, int impl, void* stream) {
  EMRT_REQUIRE(x && dy && dw && workspace, "NULL pointer");
  EMRT_REQUIRE(B > 0 && C > 0 && L > 0, "non-positive dimension");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, Lv)) return e;
  cudaStream_t st = as_stream(stream);
  bool done = false;
  if (dtype == EMRT_BF16 && impl != 1) {
    const int e = conv3x3_bwd_weight_tc(x, dy, workspace, B, Lv, C, L, lv, st);
    if (e == EMRT_OK) done = true;
    else if (e != EMRT_ERR_UNSUPPORTED || impl == 2)
      return e == EMRT_ERR_UNSUPPORTED ? set_error(e, "conv3x3 weight gradient: shape not tiled by the tcgen05 kernel") : e;
  }
  if (!done) {
    const dim3 grid((unsigned)C, 9, (unsigned)L);
    if (dtype == EMRT_F32) conv3x3_bwd_weight_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)dy, workspace, B, Lv, C, lv);
    else if (dtype == EMRT_BF16) conv3x3_bwd_weight_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, workspace, B, Lv, C, lv);
    else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
    EMRT_LAUNCH_CHECK();
  }
  const int64_t n = (int64_t)L * C * C * 9;
  conv3x3_dw_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, dw, L, C);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
