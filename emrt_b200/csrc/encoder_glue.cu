// Glue of TransformerEncoderLayer around the MSDA module (transformer_encoder_decoder.py:184-204):
//   * y = LayerNorm(x + residual) * gamma + beta (+ post_add)        norm1 / norm2 (:199-200,159-160) and the final
//                                                                    `src + src_flatten` (:203) folded into norm2's pass
//   * 3x3 conv branch, reference (SIMT) implementation                conv{l}: Conv2D(256,256,3,pad 1,no bias) (:125-144)
//   * GroupNorm(32) + exact GELU + skip on the conv output            (:187-189)
// All of it works on the token layout [B, Lv, C] (NHWC per level), so the reference's seq2_2D / flatten / transpose /
// concat copies (:163-196) do not exist here.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace emrt {

// ---- LayerNorm(x + residual) (+ post_add): one warp per row, 16-byte vectors, N <= 1024, N % (32 * VEC) == 0 ----------
// GN = true: post_add is not a tensor but the encoder layer's conv branch evaluated on the fly,
//   post_add[row, c] = GELU(GroupNorm_l(conv)[row, c]) + skip[row, c]      (transformer_encoder_decoder.py:187-189,203)
// from the conv output, the layer input and the (sum, sum of squares) statistics of groupnorm_stats_kernel — the branch
// tensor is never written or re-read.
struct GnBranch {
  const void* conv;
  const void* skip;
  const float* stats;     // [B, L, G, 2] sums from groupnorm_stats_kernel
  const float* gamma;     // [L, C]
  const float* beta;
  int Lv, L, G;
  float eps;
  LevelTable lv;
};

template <typename T, int PER, bool GN>   // PER = 16-byte vectors per lane
__global__ void __launch_bounds__(256)
residual_layernorm_kernel(const T* __restrict__ x, const T* __restrict__ residual, const float* __restrict__ gamma,
                          const float* __restrict__ beta, const T* __restrict__ post_add, T* __restrict__ y,
                          int64_t rows, int N, float eps, const __grid_constant__ GnBranch gn) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int RPW = (PER <= 2 && !GN) ? 2 : 1;           // rows per warp in flight (more loads outstanding); the
                                                           // conv-branch variant keeps one row: registers -> occupancy
  const int64_t row0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW;
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  float v[RPW][PER][VEC];
  float s[RPW], ss[RPW];
  const float inv_n = 1.f / (float)N;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + r < rows ? row0 + r : rows - 1;
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) Vec16<T>::load(x + row * N + (i * 32 + lane) * VEC, v[r][i]);
  }
  // conv-branch variant: its two extra streams (conv output, skip) are requested now, packed, so that all four loads of a
  // row are in flight together instead of two after the LayerNorm statistics
  uint4 pre_cv[GN ? PER : 1], pre_sk[GN ? PER : 1];
  if (GN) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      pre_cv[i] = Vec16<T>::load_raw(reinterpret_cast<const T*>(gn.conv) + row0 * N + (i * 32 + lane) * VEC);
      pre_sk[i] = Vec16<T>::load_raw(reinterpret_cast<const T*>(gn.skip) + row0 * N + (i * 32 + lane) * VEC);
    }
  }
  if (residual) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int64_t row = row0 + r < rows ? row0 + r : rows - 1;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        float t[VEC];
        Vec16<T>::load(residual + row * N + (i * 32 + lane) * VEC, t);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[r][i][k] += t[k];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
#pragma unroll
    for (int i = 0; i < PER; ++i)
#pragma unroll
      for (int k = 0; k < VEC; ++k) s[r] += v[r][i][k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < RPW; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    s[r] *= inv_n;          // mean
    ss[r] = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i)
#pragma unroll
      for (int k = 0; k < VEC; ++k) { const float d = v[r][i][k] - s[r]; ss[r] = fmaf(d, d, ss[r]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < RPW; ++r) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = (i * 32 + lane) * VEC;
    float g[VEC], bt[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c + k));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c + k));
      g[k] = g4.x; g[k + 1] = g4.y; g[k + 2] = g4.z; g[k + 3] = g4.w;
      bt[k] = b4.x; bt[k + 1] = b4.y; bt[k + 2] = b4.z; bt[k + 3] = b4.w;
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (row0 + r >= rows) continue;
      const float rstd = rsqrtf(fmaf(ss[r], inv_n, eps));
      float o[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = (v[r][i][k] - s[r]) * rstd * g[k] + bt[k];
      if (GN) {
        const int64_t row = row0 + r;
        const uint32_t row32 = (uint32_t)row;                       // the launcher checks rows < 2^31
        const uint32_t b = row32 / (uint32_t)gn.Lv;
        const int t = (int)(row32 - b * (uint32_t)gn.Lv);
        int l = 0;
        while (l + 1 < gn.L && t >= gn.lv.start[l + 1]) ++l;
        const int cpg = N / gn.G;                                   // VEC <= cpg: the vector lies in one group
        const float inv_cnt = __frcp_rn((float)(gn.lv.H[l] * gn.lv.W[l] * cpg));
        const float2 st = __ldg(reinterpret_cast<const float2*>(gn.stats) + ((b * gn.L + l) * gn.G + c / cpg));
        const float mean = st.x * inv_cnt;
        const float grstd = rsqrtf(fmaxf(st.y * inv_cnt - mean * mean, 0.f) + gn.eps);
        float cv[VEC], sk[VEC], gg[VEC], gb[VEC];
        Vec16<T>::unpack(pre_cv[i], cv);        // RPW == 1 in this variant: row == row0
        Vec16<T>::unpack(pre_sk[i], sk);
#pragma unroll
        for (int k = 0; k < VEC; k += 4) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(gn.gamma + l * N + c + k));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(gn.beta + l * N + c + k));
          gg[k] = g4.x * grstd; gg[k + 1] = g4.y * grstd; gg[k + 2] = g4.z * grstd; gg[k + 3] = g4.w * grstd;
          gb[k] = b4.x; gb[k + 1] = b4.y; gb[k + 2] = b4.z; gb[k + 3] = b4.w;
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] += gelu_erf(fmaf(cv[k] - mean, gg[k], gb[k])) + sk[k];
      } else if (post_add) {
        float p[VEC];
        Vec16<T>::load(post_add + (row0 + r) * N + c, p);
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] += p[k];
      }
      Vec16<T>::store(y + (row0 + r) * N + c, o);
    }
  }
}

// ---- 3x3 conv on tokens, SIMT reference: one thread per (pixel, 4 output channels) ---------------------------------
// w: [L][9][Cout][Cin] (tap = ky*3 + kx), the packing emrt_pack_conv3x3_weight produces (TW = float or bf16).
template <typename T, typename TW>
__global__ void __launch_bounds__(256)
conv3x3_tokens_simt_kernel(const T* __restrict__ x, const TW* __restrict__ w, T* __restrict__ y, int B, int Lv, int C,
                           int L, const __grid_constant__ LevelTable lv) {
  const int cq = C / 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * Lv * cq) return;
  const int co = (int)(idx % cq) * 4;
  const int64_t tok = idx / cq;
  const int t = (int)(tok % Lv);
  const int64_t b = tok / Lv;
  int l = 0;
  while (l + 1 < L && t >= lv.start[l + 1]) ++l;
  const int H = lv.H[l], W = lv.W[l];
  const int py = (t - lv.start[l]) / W, px = (t - lv.start[l]) % W;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const T* xp = x + (b * Lv + lv.start[l] + yy * W + xx) * C;
      const TW* wp = w + (((int64_t)l * 9 + ky * 3 + kx) * C + co) * C;
      for (int ci = 0; ci < C; ++ci) {
        const float xv = to_float(xp[ci]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv, to_float(wp[(int64_t)j * C + ci]), acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) y[tok * C + co + j] = from_float<T>(acc[j]);
}

// dst[l][tap][co][ci] = src_l[co][ci][ky][kx] (Paddle Conv2D weight [Cout, Cin, 3, 3]) for one level
template <typename TD>
__global__ void pack_conv3x3_kernel(const float* __restrict__ src, TD* __restrict__ dst, int C) {
  const int64_t n = (int64_t)9 * C * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C);
    const int co = (int)((i / C) % C);
    const int tap = (int)(i / ((int64_t)C * C));
    dst[i] = from_float<TD>(src[((int64_t)co * C + ci) * 9 + tap]);
  }
}

// ---- GroupNorm statistics: sum / sum of squares per (batch, level, group) ----------------------------------------------
// grid (B * L, GN_SPLITS); block 256 threads = (C / 8 groups-of-8-channels lanes) x pixel slots.  DETERMINISTIC: every
// CTA writes its partial sums, the last CTA to finish a (batch, level) — found with a ticket counter — adds the GN_SPLITS
// partials in split order, so the result does not depend on scheduling (no floating-point atomics; a run is bit-reproducible).
// Workspace (floats): stats [B*L, G, 2] | partials [B*L, GN_SPLITS, G, 2] | tickets [B*L] (zeroed by the launcher).
constexpr int GN_SPLITS = 16;

inline int64_t gn_workspace_floats(int64_t BL, int G) { return 2 * BL * G * (1 + GN_SPLITS) + BL; }

template <typename T>
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const T* __restrict__ x, float* __restrict__ ws, int Lv, int C, int L, int G,
                       const __grid_constant__ LevelTable lv) {
  const int bl = blockIdx.x, l = bl % L;
  const int64_t b = bl / L;
  const int64_t BL = gridDim.x;
  float* stats = ws;
  float* partials = ws + 2 * BL * G;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(ws + 2 * BL * G * (1 + GN_SPLITS));
  const int npix = lv.H[l] * lv.W[l];
  const int vec_per_pix = C / 8;                           // 16-byte vectors of 8 channels (bf16) or 2 x float4
  const int slots = 256 / vec_per_pix;                     // pixels handled concurrently by the CTA
  const int v = threadIdx.x % vec_per_pix, slot = threadIdx.x / vec_per_pix;
  const int cpg = C / G;                                   // channels per group (8 for EMRT)
  float s = 0.f, ss = 0.f;
  if (slot < slots) {
    constexpr int VEC = Vec16<T>::N;
    // four pixels per trip, their loads in flight together, each with its own partial sums (combined in a fixed order:
    // the result stays independent of scheduling)
    const T* xbase = x + ((b * Lv + lv.start[l]) * C) + v * 8;
    const int step = gridDim.y * slots;
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, ss4[4] = {0.f, 0.f, 0.f, 0.f};
    int p = blockIdx.y * slots + slot;
    for (; p + 3 * step < npix; p += 4 * step) {
      uint4 raw[4][8 / VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8 / VEC; ++j) raw[u][j] = Vec16<T>::load_raw(xbase + (int64_t)(p + u * step) * C + j * VEC);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8 / VEC; ++j) {
          float a[VEC];
          Vec16<T>::unpack(raw[u][j], a);
#pragma unroll
          for (int k = 0; k < VEC; ++k) { s4[u] += a[k]; ss4[u] = fmaf(a[k], a[k], ss4[u]); }
        }
    }
    for (; p < npix; p += step) {
#pragma unroll
      for (int j = 0; j < 8 / VEC; ++j) {
        float a[VEC];
        Vec16<T>::load(xbase + (int64_t)p * C + j * VEC, a);
#pragma unroll
        for (int k = 0; k < VEC; ++k) { s4[0] += a[k]; ss4[0] = fmaf(a[k], a[k], ss4[0]); }
      }
    }
    s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    ss = (ss4[0] + ss4[1]) + (ss4[2] + ss4[3]);
  }
  // reduce over the CTA's pixel slots and over the vectors of one group, in a fixed order
  __shared__ float sh[256][2];
  __shared__ bool last;
  sh[threadIdx.x][0] = s;
  sh[threadIdx.x][1] = ss;
  __syncthreads();
  const int vpg = cpg >= 8 ? cpg / 8 : 1;                  // 8-channel vectors per group
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float gs = 0.f, gss = 0.f;
    for (int k = 0; k < slots; ++k)
      for (int j = 0; j < vpg; ++j) { gs += sh[k * vec_per_pix + g * vpg + j][0]; gss += sh[k * vec_per_pix + g * vpg + j][1]; }
    float* dst = partials + (((int64_t)bl * GN_SPLITS + blockIdx.y) * G + g) * 2;
    dst[0] = gs;
    dst[1] = gss;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + bl, 1u) == gridDim.y - 1;
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      float gs = 0.f, gss = 0.f;
      for (int k = 0; k < (int)gridDim.y; ++k) {
        const float* src = partials + (((int64_t)bl * GN_SPLITS + k) * G + g) * 2;
        gs += __ldcg(src);
        gss += __ldcg(src + 1);
      }
      stats[((int64_t)bl * G + g) * 2] = gs;
      stats[((int64_t)bl * G + g) * 2 + 1] = gss;
    }
    if (threadIdx.x == 0) tickets[bl] = 0u;                // ready for the next use of this workspace
  }
}

// y = gelu((c - mean) * rstd * gamma_l + beta_l) + x   (exact erf GELU, nn.GELU default)
// grid (ceil(Lv * C / VEC / 256), B): 32-bit index math, one (mean, rstd) per 16-byte vector (VEC <= channels per group)
template <typename T>
__global__ void __launch_bounds__(256)
groupnorm_gelu_residual_kernel(const T* __restrict__ conv, const T* __restrict__ x, const float* __restrict__ stats,
                               const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
                               int Lv, int C, int L, int G, float eps, const __grid_constant__ LevelTable lv) {
  constexpr int VEC = Vec16<T>::N;
  const uint32_t vec_per_tok = (uint32_t)C / VEC;
  const uint32_t vi = blockIdx.x * 256u + threadIdx.x;           // vector index inside this image
  if (vi >= (uint32_t)Lv * vec_per_tok) return;
  const uint32_t t = vi / vec_per_tok;
  const int c0 = (int)(vi - t * vec_per_tok) * VEC;
  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < L && (int)t >= lv.start[l + 1]) ++l;
  const int cpg = C / G;
  const float inv_cnt = 1.f / (float)(lv.H[l] * lv.W[l] * cpg);
  const float* st = stats + ((b * L + l) * G + c0 / cpg) * 2;   // VEC <= cpg: the whole vector lies in one group
  const float mean = __ldg(st) * inv_cnt;
  const float rstd = rsqrtf(fmaxf(__ldg(st + 1) * inv_cnt - mean * mean, 0.f) + eps);
  const int64_t off = ((int64_t)b * Lv * vec_per_tok + vi) * VEC;
  float cv[VEC], xv[VEC], o[VEC], g[VEC], bt[VEC];
  Vec16<T>::load(conv + off, cv);
  Vec16<T>::load(x + off, xv);
#pragma unroll
  for (int k = 0; k < VEC; k += 4) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + l * C + c0 + k));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + l * C + c0 + k));
    g[k] = g4.x; g[k + 1] = g4.y; g[k + 2] = g4.z; g[k + 3] = g4.w;
    bt[k] = b4.x; bt[k + 1] = b4.y; bt[k + 2] = b4.z; bt[k + 3] = b4.w;
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float h = (cv[k] - mean) * rstd * g[k] + bt[k];
    o[k] = gelu_erf(h) + xv[k];
  }
  Vec16<T>::store(y + off, o);
}


// ---- input_proj glue (EncoderDecoder.forward, transformer_encoder_decoder.py:417-436) ---------------------------------
// y[b, p, c] = x[b, c, p]: backbone feature maps / PSP tokens arrive NCHW ([B, C, H*W]); the path works on tokens.
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_tokens_kernel(const T* __restrict__ x, T* __restrict__ y, int C, int P) {
  __shared__ T tile[32][33];
  const int64_t b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < P) tile[i][tx] = x[(b * C + c) * P + p];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (c < C && p < P) y[(b * P + p) * C + c] = tile[tx][i];
  }
}

// bf16, P % 64 == 0, C % 64 == 0: 64 x 64 tiles moved with 16-byte global accesses on both sides (full 128-byte lines:
// a warp reads 4 channel rows x 64 pixels and writes 4 pixel rows x 64 channels); the transposition itself happens on
// 2-byte shared-memory reads (row pitch 33 words: conflict-free 32-bit stores, 2-way conflicts on the 16-bit loads).
__global__ void __launch_bounds__(256)
nchw_to_tokens_bf16_64_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int P) {
  __shared__ uint32_t tile[64 * 33];                         // [c][p / 2], pitch 33 words
  const int64_t b = blockIdx.z;
  const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int r = threadIdx.x >> 3, ch = threadIdx.x & 7;      // 32 rows x 8 16-byte chunks per pass
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = r + 32 * h;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (b * C + c0 + c) * P + p0 + ch * 8));
    uint32_t* dst = tile + c * 33 + ch * 4;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
  const unsigned short* t16 = reinterpret_cast<const unsigned short*>(tile);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pp = r + 32 * h;                               // pixel row of the output
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = t16[(ch * 8 + 2 * j) * 66 + pp], hi = t16[(ch * 8 + 2 * j + 1) * 66 + pp];
      w[j] = lo | (hi << 16);
    }
    *reinterpret_cast<uint4*>(y + (b * P + p0 + pp) * C + c0 + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// GroupNorm(groups, C) of one level's tokens x [B, P, C] (no activation), written into the level's slot of the
// concatenated token tensor: y + b * y_batch_stride + p * C.  Statistics come from groupnorm_stats_kernel (L = 1).
template <typename T>
__global__ void __launch_bounds__(256)
groupnorm_tokens_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                        const float* __restrict__ beta, T* __restrict__ y, int64_t y_batch_stride, int P, int C, int G,
                        float eps) {
  constexpr int VEC = Vec16<T>::N;
  const uint32_t vec_per_tok = (uint32_t)C / VEC;
  const uint32_t vi = blockIdx.x * 256u + threadIdx.x;
  if (vi >= (uint32_t)P * vec_per_tok) return;
  const uint32_t t = vi / vec_per_tok;
  const int c0 = (int)(vi - t * vec_per_tok) * VEC;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const float inv_cnt = 1.f / (float)(P * cpg);
  const float* st = stats + (b * G + c0 / cpg) * 2;
  const float mean = __ldg(st) * inv_cnt;
  const float rstd = rsqrtf(fmaxf(__ldg(st + 1) * inv_cnt - mean * mean, 0.f) + eps);
  float v[VEC], o[VEC];
  Vec16<T>::load(x + ((int64_t)b * P * vec_per_tok + vi) * VEC, v);
#pragma unroll
  for (int k = 0; k < VEC; ++k) o[k] = (v[k] - mean) * rstd * __ldg(gamma + c0 + k) + __ldg(beta + c0 + k);
  Vec16<T>::store(y + (int64_t)b * y_batch_stride + (int64_t)t * C + c0, o);
}

// Same, for C / VEC dividing 256: a thread keeps one 16-byte channel slice (its scale / shift computed once from the
// group statistics, gamma, beta) and walks GN_TOK_PER_CTA / (256 / (C / VEC)) tokens, four loads in flight — the
// one-vector-per-thread kernel above spends most of its time on the per-thread setup (1.9 TB/s on 302 MB).
constexpr int GN_TOK_PER_CTA = 128;
template <typename T>
__global__ void __launch_bounds__(256)
groupnorm_tokens_rows_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                             const float* __restrict__ beta, T* __restrict__ y, int64_t y_batch_stride, int P, int C, int G,
                             float eps) {
  constexpr int VEC = Vec16<T>::N;
  const int vec_per_tok = C / VEC, tok_lanes = 256 / vec_per_tok;
  const int cv = threadIdx.x % vec_per_tok, tl = threadIdx.x / vec_per_tok;
  const int c0 = cv * VEC, b = blockIdx.y, cpg = C / G;
  const float inv_cnt = 1.f / (float)(P * cpg);
  const float* st = stats + (b * G + c0 / cpg) * 2;
  const float mean = __ldg(st) * inv_cnt;
  const float rstd = rsqrtf(fmaxf(__ldg(st + 1) * inv_cnt - mean * mean, 0.f) + eps);
  float gm[VEC], bt[VEC];
#pragma unroll
  for (int k = 0; k < VEC; k += 4) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + k));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c0 + k));
    gm[k] = g4.x; gm[k + 1] = g4.y; gm[k + 2] = g4.z; gm[k + 3] = g4.w;
    bt[k] = b4.x; bt[k + 1] = b4.y; bt[k + 2] = b4.z; bt[k + 3] = b4.w;
  }
  const int t_end = min(P, (int)(blockIdx.x + 1) * GN_TOK_PER_CTA);
  const T* xb = x + (int64_t)b * P * C + c0;
  T* yb = y + (int64_t)b * y_batch_stride + c0;
  int t = blockIdx.x * GN_TOK_PER_CTA + tl;
  for (; t + 3 * tok_lanes < t_end; t += 4 * tok_lanes) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = Vec16<T>::load_raw(xb + (int64_t)(t + u * tok_lanes) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[VEC], o[VEC];
      Vec16<T>::unpack(raw[u], v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = (v[k] - mean) * rstd * gm[k] + bt[k];
      Vec16<T>::store(yb + (int64_t)(t + u * tok_lanes) * C, o);
    }
  }
  for (; t < t_end; t += tok_lanes) {
    float v[VEC], o[VEC];
    Vec16<T>::load(xb + (int64_t)t * C, v);
#pragma unroll
    for (int k = 0; k < VEC; ++k) o[k] = (v[k] - mean) * rstd * gm[k] + bt[k];
    Vec16<T>::store(yb + (int64_t)t * C, o);
  }
}

}  // namespace emrt

using namespace emrt;

namespace emrt {
int64_t conv3x3_stats_workspace_floats(int B, int Lv, int L);
int conv3x3_tokens_tc(const void* x, const void* w_packed, void* y, int B, int Lv, int C, int L, const LevelTable& lv,
                      cudaStream_t st, float* stats_ws = nullptr, int max_ctas = 0);
}

static int launch_residual_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                                     const void* post_add, void* y, int64_t rows, int N, float eps, int dtype,
                                     const GnBranch* gnb, void* stream);

extern "C" int emrt_residual_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                                       const void* post_add, void* y, int64_t rows, int N, float eps, int dtype,
                                       void* stream) {
  return launch_residual_layernorm(x, residual, gamma, beta, post_add, y, rows, N, eps, dtype, nullptr, stream);
}

static int check_gn_args(int B, int C, int groups) {
  EMRT_REQUIRE(B > 0 && C > 0 && groups > 0 && C % groups == 0 && C % 8 == 0 && C <= 2048, "bad channel / group counts");
  const int cpg = C / groups;
  EMRT_REQUIRE(cpg % 8 == 0 || 8 % cpg == 0, "channels per group must divide or be a multiple of 8");
  if (8 % cpg == 0 && cpg != 8) return set_error(EMRT_ERR_UNSUPPORTED, "channels per group < 8 not supported");
  return EMRT_OK;
}

extern "C" long long emrt_groupnorm_workspace_floats(int B, int L, int groups) {
  return gn_workspace_floats((int64_t)B * L, groups);
}

extern "C" int emrt_groupnorm_stats(const void* x, float* stats, int B, int Lv, int C, int L, int groups,
                                    const int32_t* shapes_hw_host, int dtype, void* stream) {
  EMRT_REQUIRE(x && stats, "NULL pointer");
  if (int e = check_gn_args(B, C, groups)) return e;
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, Lv)) return e;
  cudaStream_t st = as_stream(stream);
  EMRT_REQUIRE(groups <= 256, "at most 256 groups");
  EMRT_CUDA_CHECK(cudaMemsetAsync(stats + 2 * (int64_t)B * L * groups * (1 + GN_SPLITS), 0, sizeof(float) * B * L, st));
  dim3 sgrid((unsigned)(B * L), GN_SPLITS);
  if (dtype == EMRT_F32) groupnorm_stats_kernel<float><<<sgrid, 256, 0, st>>>((const float*)x, stats, Lv, C, L, groups, lv);
  else if (dtype == EMRT_BF16) groupnorm_stats_kernel<__nv_bfloat16><<<sgrid, 256, 0, st>>>((const __nv_bfloat16*)x, stats, Lv, C, L, groups, lv);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_residual_layernorm_gn(const void* x, const void* residual, const float* ln_gamma, const float* ln_beta,
                                          const void* conv, const void* skip, const float* gn_stats, const float* gn_gamma,
                                          const float* gn_beta, void* y, int B, int Lv, int C, int L, int groups,
                                          float ln_eps, float gn_eps, const int32_t* shapes_hw_host, int dtype,
                                          void* stream) {
  EMRT_REQUIRE(conv && skip && gn_stats && gn_gamma && gn_beta, "NULL pointer");
  if (int e = check_gn_args(B, C, groups)) return e;
  GnBranch g;
  if (int e = fill_levels(g.lv, L, shapes_hw_host, nullptr, Lv)) return e;
  g.conv = conv; g.skip = skip; g.stats = gn_stats; g.gamma = gn_gamma; g.beta = gn_beta;
  g.Lv = Lv; g.L = L; g.G = groups; g.eps = gn_eps;
  EMRT_REQUIRE((int64_t)B * Lv < (1LL << 31), "B * Lv must fit 31 bits");
  return launch_residual_layernorm(x, residual, ln_gamma, ln_beta, nullptr, y, (int64_t)B * Lv, C, ln_eps, dtype, &g, stream);
}

static int launch_residual_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                                     const void* post_add, void* y, int64_t rows, int N, float eps, int dtype,
                                     const GnBranch* gnb, void* stream) {
  EMRT_REQUIRE(x && gamma && beta && y && rows > 0, "bad residual_layernorm arguments");
  const int vec = dtype == EMRT_F32 ? 4 : 8;
  EMRT_REQUIRE(N > 0 && N % (32 * vec) == 0 && N / (32 * vec) <= 8, "N must be a multiple of 32 16-byte vectors, at most 8 per lane");
  cudaStream_t st = as_stream(stream);
  const int per = N / (32 * vec);
  const int64_t warps = (per <= 2 && !gnb) ? (rows + 1) / 2 : rows;       // RPW rows per warp
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  GnBranch gn_arg;
  memset(&gn_arg, 0, sizeof(gn_arg));
  if (gnb) gn_arg = *gnb;
#define EMRT_LN(T, PER)                                                                                                   \
  if (gnb) residual_layernorm_kernel<T, PER, true><<<blocks, 256, 0, st>>>((const T*)x, (const T*)residual, gamma, beta, (const T*)post_add, (T*)y, rows, N, eps, gn_arg); \
  else residual_layernorm_kernel<T, PER, false><<<blocks, 256, 0, st>>>((const T*)x, (const T*)residual, gamma, beta, (const T*)post_add, (T*)y, rows, N, eps, gn_arg)
#define EMRT_LN_PER(T)                                                                       \
  switch (per) {                                                                               \
    case 1: { EMRT_LN(T, 1); } break; case 2: { EMRT_LN(T, 2); } break; case 3: { EMRT_LN(T, 3); } break;  \
    case 4: { EMRT_LN(T, 4); } break; case 5: { EMRT_LN(T, 5); } break; case 6: { EMRT_LN(T, 6); } break;  \
    case 7: { EMRT_LN(T, 7); } break; default: { EMRT_LN(T, 8); } break;                               \
  }
  if (dtype == EMRT_F32) { EMRT_LN_PER(float) }
  else if (dtype == EMRT_BF16) { EMRT_LN_PER(__nv_bfloat16) }
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
#undef EMRT_LN_PER
#undef EMRT_LN
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_pack_conv3x3_weight(const float* src, void* dst, int C, int level, int dst_dtype, void* stream) {
  EMRT_REQUIRE(src && dst && C > 0 && level >= 0, "bad pack_conv3x3_weight arguments");
  const int64_t n = (int64_t)9 * C * C;
  cudaStream_t st = as_stream(stream);
  if (dst_dtype == EMRT_F32) pack_conv3x3_kernel<float><<<256, 256, 0, st>>>(src, (float*)dst + level * n, C);
  else if (dst_dtype == EMRT_BF16) pack_conv3x3_kernel<__nv_bfloat16><<<256, 256, 0, st>>>(src, (__nv_bfloat16*)dst + level * n, C);
  else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dst_dtype %d", dst_dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_conv3x3_tokens_fwd(const void* x, const void* w_packed, void* y, int B, int Lv, int C, int L,
                                       const int32_t* shapes_hw_host, int dtype, int w_dtype, int impl, void* stream) {
  EMRT_REQUIRE(x && w_packed && y && B > 0 && C > 0 && C % 8 == 0, "bad conv3x3_tokens arguments");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, Lv)) return e;
  cudaStream_t st = as_stream(stream);
  if (impl != 1 && dtype == EMRT_BF16 && w_dtype == EMRT_BF16) {
    const int e = conv3x3_tokens_tc(x, w_packed, y, B, Lv, C, L, lv, st);
    if (e != EMRT_ERR_UNSUPPORTED) return e;
    if (impl == 2) return set_error(EMRT_ERR_UNSUPPORTED, "tcgen05 conv3x3 does not support this shape");
  }
  const int64_t total = (int64_t)B * Lv * (C / 4);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == EMRT_F32 && w_dtype == EMRT_F32)
    conv3x3_tokens_simt_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)x, (const float*)w_packed, (float*)y, B, Lv, C, L, lv);
  else if (dtype == EMRT_BF16 && w_dtype == EMRT_BF16)
    conv3x3_tokens_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed, (__nv_bfloat16*)y, B, Lv, C, L, lv);
  else return set_error(EMRT_ERR_UNSUPPORTED, "conv3x3_tokens: x / w dtypes must both be F32 or both BF16");
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" long long emrt_conv3x3_stats_workspace_floats(int B, int Lv, int L) { return conv3x3_stats_workspace_floats(B, Lv, L); }

extern "C" int emrt_conv3x3_tokens_stats_fwd(const void* x, const void* w_packed, void* y, float* stats_workspace, int B, int Lv,
                                             int C, int L, const int32_t* shapes_hw_host, int groups, void* stream) {
  return emrt_conv3x3_tokens_stats_part_fwd(x, w_packed, y, stats_workspace, B, Lv, C, L, shapes_hw_host, groups, 0, stream);
}

extern "C" int emrt_conv3x3_tokens_stats_part_fwd(const void* x, const void* w_packed, void* y, float* stats_workspace, int B,
                                                  int Lv, int C, int L, const int32_t* shapes_hw_host, int groups, int max_ctas,
                                                  void* stream) {
  EMRT_REQUIRE(x && w_packed && y && stats_workspace && B > 0 && C > 0 && max_ctas >= 0, "bad conv3x3_tokens_stats arguments");
  if (groups != 32 || C != 256) return set_error(EMRT_ERR_UNSUPPORTED, "conv3x3 + GroupNorm statistics is built for C = 256, 32 groups");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, Lv)) return e;
  const int e = conv3x3_tokens_tc(x, w_packed, y, B, Lv, C, L, lv, as_stream(stream), stats_workspace, max_ctas);
  if (e == EMRT_ERR_UNSUPPORTED) return set_error(EMRT_ERR_UNSUPPORTED, "tcgen05 conv3x3 does not tile this shape (use emrt_conv3x3_tokens_fwd + emrt_groupnorm_stats)");
  return e;
}

extern "C" int emrt_groupnorm_gelu_residual(const void* conv, const void* x, const float* gamma, const float* beta,
                                            void* y, float* stats_workspace, int B, int Lv, int C, int L, int groups,
                                            float eps, const int32_t* shapes_hw_host, int dtype, void* stream) {
  EMRT_REQUIRE(conv && x && gamma && beta && y && stats_workspace, "NULL pointer");
  EMRT_REQUIRE(B > 0 && C > 0 && groups > 0 && C % groups == 0 && C % 8 == 0 && C <= 2048, "bad channel / group counts");
  const int cpg = C / groups;
  EMRT_REQUIRE(cpg % 8 == 0 || 8 % cpg == 0, "channels per group must divide or be a multiple of 8");
  if (8 % cpg == 0 && cpg != 8) return set_error(EMRT_ERR_UNSUPPORTED, "channels per group < 8 not supported");
  LevelTable lv;
  if (int e = fill_levels(lv, L, shapes_hw_host, nullptr, Lv)) return e;
  cudaStream_t st = as_stream(stream);
  EMRT_REQUIRE(groups <= 256, "at most 256 groups");
  EMRT_CUDA_CHECK(cudaMemsetAsync(stats_workspace + 2 * (int64_t)B * L * groups * (1 + GN_SPLITS), 0, sizeof(float) * B * L, st));
  dim3 sgrid((unsigned)(B * L), GN_SPLITS);
  const int vec = dtype == EMRT_F32 ? 4 : 8;
  EMRT_REQUIRE((int64_t)Lv * (C / vec) < (1LL << 31) && B <= 65535, "image too large for the GroupNorm apply grid");
  dim3 agrid((unsigned)(((int64_t)Lv * (C / vec) + 255) / 256), (unsigned)B);
  if (dtype == EMRT_F32) {
    groupnorm_stats_kernel<float><<<sgrid, 256, 0, st>>>((const float*)conv, stats_workspace, Lv, C, L, groups, lv);
    count_launch();
    groupnorm_gelu_residual_kernel<float><<<agrid, 256, 0, st>>>((const float*)conv, (const float*)x, stats_workspace, gamma, beta, (float*)y, Lv, C, L, groups, eps, lv);
  } else if (dtype == EMRT_BF16) {
    groupnorm_stats_kernel<__nv_bfloat16><<<sgrid, 256, 0, st>>>((const __nv_bfloat16*)conv, stats_workspace, Lv, C, L, groups, lv);
    count_launch();
    groupnorm_gelu_residual_kernel<__nv_bfloat16><<<agrid, 256, 0, st>>>((const __nv_bfloat16*)conv, (const __nv_bfloat16*)x, stats_workspace, gamma, beta, (__nv_bfloat16*)y, Lv, C, L, groups, eps, lv);
  } else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_nchw_to_tokens(const void* x, void* y, int B, int C, int P, int dtype, void* stream) {
  EMRT_REQUIRE(x && y && B > 0 && C > 0 && P > 0 && B <= 65535, "bad nchw_to_tokens arguments");
  dim3 grid((P + 31) / 32, (C + 31) / 32, B);
  cudaStream_t st = as_stream(stream);
  if (dtype == EMRT_F32) nchw_to_tokens_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, C, P);
  else if (dtype == EMRT_BF16) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (P % 64 == 0 && C % 64 == 0 && aligned && P / 64 <= 65535 && C / 64 <= 65535)
      nchw_to_tokens_bf16_64_kernel<<<dim3(P / 64, C / 64, B), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, C, P);
    else
      nchw_to_tokens_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, C, P);
  } else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

extern "C" int emrt_groupnorm_tokens(const void* x, const float* gamma, const float* beta, void* y,
                                     int64_t y_batch_stride, float* stats_workspace, int B, int P, int C, int groups,
                                     float eps, int dtype, void* stream) {
  EMRT_REQUIRE(x && gamma && beta && y && stats_workspace, "NULL pointer");
  EMRT_REQUIRE(B > 0 && B <= 65535 && P > 0 && C > 0 && groups > 0 && C % groups == 0 && (C / groups) % 8 == 0 && C <= 2048,
               "bad shape: channels per group must be a multiple of 8");
  LevelTable lv;
  const int32_t hw[2] = {P, 1};
  if (int e = fill_levels(lv, 1, hw, nullptr, P)) return e;
  cudaStream_t st = as_stream(stream);
  EMRT_REQUIRE(groups <= 256, "at most 256 groups");
  EMRT_CUDA_CHECK(cudaMemsetAsync(stats_workspace + 2 * (int64_t)B * groups * (1 + GN_SPLITS), 0, sizeof(float) * B, st));
  dim3 sgrid((unsigned)B, GN_SPLITS);
  const int vec = dtype == EMRT_F32 ? 4 : 8;
  dim3 agrid((unsigned)(((int64_t)P * (C / vec) + 255) / 256), (unsigned)B);
  // the row-walking apply kernel needs C / vec to divide 256 and 16-byte aligned gamma / beta
  const bool rows_ok = (C / vec) <= 256 && 256 % (C / vec) == 0 && ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0 &&
                       !getenv("EMRT_GN_APPLY_SIMPLE");
  dim3 rgrid((unsigned)((P + GN_TOK_PER_CTA - 1) / GN_TOK_PER_CTA), (unsigned)B);
  if (dtype == EMRT_F32) {
    groupnorm_stats_kernel<float><<<sgrid, 256, 0, st>>>((const float*)x, stats_workspace, P, C, 1, groups, lv);
    count_launch();
    if (rows_ok) groupnorm_tokens_rows_kernel<float><<<rgrid, 256, 0, st>>>((const float*)x, stats_workspace, gamma, beta, (float*)y, y_batch_stride, P, C, groups, eps);
    else groupnorm_tokens_kernel<float><<<agrid, 256, 0, st>>>((const float*)x, stats_workspace, gamma, beta, (float*)y, y_batch_stride, P, C, groups, eps);
  } else if (dtype == EMRT_BF16) {
    groupnorm_stats_kernel<__nv_bfloat16><<<sgrid, 256, 0, st>>>((const __nv_bfloat16*)x, stats_workspace, P, C, 1, groups, lv);
    count_launch();
    if (rows_ok) groupnorm_tokens_rows_kernel<__nv_bfloat16><<<rgrid, 256, 0, st>>>((const __nv_bfloat16*)x, stats_workspace, gamma, beta, (__nv_bfloat16*)y, y_batch_stride, P, C, groups, eps);
    else groupnorm_tokens_kernel<__nv_bfloat16><<<agrid, 256, 0, st>>>((const __nv_bfloat16*)x, stats_workspace, gamma, beta, (__nv_bfloat16*)y, y_batch_stride, P, C, groups, eps);
  } else return set_error(EMRT_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}
