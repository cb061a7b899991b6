// Shared helpers for the emrt_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/emrt_b200.h"

namespace emrt {

// ---- error plumbing -------------------------------------------------------------------------------------
char* last_error_buffer();                 // thread-local, 512 bytes
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define EMRT_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) return ::emrt::set_error(EMRT_ERR_INVALID_ARGUMENT, __VA_ARGS__);         \
  } while (0)

#define EMRT_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::emrt::set_error(EMRT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define EMRT_LAUNCH_CHECK()                                                                \
  do {                                                                                     \
    ::emrt::count_launch();                                                                \
    EMRT_CUDA_CHECK(cudaGetLastError());                                                   \
  } while (0)

// ---- level tables passed by value into kernels -----------------------------------------------------------
struct LevelTable {
  int32_t H[EMRT_MAX_LEVELS];
  int32_t W[EMRT_MAX_LEVELS];
  int32_t start[EMRT_MAX_LEVELS];
};

inline int fill_levels(LevelTable& t, int L, const int32_t* shapes_hw, const int32_t* level_start, int Lv) {
  if (L < 1 || L > EMRT_MAX_LEVELS) return set_error(EMRT_ERR_INVALID_ARGUMENT, "L=%d out of [1,%d]", L, EMRT_MAX_LEVELS);
  if (!shapes_hw) return set_error(EMRT_ERR_INVALID_ARGUMENT, "shapes_hw_host is NULL");
  int64_t acc = 0;
  for (int l = 0; l < L; ++l) {
    t.H[l] = shapes_hw[2 * l];
    t.W[l] = shapes_hw[2 * l + 1];
    if (t.H[l] <= 0 || t.W[l] <= 0) return set_error(EMRT_ERR_INVALID_ARGUMENT, "level %d has non-positive shape", l);
    t.start[l] = level_start ? level_start[l] : (int32_t)acc;
    if (level_start && level_start[l] != acc)
      return set_error(EMRT_ERR_INVALID_ARGUMENT, "level_start[%d]=%d but prefix sum of shapes is %lld", l, level_start[l], (long long)acc);
    acc += (int64_t)t.H[l] * t.W[l];
  }
  // the reference asserts sum(h*w) == Len_v (transformer_encoder_decoder.py:81)
  if (Lv >= 0 && acc != Lv) return set_error(EMRT_ERR_INVALID_ARGUMENT, "sum(H*W)=%lld != Lv=%d", (long long)acc, Lv);
  for (int l = L; l < EMRT_MAX_LEVELS; ++l) t.H[l] = t.W[l] = t.start[l] = 0;
  return EMRT_OK;
}

// ---- dtype helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// 16-byte vector of T, unpacked to fp32
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  // the 16 bytes as loaded (to keep a prefetched vector in 4 registers) and their later conversion
  __device__ static __forceinline__ uint4 load_raw(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
  }
};

// Exact (erf-form) GELU, nn.GELU's default: h * Phi(h), Phi(h) = erfc(-h / sqrt 2) / 2.  erfc(z), z >= 0, by Abramowitz &
// Stegun 7.1.26 (|error| <= 1.5e-7 absolute) — one branch-free sequence with one MUFU.RCP and one MUFU.EX2 instead of
// erff's two divergent branches; the negative side is evaluated as erfc directly, so there is no 1 - erf cancellation.
__device__ __forceinline__ float gelu_erf(float h) {
  const float z = fabsf(h) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));        // 1 ulp: far below 1.5e-7
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * (z * -1.4426950408889634f)));   // exp(-z^2)
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float hh = h * (0.5f * p * t * e);                    // h * erfc(z) / 2
  return h >= 0.f ? h - hh : hh;
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction and lane) ---------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2_t pk2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// gelu_erf of two values, the same operations in the same order (bit-equal per element), 21 instructions per pair instead of 34
__device__ __forceinline__ f32x2_t gelu_erf2(f32x2_t h) {
  float h0, h1;
  upk2(h, h0, h1);
  const f32x2_t z = pk2(fabsf(h0) * 0.70710678118654752f, fabsf(h1) * 0.70710678118654752f);
  float d0, d1, t0, t1, a0, a1, e0, e1;
  upk2(fma2(pk2(0.3275911f), z, pk2(1.f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  upk2(mul2(z, mul2(z, pk2(-1.4426950408889634f))), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f32x2_t t = pk2(t0, t1);
  f32x2_t p = fma2(t, pk2(1.061405429f), pk2(-1.453152027f));
  p = fma2(t, p, pk2(1.421413741f));
  p = fma2(t, p, pk2(-0.284496736f));
  p = fma2(t, p, pk2(0.254829592f));
  float r0, r1;
  upk2(mul2(h, mul2(mul2(mul2(pk2(0.5f), p), t), pk2(e0, e1))), r0, r1);          // h * erfc(z) / 2
  return pk2(h0 >= 0.f ? h0 - r0 : r0, h1 >= 0.f ? h1 - r1 : r1);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int num_sms() {
  // per device ordinal (a process may drive several GPUs); benign race: every thread writes the same value
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  if (n[slot] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[slot] = v > 0 ? v : 148;
  }
  return n[slot];
}

}  // namespace emrt
