// Forward sampling-gather, specialised for EMRT's configuration (bf16 values, head dim 32, L*P = 18 points).
//
// ncu on the generic kernel (profiles/r1_gather_v0.md) shows it is INSTRUCTION-ISSUE bound (issue slots 82 % busy,
// 171 instructions per point per lane, DRAM 6 %): the four lanes that share an item each recompute the same
// bilinear footprint, and 64-bit address arithmetic + bf16 unpacking dominate.  This kernel therefore
//   1. computes every footprint ONCE (one lane per (query, point)), pre-multiplies the four bilinear weights by
//      the attention weight, and stages {4 byte offsets, 4 weights} in shared memory;
//   2. runs a branch-free inner loop per point: 2 LDS.128 + 4 LDG.128 (bf16x8) + unpack + packed FFMA2;
//   3. maps a warp to 8 CONSECUTIVE QUERIES OF ONE HEAD, so with the head-major value layout [B,M,Lv,D] the eight
//      64-byte corner fetches of one LDG fall into ~4 neighbouring 128-byte lines instead of 8 scattered ones.
// One CTA = 64 consecutive queries x 1 head (8 warps x 8 queries).
#include "msda_common.cuh"

namespace emrt {

constexpr int V1_D = 32;
constexpr int V1_QPW = 8;          // queries per warp
constexpr int V1_WARPS = 8;
constexpr int V1_QPB = V1_QPW * V1_WARPS;

__device__ __forceinline__ void ffma2(float2& acc, float2 v, const float w) {
  // Blackwell packed fp32 FMA (SASS FFMA2): two channels per issue slot
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
  float2 ww = make_float2(w, w);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(a)
      : "l"(*reinterpret_cast<unsigned long long*>(&v)), "l"(*reinterpret_cast<unsigned long long*>(&ww)));
  acc = *reinterpret_cast<float2*>(&a);
}

template <typename TL, int MODE, bool HEAD_MAJOR, int L, int P>
__global__ void __launch_bounds__(V1_WARPS * 32)
msda_gather_fwd_v1_kernel(const __nv_bfloat16* __restrict__ value, const TL* __restrict__ loc,
                          const TL* __restrict__ attn, const float* __restrict__ ref, int64_t ref_bs,
                          __nv_bfloat16* __restrict__ out, int Lq, int Lv, int M, int chunks,
                          const __grid_constant__ LevelTable lv) {
  constexpr int LP = L * P;
  __shared__ uint4 s_off[V1_WARPS][V1_QPW][LP];
  __shared__ float4 s_w[V1_WARPS][V1_QPW][LP];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = (b * chunks + chunk) * M + m : the M heads of a query chunk run back to back (they share loc rows)
  const int m = blockIdx.x % M;
  const int bc = blockIdx.x / M;
  const int chunk = bc % chunks;
  const int b = bc / chunks;
  const int q0 = chunk * V1_QPB + warp * V1_QPW;

  // bytes between consecutive pixels of one head's plane
  const uint32_t pix_bytes = HEAD_MAJOR ? V1_D * 2 : (uint32_t)M * V1_D * 2;

  // ---- stage A: one footprint per (query, point), computed once ----------------------------------------------
#pragma unroll
  for (int r = 0; r < (V1_QPW * LP + 31) / 32; ++r) {
    const int j = r * 32 + lane;
    if (j < V1_QPW * LP) {
      const int qi = j / LP, pt = j % LP, l = pt / P;
      const int q = q0 + qi;
      uint4 o = make_uint4(0, 0, 0, 0);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < Lq) {
        const int64_t item = ((int64_t)b * Lq + q) * M + m;
        const float2 xy = Pair<TL>::load(loc + (item * LP + pt) * 2);
        const float aw = load1<TL>(attn + item * LP + pt);
        const int H = lv.H[l], W = lv.W[l];
        float x, y;
        if (MODE == EMRT_LOC_PIXEL_OFFSET) {
          const float2 rf = __ldg(reinterpret_cast<const float2*>(ref + b * ref_bs + ((int64_t)q * L + l) * 2));
          x = rf.x * (float)W - 0.5f + xy.x;
          y = rf.y * (float)H - 0.5f + xy.y;
        } else {
          x = xy.x * (float)W - 0.5f;
          y = xy.y * (float)H - 0.5f;
        }
        const Footprint f = make_footprint(x, y, H, W);
        const uint32_t s = (uint32_t)lv.start[l];
        o = make_uint4((s + f.i00) * pix_bytes, (s + f.i01) * pix_bytes, (s + f.i10) * pix_bytes, (s + f.i11) * pix_bytes);
        w = make_float4(f.w00 * aw, f.w01 * aw, f.w10 * aw, f.w11 * aw);
      }
      s_off[warp][qi][pt] = o;
      s_w[warp][qi][pt] = w;
    }
  }
  __syncwarp();

  // ---- stage B: 4 lanes per query, 8 channels (16 bytes) per lane ---------------------------------------------
  const int qi = lane >> 2, sub = lane & 3;
  const int q = q0 + qi;
  const int64_t plane = HEAD_MAJOR ? ((int64_t)b * M + m) * Lv * V1_D : ((int64_t)b * Lv * M + m) * V1_D;
  const char* base = reinterpret_cast<const char*>(value + plane) + sub * 16;
  float2 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);

#pragma unroll 3
  for (int pt = 0; pt < LP; ++pt) {
    const uint4 o = s_off[warp][qi][pt];
    const float4 w = s_w[warp][qi][pt];
    const uint4 c00 = __ldg(reinterpret_cast<const uint4*>(base + o.x));
    const uint4 c01 = __ldg(reinterpret_cast<const uint4*>(base + o.y));
    const uint4 c10 = __ldg(reinterpret_cast<const uint4*>(base + o.z));
    const uint4 c11 = __ldg(reinterpret_cast<const uint4*>(base + o.w));
    const uint32_t a00[4] = {c00.x, c00.y, c00.z, c00.w}, a01[4] = {c01.x, c01.y, c01.z, c01.w};
    const uint32_t a10[4] = {c10.x, c10.y, c10.z, c10.w}, a11[4] = {c11.x, c11.y, c11.z, c11.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ffma2(acc[i], make_float2(__uint_as_float(a00[i] << 16), __uint_as_float(a00[i] & 0xffff0000u)), w.x);
      ffma2(acc[i], make_float2(__uint_as_float(a01[i] << 16), __uint_as_float(a01[i] & 0xffff0000u)), w.y);
      ffma2(acc[i], make_float2(__uint_as_float(a10[i] << 16), __uint_as_float(a10[i] & 0xffff0000u)), w.z);
      ffma2(acc[i], make_float2(__uint_as_float(a11[i] << 16), __uint_as_float(a11[i] & 0xffff0000u)), w.w);
    }
  }
  if (q < Lq) {
    uint4 r;
    uint32_t* rw = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(acc[i].x, acc[i].y);
      rw[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(out + (((int64_t)b * Lq + q) * M + m) * V1_D + sub * 8) = r;
  }
}

template <typename TL, int MODE, bool HM>
static int launch_v1(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out,
                     int B, int Lq, int Lv, int M, const LevelTable& lv, cudaStream_t st) {
  const int chunks = (Lq + V1_QPB - 1) / V1_QPB;
  const int64_t grid = (int64_t)B * chunks * M;
  if (grid > 0x7fffffffLL) return EMRT_ERR_UNSUPPORTED;
  msda_gather_fwd_v1_kernel<TL, MODE, HM, 3, 6><<<(unsigned)grid, V1_WARPS * 32, 0, st>>>(
      (const __nv_bfloat16*)value, (const TL*)loc, (const TL*)attn, ref, ref_bs, (__nv_bfloat16*)out, Lq, Lv, M, chunks, lv);
  EMRT_LAUNCH_CHECK();
  return EMRT_OK;
}

template <typename TL>
static int dispatch_v1(bool px, bool hm, const void* value, const void* loc, const void* attn, const float* ref,
                       int64_t ref_bs, void* out, int B, int Lq, int Lv, int M, const LevelTable& lv, cudaStream_t st) {
  if (px) return hm ? launch_v1<TL, 1, true>(value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st)
                    : launch_v1<TL, 1, false>(value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st);
  return hm ? launch_v1<TL, 0, true>(value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st)
            : launch_v1<TL, 0, false>(value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st);
}

int gather_fwd_v1(const void* value, const void* loc, const void* attn, const float* ref, int64_t ref_bs, void* out,
                  int B, int Lq, int Lv, int M, int D, int L, int P, const LevelTable& lv, int loc_dtype, int mode,
                  cudaStream_t st) {
  if (D != V1_D || L != 3 || P != 6) return EMRT_ERR_UNSUPPORTED;
  // per-plane byte offsets are 32-bit
  if ((int64_t)Lv * M * V1_D * 2 >= (1LL << 31)) return EMRT_ERR_UNSUPPORTED;
  const bool px = (mode & EMRT_LOC_PIXEL_OFFSET) != 0, hm = (mode & EMRT_VALUE_HEAD_MAJOR) != 0;
  switch (loc_dtype) {
    case EMRT_F32: return dispatch_v1<float>(px, hm, value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st);
    case EMRT_F16: return dispatch_v1<__half>(px, hm, value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st);
    case EMRT_BF16: return dispatch_v1<__nv_bfloat16>(px, hm, value, loc, attn, ref, ref_bs, out, B, Lq, Lv, M, lv, st);
    default: return EMRT_ERR_UNSUPPORTED;
  }
}

}  // namespace emrt
