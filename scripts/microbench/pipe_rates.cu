// Issue-rate microbenchmark for the instructions the gather kernels are built from (sm_100a).
// Each kernel runs ITER x 32 independent-chain instructions per thread; 8 warps per SMSP (1024 threads/SM), 1 CTA/SM.
// Prints warp-instructions per clock per SMSP.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_rates.cu -o pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

template <int KIND>
__global__ void __launch_bounds__(1024) k(float* out, uint32_t seed, long long* cyc) {
  float acc[16];
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x * 0.5f + i; w[i] = seed + i * 0x01010101u + threadIdx.x; }
  const uint32_t wp = seed ^ 0x3f803f80u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (KIND == 0) {          // FFMA
        acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
        acc[i] = fmaf(acc[i], 0.9999f, 0.25f);
      } else if (KIND == 1) {   // FHFMA.BF16 (fma.rn.f32.bf16)
        unsigned short a_lo, a_hi, b_lo, b_hi;
        asm("mov.b32 {%0,%1}, %2;" : "=h"(a_lo), "=h"(a_hi) : "r"(w[i]));
        asm("mov.b32 {%0,%1}, %2;" : "=h"(b_lo), "=h"(b_hi) : "r"(wp));
        asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(a_lo), "h"(b_lo));
        asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(a_hi), "h"(b_hi));
      } else if (KIND == 2) {   // FFMA2 (fma.rn.f32x2): 2 instr on 16 x 64-bit accumulators -> use pairs
        unsigned long long a = ((unsigned long long)__float_as_uint(acc[i]) << 32) | __float_as_uint(acc[i]);
        unsigned long long m = 0x3f8000003f800000ull, c = 0x3f0000003f000000ull;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(m), "l"(c));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(m), "l"(c));
        acc[i] = __uint_as_float((uint32_t)a) + __uint_as_float((uint32_t)(a >> 32));
      } else if (KIND == 3) {   // LOP3
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(wp), "r"(seed));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(seed), "r"(wp));
      } else if (KIND == 4) {   // PRMT
        asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(w[i]) : "r"(wp));
        asm volatile("prmt.b32 %0, %0, %1, 0x1054;" : "+r"(w[i]) : "r"(wp));
      } else if (KIND == 5) {   // IMAD (shift-left as multiply)
        w[i] = w[i] * 65536u + seed;
        w[i] = w[i] * 3u + wp;
      } else if (KIND == 6) {   // HFMA2.BF16 packed
        asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(wp), "r"(seed));
        asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(wp), "r"(seed));
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += acc[i]; x ^= w[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND> void run(const char* name, float* out, long long* cyc) {
  k<KIND><<<148, 1024>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  k<KIND><<<148, 1024>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  // per SMSP: 8 warps x ITER x 32 instr
  printf("%-28s %.3f warp-instr/clk/SMSP (%.0f cycles)%s\n", name, 8.0 * ITER * 32 / avg, avg, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0>("FFMA", out, cyc);
  run<1>("FHFMA.BF16 (f32 += bf16*bf16)", out, cyc);
  run<2>("FFMA2 (f32x2)", out, cyc);
  run<3>("LOP3", out, cyc);
  run<4>("PRMT", out, cyc);
  run<5>("IMAD", out, cyc);
  run<6>("HFMA2.BF16 (bf16x2)", out, cyc);
  return 0;
}
