// What one tcgen05.mma costs as a function of N and of cta_group — the measurement behind DESIGN.md's "operand fetch" rule.
//
// A cta_group::1 MMA (M = 128, K = 16, bf16) reads its A tile (4 KB) and its B tile (32 N bytes) from shared memory for every
// instruction.  If the fetch path delivers 64 B/clk, an instruction costs (4096 + 32 N) / 64 = 64 + N/2 clocks while its
// arithmetic needs N/2: N = 256 runs at 67 % of the nominal rate, N = 128 at 50 %, N = 64 at 33 %.  With cta_group::2 (M = 256
// over a CTA pair) each SM fetches its own 128 rows of A and only HALF of B: (4096 + 16 N) / 64 clocks for the same N/2 clocks
// of arithmetic per SM — N = 256 then runs at the full rate.  This program measures clocks per instruction for both forms
// (back-to-back MMAs on resident operands, one issuing thread, one commit at the end) and checks the accumulator of the
// 2-CTA form against the closed form (operands constant along k, so the shared-memory swizzle does not matter).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate scripts/microbench/umma_operand_rate.cu && /tmp/umma_rate
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {   // K-major SWIZZLE_128B, SBO 1024 (tc_common.cuh)
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int KB = 4;             // k-blocks of each operand resident in shared memory (A: 64 KB, B: up to 128 KB)
struct Smem {
  __nv_bfloat16 a[KB][128 * 64];  // this CTA's 128 rows of A
  __nv_bfloat16 b[KB][256 * 64];  // B rows (all N for cta_group::1; this CTA's N/2 for cta_group::2)
  uint64_t done, dummy;
  uint32_t tmem_base;
};

// CG = 1: grid of independent CTAs.  CG = 2: clusters of two, rank 0 issues for the pair.
// mode bit 0: walk the KB resident k-blocks (distinct operand tiles per group of four MMAs) instead of re-issuing k-block 0;
// mode bit 1: one tcgen05.commit (to a barrier nobody waits on) after every k-block, as a smem-ring pipeline issues them
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(float* out, long long* cycles, int iters, int mode) {
  extern __shared__ __align__(1024) uint8_t raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
  // operands constant along k: A[r][:] = 1 + (global row % 7), B[n][:] = 1 + (global n % 5)
  const int b_rows = CG == 2 ? N / 2 : N;
  for (int kb = 0; kb < KB; ++kb) {
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) s.a[kb][i] = __float2bfloat16((float)(1 + ((int)rank * 128 + i / 64) % 7));
    for (int i = threadIdx.x; i < b_rows * 64; i += blockDim.x) s.b[kb][i] = __float2bfloat16((float)(1 + ((int)rank * b_rows + i / 64) % 5));
  }
  if (threadIdx.x == 0) {
    mbar_init(&s.done, 1);
    mbar_init(&s.dummy, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s.tmem_base;

  if (warp == 0 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc(128 * CG, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int kb = (mode & 1) ? (it & (KB - 1)) : 0;
      const uint64_t da = make_desc(smem_u32(s.a[kb])), db = make_desc(smem_u32(s.b[kb]));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (it | k) != 0;
        if (CG == 1)
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                       "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc)
                       : "memory");
        else
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                       "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc)
                       : "memory");
      }
      if (mode & 2) {
        if (CG == 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s.dummy)) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                           smem_u32(&s.dummy)), "h"((uint16_t)3) : "memory");
      }
    }
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s.done)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&s.done)),
                   "h"((uint16_t)3)
                   : "memory");
    mbar_wait(&s.done, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  // every warp reads its lane quarter of this CTA's accumulator rows: D[r][n] = 64 * iters * a_r * b_n
  mbar_wait(&s.done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (out) {
    uint32_t r[8];
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 8) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) out[((size_t)blockIdx.x * 128 + row) * N + c + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
  }
}

template <int CG, int N>
static void run(int sms, int iters, int mode) {
  const int grid = CG == 2 ? (sms / 2) * 2 : sms;
  const size_t smem = sizeof(Smem) + 1024;
  auto kern = umma_rate_kernel<CG, N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, (size_t)grid * 128 * N * sizeof(float)));
  CK(cudaMalloc(&cyc, grid * sizeof(long long)));
  CK(cudaMemset(cyc, 0, grid * sizeof(long long)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // correctness with few iterations (exact in fp32), then timing
  const int check_iters = 3;
  CK(cudaLaunchKernelEx(&cfg, kern, out, cyc, check_iters, mode));
  CK(cudaDeviceSynchronize());
  std::vector<float> h((size_t)grid * 128 * N);
  CK(cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int b = 0; b < grid; ++b)
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        const int grow = CG == 2 ? (b % 2) * 128 + r : r;
        const float want = 64.f * check_iters * (float)(1 + grow % 7) * (float)(1 + n % 5);
        if (h[((size_t)b * 128 + r) * N + n] != want) ++bad;
      }
  CK(cudaLaunchKernelEx(&cfg, kern, (float*)nullptr, cyc, iters, mode));
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double sum = 0;
  int n = 0;
  for (int b = 0; b < grid; ++b) if (c[b] > 0) { sum += (double)c[b]; ++n; }
  const double per = sum / n / (4.0 * iters);
  const double ideal = N / 2.0;      // arithmetic clocks per SM per instruction (4096 MAC/clk/SM)
  const double model = (4096.0 + (CG == 2 ? 16.0 : 32.0) * N) / 64.0;
  printf("%-34s cta_group::%d  M = %3d  N = %3d : %6.1f clk per tcgen05.mma  (arithmetic %5.1f, 64 B/clk operand model %5.1f)  -> %4.0f %% of the nominal rate   accumulator check: %s\n",
         mode == 0 ? "same k-block, one commit" : mode == 1 ? "4 k-blocks in turn, one commit" : mode == 2 ? "same k-block, commit per k-block" : "4 k-blocks, commit per k-block",
         CG, 128 * CG, N, per, ideal, model, 100.0 * ideal / per, bad ? "MISMATCH" : "exact");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 4000;
  printf("%d SMs, %d x 4 back-to-back MMAs (K = 16 each) per issuing thread, every SM busy\n", sms, iters);
  for (int mode = 0; mode < 4; ++mode) {
    run<1, 64>(sms, iters, mode);
    run<1, 128>(sms, iters, mode);
    run<1, 256>(sms, iters, mode);
    run<2, 64>(sms, iters, mode);
    run<2, 128>(sms, iters, mode);
    run<2, 256>(sms, iters, mode);
  }
  return 0;
}
