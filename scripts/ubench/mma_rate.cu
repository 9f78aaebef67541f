// Micro-benchmark: tcgen05.mma.kind::tf32 issue rate on B200 as a function of N, operand layout and number of issuing warps.
// Operands are whatever shared memory holds (timing only).  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../deepqlearning.jl_b200/csrc/igemm.cuh"
using namespace dqn;
#define TC_KERNEL_ONLY
#include "../../deepqlearning.jl_b200/csrc/tc_gemm_impl.cuh"
using namespace tc;

// mode 0: K-major no-swizzle (LBO padded), 1: K-major no-swizzle (LBO unpadded), 2: K-major SWIZZLE_128B, 3: A from TMEM
template <int N>
__global__ void mma_rate(int iters, int issuers, int mode, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[4]; __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tslot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  constexpr uint32_t idesc = make_idesc2(128, N, false, false);
  if (warp < issuers) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint32_t lboA, sboA, lboB, sboB, lt = 0, kstepA, kstepB;
      if (mode == 0 || mode == 3) { sboA = sboB = 128; lboA = 16 * 128 + 16; lboB = (N / 8) * 128 + 16; kstepA = 2 * lboA; kstepB = 2 * lboB; }
      else if (mode == 1) { sboA = sboB = 128; lboA = 16 * 128; lboB = (N / 8) * 128; kstepA = 2 * lboA; kstepB = 2 * lboB; }
      else { sboA = sboB = 1024; lboA = lboB = 16; lt = 2; kstepA = kstepB = 32; }       // SWIZZLE_128B K-major: rows 128 B, 8-row groups 1024 B apart
      const uint32_t a0 = sbase + warp * 20480, b0 = sbase + 65536 + warp * 20480;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t da = make_desc(a0 + j * kstepA, lboA, sboA, lt), db = make_desc(b0 + j * kstepB, lboB, sboB, lt);
          const uint32_t acc = tmem + (uint32_t)(((warp * nacc) + ((it * 4 + j) % nacc)) * N) % 512;
          if (mode == 3) umma_tf32_ts(acc, tmem + 448 + (uint32_t)(j * 8), db, idesc, 1u);
          else umma_tf32(acc, da, db, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar[warp]));
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[warp * 2] = t1 - t0; }
    t1 = __shfl_sync(0xffffffffu, t0, 0);
    // the elected lane is not necessarily lane 0: take the max over lanes of t0 (others hold 0)
    long long tm = t0;
    for (int o = 16; o > 0; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, tm, o); tm = v > tm ? v : tm; }
    if (lane == 0 && blockIdx.x == 0) out[warp * 2 + 1] = t2 - tm;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N> void run(int issuers, int mode, int nacc, int ctas, int iters = 2000) {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_rate<N><<<ctas, 128, 200 * 1024>>>(iters, issuers, mode, nacc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  if (iters < 100) { printf("N=%3d issuers=%d mode=%d iters=%d: issue %lld clk, issue+commit->barrier complete %lld clk\n", N, issuers, mode, iters, h[0], h[1]); cudaFree(d); return; }
  printf("N=%3d issuers=%d mode=%d nacc=%d ctas=%3d: %s  per-MMA (all issuers): %.1f clk  [issue-only %.1f clk]\n", N, issuers, mode, nacc, ctas,
         cudaGetErrorString(e), (double)h[1] / (iters * 4.0 * issuers), (double)h[0] / (iters * 4.0));
  cudaFree(d);
}

int main() {
  for (int it = 0; it < 5; ++it) { run<64>(1, 3, 1, 1, it); run<64>(3, 3, 1, 1, it); }
  for (int mode = 0; mode < 4; mode += 3) {
    run<32>(1, mode, 1, 1); run<32>(2, mode, 1, 1); run<32>(3, mode, 1, 1); run<32>(3, mode, 1, 148);
    run<64>(1, mode, 1, 1); run<64>(2, mode, 1, 1); run<64>(3, mode, 1, 1); run<64>(3, mode, 2, 1);
    run<128>(1, mode, 1, 1); run<128>(2, mode, 1, 1); run<128>(3, mode, 1, 1);
  }
  return 0;
}
