// Micro-benchmark: global(L2-resident) -> shared-memory streaming rate per SM on B200 for the candidate producer paths.
//   mode 0: cp.async.ca 16 B   1: cp.async.cg 16 B   2: ld.global.nc.v4 -> st.shared.v4   3: cp.async.bulk (1-D TMA) rows of `rowb` bytes
// Every CTA (one per SM, 256 loader threads) streams its own slice of a buffer that fits in L2, into a ring of shared memory.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
constexpr int STAGE = 24576, NST = 6;       // 24 KB stages (the GEMM's A+B stage), 6-deep ring
__global__ void __launch_bounds__(256, 1) g2s(const uint8_t* __restrict__ src, long long per_cta, int stages, int mode, int rowb, int depth, unsigned* sink, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bars[NST];
  const int tid = threadIdx.x;
  const uint8_t* base = src + (long long)blockIdx.x * per_cta;
  if (tid == 0) { for (int i = 0; i < NST; ++i) mbar_init(s32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  unsigned acc = 0;
  const long long t0 = clock64();
  if (mode <= 1) {
    for (int s = 0; s < stages + depth; ++s) {
      if (s < stages) {
        const uint8_t* g = base + ((long long)s * STAGE) % per_cta;
        uint8_t* d = sm + (s % NST) * STAGE;
#pragma unroll
        for (int i = 0; i < STAGE / 16 / 256; ++i) {
          const int c = tid + i * 256;
          if (mode == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s32(d + c * 16)), "l"(g + c * 16) : "memory");
          else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d + c * 16)), "l"(g + c * 16) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (s >= depth) {
        if (depth == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (depth == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (depth == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else asm volatile("cp.async.wait_group 4;" ::: "memory");
        acc += *reinterpret_cast<unsigned*>(sm + ((s - depth) % NST) * STAGE + tid * 16);
      }
    }
  } else if (mode == 2) {
    for (int s = 0; s < stages; ++s) {
      const uint8_t* g = base + ((long long)s * STAGE) % per_cta;
      uint8_t* d = sm + (s % NST) * STAGE;
      uint4 v[STAGE / 16 / 256];
#pragma unroll
      for (int i = 0; i < STAGE / 16 / 256; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(g + (tid + i * 256) * 16));
#pragma unroll
      for (int i = 0; i < STAGE / 16 / 256; ++i) *reinterpret_cast<uint4*>(d + (tid + i * 256) * 16) = v[i];
      acc += v[0].x;
    }
  } else {
    // 1-D bulk copies: one thread issues STAGE/rowb copies of rowb bytes per stage, completion on the stage's mbarrier
    if (tid < 32) {
      for (int s = 0; s < stages + depth; ++s) {
        if (s < stages) {
          const uint8_t* g = base + ((long long)s * STAGE) % per_cta;
          const uint32_t d = s32(sm + (s % NST) * STAGE), b = s32(&bars[s % NST]);
          if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(STAGE) : "memory");
          __syncwarp();
          for (int r = tid; r < STAGE / rowb; r += 32)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d + r * rowb), "l"(g + (long long)r * rowb), "r"(rowb), "r"(b) : "memory");
        }
        if (s >= depth) { mbar_wait(s32(&bars[(s - depth) % NST]), ((s - depth) / NST) & 1); acc += *reinterpret_cast<unsigned*>(sm + ((s - depth) % NST) * STAGE + tid * 16); }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) *sink = acc;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int nsm = pr.multiProcessorCount;
  const long long per_cta = 24576LL * 12;            // 288 KB per CTA -> 42 MB total: L2-resident after the first pass
  uint8_t* src; cudaMalloc(&src, per_cta * nsm); cudaMemset(src, 1, per_cta * nsm);
  unsigned* sink; cudaMalloc(&sink, 4); long long* cyc; cudaMalloc(&cyc, 8 * nsm);
  cudaFuncSetAttribute(g2s, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * STAGE + 1024);
  const int stages = 600;
  struct { int mode, rowb, depth; const char* name; } cfg[] = {
    {0, 0, 1, "cp.async.ca 16B depth1"}, {0, 0, 3, "cp.async.ca 16B depth3"}, {0, 0, 4, "cp.async.ca 16B depth4"}, {1, 0, 3, "cp.async.cg 16B depth3"},
    {2, 0, 0, "ldg.nc.v4 + sts.v4"}, {3, 128, 3, "bulk 128B rows depth3"}, {3, 256, 3, "bulk 256B rows depth3"}, {3, 1024, 3, "bulk 1KB rows depth3"},
    {3, 16, 3, "bulk 16B rows depth3"}, {3, 64, 3, "bulk 64B rows depth3"}, {3, 24576, 3, "bulk 24KB depth3"}};
  for (auto& c : cfg) {
    for (int rep = 0; rep < 2; ++rep) g2s<<<nsm, 256, NST * STAGE + 1024>>>(src, per_cta, stages, c.mode, c.rowb, c.depth, sink, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    g2s<<<nsm, 256, NST * STAGE + 1024>>>(src, per_cta, stages, c.mode, c.rowb, c.depth, sink, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[256]; cudaMemcpy(h, cyc, 8 * nsm, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    printf("%-26s %s  %.1f B/clk/SM  (%.0f clk per 24 KB stage)  chip %.2f TB/s\n", c.name, cudaGetErrorString(e), (double)stages * STAGE / avg, avg / stages,
           (double)stages * STAGE * nsm / (ms * 1e-3) / 1e12);
  }
  return 0;
}
