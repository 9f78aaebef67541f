// Micro-benchmark for the next round's producer path: cp.async.bulk.tensor (TMA) streaming rate per SM for the GEMM's A stage,
// a 2-D box of 32 fp32 columns x 128 rows (16 KB, SWIZZLE_128B) out of a row-major [rows][cols] matrix that sits in L2,
// plus a correctness check of the swizzled layout the converters would read back (16-byte chunk c of row r at c ^ (r & 7)).
// One elected thread per CTA issues one instruction per stage; compare with scripts/ubench/g2s_rate.cu (cp.async 16 B: 47 B/clk/SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tma_rate tma_rate.cu   (driver entry point through the runtime)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
constexpr int ROWS = 128, COLS = 32, STAGE = ROWS * COLS * 4, NST = 8;
// every CTA streams `stages` boxes: box s covers rows [128*(blockIdx.x*tiles_per + (s/kb)%tiles_per) ...), columns [32*(s%kb) ...)
__global__ void __launch_bounds__(128, 1) tma_stream(const __grid_constant__ CUtensorMap tm, int stages, int kb, int tiles_per, int depth,
                                                     float* check, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[NST];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < NST; ++i) mbar_init(s32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int s = 0; s < stages + depth; ++s) {
    if (s < stages && tid == 0) {
      const int tile = blockIdx.x * tiles_per + (s / kb) % tiles_per, kc = s % kb;
      const uint32_t d = s32(sm + (s % NST) * STAGE), b = s32(&bars[s % NST]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(STAGE) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(d), "l"(&tm), "r"(kc * COLS), "r"(tile * ROWS), "r"(b) : "memory");
    }
    if (s >= depth) {
      const int c = s - depth;
      mbar_wait(s32(&bars[c % NST]), (c / NST) & 1);
      // converter-style read-back: thread = row, 16-byte chunk j of the row lives at chunk j ^ (row & 7)
      const float* rowp = reinterpret_cast<const float*>(sm + (c % NST) * STAGE + tid * 128);
      const float4 v = *reinterpret_cast<const float4*>(rowp + ((0 ^ (tid & 7)) * 4));
      acc += v.x;
      if (check && blockIdx.x == 0 && c == 1) {                 // second box of CTA 0: dump it un-swizzled
        for (int j = 0; j < 8; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(rowp + ((j ^ (tid & 7)) * 4));
          check[tid * 32 + 4 * j] = w.x; check[tid * 32 + 4 * j + 1] = w.y; check[tid * 32 + 4 * j + 2] = w.z; check[tid * 32 + 4 * j + 3] = w.w;
        }
      }
      __syncthreads();                                          // slot free for the copy issued NST - depth stages later
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) check[0] = acc;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int nsm = pr.multiProcessorCount;
  const int kb = 16, tiles_per = 4;                             // matrix: (nsm*4*128) rows x 512 columns fp32 = 155 MB?  keep it in L2: 148*4*128*512*4 = 155 MB -> use 2 tiles
  const int tp = 1;                                             // one 128-row tile per CTA x 512 columns = 256 KB per CTA, 38 MB total: L2 resident
  (void)tiles_per;
  const long long rows = (long long)nsm * tp * ROWS, cols = (long long)kb * COLS;
  std::vector<float> h((size_t)rows * cols);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973);
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  if (!fn) { printf("no cuTensorMapEncodeTiled entry point\n"); return 1; }
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {COLS, ROWS}, estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  float* chk; cudaMalloc(&chk, ROWS * COLS * 4); long long* cyc; cudaMalloc(&cyc, 8 * nsm);
  cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * STAGE + 1024);
  const int stages = 960;
  for (int depth : {1, 2, 4, 6}) {
    for (int rep = 0; rep < 2; ++rep) tma_stream<<<nsm, 128, NST * STAGE + 1024>>>(tm, stages, kb, tp, depth, rep ? nullptr : chk, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    tma_stream<<<nsm, 128, NST * STAGE + 1024>>>(tm, stages, kb, tp, depth, nullptr, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> hc(nsm); cudaMemcpy(hc.data(), cyc, 8 * nsm, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += hc[i]; avg /= nsm;
    printf("TMA 2-D box 32x128 fp32 SWIZZLE_128B, %d in flight: %s  %.1f B/clk/SM (%.0f clk per 16 KB stage)  chip %.2f TB/s\n", depth, cudaGetErrorString(e),
           (double)stages * STAGE / avg, avg / stages, (double)stages * STAGE * nsm / (ms * 1e-3) / 1e12);
  }
  std::vector<float> hk(ROWS * COLS); cudaMemcpy(hk.data(), chk, hk.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int rr = 0; rr < ROWS; ++rr) for (int c = 0; c < COLS; ++c) if (hk[rr * COLS + c] != h[(size_t)rr * cols + 1 * COLS + c]) ++bad;   // box 1 of CTA 0: rows 0..127, columns 32..63
  printf("swizzled read-back (chunk j of row r at j ^ (r & 7)): %s (%d mismatches)\n", bad ? "MISMATCH" : "ok", bad);
  return 0;
}
