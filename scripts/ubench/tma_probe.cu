// Probe of the TMA features the round-2 producer path relies on (run on a B200; prints tables that are read offline):
//   A. shared-memory image of a tiled 2-D box under every CUtensorMapSwizzle mode (which 16-byte chunk lands where)
//   B. im2col-mode semantics: which input pixels a cp.async.bulk.tensor.4d...im2col load returns for given start coordinates / offsets
//   C. streaming rates per SM: im2col boxes (128 pixels x 32 fp32 channels), byte boxes with 32-byte rows, whole-patch boxes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_im2col4d(uint32_t dst, const CUtensorMap* tm, int c, int w, int h, int n, uint16_t ow, uint16_t oh, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
               ::"r"(dst), "l"(tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(bar), "h"(ow), "h"(oh) : "memory");
}

// ---- A / B: one load, dump the shared-memory image ------------------------------------------------
__global__ void dump_tiled(const __grid_constant__ CUtensorMap tm, int c0, int c1, int bytes, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) { mbar_init(s32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect(s32(&bar), bytes); tma2d(s32(sm), &tm, c0, c1, s32(&bar)); }
  mbar_wait(s32(&bar), 0);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(sm)[i];
}
__global__ void dump_im2col(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int ow, int oh, int bytes, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) { mbar_init(s32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect(s32(&bar), bytes); tma_im2col4d(s32(sm), &tm, c, w, h, n, (uint16_t)ow, (uint16_t)oh, s32(&bar)); }
  mbar_wait(s32(&bar), 0);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(sm)[i];
}

// ---- C: streaming rate ------------------------------------------------------------------------------
// mode 0: im2col 4-D (pixel groups walk the batch), 1: tiled 2-D (c0 = column block, c1 = row block), 2: tiled 3-D patch (w, h0, n)
constexpr int NST = 6;
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int mode, int stages, int stage_bytes, int p0, int p1, int p2, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[NST];
  const int tid = threadIdx.x;
  const int sb = (stage_bytes + 1023) / 1024 * 1024;
  if (tid == 0) { for (int i = 0; i < NST; ++i) mbar_init(s32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int depth = 4;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int s = 0; s < stages + depth; ++s) {
    if (s < stages && tid == 0) {
      const uint32_t d = s32(sm + (s % NST) * sb), b = s32(&bars[s % NST]);
      mbar_expect(b, stage_bytes);
      if (mode == 0) {
        // p0 = images, p1 = taps per side (KH = KW = p1), p2 = stride; tile index walks 128-pixel groups of this CTA's images
        const int taps = p1 * p1, tile = (blockIdx.x * 7 + s / taps) % (p0 * 81 / 128), tap = s % taps;
        const int pix = tile * 128, n = pix / 81, rem = pix % 81, oh = rem / 9, ow = rem % 9;
        tma_im2col4d(d, &tm, 0, ow * p2, oh * p2, n, (uint16_t)(tap % p1), (uint16_t)(tap / p1), b);
      } else if (mode == 1) {
        tma2d(d, &tm, (s % p0) * p2, ((blockIdx.x * 5 + s / p0) % p1) * 128, b);
      } else {
        tma3d(d, &tm, 0, ((s * 3) % p1) * 4, (blockIdx.x * 3 + s) % p0, b);
      }
    }
    if (s >= depth) {
      const int c = s - depth;
      mbar_wait(s32(&bars[c % NST]), (c / NST) & 1);
      acc += reinterpret_cast<const float*>(sm + (c % NST) * sb)[tid];
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) cyc[0] = 0;
}

typedef CUresult (*EncTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int nsm = pr.multiProcessorCount;
  void* f1 = nullptr; void* f2 = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &qr);
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &qr);
  if (!f1 || !f2) { printf("entry points missing\n"); return 1; }
  EncTiled encT = (EncTiled)f1; EncIm2col encI = (EncIm2col)f2;
  uint32_t* dout; cudaMalloc(&dout, 1 << 20);
  cudaFuncSetAttribute(dump_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(dump_im2col, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);

  // ---------------- A: swizzle images ----------------
  {
    const int R = 64, Ccols = 64;
    std::vector<uint32_t> h(R * Ccols);
    for (int r = 0; r < R; ++r) for (int c = 0; c < Ccols; ++c) h[r * Ccols + c] = (uint32_t)(r * 1000 + c);   // chunk id = c / 4
    uint32_t* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const char* names[] = {"NONE", "32B", "64B", "128B", "128B_ATOM_32B", "128B_ATOM_32B_FLIP_8B", "128B_ATOM_64B"};
    const CUtensorMapSwizzle modes[] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_64B};
    for (int mi = 3; mi < 5; ++mi) {     // (NONE / 32B / 64B / FLIP_8B were read in an earlier run; ATOM_64B is store-only: a load traps)
      const int bw = mi == 1 ? 8 : mi == 2 ? 16 : 32;            // box width in fp32 = swizzle span
      CUtensorMap tm; cuuint64_t gdim[2] = {(cuuint64_t)Ccols, (cuuint64_t)R}, gstr[1] = {(cuuint64_t)Ccols * 4}; cuuint32_t box[2] = {(cuuint32_t)bw, 16}, es[2] = {1, 1};
      CUresult r = encT(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[mi], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("A %s: encode failed %d\n", names[mi], (int)r); continue; }
      const int bytes = bw * 4 * 16;
      dump_tiled<<<1, 128, 48 * 1024>>>(tm, 0, 0, bytes, dout);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint32_t> o(bytes / 4); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
      printf("A swizzle %-22s (%s) box %d fp32 x 16 rows: physical 16B-chunk p of smem row-slot holds logical (row,chunk):\n", names[mi], cudaGetErrorString(e), bw);
      const int cpr = bw / 4;                                   // chunks per box row
      for (int pr_ = 0; pr_ < 16; ++pr_) {
        printf("   slot %2d:", pr_);
        for (int pc = 0; pc < cpr; ++pc) { const uint32_t v = o[(pr_ * cpr + pc) * 4]; printf(" (%2u,%u)", v / 1000, (v % 1000) / 4); }
        // within-chunk order check
        bool ok = true; for (int pc = 0; pc < cpr; ++pc) for (int j = 1; j < 4; ++j) if (o[(pr_ * cpr + pc) * 4 + j] != o[(pr_ * cpr + pc) * 4] + j) ok = false;
        printf("%s\n", ok ? "" : "  [words inside a chunk permuted]");
      }
    }
    cudaFree(d);
  }

  // ---------------- B: im2col semantics ----------------
  const int N = 4, H = 20, W = 20, C = 32, KS = 4, S = 2;
  std::vector<float> hx((size_t)N * H * W * C);
  for (int n = 0; n < N; ++n) for (int hh = 0; hh < H; ++hh) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c)
    hx[(((size_t)n * H + hh) * W + w) * C + c] = (float)(n * 1000000 + hh * 10000 + w * 100 + c);
  float* dx; cudaMalloc(&dx, hx.size() * 4); cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lo[2] = {0, 0}, up[2] = {-(KS - 1), -(KS - 1)};
    cuuint32_t es[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
    CUresult r = encI(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, gdim, gstr, lo, up, 32, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("B im2col encode (C=32,W=20,H=20,N=4; corners lo {0,0} up {-3,-3}; 32 ch x 128 pixels; strides {1,2,2,1}): %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      const int trials[][6] = {{0, 0, 0, 0, 0, 0}, {0, 2, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 2, 0, 0, 0}, {0, 0, 0, 0, 1, 2}, {0, 8, 4, 1, 3, 3}, {0, 16, 16, 0, 0, 0}, {0, 16, 16, 3, 0, 0}};
      for (auto& t : trials) {
        dump_im2col<<<1, 128, 48 * 1024>>>(tm, t[0], t[1], t[2], t[3], t[4], t[5], 128 * 128, dout);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o(128 * 32); cudaMemcpy(o.data(), dout, 128 * 128, cudaMemcpyDeviceToHost);
        printf("B coords {c=%d,w=%d,h=%d,n=%d} offsets {%d,%d}: %s\n   rows (n,h,w|c of chunk0):", t[0], t[1], t[2], t[3], t[4], t[5], cudaGetErrorString(e));
        for (int rr = 0; rr < 128; ++rr) {
          const float v = o[rr * 32 + ((0 ^ (rr & 7)) * 4)];      // logical chunk 0 of row rr under SWIZZLE_128B
          const long long q = (long long)v;
          if (rr < 22 || rr > 120 || (rr % 9) == 0) printf(" %d:(%lld,%lld,%lld|%lld)", rr, q / 1000000, (q / 10000) % 100, (q / 100) % 100, q % 100);
        }
        printf("\n");
      }
    }
    // ---------------- C: im2col streaming rate over a 512-image batch ----------------
    const int NB = 512;
    float* dbig; cudaMalloc(&dbig, (size_t)NB * H * W * C * 4); cudaMemset(dbig, 0, (size_t)NB * H * W * C * 4);
    cuuint64_t gdimB[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    r = encI(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dbig, gdimB, gstr, lo, up, 32, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    long long* cyc; cudaMalloc(&cyc, 8 * nsm);
    if (r == CUDA_SUCCESS) {
      const int stages = 640;
      for (int rep = 0; rep < 2; ++rep) stream_kernel<<<nsm, 128, NST * 16384 + 1024>>>(tm, 0, stages, 16384, NB, KS, S, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<long long> hc(nsm); cudaMemcpy(hc.data(), cyc, 8 * nsm, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < nsm; ++i) avg += hc[i]; avg /= nsm;
      printf("C im2col box 128 pixels x 32 fp32 (conv2 geometry, 26 MB input in L2): %s  %.0f clk per 16 KB stage = %.1f B/clk/SM\n", cudaGetErrorString(e), avg / stages, stages * 16384.0 / avg);
    }
    cudaFree(dbig);
    // tiled boxes with short rows out of a byte matrix (conv1-like: 32-byte runs), and fp32 32x128 for reference
    {
      const long long rows = 128LL * 1024, cols = 512;               // 64 MB of bytes, partly L2 resident
      uint8_t* db; cudaMalloc(&db, rows * cols); cudaMemset(db, 1, rows * cols);
      for (int bw : {16, 32, 64, 128}) {
        CUtensorMap t2; cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gs[1] = {(cuuint64_t)cols}; cuuint32_t bx[2] = {(cuuint32_t)bw, 128}, e2[2] = {1, 1};
        CUresult rr = encT(&t2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, db, gd, gs, bx, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) { printf("C byte box %d: encode failed %d\n", bw, (int)rr); continue; }
        const int stages = 640, sbytes = bw * 128;
        for (int rep = 0; rep < 2; ++rep) stream_kernel<<<nsm, 128, NST * 16384 + 1024>>>(t2, 1, stages, sbytes, (int)(cols / bw), (int)(rows / 128), bw, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> hc(nsm); cudaMemcpy(hc.data(), cyc, 8 * nsm, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < nsm; ++i) avg += hc[i]; avg /= nsm;
        printf("C tiled byte box %3d B x 128 rows (row pitch 512 B): %s  %.0f clk per box = %.2f clk/row = %.1f B/clk/SM\n", bw, cudaGetErrorString(e), avg / stages, avg / stages / 128, stages * (double)sbytes / avg);
      }
      cudaFree(db);
    }
    // whole-patch 3-D boxes: 84 u32 (one pixel = 4 channel bytes) x 36 rows of one image, the first conv layer's tile input
    {
      const int NI = 512;
      uint8_t* db; cudaMalloc(&db, (size_t)NI * 84 * 84 * 4); cudaMemset(db, 1, (size_t)NI * 84 * 84 * 4);
      CUtensorMap t3; cuuint64_t gd[3] = {84, 84, (cuuint64_t)NI}, gs[2] = {84 * 4, 84 * 84 * 4}; cuuint32_t bx[3] = {84, 36, 1}, e3[3] = {1, 1, 1};
      CUresult rr = encT(&t3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, db, gd, gs, bx, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rr != CUDA_SUCCESS) printf("C patch box: encode failed %d\n", (int)rr);
      else {
        const int stages = 640, sbytes = 84 * 4 * 36;
        for (int rep = 0; rep < 2; ++rep) stream_kernel<<<nsm, 128, NST * 16384 + 1024>>>(t3, 2, stages, sbytes, NI, 12, 0, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> hc(nsm); cudaMemcpy(hc.data(), cyc, 8 * nsm, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < nsm; ++i) avg += hc[i]; avg /= nsm;
        printf("C patch box 84 px x 36 rows (12096 B, rows of 336 B): %s  %.0f clk per box = %.1f B/clk/SM\n", cudaGetErrorString(e), avg / stages, stages * (double)sbytes / avg);
      }
      cudaFree(db);
    }
  }
  return 0;
}
