// GPU self-test of the tcgen05 kernel in isolation: every operand-layout combination against the CPU executor of the
// same functor.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_selftest tests/csrc/tc_selftest.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <cuda_runtime.h>
#include "../../deepqlearning.jl_b200/csrc/igemm.cuh"
using namespace dqn;
#define TC_KERNEL_ONLY
#include "../../deepqlearning.jl_b200/csrc/tc_gemm_impl.cuh"
#include "../../deepqlearning.jl_b200/csrc/conv1_tc.cuh"

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static unsigned long long s_ = 88172645463325252ull;
static float rnd() { s_ ^= s_ << 13; s_ ^= s_ >> 7; s_ ^= s_ << 17; return (float)((s_ >> 11) % 20001) / 10000.f - 1.f; }

struct Arena {
  float* d = nullptr; long long plane = 0; std::vector<float> h; long long used = 0;
  void init(long long n) { plane = n; h.assign(2 * n, 0.f); CKC(cudaMalloc(&d, 2 * n * sizeof(float))); }
  long long put(const std::vector<float>& v, bool single = false) {      // returns offset; writes hi/lo split planes
    long long o = used; used += ((long long)v.size() + 63) / 64 * 64;
    for (size_t i = 0; i < v.size(); ++i) h[o + i] = v[i];      // raw fp32: the kernel splits into TF32 hi/lo terms in shared memory
    (void)single;
    return o;
  }
  long long reserve(long long n) { long long o = used; used += (n + 63) / 64 * 64; return o; }
  void upload() { CKC(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice)); }
};

#ifndef TEST_R
#define TEST_R 2
#endif
static int g_nsm = 148;
static int g_feed = 0;            // 0: cp.async loaders, 1: TMA producer (where the operands are boxes; other launches fall back and say so)
template <int BN, bool TMA, class Op>
static void launch_tc_feed(const Op& op, int nz, int nsplit, float* ws, long long ws_stride, const float* zero, const tc::TmaMaps& tm) {
  constexpr int R = BN == 32 ? 2 : TEST_R, NBUF = BN == 32 ? 2 : (TMA ? TC_TMA64_NBUF : (TEST_R == 1 ? 2 : 1));     // as launch_tc instantiates them
  using L = tc::Lay<BN, R, NBUF, Op::A_MCONTIG, !Op::B_KCONTIG, false, TMA>;
  static bool attr = false;
  if (!attr) { CKC(cudaFuncSetAttribute(tc::tc_gemm_kernel<BN, R, NBUF, false, Op, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM)); attr = true; }
  Op o0 = op; if (Op::Z_IS_CLASS) o0.set_class(0);
  const int MT = (o0.M + 127) / 128, NT = (o0.N + BN - 1) / BN, ntiles = MT * NT * nz * nsplit;
  const int grid = ntiles < g_nsm ? ntiles : g_nsm;
  tc::tc_gemm_kernel<BN, R, NBUF, false, Op, TMA><<<grid, tc::THREADS, L::SMEM>>>(op, op, op, op, nsplit, ws, ws_stride, zero, MT, NT, ntiles, 0, 0, tm);
}
static bool g_last_tma = false;
template <int BN, class Op>
static void launch_tc_raw(const Op& op, int nz, int nsplit, float* ws, long long ws_stride, const float* zero) {
  tc::TmaMaps tm; memset(&tm, 0, sizeof tm);
  g_last_tma = g_feed == 1 && tc::tma_build(&op, 1, BN, tm);
  if (g_last_tma) launch_tc_feed<BN, true>(op, nz, nsplit, ws, ws_stride, zero, tm);
  else launch_tc_feed<BN, false>(op, nz, nsplit, ws, ws_stride, zero, tm);
}
template <int BN, class Op>
static void run_tc(const Op& op, int nz, int nsplit, float* ws, long long ws_stride, const float* zero) {
  launch_tc_raw<BN>(op, nz, nsplit, ws, ws_stride, zero);
  CKC(cudaGetLastError());
  CKC(cudaDeviceSynchronize());
}

static int fails = 0;
static double g_tol = 2e-5;      // relative to the largest reference magnitude; long contractions (K ~ 40k per output, no k split here) get 1e-4
static void report(const char* name, const std::vector<float>& got, const std::vector<float>& ref) {
  double mx = 0, mr = 0; size_t at = 0;
  for (size_t i = 0; i < ref.size(); ++i) { double d = std::fabs((double)got[i] - ref[i]); if (d > mx) { mx = d; at = i; } mr = std::fmax(mr, std::fabs((double)ref[i])); }
  const bool ok = mx <= g_tol * mr;
  printf("%s %-5s %-28s maxdiff %.3e (ref max %.3e) at %zu: got %g ref %g\n", ok ? "ok  " : "FAIL", g_last_tma ? "[tma]" : "[cpa]", name, mx, mr, at, got[at], ref[at]);
  if (!ok) ++fails;
}

template <int BN> static void test_dense(int M, int N, int K) {
  Arena ar; ar.init(4 << 20);
  std::vector<float> X(M * K), W((K + 1) * N), D(M * N), Y(M * N), C(M * N, 0.f);
  for (auto& v : X) v = rnd(); for (auto& v : W) v = rnd() * 0.1f; for (auto& v : D) v = rnd(); for (auto& v : Y) v = rnd();
  long long oX = ar.put(X), oW = ar.put(W), oD = ar.put(D), oOnes = ar.put(std::vector<float>{1.f, 0.f, 0.f, 0.f}, true);
  ar.upload();
  float *dX, *dW, *dD, *dY, *dC, *dG, *dDX;
  CKC(cudaMalloc(&dX, X.size() * 4)); CKC(cudaMalloc(&dW, W.size() * 4)); CKC(cudaMalloc(&dD, D.size() * 4)); CKC(cudaMalloc(&dY, Y.size() * 4));
  CKC(cudaMalloc(&dC, C.size() * 4)); CKC(cudaMalloc(&dG, W.size() * 4)); CKC(cudaMalloc(&dDX, X.size() * 4));
  CKC(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice));
  {   // forward: A K-major, B MN-major
    DenseFwdOp op{}; op.X = X.data(); op.ldx = K; op.W = W.data(); op.C = C.data(); op.ldc = N; op.act = ACT_RELU; op.M = M; op.N = N; op.K = K; op.vecA = op.vecB = 1;
    std::vector<float> ref(M * N); op.C = ref.data(); igemm_host(op);
    DenseFwdOp g = op; g.X = dX; g.W = dW; g.C = dC; g.Xs = ar.d + oX; g.Ws = ar.d + oW; g.Cs = nullptr; g.lo_delta = ar.plane;
    run_tc<BN>(g, 1, 1, nullptr, 0, ar.d);
    std::vector<float> got(M * N); CKC(cudaMemcpy(got.data(), dC, got.size() * 4, cudaMemcpyDeviceToHost));
    report("dense_fwd  (A:K  B:MN)", got, ref);
  }
  {   // dgrad: both K-major
    DenseDgradOp op{}; op.D = D.data(); op.ldd = N; op.W = W.data(); op.ldx = K; op.Y = X.data(); op.ldy = K; op.act = ACT_TANH; op.accumulate = 0; op.apply_act = 1;
    op.M = M; op.N = K; op.K = N; op.vecA = op.vecB = 1;
    std::vector<float> ref(M * K); op.dX = ref.data(); igemm_host(op);
    DenseDgradOp g = op; g.D = dD; g.W = dW; g.dX = dDX; g.Y = dX; g.Ds = ar.d + oD; g.Ws = ar.d + oW; g.dXs = nullptr; g.lo_delta = ar.plane;
    run_tc<BN>(g, 1, 1, nullptr, 0, ar.d);
    std::vector<float> got(M * K); CKC(cudaMemcpy(got.data(), dDX, got.size() * 4, cudaMemcpyDeviceToHost));
    report("dense_dgrad(A:K  B:K )", got, ref);
  }
  {   // wgrad: both MN-major, with the ones column
    DenseWgradOp op{}; op.X = X.data(); op.ldx = K; op.D = D.data(); op.ldd = N; op.M = K + 1; op.N = N; op.K = M; op.vecA = op.vecB = 1;
    if (g_feed == 1) { op.no_bias = 1; op.M = K; }            // TMA feed: no ones row (the bias gradient is a column sum elsewhere)
    std::vector<float> ref((K + 1) * N, 0.f); op.dW = ref.data(); igemm_host(op);
    DenseWgradOp g = op; g.X = dX; g.D = dD; g.dW = dG; g.Xs = ar.d + oX; g.Ds = ar.d + oD; g.ones = ar.d + oOnes; g.lo_delta = ar.plane;
    run_tc<BN>(g, 1, 1, nullptr, 0, ar.d);
    std::vector<float> got((K + 1) * N, 0.f); CKC(cudaMemcpy(got.data(), dG, (size_t)op.M * N * 4, cudaMemcpyDeviceToHost));
    report("dense_wgrad(A:MN B:MN)", got, ref);
  }
  cudaFree(dX); cudaFree(dW); cudaFree(dD); cudaFree(dY); cudaFree(dC); cudaFree(dG); cudaFree(dDX); cudaFree(ar.d);
}

template <int BN> static void test_conv(int nimg, int IH, int IW, int Cin, int Cout, int KH, int KW, int S) {
  ConvGeom g{}; g.IH = IH; g.IW = IW; g.Cin = Cin; g.OH = (IH - KH) / S + 1; g.OW = (IW - KW) / S + 1; g.Cout = Cout; g.KH = KH; g.KW = KW; g.S = S; g.init();
  const int K = KH * KW * Cin, P = nimg * g.OH * g.OW;
  Arena ar; ar.init(8 << 20);
  std::vector<float> X((size_t)nimg * IH * IW * Cin), W((K + 1) * Cout), D((size_t)P * Cout);
  for (auto& v : X) v = rnd(); for (auto& v : W) v = rnd() * 0.1f; for (auto& v : D) v = rnd();
  long long oX = ar.put(X), oW = ar.put(W), oD = ar.put(D), oOnes = ar.put(std::vector<float>{1.f, 0.f, 0.f, 0.f}, true);
  ar.upload();
  float *dX, *dW, *dD, *dY, *dG, *dDX;
  CKC(cudaMalloc(&dX, X.size() * 4)); CKC(cudaMalloc(&dW, W.size() * 4)); CKC(cudaMalloc(&dD, D.size() * 4)); CKC(cudaMalloc(&dY, D.size() * 4));
  CKC(cudaMalloc(&dG, W.size() * 4)); CKC(cudaMalloc(&dDX, X.size() * 4));
  CKC(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  {
    ConvFwdOp op{}; op.X = X.data(); op.W = W.data(); op.act = ACT_RELU; op.nimg = nimg; op.g = g; op.M = P; op.N = Cout; op.K = K; op.vecA = op.vecB = 1;
    std::vector<float> ref((size_t)P * Cout); op.Y = ref.data(); igemm_host(op);
    ConvFwdOp q = op; q.X = dX; q.W = dW; q.Y = dY; q.Xs = ar.d + oX; q.Ws = ar.d + oW; q.Ys = nullptr; q.lo_delta = ar.plane;
    run_tc<BN>(q, 1, 1, nullptr, 0, ar.d);
    std::vector<float> got(ref.size()); CKC(cudaMemcpy(got.data(), dY, got.size() * 4, cudaMemcpyDeviceToHost));
    report("conv_fwd   (A:K  B:MN)", got, ref);
  }
  {
    ConvWgradOp op{}; op.X = X.data(); op.D = D.data(); op.nimg = nimg; op.g = g; op.M = K + 1; op.N = Cout; op.K = P; op.vecA = op.vecB = 1;
    if (g_feed == 1 && Cin % 32 == 0) { op.no_bias = 1; op.M = K; }     // TMA feed: no ones row (bias gradient = column sums, elsewhere)
    std::vector<float> ref((K + 1) * Cout, 0.f); op.dW = ref.data(); igemm_host(op);
    ConvWgradOp q = op; q.X = dX; q.D = dD; q.dW = dG; q.Xs = ar.d + oX; q.Ds = ar.d + oD; q.ones = ar.d + oOnes; q.lo_delta = ar.plane;
    run_tc<BN>(q, 1, 1, nullptr, 0, ar.d);
    std::vector<float> got(ref.size(), 0.f); CKC(cudaMemcpy(got.data(), dG, (size_t)op.M * Cout * 4, cudaMemcpyDeviceToHost));
    report("conv_wgrad (A:MN B:MN)", got, ref);
  }
  {
    std::vector<float> Yp(X.size()); for (auto& v : Yp) v = rnd();
    float* dYp; CKC(cudaMalloc(&dYp, Yp.size() * 4)); CKC(cudaMemcpy(dYp, Yp.data(), Yp.size() * 4, cudaMemcpyHostToDevice));
    ConvDgradOp op{}; op.D = D.data(); op.W = W.data(); op.Yprev = Yp.data(); op.act = ACT_RELU; op.apply_act = 1; op.nimg = nimg; op.g = g; op.vecA = op.vecB = 1;
    std::vector<float> ref(X.size(), 0.f); op.dX = ref.data(); for (int z = 0; z < S * S; ++z) igemm_host(op, z);
    ConvDgradOp q = op; q.D = dD; q.W = dW; q.dX = dDX; q.Yprev = dYp; q.Ds = ar.d + oD; q.Ws = ar.d + oW; q.dXs = nullptr; q.lo_delta = ar.plane;
    CKC(cudaMemset(dDX, 0, X.size() * 4));
    run_tc<BN>(q, S * S, 1, nullptr, 0, ar.d);
    std::vector<float> got(ref.size()); CKC(cudaMemcpy(got.data(), dDX, got.size() * 4, cudaMemcpyDeviceToHost));
    report("conv_dgrad (A:K  B:K )", got, ref);
    if (ConvDgradMergedOp::geometry_ok(g)) {       // the same gradient through the class-merged contraction (rearranged weights)
      const int TH = KH / S, TW = KW / S, Nm = S * S * Cin, Km = TH * TW * Cout;
      std::vector<float> Wm((size_t)Km * Nm);
      for (int k = 0; k < Km; ++k) for (int n = 0; n < Nm; ++n) {
        const int ci = n % Cin, cls = n / Cin, ph = cls / S, pw = cls % S, co = k % Cout, tap = k / Cout, th = tap / TW, tw = tap % TW;
        Wm[(size_t)k * Nm + n] = W[((size_t)((ph + S * th) * KW + (pw + S * tw)) * Cin + ci) * Cout + co];
      }
      float* dWm; CKC(cudaMalloc(&dWm, Wm.size() * 4)); CKC(cudaMemcpy(dWm, Wm.data(), Wm.size() * 4, cudaMemcpyHostToDevice));
      ConvDgradMergedOp mg{}; mg.D = dD; mg.Wm = dWm; mg.dX = dDX; mg.Yprev = dYp; mg.act = ACT_RELU; mg.apply_act = 1; mg.nimg = nimg; mg.g = g; mg.vecA = mg.vecB = 1;
      mg.Ds = dD; mg.Ws = dWm; mg.init();
      CKC(cudaMemset(dDX, 0, X.size() * 4));
      if (mg.N >= 64) run_tc<64>(mg, 1, 1, nullptr, 0, ar.d); else run_tc<32>(mg, 1, 1, nullptr, 0, ar.d);
      std::vector<float> got2(ref.size()); CKC(cudaMemcpy(got2.data(), dDX, got2.size() * 4, cudaMemcpyDeviceToHost));
      report("conv_dgrad merged classes", got2, ref);
      cudaFree(dWm);
    }
    cudaFree(dYp);
  }
  cudaFree(dX); cudaFree(dW); cudaFree(dD); cudaFree(dY); cudaFree(dG); cudaFree(dDX); cudaFree(ar.d);
}

// the dedicated first-layer kernel (conv1_tc.cuh) on byte observations against the CPU executor of ConvFwdOp (x_u8: k/255f0 operands)
static void test_conv1_bytes(int nimg, int IH, int IW, int iters) {
  ConvGeom g{}; g.IH = IH; g.IW = IW; g.Cin = 4; g.KH = 8; g.KW = 8; g.S = 4; g.OH = (IH - 8) / 4 + 1; g.OW = (IW - 8) / 4 + 1; g.Cout = 32; g.init();
  const int K = 256, N = 32, P = nimg * g.OH * g.OW;
  if (!c1::geometry_ok(g)) { printf("FAIL conv1 geometry rejected\n"); ++fails; return; }
  std::vector<uint8_t> X((size_t)nimg * IH * IW * 4);
  for (auto& v : X) { s_ ^= s_ << 13; s_ ^= s_ >> 7; s_ ^= s_ << 17; v = (uint8_t)(s_ >> 24); }
  std::vector<float> W((K + 1) * N), Ws(K * N);
  for (auto& v : W) v = rnd() * 0.1f;
  for (int i = 0; i < K * N; ++i) Ws[i] = W[i] * (1.0f / 255.0f);
  uint8_t* dX; float *dW, *dWs, *dY;
  CKC(cudaMalloc(&dX, X.size())); CKC(cudaMalloc(&dW, W.size() * 4)); CKC(cudaMalloc(&dWs, Ws.size() * 4)); CKC(cudaMalloc(&dY, (size_t)P * N * 4));
  CKC(cudaMemcpy(dX, X.data(), X.size(), cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dWs, Ws.data(), Ws.size() * 4, cudaMemcpyHostToDevice)); CKC(cudaMemset(dY, 0xFF, (size_t)P * N * 4));
  ConvFwdOp op{}; op.X = dX; op.x_u8 = 1; op.W = dW; op.Ws = dWs; op.Y = dY; op.act = ACT_RELU; op.nimg = nimg; op.g = g; op.M = P; op.N = N; op.K = K; op.a8 = 1;
  CUtensorMap tm;
  if (!c1::make_map(&tm, dX, nimg, g)) { printf("FAIL conv1 tensor map\n"); ++fails; return; }
  const c1::Params pr = c1::make_params(op);
  const int smem = c1::smem_bytes(pr.patch_slot), ntiles = (P + 127) / 128, grid = ntiles < g_nsm ? ntiles : g_nsm;
  CKC(cudaFuncSetAttribute(c1::conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  c1::conv1_fwd_kernel<<<grid, c1::C1_THREADS, smem>>>(tm, tm, pr, ntiles);
  CKC(cudaGetLastError()); CKC(cudaDeviceSynchronize());
  if (iters > 0) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) c1::conv1_fwd_kernel<<<grid, c1::C1_THREADS, smem>>>(tm, tm, pr, ntiles);
    cudaEventRecord(b); CKC(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double us = 1e3 * ms / iters;
    printf("bench conv1 dedicated kernel %d img %dx%dx4 (M=%d): %.1f us  %.1f TFLOP/s algorithmic  grid %d\n", nimg, IH, IW, P, us, 2.0 * P * N * K / (us * 1e-6) / 1e12, grid);
  } else {
    ConvFwdOp h = op; h.X = X.data(); h.W = W.data(); h.a8 = 0; h.vecA = h.vecB = 1;
    std::vector<float> ref((size_t)P * N); h.Y = ref.data(); igemm_host(h);
    std::vector<float> got(ref.size()); CKC(cudaMemcpy(got.data(), dY, got.size() * 4, cudaMemcpyDeviceToHost));
    g_last_tma = true;
    report("conv1_fwd bytes (dedicated)", got, ref);
  }
  cudaFree(dX); cudaFree(dW); cudaFree(dWs); cudaFree(dY);
}

// timing of the conv input-gradient contractions at the benchmarked sizes (class-wise and class-merged), with and without the act' epilogue
static void bench_conv_dgrad(int nimg, int IH, int IW, int Cin, int Cout, int KH, int S, int iters) {
  ConvGeom g{}; g.IH = IH; g.IW = IW; g.Cin = Cin; g.OH = (IH - KH) / S + 1; g.OW = (IW - KH) / S + 1; g.Cout = Cout; g.KH = KH; g.KW = KH; g.S = S; g.init();
  const int K = KH * KH * Cin, P = nimg * g.OH * g.OW;
  float *dW, *dD, *dDX, *dYp, *dWm, *zero;
  const size_t nx = (size_t)nimg * IH * IW * Cin;
  CKC(cudaMalloc(&dW, (size_t)(K + 1) * Cout * 4)); CKC(cudaMalloc(&dD, (size_t)P * Cout * 4)); CKC(cudaMalloc(&dDX, nx * 4)); CKC(cudaMalloc(&dYp, nx * 4));
  CKC(cudaMalloc(&dWm, (size_t)K * Cout * 4)); CKC(cudaMalloc(&zero, 256));
  CKC(cudaMemset(dW, 0, (size_t)(K + 1) * Cout * 4)); CKC(cudaMemset(dD, 0, (size_t)P * Cout * 4)); CKC(cudaMemset(dYp, 0, nx * 4)); CKC(cudaMemset(dWm, 0, (size_t)K * Cout * 4)); CKC(cudaMemset(zero, 0, 256));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int apply = 1; apply >= 0; --apply) {
    ConvDgradOp q{}; q.D = dD; q.W = dW; q.dX = dDX; q.Yprev = dYp; q.act = ACT_RELU; q.apply_act = apply; q.nimg = nimg; q.g = g; q.vecA = q.vecB = 1; q.Ds = dD; q.Ws = dW;
    for (int rep = 0; rep < 2; ++rep) {
      if (rep) cudaEventRecord(a);
      for (int i = 0; i < (rep ? iters : 1); ++i) { if (Cin >= 64) launch_tc_raw<64>(q, S * S, 1, nullptr, 0, zero); else launch_tc_raw<32>(q, S * S, 1, nullptr, 0, zero); }
      if (rep) cudaEventRecord(b);
      CKC(cudaDeviceSynchronize());
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("bench conv dgrad class-wise %s %dx%dx%d k%d s%d act'=%d: %.1f us\n", g_last_tma ? "[tma]" : "[cpa]", IH, IW, Cin, KH, S, apply, 1e3 * ms / iters);
    if (ConvDgradMergedOp::geometry_ok(g)) {
      ConvDgradMergedOp mg{}; mg.D = dD; mg.Wm = dWm; mg.dX = dDX; mg.Yprev = dYp; mg.act = ACT_RELU; mg.apply_act = apply; mg.nimg = nimg; mg.g = g; mg.vecA = mg.vecB = 1; mg.Ds = dD; mg.Ws = dWm; mg.init();
      for (int rep = 0; rep < 2; ++rep) {
        if (rep) cudaEventRecord(a);
        for (int i = 0; i < (rep ? iters : 1); ++i) launch_tc_raw<64>(mg, 1, 1, nullptr, 0, zero);
        if (rep) cudaEventRecord(b);
        CKC(cudaDeviceSynchronize());
      }
      cudaEventElapsedTime(&ms, a, b);
      printf("bench conv dgrad merged     %s %dx%dx%d k%d s%d act'=%d: %.1f us  (M=%d N=%d K=%d)\n", g_last_tma ? "[tma]" : "[cpa]", IH, IW, Cin, KH, S, apply, 1e3 * ms / iters, mg.M, mg.N, mg.K);
    }
  }
  cudaFree(dW); cudaFree(dD); cudaFree(dDX); cudaFree(dYp); cudaFree(dWm); cudaFree(zero);
}

// timing of one big dense forward (conv2-like and fc1-like shapes) - tuning aid
template <int BN> static void bench_dense(int M, int N, int K, int iters) {
  Arena ar; ar.init(64 << 20);
  std::vector<float> X((size_t)M * K, 0.5f), W((size_t)(K + 1) * N, 0.25f);
  long long oX = ar.put(X), oW = ar.put(W);
  ar.upload();
  float *dW, *dC; CKC(cudaMalloc(&dW, W.size() * 4)); CKC(cudaMalloc(&dC, (size_t)M * N * 4));
  CKC(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  DenseFwdOp g{}; g.X = nullptr; g.ldx = K; g.W = dW; g.C = dC; g.ldc = N; g.act = ACT_RELU; g.M = M; g.N = N; g.K = K; g.vecA = g.vecB = 1;
  g.Xs = ar.d + oX; g.Ws = ar.d + oW; g.Cs = nullptr; g.lo_delta = ar.plane;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  run_tc<BN>(g, 1, 1, nullptr, 0, ar.d);
  using L = tc::Lay<BN, BN == 32 ? 2 : TEST_R, BN == 32 ? 2 : (TEST_R == 1 ? 2 : 1), false, true>;
  dim3 grid((M + 127) / 128, (N + BN - 1) / BN, 1);
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) launch_tc_raw<BN>(g, 1, 1, nullptr, 0, ar.d);
  cudaEventRecord(b); CKC(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double us = 1e3 * ms / iters, tf = 2.0 * M * N * K / (us * 1e-6) / 1e12;
#ifdef TC_TRACE
  {
    std::vector<long long> tr(16384);
    CKC(cudaMemcpyFromSymbol(tr.data(), tc::tc_trace, sizeof(long long) * 16384));
    printf("  trace CTA0, cycles since launch; per stage: P.top P.slot-free P.issued | C.landed C.tmem-free C.st-issued C.st-done C.arrived | M.top M.full M.issued M.acc-free(1st stage of a tile) | per tile index: E.top E.acc-full E.done\n");
    const long long t0 = tr[0];
    for (int it = 0; it < 40; ++it) {
      printf("  %2d:", it);
      for (int j = 0; j < 15; ++j) { if (j == 3 || j == 8 || j == 12) printf(" |"); printf(" %6lld", tr[it * 16 + j] ? tr[it * 16 + j] - t0 : -1); }
      printf("\n");
    }
    printf("  epilogue of tile 0, per 16-column chunk: ld-issued ld-done summed stored\n");
    for (int c = 0; c < 4; ++c) { printf("   c%d:", c); for (int k = 0; k < 4; ++k) printf(" %6lld", tr[16000 + c * 4 + k] ? tr[16000 + c * 4 + k] - t0 : -1); printf("\n"); }
    std::vector<long long> z(16384, 0); CKC(cudaMemcpyToSymbol(tc::tc_trace, z.data(), sizeof(long long) * 16384));
  }
#endif
  printf("bench dense M=%d N=%d K=%d BN=%d stages=%d: %.1f us  %.1f TFLOP/s algorithmic (x3 = %.0f TF32)  grid %d\n", M, N, K, BN, L::STAGES, us, tf, 3 * tf, grid.x * grid.y);
  cudaFree(dW); cudaFree(dC); cudaFree(ar.d);
}

// timing of a dense weight gradient [x]^T delta (both operands MN-major) at the fc1 shape: features x units x batch rows - tuning aid
template <int BN> static void bench_dense_wgrad(int feat, int units, int rows, int iters) {
  Arena ar; ar.init(64 << 20);
  std::vector<float> X((size_t)rows * feat, 0.5f), D((size_t)rows * units, 0.25f);
  long long oX = ar.put(X), oD = ar.put(D), oOnes = ar.put(std::vector<float>{1.f, 0.f, 0.f, 0.f}, true);
  ar.upload();
  float* dG; CKC(cudaMalloc(&dG, (size_t)(feat + 1) * units * 4));
  DenseWgradOp g{}; g.X = nullptr; g.ldx = feat; g.D = nullptr; g.ldd = units; g.M = feat + 1; g.N = units; g.K = rows; g.vecA = g.vecB = 1;
  if (g_feed == 1) { g.no_bias = 1; g.M = feat; }
  g.dW = dG; g.Xs = ar.d + oX; g.Ds = ar.d + oD; g.ones = ar.d + oOnes; g.lo_delta = ar.plane;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  run_tc<BN>(g, 1, 1, nullptr, 0, ar.d);
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) launch_tc_raw<BN>(g, 1, 1, nullptr, 0, ar.d);
  cudaEventRecord(b); CKC(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double us = 1e3 * ms / iters, tf = 2.0 * g.M * g.N * g.K / (us * 1e-6) / 1e12;
#ifdef TC_TRACE
  {
    std::vector<long long> tr(16384);
    CKC(cudaMemcpyFromSymbol(tr.data(), tc::tc_trace, sizeof(long long) * 16384));
    const long long t0 = tr[0];
    for (int it = 0; it < 20; ++it) {
      printf("  %2d:", it);
      for (int j = 0; j < 15; ++j) { if (j == 3 || j == 8 || j == 12) printf(" |"); printf(" %6lld", tr[it * 16 + j] ? tr[it * 16 + j] - t0 : -1); }
      printf("\n");
    }
    std::vector<long long> z(16384, 0); CKC(cudaMemcpyToSymbol(tc::tc_trace, z.data(), sizeof(long long) * 16384));
  }
#endif
  printf("bench dense wgrad %s feat=%d units=%d rows=%d BN=%d: %.1f us  %.1f TFLOP/s algorithmic  tiles %d\n", g_last_tma ? "[tma]" : "[cpa]", feat, units, rows, BN, us, tf,
         ((g.M + 127) / 128) * ((units + BN - 1) / BN));
  cudaFree(dG); cudaFree(ar.d);
}

int main(int argc, char** argv) {
  tc::tma_conv_dgrad_enabled() = 1;      // keep the TMA-fed conv input gradients covered
  { const char* v = getenv("DQN_TC_AHELP"); if (v) tc::a_helper_enabled() = atoi(v); }
  { cudaDeviceProp pr; CKC(cudaGetDeviceProperties(&pr, 0)); g_nsm = pr.multiProcessorCount; }
  if (argc > 1) {
    const int it = argc > 2 ? atoi(argv[2]) : 20;
    if (argc > 3 && argv[3][0] == 'w') {
      for (g_feed = 0; g_feed < 2; ++g_feed) bench_dense_wgrad<64>(3136, 512, 256, it);
      return 0;
    }
    if (argc > 3 && argv[3][0] == 'd') {
      for (g_feed = 0; g_feed < 2; ++g_feed) { bench_conv_dgrad(256, 20, 20, 32, 64, 4, 2, it); bench_conv_dgrad(256, 9, 9, 64, 64, 3, 1, it); }
      return 0;
    }
    if (argc > 3) { test_conv1_bytes(512, 84, 84, it); test_conv1_bytes(256, 84, 84, it); return 0; }
    for (g_feed = 0; g_feed < 2; ++g_feed) {
      printf("== feed: %s\n", g_feed ? "TMA" : "cp.async");
      bench_dense<64>(41472, 64, 512, it);       // conv2 forward shape
      bench_dense<64>(37888, 64, 512, it);       // same, exactly one wave of 296 CTAs
      bench_dense<32>(204800, 32, 256, it);      // conv1 forward shape (with a lo plane here)
      bench_dense<64>(512, 1024, 3136, it);      // fc1 forward, both towers (64 tiles, no k split here)
      bench_dense<64>(18944, 128, 512, it);      // two waves of BN=64 tiles
    }
    return 0;
  }
  for (g_feed = 0; g_feed < 2; ++g_feed) {
    printf("==== feed: %s\n", g_feed ? "TMA where the operands are boxes" : "cp.async loaders");
    printf("-- dense M=256 N=64 K=96 (BN=64)\n");  test_dense<64>(256, 64, 96);
    printf("-- dense M=200 N=128 K=160 (BN=64, two n tiles)\n"); test_dense<64>(200, 128, 160);
    printf("-- dense M=130 N=32 K=64 (BN=32)\n");   test_dense<32>(130, 32, 64);
    printf("-- conv 4 img 20x20x8 -> 16, 4x4 s2 (BN=32)\n"); test_conv<32>(4, 20, 20, 8, 16, 4, 4, 2);
    printf("-- conv 3 img 9x9x32 -> 64, 3x3 s1 (BN=64)\n");  test_conv<64>(3, 9, 9, 32, 64, 3, 3, 1);
    printf("-- conv 5 img 20x20x32 -> 64, 4x4 s2 (BN=64, conv2 geometry, tiles cross images)\n");  test_conv<64>(5, 20, 20, 32, 64, 4, 4, 2);
    printf("-- conv 5 img 9x9x64 -> 64, 3x3 s1 (BN=64, conv3 geometry, two k stages per tap)\n");  test_conv<64>(5, 9, 9, 64, 64, 3, 3, 1);
    printf("-- dense M=39685 N=64 K=96 (BN=64, 311 tiles: several tiles per persistent CTA)\n"); g_tol = 1e-4; test_dense<64>(39685, 64, 96); g_tol = 2e-5;
    printf("-- conv 2 img 84x84x4 -> 32, 8x8 s4 (BN=32, conv1 geometry)\n"); test_conv<32>(2, 84, 84, 4, 32, 8, 8, 4);
  }
  printf("==== dedicated first-layer kernel\n");
  test_conv1_bytes(5, 84, 84, 0);        // tiles that cross images, a partial last tile
  test_conv1_bytes(1, 84, 84, 0);
  test_conv1_bytes(3, 80, 100, 0);       // another geometry of the same family (OW = 24)
  test_conv1_bytes(300, 84, 84, 0);      // several tiles per persistent CTA: both patch / accumulator buffers wrap
  printf(fails ? "SELFTEST FAILED %d\n" : "SELFTEST OK\n", fails);
  return fails ? 1 : 0;
}
