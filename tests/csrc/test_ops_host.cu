// CPU check of the implicit-GEMM operand functors (index algebra only; no GPU, no CUDA calls).
// Each op is run through dqn::igemm_host - the same loadA/loadB/store calls the kernel makes - and compared
// with a direct formula written independently here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepqlearning.jl_b200/csrc/igemm.cuh"
using namespace dqn;

static unsigned long long s_ = 88172645463325252ull;
static float rnd() { s_ ^= s_ << 13; s_ ^= s_ >> 7; s_ ^= s_ << 17; return (float)((s_ >> 11) % 2001) / 1000.f - 1.f; }
static int fails = 0;
static void cmp(const char* name, const std::vector<float>& a, const std::vector<float>& b, float tol = 2e-4f) {
  double mx = 0; size_t at = 0;
  if (a.size() != b.size()) { printf("FAIL %s size\n", name); ++fails; return; }
  for (size_t i = 0; i < a.size(); ++i) { double d = std::fabs((double)a[i] - b[i]); if (d > mx) { mx = d; at = i; } }
  if (mx > tol || mx != mx) { printf("FAIL %s maxdiff %g at %zu (%g vs %g)\n", name, mx, at, a[at], b[at]); ++fails; }
  else printf("ok   %s maxdiff %g\n", name, mx);
}

static void test_dense(int M, int K, int N, int act, bool u8) {
  std::vector<float> X(M * K), W((K + 1) * N), C(M * N), R(M * N);
  std::vector<uint8_t> X8(M * K);
  for (auto& v : X) v = rnd();
  for (auto& v : X8) v = (uint8_t)(rand() & 255);
  for (auto& v : W) v = rnd();
  auto x = [&](int m, int k) { return u8 ? (float)X8[m * K + k] / 255.f : X[m * K + k]; };
  DenseFwdOp op{}; op.X = u8 ? (const void*)X8.data() : (const void*)X.data(); op.ldx = K; op.x_u8 = u8; op.W = W.data(); op.C = C.data(); op.ldc = N; op.act = act;
  op.M = M; op.N = N; op.K = K; op.vecA = K % 4 == 0; op.vecB = N % 4 == 0;
  igemm_host(op);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = W[K * N + n]; for (int k = 0; k < K; ++k) s += (double)x(m, k) * W[k * N + n]; R[m * N + n] = act_apply((float)s, act); }
  cmp("dense_fwd", C, R);
  // wgrad
  std::vector<float> D(M * N), dW((K + 1) * N), RW((K + 1) * N);
  for (auto& v : D) v = rnd();
  DenseWgradOp wg{}; wg.X = op.X; wg.ldx = K; wg.x_u8 = u8; wg.D = D.data(); wg.ldd = N; wg.dW = dW.data(); wg.M = K + 1; wg.N = N; wg.K = M; wg.vecA = K % 4 == 0; wg.vecB = N % 4 == 0;
  igemm_host(wg);
  for (int k = 0; k <= K; ++k) for (int n = 0; n < N; ++n) { double s = 0; for (int m = 0; m < M; ++m) s += (double)(k < K ? x(m, k) : 1.f) * D[m * N + n]; RW[k * N + n] = (float)s; }
  cmp("dense_wgrad", dW, RW);
  // dgrad (with act' of a previous layer output Y and accumulate)
  std::vector<float> Y(M * K), dX(M * K), RX(M * K);
  for (auto& v : Y) v = rnd();
  for (auto& v : dX) v = rnd();
  RX = dX;
  DenseDgradOp dg{}; dg.D = D.data(); dg.ldd = N; dg.W = W.data(); dg.dX = dX.data(); dg.ldx = K; dg.Y = Y.data(); dg.ldy = K; dg.act = ACT_TANH; dg.accumulate = 1; dg.apply_act = 1;
  dg.M = M; dg.N = K; dg.K = N; dg.vecA = N % 4 == 0; dg.vecB = N % 4 == 0;
  igemm_host(dg);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { double s = 0; for (int n = 0; n < N; ++n) s += (double)D[m * N + n] * W[k * N + n]; RX[m * K + k] = (float)((s + RX[m * K + k]) * act_deriv(Y[m * K + k], ACT_TANH)); }
  cmp("dense_dgrad", dX, RX);
}

static void test_dgrad_two_towers(int M, int Kin, int N0, int N1) {
  std::vector<float> D0(M * N0), D1(M * N1), W0((Kin + 1) * N0), W1((Kin + 1) * N1), Y(M * Kin), dX(M * Kin, 0.f), R(M * Kin);
  for (auto& v : D0) v = rnd(); for (auto& v : D1) v = rnd(); for (auto& v : W0) v = rnd(); for (auto& v : W1) v = rnd(); for (auto& v : Y) v = rnd();
  DenseDgradOp op{}; op.D = D0.data(); op.ldd = N0; op.W = W0.data(); op.dX = dX.data(); op.ldx = Kin; op.Y = Y.data(); op.ldy = Kin; op.act = ACT_RELU; op.apply_act = 1;
  op.M = M; op.N = Kin; op.K = N0 + N1; op.K1 = N0; op.D2 = D1.data(); op.ldd2 = N1; op.W2 = W1.data(); op.vecA = op.vecB = (N0 % 4 == 0 && N1 % 4 == 0);
  igemm_host(op);
  for (int m = 0; m < M; ++m) for (int k = 0; k < Kin; ++k) {
    double s = 0; for (int n = 0; n < N0; ++n) s += (double)D0[m * N0 + n] * W0[k * N0 + n]; for (int n = 0; n < N1; ++n) s += (double)D1[m * N1 + n] * W1[k * N1 + n];
    R[m * Kin + k] = (float)(s * act_deriv(Y[m * Kin + k], ACT_RELU));
  }
  cmp("dense_dgrad_2towers", dX, R);
}

static void test_conv(int nimg, int IH, int IW, int Cin, int Cout, int KH, int KW, int S, bool u8) {
  ConvGeom g{}; g.IH = IH; g.IW = IW; g.Cin = Cin; g.OH = (IH - KH) / S + 1; g.OW = (IW - KW) / S + 1; g.Cout = Cout; g.KH = KH; g.KW = KW; g.S = S; g.init();
  const int K = KH * KW * Cin, P = nimg * g.OH * g.OW;
  std::vector<float> X(nimg * IH * IW * Cin), W((K + 1) * Cout), Y(P * Cout), R(P * Cout);
  std::vector<uint8_t> X8(X.size());
  for (auto& v : X) v = rnd();
  for (auto& v : X8) v = (uint8_t)(rand() & 255);
  for (auto& v : W) v = rnd();
  auto x = [&](int n, int h, int w, int c) { size_t o = ((size_t)(n * IH + h) * IW + w) * Cin + c; return u8 ? (float)X8[o] / 255.f : X[o]; };
  ConvFwdOp op{}; op.X = u8 ? (const void*)X8.data() : (const void*)X.data(); op.x_u8 = u8; op.W = W.data(); op.Y = Y.data(); op.act = ACT_RELU; op.nimg = nimg; op.g = g;
  op.M = P; op.N = Cout; op.K = K; op.vecA = Cin % 4 == 0; op.vecB = Cout % 4 == 0;
  igemm_host(op);
  for (int n = 0; n < nimg; ++n) for (int oh = 0; oh < g.OH; ++oh) for (int ow = 0; ow < g.OW; ++ow) for (int co = 0; co < Cout; ++co) {
    double s = W[K * Cout + co];
    for (int kh = 0; kh < KH; ++kh) for (int kw = 0; kw < KW; ++kw) for (int ci = 0; ci < Cin; ++ci)
      s += (double)x(n, oh * S + kh, ow * S + kw, ci) * W[((kh * KW + kw) * Cin + ci) * Cout + co];
    R[((n * g.OH + oh) * g.OW + ow) * Cout + co] = act_apply((float)s, ACT_RELU);
  }
  cmp("conv_fwd", Y, R);
  // wgrad
  std::vector<float> D(P * Cout), dW((K + 1) * Cout), RW((K + 1) * Cout, 0.f);
  for (auto& v : D) v = rnd();
  ConvWgradOp wg{}; wg.X = op.X; wg.x_u8 = u8; wg.D = D.data(); wg.dW = dW.data(); wg.nimg = nimg; wg.g = g; wg.M = K + 1; wg.N = Cout; wg.K = P; wg.vecA = Cin % 4 == 0; wg.vecB = Cout % 4 == 0;
  igemm_host(wg);
  {
    std::vector<double> acc((K + 1) * Cout, 0.0);
    for (int n = 0; n < nimg; ++n) for (int oh = 0; oh < g.OH; ++oh) for (int ow = 0; ow < g.OW; ++ow) for (int co = 0; co < Cout; ++co) {
      const double d = D[((n * g.OH + oh) * g.OW + ow) * Cout + co];
      for (int kh = 0; kh < KH; ++kh) for (int kw = 0; kw < KW; ++kw) for (int ci = 0; ci < Cin; ++ci)
        acc[((kh * KW + kw) * Cin + ci) * Cout + co] += d * x(n, oh * S + kh, ow * S + kw, ci);
      acc[K * Cout + co] += d;
    }
    for (size_t i = 0; i < acc.size(); ++i) RW[i] = (float)acc[i];
  }
  cmp("conv_wgrad", dW, RW, 1e-3f);
  // dgrad over all parity classes
  if (!u8) {
    std::vector<float> dX(X.size(), 123.f), RX(X.size()), Yp(X.size());
    for (auto& v : Yp) v = rnd();
    ConvDgradOp dg{}; dg.D = D.data(); dg.W = W.data(); dg.dX = dX.data(); dg.Yprev = Yp.data(); dg.act = ACT_RELU; dg.apply_act = 1; dg.nimg = nimg; dg.g = g;
    dg.vecA = Cout % 4 == 0; dg.vecB = Cout % 4 == 0;
    for (int z = 0; z < S * S; ++z) igemm_host(dg, z);
    std::vector<double> acc(X.size(), 0.0);
    for (int n = 0; n < nimg; ++n) for (int oh = 0; oh < g.OH; ++oh) for (int ow = 0; ow < g.OW; ++ow) for (int co = 0; co < Cout; ++co) {
      const double d = D[((n * g.OH + oh) * g.OW + ow) * Cout + co];
      for (int kh = 0; kh < KH; ++kh) for (int kw = 0; kw < KW; ++kw) for (int ci = 0; ci < Cin; ++ci)
        acc[((size_t)(n * IH + oh * S + kh) * IW + ow * S + kw) * Cin + ci] += d * W[((kh * KW + kw) * Cin + ci) * Cout + co];
    }
    for (size_t i = 0; i < acc.size(); ++i) RX[i] = (float)(acc[i] * act_deriv(Yp[i], ACT_RELU));
    cmp("conv_dgrad", dX, RX, 1e-3f);
    if (ConvDgradMergedOp::geometry_ok(g)) {       // the stride-parity classes merged into the column dimension, over rearranged weights
      const int TH = KH / S, TW = KW / S, Nm = S * S * Cin, Km = TH * TW * Cout;
      std::vector<float> Wm((size_t)Km * Nm), dX2(X.size(), 123.f);
      for (int k = 0; k < Km; ++k) for (int n = 0; n < Nm; ++n) {
        const int ci = n % Cin, cls = n / Cin, ph = cls / S, pw = cls % S, co = k % Cout, tap = k / Cout, th = tap / TW, tw = tap % TW;
        Wm[(size_t)k * Nm + n] = W[((size_t)((ph + S * th) * KW + (pw + S * tw)) * Cin + ci) * Cout + co];
      }
      ConvDgradMergedOp mg{}; mg.D = D.data(); mg.Wm = Wm.data(); mg.dX = dX2.data(); mg.Yprev = Yp.data(); mg.act = ACT_RELU; mg.apply_act = 1; mg.nimg = nimg; mg.g = g;
      mg.vecA = mg.vecB = 1; mg.init();
      igemm_host(mg);
      cmp("conv_dgrad merged classes", dX2, RX, 1e-3f);
    }
  }
}

static void test_fastdiv_and_u8() {
  for (uint32_t d : {1u, 2u, 3u, 4u, 5u, 7u, 9u, 20u, 32u, 49u, 64u, 81u, 84u, 400u, 3136u, 102400u}) {
    FastDiv f; f.init(d);
    for (uint32_t n : {0u, 1u, d - 1, d, d + 1, 12345u, 204799u, 1000003u, 0x7fffffffu, 0x7ffffffeu, 7u * d, 7u * d - 1}) {
      if (n > 0x7fffffffu) continue;
      uint32_t q, r; f.divmod(n, q, r);
      if (q != n / d || r != n % d) { printf("FAIL fastdiv %u / %u\n", n, d); ++fails; }
    }
  }
  for (int k = 0; k < 256; ++k) if (u8_to_f32((uint8_t)k) != (float)k / 255.f) { printf("FAIL u8_to_f32 %d\n", k); ++fails; }
  printf("ok   fastdiv/u8\n");
}

// The 8-column epilogue store of an op writes exactly what its two 4-column stores write (host path of st_global_v8: two float4 stores).
static void test_store8() {
  const int M = 5, N = 16, K = 8;
  {
    std::vector<float> W((K + 1) * N), c4(M * N, -7.f), c8(M * N, -7.f);
    for (size_t i = 0; i < W.size(); ++i) W[i] = 0.01f * (float)i - 0.3f;
    alignas(32) static float buf4[5 * 16], buf8[5 * 16];
    DenseFwdOp o{}; o.W = W.data(); o.ldc = N; o.act = ACT_RELU; o.M = M; o.N = N; o.K = K;
    for (int pass = 0; pass < 2; ++pass) {
      o.C = pass ? buf8 : buf4;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; n += 8) {
          const float4 v = make4(m - 0.25f * n, 0.5f * n - m, 1.f + n, -2.f * m), u = make4(0.125f * (m + n), 3.f - m, n * 0.75f, -1.f - m);
          const float4 x = o.epi_aux4(m, n), y = o.epi_aux4(m, n + 4);
          if (pass) o.store8x(m, n, v, u, x, y); else { o.store4x(m, n, v, x); o.store4x(m, n + 4, u, y); }
        }
    }
    double d = 0; for (int i = 0; i < M * N; ++i) d = std::fmax(d, std::fabs((double)buf4[i] - buf8[i]));
    if (!o.can_store8() || d != 0) { printf("FAIL store8 dense fwd %g\n", d); ++fails; } else printf("ok   store8 dense fwd\n");
  }
  {
    ConvGeom g{}; g.IH = 6; g.IW = 6; g.Cin = 8; g.KH = 4; g.KW = 4; g.S = 2; g.OH = 2; g.OW = 2; g.Cout = 8; g.init();
    alignas(32) static float x4[6 * 6 * 8], x8[6 * 6 * 8], yp[6 * 6 * 8];
    for (int i = 0; i < 6 * 6 * 8; ++i) { x4[i] = x8[i] = -7.f; yp[i] = (i % 3 == 0) ? 0.f : 0.5f; }
    ConvDgradMergedOp o{}; o.dX = x4; o.Yprev = yp; o.act = ACT_RELU; o.apply_act = 1; o.nimg = 1; o.g = g; o.init();
    for (int pass = 0; pass < 2; ++pass) {
      o.dX = pass ? x8 : x4;
      for (int m = 0; m < o.M; ++m)
        for (int n = 0; n < o.N; n += 8) {
          const float4 v = make4(m + 1.f, n - 2.f, 0.5f * m, 0.25f * n), u = make4(-1.f * m, 2.f + n, 1.5f, m * n * 0.125f);
          const float4 x = o.epi_aux4(m, n), y = o.epi_aux4(m, n + 4);
          if (pass) o.store8x(m, n, v, u, x, y); else { o.store4x(m, n, v, x); o.store4x(m, n + 4, u, y); }
        }
    }
    double d = 0; for (int i = 0; i < 6 * 6 * 8; ++i) d = std::fmax(d, std::fabs((double)x4[i] - x8[i]));
    if (!o.can_store8() || d != 0) { printf("FAIL store8 merged conv dgrad %g\n", d); ++fails; } else printf("ok   store8 merged conv dgrad\n");
  }
}

int main() {
  test_fastdiv_and_u8();
  test_store8();
  test_dgrad_two_towers(9, 20, 8, 12);
  test_dgrad_two_towers(5, 7, 4, 8);
  test_dense(7, 12, 8, ACT_RELU, false);
  test_dense(5, 2, 32, ACT_IDENTITY, false);     // README net first layer
  test_dense(9, 32, 1, ACT_IDENTITY, false);     // value head
  test_dense(6, 13, 5, ACT_TANH, false);         // nothing aligned
  test_dense(4, 16, 6, ACT_SIGMOID, true);       // u8 observations
  test_conv(2, 20, 20, 4, 8, 8, 8, 4, true);     // conv1-like, u8 input
  test_conv(2, 9, 9, 8, 12, 4, 4, 2, false);     // conv2-like, stride 2
  test_conv(2, 10, 10, 8, 12, 4, 4, 2, false);   // conv2 geometry family (even extents): also the class-merged input gradient
  test_conv(1, 12, 12, 4, 8, 6, 6, 3, false);    // stride 3, nine classes merged
  test_conv(2, 7, 7, 8, 8, 3, 3, 1, false);      // conv3-like
  test_conv(1, 12, 11, 3, 5, 4, 3, 2, false);    // odd everything (scalar paths, uneven parity classes)
  test_conv(1, 10, 10, 4, 4, 3, 3, 2, false);    // KH % S != 0
  printf(fails ? "FAILED %d\n" : "ALL OK\n", fails);
  return fails ? 1 : 0;
}
