// igemm.cuh - fp32 CUDA-core implicit-GEMM family (DQN_MATH_FP32 path and the small/odd-shaped layers).
//
// One tiled kernel body, six operand functors ("ops").  Every contraction of the Q-network forward and
// reverse pass is phrased as C[M][N] = sum_k A(m,k) * B(k,n) with the operands gathered on the fly:
//
//   DenseFwdOp    y = act(x W + b)                         Flux Dense forward
//   DenseDgradOp  dx = (delta W^T) .* act'(x)              reverse pass, SURVEY App. A step 9
//   DenseWgradOp  [dW; db] = [x 1]^T delta
//   ConvFwdOp     NHWC implicit im2col, true convolution   Flux Conv forward (App. B.2; the kernel flip is
//                                                          folded into the stored weight order at import)
//   ConvDgradOp   stride-parity classes, only live taps
//   ConvWgradOp   split-K over pixels, deterministic two-pass reduction (no atomics)
//
// Weight matrices are stored augmented: W is [(K+1)][N] row-major with the bias as row K, so that the
// gradient of weight and bias is one contraction against [x 1].
//
// The functors are __host__ __device__: tests/csrc/test_ops_host.cpp runs each of them through a plain
// CPU triple loop against direct convolution / matmul formulas, which checks the index algebra without
// a GPU.  The kernel body itself is checked on the GPU against oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef DQN_HD
#define DQN_HD __host__ __device__ __forceinline__
#endif

namespace dqn {

enum { ACT_IDENTITY = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_SIGMOID = 3 };

DQN_HD float act_apply(float z, int act) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? z : 0.f;
    case ACT_TANH: return tanhf(z);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
// sigma'(z) through the stored output y = sigma(z)
DQN_HD float act_deriv(float y, int act) {
  switch (act) {
    case ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case ACT_TANH: return 1.f - y * y;
    case ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

DQN_HD float u8_to_f32(uint8_t k) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn((float)k, 255.f);       // exactly Float32(k)/255f0, never a reciprocal multiply
#else
  return (float)k / 255.f;
#endif
}

DQN_HD float4 make4(float a, float b, float c, float d) { float4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }

// Load 4 consecutive elements p[0..3] of an fp32 or u8 array with a count guard (cnt = how many are in range).
DQN_HD float4 load4_f32(const float* p, int cnt, bool vec_ok) {
  if (cnt >= 4 && vec_ok) return *reinterpret_cast<const float4*>(p);
  float4 v = make4(0.f, 0.f, 0.f, 0.f);
  if (cnt > 0) v.x = p[0];
  if (cnt > 1) v.y = p[1];
  if (cnt > 2) v.z = p[2];
  if (cnt > 3) v.w = p[3];
  return v;
}
DQN_HD float4 load4_u8(const uint8_t* p, int cnt, bool vec_ok) {
  if (cnt >= 4 && vec_ok) {
    uchar4 q = *reinterpret_cast<const uchar4*>(p);
    return make4(u8_to_f32(q.x), u8_to_f32(q.y), u8_to_f32(q.z), u8_to_f32(q.w));
  }
  float4 v = make4(0.f, 0.f, 0.f, 0.f);
  if (cnt > 0) v.x = u8_to_f32(p[0]);
  if (cnt > 1) v.y = u8_to_f32(p[1]);
  if (cnt > 2) v.z = u8_to_f32(p[2]);
  if (cnt > 3) v.w = u8_to_f32(p[3]);
  return v;
}

struct ACtx { long long base; int i0, i1; int valid; };

// ------------------------------------------------------------------------------------------------
// Dense forward:  C[m][n] = act( sum_k X[m][k] W[k][n] + W[K][n] )
struct DenseFwdOp {
  static constexpr bool A_MCONTIG = false, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; long long ldx; int x_u8;
  const float* W;            // [(K+1)][N]
  float* C; long long ldc; int act;
  int M, N, K;
  int vecA, vecB;
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const { ACtx c; c.base = (long long)m * ldx; c.valid = m < M; c.i0 = c.i1 = 0; return c; }
  DQN_HD float4 loadA(const ACtx& c, int, int k) const {
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    if (x_u8) return load4_u8((const uint8_t*)X + c.base + k, K - k, vecA);
    return load4_f32((const float*)X + c.base + k, K - k, vecA);
  }
  DQN_HD float4 loadB(int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(W + (long long)k * N + n, N - n, vecB);
  }
  DQN_HD void store(int m, int n, float v) const {
    C[(long long)m * ldc + n] = act_apply(v + W[(long long)K * N + n], act);
  }
};

// Dense dgrad:  dX[m][n] (+)= sum_k D[m][k] W[n][k]  ;  times act'(Y[m][n]) when apply_act
struct DenseDgradOp {
  static constexpr bool A_MCONTIG = false, B_KCONTIG = true, Z_IS_CLASS = false;
  const float* D; long long ldd;
  const float* W;            // [(Kin+1)][Nout]; here GEMM-N = Kin, GEMM-K = Nout
  float* dX; long long ldx;
  const float* Y; long long ldy; int act; int accumulate; int apply_act;
  int M, N, K;
  int vecA, vecB;
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const { ACtx c; c.base = (long long)m * ldd; c.valid = m < M; c.i0 = c.i1 = 0; return c; }
  DQN_HD float4 loadA(const ACtx& c, int, int k) const {
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    return load4_f32(D + c.base + k, K - k, vecA);
  }
  DQN_HD float4 loadB(int k, int n) const {      // 4 consecutive k at column n
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(W + (long long)n * K + k, K - k, vecB);
  }
  DQN_HD void store(int m, int n, float v) const {
    long long o = (long long)m * ldx + n;
    if (accumulate) v += dX[o];
    if (apply_act) v *= act_deriv(Y[(long long)m * ldy + n], act);
    dX[o] = v;
  }
};

// Dense wgrad:  dW[m][n] = sum_k [X 1][k][m] D[k][n],  m in [0, Kin], k over the batch rows
struct DenseWgradOp {
  static constexpr bool A_MCONTIG = true, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; long long ldx; int x_u8;
  const float* D; long long ldd;
  float* dW;                 // [(Kin+1)][N]
  int M, N, K;               // M = Kin+1, K = batch rows
  int vecA, vecB;
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const { ACtx c; c.base = m; c.valid = m < M; c.i0 = c.i1 = 0; return c; }
  DQN_HD float4 loadA(const ACtx& c, int m, int k) const {   // 4 consecutive m at batch row k
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    const int kin = M - 1;
    int cnt = kin - m;                                      // real features left
    float4 v;
    if (cnt <= 0) v = make4(0, 0, 0, 0);
    else if (x_u8) v = load4_u8((const uint8_t*)X + (long long)k * ldx + m, cnt, vecA);
    else v = load4_f32((const float*)X + (long long)k * ldx + m, cnt, vecA);
    if (cnt >= 0 && cnt < 4) { float* f = &v.x; f[cnt] = 1.f; }   // the ones column that yields db
    return v;
  }
  DQN_HD float4 loadB(int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(D + (long long)k * ldd + n, N - n, vecB);
  }
  DQN_HD void store(int m, int n, float v) const { dW[(long long)m * N + n] = v; }
};

// ------------------------------------------------------------------------------------------------
struct ConvGeom { int IH, IW, Cin, OH, OW, Cout, KH, KW, S; };

// Conv forward (NHWC, weights [(KH*KW*Cin+1)][Cout], taps already flipped to cross-correlation order)
struct ConvFwdOp {
  static constexpr bool A_MCONTIG = false, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; int x_u8;
  const float* W; float* Y; int act; int nimg; ConvGeom g;
  int M, N, K;
  int vecA, vecB;
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const {
    ACtx c; c.valid = m < M; c.i0 = c.i1 = 0; c.base = 0;
    if (c.valid) {
      int ow = m % g.OW; int t = m / g.OW; int oh = t % g.OH; int n = t / g.OH;
      c.base = (((long long)n * g.IH + oh * g.S) * g.IW + ow * g.S) * g.Cin;
    }
    return c;
  }
  DQN_HD long long koff(int k) const {
    int ci = k % g.Cin; int t = k / g.Cin; int kw = t % g.KW; int kh = t / g.KW;
    return ((long long)kh * g.IW + kw) * g.Cin + ci;
  }
  DQN_HD float ld1(long long o) const { return x_u8 ? u8_to_f32(((const uint8_t*)X)[o]) : ((const float*)X)[o]; }
  DQN_HD float4 loadA(const ACtx& c, int, int k) const {
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    if (vecA) {     // Cin % 4 == 0: the four k are four channels of one tap
      long long o = c.base + koff(k);
      return x_u8 ? load4_u8((const uint8_t*)X + o, 4, true) : load4_f32((const float*)X + o, 4, true);
    }
    float4 v = make4(0, 0, 0, 0); float* f = &v.x;
    for (int j = 0; j < 4 && k + j < K; ++j) f[j] = ld1(c.base + koff(k + j));
    return v;
  }
  DQN_HD float4 loadB(int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(W + (long long)k * N + n, N - n, vecB);
  }
  DQN_HD void store(int m, int n, float v) const {
    Y[(long long)m * N + n] = act_apply(v + W[(long long)K * N + n], act);
  }
};

// Conv wgrad: dW[m][co] = sum_pix [im2col(X) 1][pix][m] D[pix][co];  m = (kh,kw,ci) or the bias row
struct ConvWgradOp {
  static constexpr bool A_MCONTIG = true, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; int x_u8;
  const float* D; float* dW; int nimg; ConvGeom g;
  int M, N, K;               // M = KH*KW*Cin + 1, N = Cout, K = nimg*OH*OW
  int vecA, vecB;
  DQN_HD void set_class(int) {}
  DQN_HD long long moff(int m) const {
    int ci = m % g.Cin; int t = m / g.Cin; int kw = t % g.KW; int kh = t / g.KW;
    return ((long long)kh * g.IW + kw) * g.Cin + ci;
  }
  DQN_HD ACtx prepA(int m) const { ACtx c; c.valid = m < M; c.i0 = c.i1 = 0; c.base = (c.valid && m < M - 1) ? moff(m) : 0; return c; }
  DQN_HD float ld1(long long o) const { return x_u8 ? u8_to_f32(((const uint8_t*)X)[o]) : ((const float*)X)[o]; }
  DQN_HD float4 loadA(const ACtx& c, int m, int k) const {   // 4 consecutive m at pixel k
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    int ow = k % g.OW; int t = k / g.OW; int oh = t % g.OH; int n = t / g.OH;
    long long pb = (((long long)n * g.IH + oh * g.S) * g.IW + ow * g.S) * g.Cin;
    const int kk = M - 1;
    int cnt = kk - m;
    float4 v = make4(0, 0, 0, 0); float* f = &v.x;
    if (cnt >= 4 && vecA) {
      v = x_u8 ? load4_u8((const uint8_t*)X + pb + c.base, 4, true) : load4_f32((const float*)X + pb + c.base, 4, true);
    } else {
      for (int j = 0; j < 4 && j < cnt; ++j) f[j] = ld1(pb + moff(m + j));
    }
    if (cnt >= 0 && cnt < 4) f[cnt] = 1.f;
    return v;
  }
  DQN_HD float4 loadB(int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(D + (long long)k * N + n, N - n, vecB);
  }
  DQN_HD void store(int m, int n, float v) const { dW[(long long)m * N + n] = v; }
};

// Conv dgrad by stride-parity class (ph,pw): rows are the input pixels with ih%S==ph, iw%S==pw, and only
// the taps kh = ph + S*th, kw = pw + S*tw can reach them:  oh = ih/S - th, ow = iw/S - tw.
struct ConvDgradOp {
  static constexpr bool A_MCONTIG = false, B_KCONTIG = true, Z_IS_CLASS = true;
  const float* D; const float* W; float* dX; const float* Yprev; int act; int apply_act; int nimg; ConvGeom g;
  int ph, pw, AH, BW, TH, TW;
  int M, N, K;               // set by set_class: M = nimg*AH*BW, N = Cin, K = TH*TW*Cout
  int vecA, vecB;
  DQN_HD void set_class(int z) {
    ph = z / g.S; pw = z % g.S;
    AH = (g.IH - ph + g.S - 1) / g.S; BW = (g.IW - pw + g.S - 1) / g.S;
    TH = (g.KH - ph + g.S - 1) / g.S; TW = (g.KW - pw + g.S - 1) / g.S;
    if (TH < 0) TH = 0; if (TW < 0) TW = 0;
    M = nimg * AH * BW; N = g.Cin; K = TH * TW * g.Cout;
  }
  DQN_HD ACtx prepA(int m) const {
    ACtx c; c.valid = m < M; c.base = 0; c.i0 = c.i1 = 0;
    if (c.valid) { int b = m % BW; int t = m / BW; int a = t % AH; int n = t / AH; c.i0 = a; c.i1 = b; c.base = n; }
    return c;
  }
  DQN_HD float4 loadA(const ACtx& c, int, int k) const {     // 4 consecutive co of one tap
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    float4 v = make4(0, 0, 0, 0); float* f = &v.x;
    if (vecA) {
      int co = k % g.Cout; int t = k / g.Cout; int tw = t % TW; int th = t / TW;
      int oh = c.i0 - th, ow = c.i1 - tw;
      if (oh < 0 || oh >= g.OH || ow < 0 || ow >= g.OW) return v;
      return load4_f32(D + (((long long)c.base * g.OH + oh) * g.OW + ow) * g.Cout + co, 4, true);
    }
    for (int j = 0; j < 4 && k + j < K; ++j) {
      int co = (k + j) % g.Cout; int t = (k + j) / g.Cout; int tw = t % TW; int th = t / TW;
      int oh = c.i0 - th, ow = c.i1 - tw;
      if (oh >= 0 && oh < g.OH && ow >= 0 && ow < g.OW) f[j] = D[(((long long)c.base * g.OH + oh) * g.OW + ow) * g.Cout + co];
    }
    return v;
  }
  DQN_HD float4 loadB(int k, int n) const {                  // 4 consecutive k (= co) at ci = n
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    float4 v = make4(0, 0, 0, 0); float* f = &v.x;
    if (vecB) {
      int co = k % g.Cout; int t = k / g.Cout; int tw = t % TW; int th = t / TW;
      int kh = ph + th * g.S, kw = pw + tw * g.S;
      return load4_f32(W + (((long long)kh * g.KW + kw) * g.Cin + n) * g.Cout + co, 4, true);
    }
    for (int j = 0; j < 4 && k + j < K; ++j) {
      int co = (k + j) % g.Cout; int t = (k + j) / g.Cout; int tw = t % TW; int th = t / TW;
      int kh = ph + th * g.S, kw = pw + tw * g.S;
      f[j] = W[(((long long)kh * g.KW + kw) * g.Cin + n) * g.Cout + co];
    }
    return v;
  }
  DQN_HD void store(int m, int n, float v) const {
    int b = m % BW; int t = m / BW; int a = t % AH; int img = t / AH;
    long long o = (((long long)img * g.IH + a * g.S + ph) * g.IW + b * g.S + pw) * g.Cin + n;
    if (apply_act) v *= act_deriv(Yprev[o], act);
    dX[o] = v;
  }
};

// ------------------------------------------------------------------------------------------------
// The tiled kernel.  256 threads, BK = 16, register-staged double buffering (one barrier per k-tile).
constexpr int IGEMM_THREADS = 256;
constexpr int IGEMM_BK = 16;

#ifdef __CUDACC__
template <int BM, int BN, int TM, int TN, class Op>
__global__ void __launch_bounds__(IGEMM_THREADS, 2)
igemm_kernel(Op opa, Op opb, int nsplit, float* __restrict__ ws, long long ws_stride) {
  constexpr int BK = IGEMM_BK;
  constexpr int NT = IGEMM_THREADS;
  static_assert((BM / TM) * (BN / TN) == NT, "thread tile");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 fragments");
  constexpr int LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int zi = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  Op op = (Op::Z_IS_CLASS || zi == 0) ? opa : opb;
  if (Op::Z_IS_CLASS) op.set_class(zi);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= op.M || n0 >= op.N) return;

  // k range of this split, in whole BK tiles
  const int ktiles = (op.K + BK - 1) / BK;
  const int per = (ktiles + nsplit - 1) / nsplit;
  const int kt0 = split * per, kt1 = min(ktiles, kt0 + per);

  const int tid = threadIdx.x;
  constexpr int A_F4 = BM * BK / 4, B_F4 = BK * BN / 4;
  constexpr int A_PER = (A_F4 + NT - 1) / NT, B_PER = (B_F4 + NT - 1) / NT;

  // per-thread A contexts (row decode hoisted out of the k loop)
  ACtx actx[A_PER]; int a_m[A_PER], a_k[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    int e = tid + i * NT;
    if (Op::A_MCONTIG) { a_m[i] = (e % (BM / 4)) * 4; a_k[i] = e / (BM / 4); }
    else               { a_k[i] = (e % (BK / 4)) * 4; a_m[i] = e / (BK / 4); }
    actx[i] = op.prepA(m0 + a_m[i]);
    if (e >= A_F4) actx[i].valid = 0;
  }
  int b_k[B_PER], b_n[B_PER];
#pragma unroll
  for (int i = 0; i < B_PER; ++i) {
    int e = tid + i * NT;
    if (Op::B_KCONTIG) { b_k[i] = (e % (BK / 4)) * 4; b_n[i] = e / (BK / 4); }
    else               { b_n[i] = (e % (BN / 4)) * 4; b_k[i] = e / (BN / 4); }
  }

  float4 ra[A_PER], rb[B_PER];
  auto gload = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) ra[i] = op.loadA(actx[i], m0 + a_m[i], k0 + a_k[i]);
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int e = tid + i * NT;
      rb[i] = (e < B_F4) ? op.loadB(k0 + b_k[i], n0 + b_n[i]) : make4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int e = tid + i * NT;
      if (e < A_F4) {
        if (Op::A_MCONTIG) *reinterpret_cast<float4*>(&As[buf][a_k[i]][a_m[i]]) = ra[i];
        else { As[buf][a_k[i] + 0][a_m[i]] = ra[i].x; As[buf][a_k[i] + 1][a_m[i]] = ra[i].y;
               As[buf][a_k[i] + 2][a_m[i]] = ra[i].z; As[buf][a_k[i] + 3][a_m[i]] = ra[i].w; }
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int e = tid + i * NT;
      if (e < B_F4) {
        if (Op::B_KCONTIG) { Bs[buf][b_k[i] + 0][b_n[i]] = rb[i].x; Bs[buf][b_k[i] + 1][b_n[i]] = rb[i].y;
                             Bs[buf][b_k[i] + 2][b_n[i]] = rb[i].z; Bs[buf][b_k[i] + 3][b_n[i]] = rb[i].w; }
        else *reinterpret_cast<float4*>(&Bs[buf][b_k[i]][b_n[i]]) = rb[i];
      }
    }
  };

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (kt0 < kt1) {
    gload(kt0);
    sstore(0);
    __syncthreads();
    for (int kt = kt0; kt < kt1; ++kt) {
      const int buf = (kt - kt0) & 1;
      if (kt + 1 < kt1) gload(kt + 1);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) *reinterpret_cast<float4*>(&a[i]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
#pragma unroll
        for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (kt + 1 < kt1) { sstore(buf ^ 1); __syncthreads(); }
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= op.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= op.N) continue;
      if (nsplit > 1) ws[(long long)blockIdx.z * ws_stride + (long long)m * op.N + n] = acc[i][j];
      else op.store(m, n, acc[i][j]);
    }
  }
}

// out[i] = sum_s ws[(z*nsplit + s)][i], s ascending (deterministic), for the split-K weight gradients
template <class Op>
__global__ void splitk_reduce_kernel(Op opa, Op opb, int nsplit, const float* __restrict__ ws, long long ws_stride) {
  const int zi = blockIdx.y;
  Op op = zi == 0 ? opa : opb;
  const long long total = (long long)op.M * op.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += ws[((long long)(zi * nsplit + k)) * ws_stride + i];
    op.store((int)(i / op.N), (int)(i % op.N), s);
  }
}
#endif  // __CUDACC__

// CPU reference executor of an op (index-algebra tests, no GPU): exactly the calls the kernel makes.
template <class Op>
inline void igemm_host(Op op, int zclass = 0) {
  if (Op::Z_IS_CLASS) op.set_class(zclass);
  for (int m = 0; m < op.M; ++m) {
    for (int n = 0; n < op.N; ++n) {
      double acc = 0.0;
      for (int k4 = 0; k4 < op.K; k4 += 4) {
        float av[4], bv[4];
        if (Op::A_MCONTIG) {
          for (int j = 0; j < 4; ++j) {
            int mb = (m / 4) * 4; ACtx c = op.prepA(mb); float4 v = op.loadA(c, mb, k4 + j); av[j] = (&v.x)[m - mb];
          }
        } else { ACtx c = op.prepA(m); float4 v = op.loadA(c, m, k4); av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w; }
        if (Op::B_KCONTIG) { float4 v = op.loadB(k4, n); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
        else { for (int j = 0; j < 4; ++j) { int nb = (n / 4) * 4; float4 v = op.loadB(k4 + j, nb); bv[j] = (&v.x)[n - nb]; } }
        for (int j = 0; j < 4; ++j) acc += (double)av[j] * (double)bv[j];
      }
      op.store(m, n, (float)acc);
    }
  }
}

}  // namespace dqn
