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
#include <string.h>

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

// Float32(k)/255f0 without a division: q = k*r, one Newton correction q + fma(-q,255,k)*r.  Exhaustively equal to
// the IEEE quotient for all 256 byte values (checked on the CPU in tests/test_oracle_cpu.py and on the GPU by the
// bit-exact get_batch test).
DQN_HD float u8_to_f32(uint8_t k) {
  const float kf = (float)k, r = 0.00392156885936856269836425781250f;   // fl(1/255)
  const float q = kf * r;
#ifdef __CUDA_ARCH__
  return __fmaf_rn(__fmaf_rn(-q, 255.f, kf), r, q);
#else
  return (float)((double)(float)((double)(-q) * 255.0 + (double)kf) * (double)r + (double)q);
#endif
}

DQN_HD float4 make4(float a, float b, float c, float d) { float4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }

// division by a runtime-constant divisor: q = umulhi(n, mul) >> shr, exact for n < 2^31 (mul = ceil(2^p/d), p = 31 + ceil(log2 d))
struct FastDiv {
  uint32_t d, mul, shr;
  void init(uint32_t dd) {
    d = dd; mul = 0; shr = 0;
    if (dd <= 1) return;
    uint32_t lg = 0; while ((1u << lg) < dd) ++lg;
    const uint32_t p = 31 + lg;
    mul = (uint32_t)(((1ull << p) + dd - 1) / dd);
    shr = p - 32;
  }
  DQN_HD uint32_t div(uint32_t n) const {
    if (d <= 1) return n;
#ifdef __CUDA_ARCH__
    return __umulhi(n, mul) >> shr;
#else
    return (uint32_t)(((uint64_t)n * mul) >> 32) >> shr;
#endif
  }
  DQN_HD void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

// read-only global loads (operands are never written by the kernel that gathers them)
DQN_HD float4 ldg4(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}
DQN_HD float ldg1(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
DQN_HD uchar4 ldg4u(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  return __ldg(reinterpret_cast<const uchar4*>(p));
#else
  return *reinterpret_cast<const uchar4*>(p);
#endif
}
DQN_HD uint8_t ldg1u(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// Load 4 consecutive elements p[0..3] of an fp32 or u8 array with a count guard (cnt = how many are in range).
DQN_HD float4 load4_f32(const float* p, int cnt, bool vec_ok) {
  if (cnt >= 4 && vec_ok) return ldg4(p);
  float4 v = make4(0.f, 0.f, 0.f, 0.f);
  if (cnt > 0) v.x = ldg1(p);
  if (cnt > 1) v.y = ldg1(p + 1);
  if (cnt > 2) v.z = ldg1(p + 2);
  if (cnt > 3) v.w = ldg1(p + 3);
  return v;
}
DQN_HD float4 load4_u8(const uint8_t* p, int cnt, bool vec_ok) {
  if (cnt >= 4 && vec_ok) {
    const uchar4 q = ldg4u(p);
    return make4(u8_to_f32(q.x), u8_to_f32(q.y), u8_to_f32(q.z), u8_to_f32(q.w));
  }
  float4 v = make4(0.f, 0.f, 0.f, 0.f);
  if (cnt > 0) v.x = u8_to_f32(ldg1u(p));
  if (cnt > 1) v.y = u8_to_f32(ldg1u(p + 1));
  if (cnt > 2) v.z = u8_to_f32(ldg1u(p + 2));
  if (cnt > 3) v.w = u8_to_f32(ldg1u(p + 3));
  return v;
}
// set element j (0..3) of a float4 without dynamic indexing (keeps the value in registers)
DQN_HD void set4(float4& v, int j, float x) { if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else if (j == 3) v.w = x; }

// x = hi + lo with hi, lo both exactly representable in TF32 (round-to-nearest, ties away: cvt.rna.tf32.f32)
DQN_HD float tf32_rna(float x) {
#ifdef __CUDA_ARCH__
  uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x)); return __uint_as_float(h);
#else
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r;
#endif
}
DQN_HD void split_tf32(float x, float& hi, float& lo) { hi = tf32_rna(x); lo = tf32_rna(x - hi); }
DQN_HD void split4(const float4& v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}
// Pre-split tensors live in one arena: the lo plane of every tensor sits `lo_delta` floats after its hi plane, so a
// single pointer names a 16-byte chunk of both planes.
DQN_HD void store_split4(float* hi_ptr, long long lo_delta, const float4& v) {
  float4 hi, lo; split4(v, hi, lo);
  *reinterpret_cast<float4*>(hi_ptr) = hi;
  *reinterpret_cast<float4*>(hi_ptr + lo_delta) = lo;
}

// Per-row (m) and per-column-of-A (k) decode contexts, hoisted out of the inner loops by the kernels.
struct ACtx { long long base; int i0, i1; int valid; };
struct KCtx { long long off; int t0, t1, t2; long long offb; };

// ------------------------------------------------------------------------------------------------
// Dense forward:  C[m][n] = act( sum_k X[m][k] W[k][n] + W[K][n] )
// 256-bit global store (sm_100: STG.E.ENL2.256): a lane that owns 8 consecutive fp32 of an output row writes one whole 32-byte sector.
// The tensor-core epilogue has one row per lane, so with 128-bit stores every store instruction touched 32 rows with half a sector each.
DQN_HD void st_global_v8(float* p, const float4& a, const float4& b) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
  reinterpret_cast<float4*>(p)[0] = a; reinterpret_cast<float4*>(p)[1] = b;
#endif
}
DQN_HD bool al32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

struct DenseFwdOp {
  static constexpr bool HAS_A8 = false;      // can feed the tensor-core kernel from a byte tensor
  static constexpr bool A_MCONTIG = false, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; long long ldx; int x_u8;
  const float* W;            // [(K+1)][N]
  float* C; long long ldc; int act;
  int M, N, K;
  int vecA, vecB;
  // tensor-core path: pre-split operands (hi plane pointers into the arena; lo = hi + lo_delta) and split output
  const float* Xs; const float* Ws; float* Cs; long long lo_delta; int a_single;
  DQN_HD bool tc_ready() const { return Xs && Ws && (K % 4 == 0) && (N % 4 == 0) && (ldx % 4 == 0); }
  DQN_HD const float* ptrA(const ACtx& c, const KCtx&, int, int k) const { return (c.valid && k < K) ? Xs + c.base + k : nullptr; }
  DQN_HD const float* ptrB(const KCtx&, int k, int n) const { return (k < K && n < N) ? Ws + (long long)k * N + n : nullptr; }
  // unchecked forms for stages that lie entirely inside the operand (interiorA / interiorB say so)
  DQN_HD bool interiorA(int m0, int k0, int bm, int bk) const { return m0 + bm <= M && k0 + bk <= K; }
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx& c, const KCtx&, int, int k) const { return Xs + c.base + k; }
  DQN_HD const float* ptrB_u(const KCtx&, int k, int n) const { return Ws + (k * N + n); }
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const { ACtx c; c.base = (long long)m * ldx; c.valid = m < M; c.i0 = c.i1 = 0; return c; }
  DQN_HD KCtx prepK(int k) const { KCtx c; c.off = k; c.t0 = c.t1 = c.t2 = 0; c.offb = 0; return c; }
  DQN_HD float4 loadA(const ACtx& c, const KCtx&, int, int k) const {
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    if (x_u8) return load4_u8((const uint8_t*)X + c.base + k, K - k, vecA);
    return load4_f32((const float*)X + c.base + k, K - k, vecA);
  }
  DQN_HD float4 loadB(const KCtx&, int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(W + (long long)k * N + n, N - n, vecB);
  }
  DQN_HD void store(int m, int n, float v) const {
    C[(long long)m * ldc + n] = act_apply(v + W[(long long)K * N + n], act);
  }
  DQN_HD bool can_store4() const { return (N % 4 == 0) && (ldc % 4 == 0); }
  DQN_HD void store4(int m, int n, float4 v) const {
    const float4 b = ldg4(W + (long long)K * N + n);
    const float4 y = make4(act_apply(v.x + b.x, act), act_apply(v.y + b.y, act), act_apply(v.z + b.z, act), act_apply(v.w + b.w, act));
    *reinterpret_cast<float4*>(C + (long long)m * ldc + n) = y;
    if (Cs) store_split4(Cs + (long long)m * ldc + n, lo_delta, y);
  }
  // epilogue operand fetched ahead of the accumulator (tensor-core kernel): the bias of columns n..n+3
  DQN_HD float4 epi_aux4(int, int n) const { return ldg4(W + (long long)K * N + n); }
  DQN_HD void store4x(int m, int n, float4 v, const float4& b) const {
    *reinterpret_cast<float4*>(C + (long long)m * ldc + n) = make4(act_apply(v.x + b.x, act), act_apply(v.y + b.y, act), act_apply(v.z + b.z, act), act_apply(v.w + b.w, act));
  }
  // eight consecutive columns (n % 8 == 0) in one 256-bit store
  DQN_HD bool can_store8() const { return (N % 8 == 0) && (ldc % 8 == 0) && al32(C); }
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4& b, const float4& c) const {
    st_global_v8(C + (long long)m * ldc + n, make4(act_apply(v.x + b.x, act), act_apply(v.y + b.y, act), act_apply(v.z + b.z, act), act_apply(v.w + b.w, act)),
                 make4(act_apply(u.x + c.x, act), act_apply(u.y + c.y, act), act_apply(u.z + c.z, act), act_apply(u.w + c.w, act)));
  }
};

// Dense dgrad:  dX[m][n] (+)= sum_k D[m][k] W[n][k]  ;  times act'(Y[m][n]) when apply_act
struct DenseDgradOp {
  static constexpr bool HAS_A8 = false;      // can feed the tensor-core kernel from a byte tensor
  static constexpr bool A_MCONTIG = false, B_KCONTIG = true, Z_IS_CLASS = false;
  const float* D; long long ldd;
  const float* W;            // [(Kin+1)][Nout]; here GEMM-N = Kin, GEMM-K = Nout
  float* dX; long long ldx;
  const float* Y; long long ldy; int act; int accumulate; int apply_act;
  int M, N, K;
  int vecA, vecB;
  const float* Ds; const float* Ws; float* dXs; long long lo_delta; int a_single;
  // optional second K segment (the other dueling tower): k in [K1, K) reads D2 / W2, so both towers' contributions to the
  // trunk gradient are ONE contraction instead of two launches with a read-modify-write in between.  K1 == 0: single segment.
  int K1; const float* D2; long long ldd2; const float* W2; const float* Ds2; const float* Ws2;
  DQN_HD int seg0() const { return K1 > 0 ? K1 : K; }
  DQN_HD bool tc_ready() const {
    return Ds && Ws && (K % 4 == 0) && (ldd % 4 == 0) && (K1 == 0 || (Ds2 && Ws2 && K1 % 4 == 0 && ldd2 % 4 == 0 && (long long)M * ldd2 < (1LL << 31))) && (long long)N * K < (1LL << 31);
  }
  // k context: kc.t0 = segment (0: D/W, 1: D2/W2), kc.off = k inside the segment, kc.t1 = the segment's width; row context: c.i0 = m * ldd2
  DQN_HD const float* ptrA(const ACtx& c, const KCtx& kc, int, int k) const {
    if (!c.valid || k >= K) return nullptr;
    return (kc.t0 ? Ds2 + c.i0 : Ds + c.base) + kc.off;
  }
  DQN_HD const float* ptrB(const KCtx& kc, int k, int n) const {
    if (k >= K || n >= N) return nullptr;
    return (kc.t0 ? Ws2 : Ws) + (n * kc.t1 + (int)kc.off);
  }
  DQN_HD bool interiorA(int m0, int k0, int bm, int bk) const { return m0 + bm <= M && k0 + bk <= K; }
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx& c, const KCtx& kc, int, int) const { return (kc.t0 ? Ds2 + c.i0 : Ds + c.base) + kc.off; }
  DQN_HD const float* ptrB_u(const KCtx& kc, int, int n) const { return (kc.t0 ? Ws2 : Ws) + (n * kc.t1 + (int)kc.off); }
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const { ACtx c; c.base = (long long)m * ldd; c.valid = m < M; c.i0 = (int)((long long)m * ldd2); c.i1 = 0; return c; }
  DQN_HD KCtx prepK(int k) const {
    KCtx c; c.t2 = 0; c.offb = 0;
    c.t0 = k < seg0() ? 0 : 1; c.off = c.t0 ? k - K1 : k; c.t1 = c.t0 ? K - K1 : seg0();
    return c;
  }
  DQN_HD float4 loadA(const ACtx& c, const KCtx&, int m, int k) const {
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    if (k < seg0()) return load4_f32(D + c.base + k, seg0() - k, vecA);
    return load4_f32(D2 + (long long)m * ldd2 + (k - K1), K - k, vecA);
  }
  DQN_HD float4 loadB(const KCtx&, int k, int n) const {      // 4 consecutive k at column n
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    if (k < seg0()) return load4_f32(W + (long long)n * seg0() + k, seg0() - k, vecB);
    return load4_f32(W2 + (long long)n * (K - K1) + (k - K1), K - k, vecB);
  }
  DQN_HD void store(int m, int n, float v) const {
    long long o = (long long)m * ldx + n;
    if (accumulate) v += dX[o];
    if (apply_act) v *= act_deriv(Y[(long long)m * ldy + n], act);
    dX[o] = v;
  }
  DQN_HD bool can_store4() const { return (ldx % 4 == 0) && (ldy % 4 == 0) && (N % 4 == 0); }
  DQN_HD void store4(int m, int n, float4 v) const {
    float4* o = reinterpret_cast<float4*>(dX + (long long)m * ldx + n);
    if (accumulate) { const float4 p = *o; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    if (apply_act) {
      const float4 y = *reinterpret_cast<const float4*>(Y + (long long)m * ldy + n);
      v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act);
    }
    *o = v;
    if (dXs) store_split4(dXs + (long long)m * ldx + n, lo_delta, v);
  }
  // epilogue operand fetched ahead of the accumulator: the stored layer output whose act' multiplies the gradient
  DQN_HD float4 epi_aux4(int m, int n) const { return apply_act ? ldg4(Y + (long long)m * ldy + n) : make4(1.f, 1.f, 1.f, 1.f); }
  DQN_HD void store4x(int m, int n, float4 v, const float4& y) const {
    float4* o = reinterpret_cast<float4*>(dX + (long long)m * ldx + n);
    if (accumulate) { const float4 p = *o; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    if (apply_act) { v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act); }
    *o = v;
  }
  DQN_HD bool can_store8() const { return !accumulate && (ldx % 8 == 0) && (N % 8 == 0) && al32(dX); }
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4& y, const float4& z) const {
    if (apply_act) {
      v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act);
      u.x *= act_deriv(z.x, act); u.y *= act_deriv(z.y, act); u.z *= act_deriv(z.z, act); u.w *= act_deriv(z.w, act);
    }
    st_global_v8(dX + (long long)m * ldx + n, v, u);
  }
};

// Dense wgrad:  dW[m][n] = sum_k [X 1][k][m] D[k][n],  m in [0, Kin], k over the batch rows
struct DenseWgradOp {
  static constexpr bool HAS_A8 = false;      // can feed the tensor-core kernel from a byte tensor
  static constexpr bool A_MCONTIG = true, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; long long ldx; int x_u8;
  const float* D; long long ldd;
  float* dW;                 // [(Kin+1)][N]
  int M, N, K;               // M = Kin+1, K = batch rows
  int vecA, vecB;
  const float* Xs; const float* Ds; const float* ones; long long lo_delta; int a_single; float out_scale;   // out_scale: 1/255 when X holds raw bytes
  int no_bias;               // 1: M = Kin, the bias gradient (column sums of D) is produced by colsum_kernel instead of a ones row
  DQN_HD int kin() const { return no_bias ? M : M - 1; }
  DQN_HD bool tc_ready() const { return Xs && Ds && ones && (kin() % 4 == 0) && (N % 4 == 0) && (ldx % 4 == 0) && (ldd % 4 == 0); }
  DQN_HD const float* ptrA(const ACtx& c, const KCtx& kc, int m, int k) const {     // 4 consecutive m at batch row k
    if (!c.valid || k >= K) return nullptr;
    const int cnt = kin() - m;
    return cnt >= 4 ? Xs + kc.off + m : (cnt == 0 ? ones : nullptr);
  }
  DQN_HD const float* ptrB(const KCtx&, int k, int n) const { return (k < K && n < N) ? Ds + (long long)k * ldd + n : nullptr; }
  DQN_HD bool interiorA(int m0, int k0, int bm, int bk) const { return m0 + bm <= kin() && k0 + bk <= K; }       // excludes the ones row
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx&, const KCtx& kc, int m, int) const { return Xs + kc.off + m; }
  DQN_HD const float* ptrB_u(const KCtx&, int k, int n) const { return Ds + ((long long)k * ldd + n); }
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const { ACtx c; c.base = m; c.valid = m < M; c.i0 = c.i1 = 0; return c; }
  DQN_HD KCtx prepK(int k) const { KCtx c; c.off = (long long)k * ldx; c.t0 = c.t1 = c.t2 = 0; c.offb = 0; return c; }
  DQN_HD float4 loadA(const ACtx& c, const KCtx& kc, int m, int k) const {   // 4 consecutive m at batch row k
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    const int cnt = kin() - m;                              // real features left
    float4 v;
    if (cnt <= 0) v = make4(0, 0, 0, 0);
    else if (x_u8) v = load4_u8((const uint8_t*)X + kc.off + m, cnt, vecA);
    else v = load4_f32((const float*)X + kc.off + m, cnt, vecA);
    if (cnt >= 0 && cnt < 4) set4(v, cnt, 1.f);             // the ones column that yields db
    return v;
  }
  DQN_HD float4 loadB(const KCtx&, int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(D + (long long)k * ldd + n, N - n, vecB);
  }
  DQN_HD float oscale(int m) const { return (out_scale != 0.f && m < kin()) ? out_scale : 1.f; }
  DQN_HD void store(int m, int n, float v) const { dW[(long long)m * N + n] = v * oscale(m); }
  DQN_HD bool can_store4() const { return N % 4 == 0; }
  DQN_HD void store4(int m, int n, float4 v) const {
    const float sc = oscale(m);
    *reinterpret_cast<float4*>(dW + (long long)m * N + n) = make4(v.x * sc, v.y * sc, v.z * sc, v.w * sc);
  }
  DQN_HD float4 epi_aux4(int, int) const { return make4(0.f, 0.f, 0.f, 0.f); }
  DQN_HD void store4x(int m, int n, float4 v, const float4&) const { store4(m, n, v); }
  DQN_HD bool can_store8() const { return (N % 8 == 0) && al32(dW); }
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4&, const float4&) const {
    const float sc = oscale(m);
    st_global_v8(dW + (long long)m * N + n, make4(v.x * sc, v.y * sc, v.z * sc, v.w * sc), make4(u.x * sc, u.y * sc, u.z * sc, u.w * sc));
  }
};

// ------------------------------------------------------------------------------------------------
struct ConvGeom {
  int IH, IW, Cin, OH, OW, Cout, KH, KW, S;
  FastDiv fCin, fKW, fOW, fOH, fCout;          // filled by init()
  FastDiv fBW[4], fAH[4], fTW[4];              // dgrad parity classes (S <= 4)
  void init() {
    fCin.init(Cin); fKW.init(KW); fOW.init(OW); fOH.init(OH); fCout.init(Cout);
    for (int p = 0; p < 4; ++p) {
      const int bw = p < S ? (IW - p + S - 1) / S : 1, ah = p < S ? (IH - p + S - 1) / S : 1, tw = p < S ? (KW - p + S - 1) / S : 1;
      fBW[p].init(bw > 0 ? bw : 1); fAH[p].init(ah > 0 ? ah : 1); fTW[p].init(tw > 0 ? tw : 1);
    }
  }
};

// Conv forward (NHWC, weights [(KH*KW*Cin+1)][Cout], taps already flipped to cross-correlation order)
struct ConvFwdOp {
  static constexpr bool HAS_A8 = true;      // can feed the tensor-core kernel from a byte tensor
  static constexpr bool A_MCONTIG = false, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; int x_u8;
  const float* W; float* Y; int act; int nimg; ConvGeom g;
  int M, N, K;
  int vecA, vecB;
  const float* Xs; const float* Ws; float* Ys; long long lo_delta; int a_single;
  int a8;                    // tensor-core path reads the byte tensor X itself (16 consecutive k per chunk) instead of an fp32 copy Xs
  // byte chunks: 16 consecutive k are 16 contiguous, 16-byte aligned bytes iff a tap row (KW*Cin) is a multiple of 32 and pixel
  // strides (S*Cin, IW*Cin, image size) are multiples of 16
  DQN_HD bool tc8_ready() const {
    return a8 && x_u8 && X && Ws && (N % 4 == 0) && ((g.KW * g.Cin) % 32 == 0) && ((g.S * g.Cin) % 16 == 0) && ((g.IW * g.Cin) % 16 == 0) &&
           (((long long)g.IH * g.IW * g.Cin) % 16 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  }
  DQN_HD bool tc_ready() const { return a8 ? tc8_ready() : (Xs && Ws && (g.Cin % 4 == 0) && (N % 4 == 0)); }
  DQN_HD const uint8_t* ptrA8(const ACtx& c, const KCtx& kc, int, int k) const { return (c.valid && k < K) ? (const uint8_t*)X + c.base + kc.off : nullptr; }
  DQN_HD const float* ptrA(const ACtx& c, const KCtx& kc, int, int k) const { return (c.valid && k < K) ? Xs + c.base + kc.off : nullptr; }
  DQN_HD const float* ptrB(const KCtx&, int k, int n) const { return (k < K && n < N) ? Ws + (long long)k * N + n : nullptr; }
  DQN_HD bool interiorA(int m0, int k0, int bm, int bk) const { return m0 + bm <= M && k0 + bk <= K; }
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx& c, const KCtx& kc, int, int) const { return Xs + c.base + kc.off; }
  DQN_HD const float* ptrB_u(const KCtx&, int k, int n) const { return Ws + (k * N + n); }
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const {
    ACtx c; c.valid = m < M; c.i0 = c.i1 = 0; c.base = 0;
    if (c.valid) {
      uint32_t t, ow, n, oh;
      g.fOW.divmod((uint32_t)m, t, ow); g.fOH.divmod(t, n, oh);
      c.base = (((long long)n * g.IH + oh * g.S) * g.IW + ow * g.S) * g.Cin;
    }
    return c;
  }
  DQN_HD long long koff(int k) const {
    uint32_t t, ci, kh, kw;
    g.fCin.divmod((uint32_t)k, t, ci); g.fKW.divmod(t, kh, kw);
    return ((long long)kh * g.IW + kw) * g.Cin + ci;
  }
  DQN_HD KCtx prepK(int k) const { KCtx c; c.off = k < K ? koff(k) : 0; c.t0 = c.t1 = c.t2 = 0; c.offb = 0; return c; }
  DQN_HD float ld1(long long o) const { return x_u8 ? u8_to_f32(ldg1u((const uint8_t*)X + o)) : ldg1((const float*)X + o); }
  DQN_HD float4 loadA(const ACtx& c, const KCtx& kc, int, int k) const {
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    if (vecA) {     // Cin % 4 == 0: the four k are four channels of one tap
      const long long o = c.base + kc.off;
      return x_u8 ? load4_u8((const uint8_t*)X + o, 4, true) : ldg4((const float*)X + o);
    }
    float4 v = make4(0, 0, 0, 0);
    for (int j = 0; j < 4 && k + j < K; ++j) set4(v, j, ld1(c.base + koff(k + j)));
    return v;
  }
  DQN_HD float4 loadB(const KCtx&, int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(W + (long long)k * N + n, N - n, vecB);
  }
  DQN_HD void store(int m, int n, float v) const {
    Y[(long long)m * N + n] = act_apply(v + W[(long long)K * N + n], act);
  }
  DQN_HD bool can_store4() const { return N % 4 == 0; }
  DQN_HD void store4(int m, int n, float4 v) const {
    const float4 b = ldg4(W + (long long)K * N + n);
    const float4 y = make4(act_apply(v.x + b.x, act), act_apply(v.y + b.y, act), act_apply(v.z + b.z, act), act_apply(v.w + b.w, act));
    *reinterpret_cast<float4*>(Y + (long long)m * N + n) = y;
    if (Ys) store_split4(Ys + (long long)m * N + n, lo_delta, y);
  }
  DQN_HD float4 epi_aux4(int, int n) const { return ldg4(W + (long long)K * N + n); }
  DQN_HD void store4x(int m, int n, float4 v, const float4& b) const {
    *reinterpret_cast<float4*>(Y + (long long)m * N + n) = make4(act_apply(v.x + b.x, act), act_apply(v.y + b.y, act), act_apply(v.z + b.z, act), act_apply(v.w + b.w, act));
  }
  DQN_HD bool can_store8() const { return (N % 8 == 0) && al32(Y); }
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4& b, const float4& c) const {
    st_global_v8(Y + (long long)m * N + n, make4(act_apply(v.x + b.x, act), act_apply(v.y + b.y, act), act_apply(v.z + b.z, act), act_apply(v.w + b.w, act)),
                 make4(act_apply(u.x + c.x, act), act_apply(u.y + c.y, act), act_apply(u.z + c.z, act), act_apply(u.w + c.w, act)));
  }
};

// Conv wgrad: dW[m][co] = sum_pix [im2col(X) 1][pix][m] D[pix][co];  m = (kh,kw,ci) or the bias row
struct ConvWgradOp {
  static constexpr bool HAS_A8 = true;      // can feed the tensor-core kernel from a byte tensor
  static constexpr bool A_MCONTIG = true, B_KCONTIG = false, Z_IS_CLASS = false;
  const void* X; int x_u8;
  const float* D; float* dW; int nimg; ConvGeom g;
  int M, N, K;               // M = KH*KW*Cin + 1 (or without the + 1, see no_bias), N = Cout, K = nimg*OH*OW
  int vecA, vecB;
  const float* Xs; const float* Ds; const float* ones; long long lo_delta; int a_single; float out_scale;
  int no_bias;               // 1: M = KH*KW*Cin, the bias gradient (column sums of D) is produced by colsum_kernel instead of a ones row
  int a8;                    // tensor-core path reads the byte tensor X itself (16 consecutive m per chunk); needs no_bias
  DQN_HD int kin() const { return no_bias ? M : M - 1; }
  DQN_HD bool tc8_ready() const {
    return a8 && no_bias && x_u8 && X && Ds && (N % 4 == 0) && (M % 16 == 0) && ((g.KW * g.Cin) % 16 == 0) && ((g.S * g.Cin) % 16 == 0) &&
           ((g.IW * g.Cin) % 16 == 0) && (((long long)g.IH * g.IW * g.Cin) % 16 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  }
  DQN_HD bool tc_ready() const { return a8 ? tc8_ready() : (Xs && Ds && ones && (g.Cin % 4 == 0) && (N % 4 == 0)); }
  DQN_HD const uint8_t* ptrA8(const ACtx& c, const KCtx& kc, int m, int k) const {     // 16 consecutive m (taps x channels) at pixel k
    return (c.valid && k < K && m + 16 <= kin()) ? (const uint8_t*)X + kc.off + c.base : nullptr;
  }
  DQN_HD const float* ptrA(const ACtx& c, const KCtx& kc, int m, int k) const {     // 4 consecutive m (channels of one tap) at pixel k
    if (!c.valid || k >= K) return nullptr;
    const int cnt = kin() - m;
    return cnt >= 4 ? Xs + kc.off + c.base : (cnt == 0 ? ones : nullptr);
  }
  DQN_HD const float* ptrB(const KCtx&, int k, int n) const { return (k < K && n < N) ? Ds + (long long)k * N + n : nullptr; }
  DQN_HD bool interiorA(int m0, int k0, int bm, int bk) const { return m0 + bm <= kin() && k0 + bk <= K; }          // excludes the ones row
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx& c, const KCtx& kc, int, int) const { return Xs + kc.off + c.base; }
  DQN_HD const float* ptrB_u(const KCtx&, int k, int n) const { return Ds + ((long long)k * N + n); }
  DQN_HD void set_class(int) {}
  DQN_HD long long moff(int m) const {
    uint32_t t, ci, kh, kw;
    g.fCin.divmod((uint32_t)m, t, ci); g.fKW.divmod(t, kh, kw);
    return ((long long)kh * g.IW + kw) * g.Cin + ci;
  }
  DQN_HD ACtx prepA(int m) const { ACtx c; c.valid = m < M; c.i0 = c.i1 = 0; c.base = (c.valid && m < kin()) ? moff(m) : 0; return c; }
  DQN_HD KCtx prepK(int k) const {          // pixel -> offset of its receptive field origin
    KCtx c; c.t0 = c.t1 = c.t2 = 0; c.off = 0; c.offb = 0;
    if (k < K) {
      uint32_t t, ow, n, oh;
      g.fOW.divmod((uint32_t)k, t, ow); g.fOH.divmod(t, n, oh);
      c.off = (((long long)n * g.IH + oh * g.S) * g.IW + ow * g.S) * g.Cin;
    }
    return c;
  }
  DQN_HD float ld1(long long o) const { return x_u8 ? u8_to_f32(ldg1u((const uint8_t*)X + o)) : ldg1((const float*)X + o); }
  DQN_HD float4 loadA(const ACtx& c, const KCtx& kc, int m, int k) const {   // 4 consecutive m at pixel k
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    const int cnt = kin() - m;
    float4 v = make4(0, 0, 0, 0);
    if (cnt >= 4 && vecA) {
      v = x_u8 ? load4_u8((const uint8_t*)X + kc.off + c.base, 4, true) : ldg4((const float*)X + kc.off + c.base);
    } else {
      for (int j = 0; j < 4 && j < cnt; ++j) set4(v, j, ld1(kc.off + moff(m + j)));
    }
    if (cnt >= 0 && cnt < 4) set4(v, cnt, 1.f);
    return v;
  }
  DQN_HD float4 loadB(const KCtx&, int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return load4_f32(D + (long long)k * N + n, N - n, vecB);
  }
  DQN_HD float oscale(int m) const { return (out_scale != 0.f && m < kin()) ? out_scale : 1.f; }
  DQN_HD void store(int m, int n, float v) const { dW[(long long)m * N + n] = v * oscale(m); }
  DQN_HD bool can_store4() const { return N % 4 == 0; }
  DQN_HD void store4(int m, int n, float4 v) const {
    const float sc = oscale(m);
    *reinterpret_cast<float4*>(dW + (long long)m * N + n) = make4(v.x * sc, v.y * sc, v.z * sc, v.w * sc);
  }
  DQN_HD float4 epi_aux4(int, int) const { return make4(0.f, 0.f, 0.f, 0.f); }
  DQN_HD void store4x(int m, int n, float4 v, const float4&) const { store4(m, n, v); }
  DQN_HD bool can_store8() const { return (N % 8 == 0) && al32(dW); }
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4&, const float4&) const {
    const float sc = oscale(m);
    st_global_v8(dW + (long long)m * N + n, make4(v.x * sc, v.y * sc, v.z * sc, v.w * sc), make4(u.x * sc, u.y * sc, u.z * sc, u.w * sc));
  }
};

// Conv dgrad by stride-parity class (ph,pw): rows are the input pixels with ih%S==ph, iw%S==pw, and only
// the taps kh = ph + S*th, kw = pw + S*tw can reach them:  oh = ih/S - th, ow = iw/S - tw.
struct ConvDgradOp {
  static constexpr bool HAS_A8 = false;      // can feed the tensor-core kernel from a byte tensor
  static constexpr bool A_MCONTIG = false, B_KCONTIG = true, Z_IS_CLASS = true;
  const float* D; const float* W; float* dX; const float* Yprev; int act; int apply_act; int nimg; ConvGeom g;
  int ph, pw, AH, BW, TH, TW;
  FastDiv fbw, fah, ftw;     // the class's divisors, picked by set_class with static indices (no local-memory copy)
  int M, N, K;               // set by set_class: M = nimg*AH*BW, N = Cin, K = TH*TW*Cout
  int vecA, vecB;
  const float* Ds; const float* Ws; float* dXs; long long lo_delta; int a_single;
  DQN_HD bool tc_ready() const { return Ds && Ws && (g.Cout % 4 == 0); }
  // row context: c.base = offset of D[img][i0][i1][0] (the tap (0,0) source pixel, possibly outside the image), i0/i1 for the bounds;
  // k context: kc.off = co - (th*OW + tw)*Cout moves from there to tap (th,tw), kc.offb = offset of W[kh][kw][0][co]
  DQN_HD bool tap_ok(const ACtx& c, const KCtx& kc) const {
    return (unsigned)(c.i0 - kc.t0) < (unsigned)g.OH && (unsigned)(c.i1 - kc.t1) < (unsigned)g.OW;
  }
  DQN_HD const float* ptrA(const ACtx& c, const KCtx& kc, int, int k) const {
    if (!c.valid || k >= K || !tap_ok(c, kc)) return nullptr;
    return Ds + c.base + kc.off;
  }
  DQN_HD const float* ptrB(const KCtx& kc, int k, int n) const {
    if (k >= K || n >= N) return nullptr;
    return Ws + kc.offb + (long long)(n * g.Cout);
  }
  DQN_HD bool interiorA(int, int, int, int) const { return false; }            // tap bounds depend on the row: always the checked form
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx& c, const KCtx& kc, int, int) const { return Ds + c.base + kc.off; }
  DQN_HD const float* ptrB_u(const KCtx& kc, int, int n) const { return Ws + kc.offb + (long long)(n * g.Cout); }
  DQN_HD void set_class(int z) {
    ph = z / g.S; pw = z - ph * g.S;
    AH = (g.IH - ph + g.S - 1) / g.S; BW = (g.IW - pw + g.S - 1) / g.S;
    TH = (g.KH - ph + g.S - 1) / g.S; TW = (g.KW - pw + g.S - 1) / g.S;
    if (TH < 0) TH = 0; if (TW < 0) TW = 0;
    fbw = pw == 0 ? g.fBW[0] : pw == 1 ? g.fBW[1] : pw == 2 ? g.fBW[2] : g.fBW[3];
    ftw = pw == 0 ? g.fTW[0] : pw == 1 ? g.fTW[1] : pw == 2 ? g.fTW[2] : g.fTW[3];
    fah = ph == 0 ? g.fAH[0] : ph == 1 ? g.fAH[1] : ph == 2 ? g.fAH[2] : g.fAH[3];
    M = nimg * AH * BW; N = g.Cin; K = TH * TW * g.Cout;
  }
  DQN_HD ACtx prepA(int m) const {
    ACtx c; c.valid = m < M; c.base = 0; c.i0 = c.i1 = 0;
    if (c.valid) {
      uint32_t t, b, n, a; fbw.divmod((uint32_t)m, t, b); fah.divmod(t, n, a); c.i0 = (int)a; c.i1 = (int)b;
      c.base = (((long long)n * g.OH + (int)a) * g.OW + (int)b) * g.Cout;
    }
    return c;
  }
  DQN_HD KCtx prepK(int k) const {      // k -> (th, tw, co)
    KCtx c; c.off = 0; c.offb = 0; c.t0 = c.t1 = c.t2 = 0;
    if (k < K) {
      uint32_t t, co, th, tw; g.fCout.divmod((uint32_t)k, t, co); ftw.divmod(t, th, tw); c.t0 = (int)th; c.t1 = (int)tw; c.t2 = (int)co;
      c.off = (long long)(int)co - (long long)((int)th * g.OW + (int)tw) * g.Cout;
      const int kh = ph + (int)th * g.S, kw = pw + (int)tw * g.S;
      c.offb = ((long long)kh * g.KW + kw) * g.Cin * g.Cout + (int)co;
    }
    return c;
  }
  DQN_HD float4 loadA(const ACtx& c, const KCtx& kc, int, int k) const {     // 4 consecutive co of one tap
    if (!c.valid || k >= K) return make4(0, 0, 0, 0);
    if (vecA) {
      if (!tap_ok(c, kc)) return make4(0, 0, 0, 0);
      return ldg4(D + c.base + kc.off);
    }
    float4 v = make4(0, 0, 0, 0);
    for (int j = 0; j < 4 && k + j < K; ++j) {
      const KCtx q = prepK(k + j);
      if (tap_ok(c, q)) set4(v, j, ldg1(D + c.base + q.off));
    }
    return v;
  }
  DQN_HD float4 loadB(const KCtx& kc, int k, int n) const {                  // 4 consecutive k (= co) at ci = n
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    if (vecB) return ldg4(W + kc.offb + (long long)(n * g.Cout));
    float4 v = make4(0, 0, 0, 0);
    for (int j = 0; j < 4 && k + j < K; ++j) {
      const KCtx q = prepK(k + j);
      set4(v, j, ldg1(W + q.offb + (long long)(n * g.Cout)));
    }
    return v;
  }
  DQN_HD long long out_off(int m, int n) const {
    uint32_t t, b, img, a; fbw.divmod((uint32_t)m, t, b); fah.divmod(t, img, a);
    return (((long long)img * g.IH + a * g.S + ph) * g.IW + b * g.S + pw) * g.Cin + n;
  }
  DQN_HD void store(int m, int n, float v) const {
    const long long o = out_off(m, n);
    if (apply_act) v *= act_deriv(Yprev[o], act);
    dX[o] = v;
  }
  DQN_HD bool can_store4() const { return g.Cin % 4 == 0; }
  DQN_HD void store4(int m, int n, float4 v) const {
    const long long o = out_off(m, n);
    if (apply_act) {
      const float4 y = *reinterpret_cast<const float4*>(Yprev + o);
      v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act);
    }
    *reinterpret_cast<float4*>(dX + o) = v;
    if (dXs) store_split4(dXs + o, lo_delta, v);
  }
  DQN_HD float4 epi_aux4(int m, int n) const { return apply_act ? ldg4(Yprev + out_off(m, n)) : make4(1.f, 1.f, 1.f, 1.f); }
  DQN_HD void store4x(int m, int n, float4 v, const float4& y) const {
    if (apply_act) { v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act); }
    *reinterpret_cast<float4*>(dX + out_off(m, n)) = v;
  }
  DQN_HD bool can_store8() const { return (g.Cin % 8 == 0) && al32(dX); }      // columns n..n+7 (n % 8 == 0) are channels of one pixel
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4& y, const float4& z) const {
    if (apply_act) {
      v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act);
      u.x *= act_deriv(z.x, act); u.y *= act_deriv(z.y, act); u.z *= act_deriv(z.z, act); u.w *= act_deriv(z.w, act);
    }
    st_global_v8(dX + out_off(m, n), v, u);
  }
};


// Conv dgrad with the stride-parity classes MERGED into the column dimension (S > 1, KH, KW, IH, IW multiples of S): every class
// (ph,pw) reads the same source pixels delta[n][a - th][b - tw][:] for its row (n, a, b) - so one contraction with
//   rows    m  = (n, a, b), a < IH/S, b < IW/S                     (one row per S x S block of input pixels)
//   k          = (th, tw, co), th < KH/S, tw < KW/S
//   columns n' = (ph, pw, ci):  dX[n][S a + ph][S b + pw][ci] = sum_k A(m,k) Wm[k][n'],  Wm[k][n'] = W[ph + S th][pw + S tw][ci][co]
// loads and splits the delta operand ONCE instead of S*S times and has S*S*Cin columns (conv2 of the Nature network: 128 instead of
// 4 launches' worth of 32).  Wm is a rearranged copy of the layer's weights, rebuilt by dgrad_merge_weights_kernel before the launch.
struct ConvDgradMergedOp {
  static constexpr bool HAS_A8 = false;
  static constexpr bool A_MCONTIG = false, B_KCONTIG = false, Z_IS_CLASS = false;
  const float* D; const float* Wm; float* dX; const float* Yprev; int act; int apply_act; int nimg; ConvGeom g;
  int AH, BW, TH, TW;
  FastDiv fbw, fah, ftw, fcin;
  int M, N, K;               // M = nimg*AH*BW, N = S*S*Cin, K = TH*TW*Cout
  int vecA, vecB;
  const float* Ds; const float* Ws; long long lo_delta; int a_single;     // Ws = Wm for the tensor-core path
  void init() {
    AH = g.IH / g.S; BW = g.IW / g.S; TH = g.KH / g.S; TW = g.KW / g.S;
    fbw.init(BW); fah.init(AH); ftw.init(TW); fcin.init(g.Cin);
    M = nimg * AH * BW; N = g.S * g.S * g.Cin; K = TH * TW * g.Cout;
  }
  static bool geometry_ok(const ConvGeom& g) {
    return g.S > 1 && g.S <= 4 && g.KH % g.S == 0 && g.KW % g.S == 0 && g.IH % g.S == 0 && g.IW % g.S == 0 && g.Cin % 4 == 0 && g.Cout % 4 == 0 &&
           (g.IH / g.S) == g.OH + g.KH / g.S - 1 && (g.IW / g.S) == g.OW + g.KW / g.S - 1;
  }
  DQN_HD bool tc_ready() const { return Ds && Ws && (g.Cout % 4 == 0) && (N % 4 == 0); }
  DQN_HD bool tap_ok(const ACtx& c, const KCtx& kc) const { return (unsigned)(c.i0 - kc.t0) < (unsigned)g.OH && (unsigned)(c.i1 - kc.t1) < (unsigned)g.OW; }
  DQN_HD const float* ptrA(const ACtx& c, const KCtx& kc, int, int k) const { return (c.valid && k < K && tap_ok(c, kc)) ? Ds + c.base + kc.off : nullptr; }
  DQN_HD const float* ptrB(const KCtx&, int k, int n) const { return (k < K && n < N) ? Ws + (long long)k * N + n : nullptr; }
  DQN_HD bool interiorA(int, int, int, int) const { return false; }
  DQN_HD bool interiorB(int n0, int k0, int bn, int bk) const { return n0 + bn <= N && k0 + bk <= K; }
  DQN_HD const float* ptrA_u(const ACtx& c, const KCtx& kc, int, int) const { return Ds + c.base + kc.off; }
  DQN_HD const float* ptrB_u(const KCtx&, int k, int n) const { return Ws + ((long long)k * N + n); }
  DQN_HD void set_class(int) {}
  DQN_HD ACtx prepA(int m) const {
    ACtx c; c.valid = m < M; c.base = 0; c.i0 = c.i1 = 0;
    if (c.valid) {
      uint32_t t, b, n, a; fbw.divmod((uint32_t)m, t, b); fah.divmod(t, n, a); c.i0 = (int)a; c.i1 = (int)b;
      c.base = (((long long)n * g.OH + (int)a) * g.OW + (int)b) * g.Cout;
    }
    return c;
  }
  DQN_HD KCtx prepK(int k) const {
    KCtx c; c.off = 0; c.offb = 0; c.t0 = c.t1 = c.t2 = 0;
    if (k < K) {
      uint32_t t, co, th, tw; g.fCout.divmod((uint32_t)k, t, co); ftw.divmod(t, th, tw); c.t0 = (int)th; c.t1 = (int)tw; c.t2 = (int)co;
      c.off = (long long)(int)co - (long long)((int)th * g.OW + (int)tw) * g.Cout;
    }
    return c;
  }
  DQN_HD float4 loadA(const ACtx& c, const KCtx& kc, int, int k) const {
    if (!c.valid || k >= K || !tap_ok(c, kc)) return make4(0, 0, 0, 0);
    return ldg4(D + c.base + kc.off);
  }
  DQN_HD float4 loadB(const KCtx&, int k, int n) const {
    if (k >= K || n >= N) return make4(0, 0, 0, 0);
    return ldg4(Wm + (long long)k * N + n);
  }
  DQN_HD long long out_off(int m, int n) const {
    uint32_t t, b, img, a, cls, ci; fbw.divmod((uint32_t)m, t, b); fah.divmod(t, img, a); fcin.divmod((uint32_t)n, cls, ci);
    const int ph = (int)cls / g.S, pw = (int)cls - ph * g.S;
    return (((long long)img * g.IH + a * g.S + ph) * g.IW + b * g.S + pw) * g.Cin + ci;
  }
  DQN_HD void store(int m, int n, float v) const {
    const long long o = out_off(m, n);
    if (apply_act) v *= act_deriv(Yprev[o], act);
    dX[o] = v;
  }
  DQN_HD bool can_store4() const { return g.Cin % 4 == 0; }
  DQN_HD void store4(int m, int n, float4 v) const { store4x(m, n, v, epi_aux4(m, n)); }
  DQN_HD float4 epi_aux4(int m, int n) const { return apply_act ? ldg4(Yprev + out_off(m, n)) : make4(1.f, 1.f, 1.f, 1.f); }
  DQN_HD void store4x(int m, int n, float4 v, const float4& y) const {
    if (apply_act) { v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act); }
    *reinterpret_cast<float4*>(dX + out_off(m, n)) = v;
  }
  DQN_HD bool can_store8() const { return (g.Cin % 8 == 0) && al32(dX); }      // columns n..n+7 (n % 8 == 0) are channels of one pixel
  DQN_HD void store8x(int m, int n, float4 v, float4 u, const float4& y, const float4& z) const {
    if (apply_act) {
      v.x *= act_deriv(y.x, act); v.y *= act_deriv(y.y, act); v.z *= act_deriv(y.z, act); v.w *= act_deriv(y.w, act);
      u.x *= act_deriv(z.x, act); u.y *= act_deriv(z.y, act); u.z *= act_deriv(z.z, act); u.w *= act_deriv(z.w, act);
    }
    st_global_v8(dX + out_off(m, n), v, u);
  }
};
#ifdef __CUDACC__
// Wm[(th*TW + tw)*Cout + co][(ph*S + pw)*Cin + ci] = W[(ph + S th)*KW + (pw + S tw)][ci][co]   (W: [(kh,kw)][Cin][Cout])
__global__ void dgrad_merge_weights_kernel(const float* __restrict__ W, float* __restrict__ Wm, int KH, int KW, int S, int Cin, int Cout) {
  const int TH = KH / S, TW = KW / S, N = S * S * Cin, K = TH * TW * Cout;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)K * N; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % N), k = (int)(e / N);
    const int ci = n % Cin, cls = n / Cin, ph = cls / S, pw = cls % S;
    const int co = k % Cout, tap = k / Cout, th = tap / TW, tw = tap % TW;
    Wm[e] = W[((long long)((ph + S * th) * KW + (pw + S * tw)) * Cin + ci) * Cout + co];
  }
}
#endif

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
    if (!Op::A_MCONTIG) {
      const KCtx kc = op.prepK(k0 + a_k[0]);                 // every A fragment of this thread sits in the same k chunk
#pragma unroll
      for (int i = 0; i < A_PER; ++i) ra[i] = op.loadA(actx[i], kc, m0 + a_m[i], k0 + a_k[i]);
    } else {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) { const KCtx kc = op.prepK(k0 + a_k[i]); ra[i] = op.loadA(actx[i], kc, m0 + a_m[i], k0 + a_k[i]); }
    }
    KCtx kb; kb.off = 0; kb.offb = 0; kb.t0 = kb.t1 = kb.t2 = 0;
    if (Op::B_KCONTIG) kb = op.prepK(k0 + b_k[0]);
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int e = tid + i * NT;
      rb[i] = (e < B_F4) ? op.loadB(kb, k0 + b_k[i], n0 + b_n[i]) : make4(0, 0, 0, 0);
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
  const bool v4 = (nsplit == 1) && op.can_store4();
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= op.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const int n = n0 + tx * TN + j;
      if (v4 && n + 3 < op.N) { op.store4(m, n, make4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3])); continue; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (n + q >= op.N) continue;
        if (nsplit > 1) ws[(long long)blockIdx.z * ws_stride + (long long)m * op.N + n + q] = acc[i][j + q];
        else op.store(m, n + q, acc[i][j + q]);
      }
    }
  }
}

// out[i] = sum_s ws[(z*nsplit + s)][i] (deterministic: fixed order), the epilogue of every split-K contraction.
// sub = 1: one thread per four columns, the splits added in ascending order.  sub = 4 (small outputs with many splits - the conv weight
// gradients: 8 CTAs x 64 dependent-ish loads was 11.6 us on the tail of the step): four adjacent lanes share an output, each adds a
// contiguous quarter of the splits in ascending order, the quarters meet as (q0 + q1) + (q2 + q3) by two butterfly shuffles.
template <class Op>
__global__ void splitk_reduce_kernel(Op opa, Op opb, int nsplit, const float* __restrict__ ws, long long ws_stride, int m_off = 0, int rows = -1, int sub = 1) {
  const int zi = blockIdx.y;
  Op op = zi == 0 ? opa : opb;
  if (Op::Z_IS_CLASS) op.set_class(0);                    // only single-class dgrads are ever split
  // rows [m_off, m_off + rows) of the output (the tail tiles of a launch, see tc_gemm_impl.cuh); default: all of it
  const long long total = (long long)(rows < 0 ? op.M : rows) * op.N;
  if (op.can_store4()) {                     // N % 4 == 0: four columns per thread, vector epilogue (also writes the split planes)
    const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const int q = (int)(gt % sub);
    const int k0 = (int)((long long)q * nsplit / sub), k1 = (int)((long long)(q + 1) * nsplit / sub);
    // (the loop bound is uniform over the `sub` lanes of an output and the shuffles are predicated per output group, not per warp)
    for (long long i = (gt / sub) * 4; i - (long long)((threadIdx.x & 31) / sub) * 4 < total; i += (nth / sub) * 4) {
      const bool valid = i < total;
      float4 s = make4(0, 0, 0, 0);
      if (valid) {
        int k = k0;
        for (; k + 8 <= k1; k += 8) {          // eight independent loads in flight (small grids are latency bound), added in ascending order
          float4 p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = __ldcg(reinterpret_cast<const float4*>(ws + ((long long)(zi * nsplit + k + j)) * ws_stride + i));
#pragma unroll
          for (int j = 0; j < 8; ++j) { s.x += p[j].x; s.y += p[j].y; s.z += p[j].z; s.w += p[j].w; }
        }
        for (; k < k1; ++k) {
          const float4 p = __ldcg(reinterpret_cast<const float4*>(ws + ((long long)(zi * nsplit + k)) * ws_stride + i));
          s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
      }
      if (sub == 4) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
          s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
        }
      }
      if (valid && q == 0) op.store4(m_off + (int)(i / op.N), (int)(i % op.N), s);
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += ws[((long long)(zi * nsplit + k)) * ws_stride + i];
    op.store(m_off + (int)(i / op.N), (int)(i % op.N), s);
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
            int mb = (m / 4) * 4; ACtx c = op.prepA(mb); KCtx kc = op.prepK(k4 + j); float4 v = op.loadA(c, kc, mb, k4 + j); av[j] = (&v.x)[m - mb];
          }
        } else { ACtx c = op.prepA(m); KCtx kc = op.prepK(k4); float4 v = op.loadA(c, kc, m, k4); av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w; }
        if (Op::B_KCONTIG) { KCtx kc = op.prepK(k4); float4 v = op.loadB(kc, k4, n); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
        else { for (int j = 0; j < 4; ++j) { int nb = (n / 4) * 4; KCtx kc = op.prepK(k4 + j); float4 v = op.loadB(kc, k4 + j, nb); bv[j] = (&v.x)[n - nb]; } }
        for (int j = 0; j < 4; ++j) acc += (double)av[j] * (double)bv[j];
      }
      op.store(m, n, (float)acc);
    }
  }
}

}  // namespace dqn
