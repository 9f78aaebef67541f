// lstm.cuh - the recurrent half of the path: EpisodeReplayBuffer gather and the LSTM cell, forward and BPTT
// (reference: src/solver.jl:239-287, src/episode_replay.jl:21-95, Flux 0.14 LSTMCell as restated in SURVEY App. B.3).
//
// The contractions of the recurrent step that are batched over time - input projection x W_i^T + b for all T*B rows, the Dense
// heads, all weight gradients - go through the same contraction kernels as the feed-forward path (igemm.cuh / tc_gemm_impl.cuh).
// What is sequential by nature stays here: one small kernel per time step for the recurrence h_{t-1} W_h^T (forward) and its
// transpose (BPTT), fused with the gate non-linearities and their derivatives.  Gate order of Flux: input | forget | cell | output.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace dqn {

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

constexpr int LSTM_TB = 8;        // batch rows per CTA of the step kernels (CTA = 32 hidden units x 8 rows)

// rows of a [rows][H] state matrix <- the (1, H) initial state (Flux.reset!: state = state0, broadcast over the batch)
__global__ void lstm_broadcast_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int H) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < (long long)rows * H) dst[i] = src[i % H];
}

// Wt[n][k] = W[k][n]  (W is [K][N]); the BPTT step reads W_h^T with the same coalesced pattern the forward step reads W_h
__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int K, int N) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) { const int k = k0 + r, n = n0 + threadIdx.x; tile[r][threadIdx.x] = (k < K && n < N) ? W[(long long)k * N + n] : 0.f; }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) { const int n = n0 + r, k = k0 + threadIdx.x; if (n < N && k < K) Wt[(long long)n * K + k] = tile[threadIdx.x][r]; }
}

// One forward time step for up to two independent chains (blockIdx.z: e.g. the s pass and the s' pass of the online network):
//   g = xproj_t + h_prev W_h   (xproj already holds x W_i^T + b);  c = sigma(f) c_prev + sigma(i) tanh(g_c);  h = sigma(o) tanh(c)
// grid = (H / 32, ceil(B / LSTM_TB), chains), block = (32, LSTM_TB): thread = (hidden unit j, batch row b), four gate dot products.
struct LstmFwdArgs {
  const float* xproj[2];    // [B][4H] of this step
  const float* h_prev[2];   // [B][H]
  const float* c_prev[2];
  float* h_out[2];
  float* c_out[2];
  float* gates[2];          // [B][4H] activated gates (sigma(i), sigma(f), tanh(g), sigma(o)) kept for BPTT, or null
  const float* Wh;          // [H][4H]
  int B, H;
};
__global__ void __launch_bounds__(32 * LSTM_TB) lstm_fwd_step_kernel(const LstmFwdArgs a) {
  extern __shared__ float hs[];                         // [LSTM_TB][H] rows of h_prev
  const int z = blockIdx.z, j = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y * LSTM_TB + threadIdx.y;
  const int H = a.H, N = 4 * H;
  for (int i = threadIdx.y * 32 + threadIdx.x; i < LSTM_TB * H; i += 32 * LSTM_TB) {
    const int bb = blockIdx.y * LSTM_TB + i / H;
    hs[i] = bb < a.B ? a.h_prev[z][(long long)bb * H + (i % H)] : 0.f;
  }
  __syncthreads();
  if (j >= H || b >= a.B) return;
  const float* xp = a.xproj[z] + (long long)b * N;
  float gi = xp[j], gf = xp[H + j], gc = xp[2 * H + j], go = xp[3 * H + j];
  const float* hrow = hs + threadIdx.y * H;
  const float* w = a.Wh + j;
#pragma unroll 4
  for (int k = 0; k < H; ++k) {
    const float hk = hrow[k];
    const float* wk = w + (long long)k * N;
    gi = fmaf(hk, __ldg(wk), gi); gf = fmaf(hk, __ldg(wk + H), gf); gc = fmaf(hk, __ldg(wk + 2 * H), gc); go = fmaf(hk, __ldg(wk + 3 * H), go);
  }
  const float si = sigmoidf_(gi), sf = sigmoidf_(gf), tg = tanhf(gc), so = sigmoidf_(go);
  const float c = sf * a.c_prev[z][(long long)b * H + j] + si * tg;
  a.c_out[z][(long long)b * H + j] = c;
  a.h_out[z][(long long)b * H + j] = so * tanhf(c);
  if (a.gates[z]) { float* g = a.gates[z] + (long long)b * N; g[j] = si; g[H + j] = sf; g[2 * H + j] = tg; g[3 * H + j] = so; }
}

// One BPTT step (time t, walking backwards):  dh = dh_out_t + dgates_{t+1} W_h^T,  then the cell's reverse pass
//   do = dh tanh(c_t);  dc = dc_next + dh o (1 - tanh(c_t)^2);  di = dc g;  df = dc c_{t-1};  dg = dc i;  dc_prev = dc f
//   pre-activation gradients: di i(1-i) | df f(1-f) | dg (1-g^2) | do o(1-o)   -> dgates_t [B][4H]
struct LstmBwdArgs {
  const float* dh_out;      // [B][H] gradient from the heads into h_t
  const float* dg_next;     // [B][4H] dgates of step t+1, or null at t = T-1
  const float* WhT;         // [4H][H]
  const float* gates;       // [B][4H] activated gates of step t
  const float* c_prev;      // c_{t-1}
  const float* c_cur;       // c_t
  float* dc;                // [B][H] in: dc_next, out: dc_prev (in place)
  float* dgates;            // [B][4H] out
  int B, H, first;          // first: dc_next is zero (t = T-1)
};
__global__ void __launch_bounds__(32 * LSTM_TB) lstm_bwd_step_kernel(const LstmBwdArgs a) {
  extern __shared__ float ds[];                         // [LSTM_TB][4H] rows of dgates_{t+1}
  const int j = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y * LSTM_TB + threadIdx.y;
  const int H = a.H, N = 4 * H;
  if (a.dg_next) {
    for (int i = threadIdx.y * 32 + threadIdx.x; i < LSTM_TB * N; i += 32 * LSTM_TB) {
      const int bb = blockIdx.y * LSTM_TB + i / N;
      ds[i] = bb < a.B ? a.dg_next[(long long)bb * N + (i % N)] : 0.f;
    }
  }
  __syncthreads();
  if (j >= H || b >= a.B) return;
  float dh = a.dh_out[(long long)b * H + j];
  if (a.dg_next) {
    const float* drow = ds + threadIdx.y * N;
    const float* w = a.WhT + j;
    float acc = 0.f;
#pragma unroll 4
    for (int n = 0; n < N; ++n) acc = fmaf(drow[n], __ldg(w + (long long)n * H), acc);
    dh += acc;
  }
  const float* g = a.gates + (long long)b * N;
  const float si = g[j], sf = g[H + j], tg = g[2 * H + j], so = g[3 * H + j];
  const float tc = tanhf(a.c_cur[(long long)b * H + j]);
  const float dc = (a.first ? 0.f : a.dc[(long long)b * H + j]) + dh * so * (1.f - tc * tc);
  float* o = a.dgates + (long long)b * N;
  o[j] = dc * tg * si * (1.f - si);
  o[H + j] = dc * a.c_prev[(long long)b * H + j] * sf * (1.f - sf);
  o[2 * H + j] = dc * si * (1.f - tg * tg);
  o[3 * H + j] = dh * tc * so * (1.f - so);
  a.dc[(long long)b * H + j] = dc * sf;
}

// EpisodeReplayBuffer.sample (src/episode_replay.jl:71-95) after the episode indices are drawn: per sampled episode its start offset
// ep_start = 1 + floor(u * len) (u = Philox(seed; slot, attempt 0x40000000, call)), then the trace copy WITH the reference's quirk -
// ep_start only shortens the trace: steps ep[1], ep[2], ... are copied for j = ep_start : min(len, T).  Unfilled steps are zeros,
// action 1, mask 0 (reset_batches!).  Batch rows are time-major: row t * B + i.   grid = (T, B), block = 128.
__global__ void episode_gather_kernel(const long long* __restrict__ idx, int B, int T, int L, long long d, uint64_t seed, const DevState* __restrict__ st,
                                      int use_call, uint64_t call_in, const int* __restrict__ ep_len, const float* __restrict__ ep_s,
                                      const float* __restrict__ ep_sp, const int* __restrict__ ep_a, const float* __restrict__ ep_r,
                                      const uint8_t* __restrict__ ep_done, float* __restrict__ xs, float* __restrict__ xsp, int* __restrict__ a_b,
                                      float* __restrict__ r_b, float* __restrict__ d_b, float* __restrict__ m_b, int* __restrict__ start_out) {
  const int t = blockIdx.x, i = blockIdx.y;
  const long long e = idx[i];
  const int len = ep_len[e];
  const uint64_t call = use_call ? call_in : st->sample_call;
  const float u = philox_uniform(seed, call, (uint32_t)i, 0x40000000u);
  int start = 1 + (int)__fmul_rn(u, (float)len);
  if (start > len) start = len;
  const int n = min(len, T) - start + 1;                // steps copied (may be <= 0)
  const bool live = t < n;
  const long long row = (long long)t * B + i;
  const float* s = ep_s + ((long long)e * L + t) * d;
  const float* sp = ep_sp + ((long long)e * L + t) * d;
  for (long long k = threadIdx.x; k < d; k += blockDim.x) {
    xs[row * d + k] = live ? s[k] : 0.f;
    xsp[row * d + k] = live ? sp[k] : 0.f;
  }
  if (threadIdx.x == 0) {
    a_b[row] = live ? ep_a[(long long)e * L + t] : 1;
    r_b[row] = live ? ep_r[(long long)e * L + t] : 0.f;
    d_b[row] = (live && ep_done[(long long)e * L + t]) ? 1.f : 0.f;
    m_b[row] = live ? 1.f : 0.f;
    if (t == 0 && start_out) start_out[i] = start;
  }
}


// =================================================================================================================================
// The whole recurrence in ONE launch (thread-block clusters + distributed shared memory).
//
// The per-step kernels above cost a launch and a cold pass over W_h per time step: 3 x 32 dependent launches of ~25-35 us bound the
// recurrent step (bench.py --workload drqn: 2.7 of 3.0 ms).  The recurrence of a batch row does not depend on other rows, and W_h is
// small, so: a cluster of 8 CTAs takes 8 batch rows through all T steps; CTA r of the cluster owns hidden units [r H/8, (r+1) H/8) -
// its slice of W_h (H x H/2 gate columns) lives in REGISTERS (thread = one gate column x one quarter of k: H/4 weights), the 8 x H
// block of h_{t-1} in shared memory.  Per step: partial dot products from registers x shared-memory broadcasts, a 4-lane butterfly,
// the cell update by the lanes that own gate i (the other three gates arrive by shuffle: the four gates of a unit sit in one
// half-warp), then every CTA writes its slice of h_t into the shared memory of ALL eight CTAs (st.shared::cluster) and the cluster
// barrier separates the steps.  c stays in registers for the whole sequence.  Same arithmetic order per output as the per-step
// kernels up to the association of the k sum (quarters, then butterfly), i.e. fp32 rounding-level differences only.
//   grid = (8, ceil(B / 8), chains), cluster (8, 1, 1), block = 2 H threads (H = 64, 128, 256).
// =================================================================================================================================
constexpr int LSTM_CL = 8;        // CTAs per cluster = slices of the hidden units
constexpr int LSTM_RB = 8;        // batch rows per cluster

__device__ __forceinline__ uint32_t lstm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void cluster_arrive_() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_() { cluster_arrive_(); cluster_wait_(); }
__device__ __forceinline__ uint32_t cluster_rank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct LstmSeqFwdArgs {
  const float* xproj[2];    // [T][B][4H]  x W_i^T + b of every step of the chain
  const float* h0[2];       // (H) initial hidden state (state0, broadcast over the batch)
  const float* c0[2];
  float* hs[2];             // [T][B][H] out: h after step t (block t)
  float* cs[2];             // [T][B][H] out: c after step t, or null (only the chain kept for BPTT needs it)
  float* gates[2];          // [T][B][4H] activated gates kept for BPTT, or null
  const float* Wh;          // [H][4H]
  int B, H, T;
};

template <int H>
__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(2 * H, 1) lstm_seq_fwd_kernel(const LstmSeqFwdArgs a) {
  constexpr int U = H / LSTM_CL;          // hidden units of this CTA
  constexpr int KQ = H / 4;               // k values per thread (one quarter of the dot product), as KQ / 4 chunks of four
  __shared__ __align__(16) float hbuf[2][LSTM_RB][H];
  const int z = blockIdx.z, tid = threadIdx.x, q = tid & 3, col = tid >> 2;        // col = unit_local * 4 + gate
  const int g = col & 3, ul = col >> 2;
  const uint32_t rank = cluster_rank_();
  const int unit = (int)rank * U + ul;
  const int gcol = g * H + unit;                                                    // column of the [.][4H] matrices (Flux order i|f|g|o)
  const int row0 = blockIdx.y * LSTM_RB;
  // this thread's weights: W_h[k][gcol] for k in chunks j = 4 i + q (interleaved so that the four quarter-lanes read 64 contiguous bytes of h)
  float w[KQ];
#pragma unroll
  for (int i = 0; i < KQ / 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) w[4 * i + e] = a.Wh[(long long)(4 * (4 * i + q) + e) * (4 * H) + gcol];
  for (int i = tid; i < LSTM_RB * H; i += 2 * H) hbuf[0][i / H][i % H] = a.h0[z][i % H];
  // rows finalised by this lane: q and q + 4 of the cluster's eight
  float c_reg[2];
  c_reg[0] = c_reg[1] = a.c0[z][unit];
  const float* xpz = a.xproj[z];
  float* hsz = a.hs[z]; float* csz = a.cs[z]; float* gtz = a.gates[z];
  __syncthreads();
  cluster_sync_();                                                                  // every CTA of the cluster is running before remote stores
  const uint32_t hb_addr = lstm_smem_u32(&hbuf[0][0][0]);
  for (int t = 0; t < a.T; ++t) {
    const float (*h)[H] = hbuf[t & 1];
    float xp[2];                                                                    // this step's input projection: in flight during the dot products
#pragma unroll
    for (int j = 0; j < 2; ++j) { const int b = row0 + q + 4 * j; xp[j] = b < a.B ? __ldg(xpz + ((long long)t * a.B + b) * (4 * H) + gcol) : 0.f; }
    if (t > 0) cluster_wait_();                                                     // h_{t-1} of every slice has landed (arrive: end of the previous step)
    float s[LSTM_RB];
#pragma unroll
    for (int r = 0; r < LSTM_RB; ++r) s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < KQ / 4; ++i) {
#pragma unroll
      for (int r = 0; r < LSTM_RB; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(&h[r][4 * (4 * i + q)]);
        s[r] = fmaf(hv.x, w[4 * i], s[r]); s[r] = fmaf(hv.y, w[4 * i + 1], s[r]); s[r] = fmaf(hv.z, w[4 * i + 2], s[r]); s[r] = fmaf(hv.w, w[4 * i + 3], s[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < LSTM_RB; ++r) { s[r] += __shfl_xor_sync(0xffffffffu, s[r], 1); s[r] += __shfl_xor_sync(0xffffffffu, s[r], 2); }
    // pre-activations of rows q and q + 4, then the four gates of the unit meet in the gate-i lanes
    float pre[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float sv = j == 0 ? (q == 0 ? s[0] : q == 1 ? s[1] : q == 2 ? s[2] : s[3]) : (q == 0 ? s[4] : q == 1 ? s[5] : q == 2 ? s[6] : s[7]);
      pre[j] = xp[j] + sv;
    }
    const int base = (threadIdx.x & 31) & ~15;                                      // first lane of this unit's half-warp
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      // every lane applies ITS gate's non-linearity (sigma for i, f, o; tanh for the cell candidate), then the gate-i lanes collect the other three
      const float act = g == 2 ? tanhf(pre[j]) : sigmoidf_(pre[j]);
      const float sf = __shfl_sync(0xffffffffu, act, base + 4 + q), tg = __shfl_sync(0xffffffffu, act, base + 8 + q), so = __shfl_sync(0xffffffffu, act, base + 12 + q);
      const int rl = q + 4 * j;
      if (g == 0) {
        const float si = act;
        const float c = sf * c_reg[j] + si * tg;
        const float hn = so * tanhf(c);
        c_reg[j] = c;
        const uint32_t off = hb_addr + (uint32_t)((((t + 1) & 1) * LSTM_RB + rl) * H + unit) * 4u;
#pragma unroll
        for (uint32_t rr = 0; rr < LSTM_CL; ++rr) st_cluster_f32(mapa_u32(off, rr), hn);
        // (stored here, before the arrive: deferring these global stores past it - a release waits for every earlier store - made the
        //  one-chain launch 5 % faster and the two-chain launch 25 % slower; measured twice, kept simple)
        const int b = row0 + rl;
        if (b < a.B) {
          const long long o = ((long long)t * a.B + b) * H + unit;
          hsz[o] = hn;
          if (csz) csz[o] = c;
          if (gtz) { float* gp = gtz + ((long long)t * a.B + b) * (4 * H); gp[unit] = si; gp[H + unit] = sf; gp[2 * H + unit] = tg; gp[3 * H + unit] = so; }
        }
      }
    }
    __syncwarp();
    cluster_arrive_();
  }
  cluster_wait_();                                                                  // nobody leaves while a peer may still write into its shared memory
}

// BPTT over the whole sequence, same decomposition: CTA r owns the hidden units [r H/8, (r+1) H/8): their dh (the column slice of
// dgates_{t+1} W_h^T: K = 4H, thread = unit x one sixteenth of k, H/4 weights in registers), their cell's reverse pass, their four gate
// gradients - written to every CTA's copy of dgates_t (the operand of step t-1) and to global memory (the weight-gradient operand).
struct LstmSeqBwdArgs {
  const float* dh_out;      // [T][B][H] gradient from the heads into h_t
  const float* Wh;          // [H][4H]
  const float* gates;       // [T][B][4H] activated gates
  const float* cs;          // [T][B][H] c after step t
  const float* c0;          // (H)
  float* dgates;            // [T][B][4H] out
  int B, H, T;
};

template <int H>
__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(2 * H, 1) lstm_seq_bwd_kernel(const LstmSeqBwdArgs a) {
  constexpr int U = H / LSTM_CL;          // units of this CTA
  constexpr int KP = 16;                  // parts of k = 4H per unit -> threads = U * 16 = 2 H
  constexpr int KQ = 4 * H / KP;          // k values per thread = H / 4
  extern __shared__ __align__(16) float dbuf_raw[];                                 // [2][LSTM_RB][4H]: 32 KB at H = 128, 64 KB at H = 256
  float (*dbuf)[LSTM_RB][4 * H] = reinterpret_cast<float (*)[LSTM_RB][4 * H]>(dbuf_raw);
  const int tid = threadIdx.x, p = tid & 15, ul = tid >> 4;
  const uint32_t rank = cluster_rank_();
  const int unit = (int)rank * U + ul;
  const int row0 = blockIdx.y * LSTM_RB;
  // W_h^T[n][unit] = W_h[unit][n] for n in chunks j = 16 i + p
  float w[KQ];
#pragma unroll
  for (int i = 0; i < KQ / 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) w[4 * i + e] = a.Wh[(long long)unit * (4 * H) + 4 * (16 * i + p) + e];
  float dc_reg = 0.f;                                                               // lanes p < 8 carry dc of row p
  cluster_sync_();
  const uint32_t db_addr = lstm_smem_u32(&dbuf[0][0][0]);
  for (int t = a.T - 1; t >= 0; --t) {
    // this lane's row operands (lanes p < 8: row p of the cluster's eight): in flight during the dot products
    float dho = 0.f, si = 0.f, sf = 0.f, tg = 0.f, so = 0.f, ccur = 0.f, cprev = 0.f;
    const int b = row0 + p;
    const bool own = p < LSTM_RB && b < a.B;
    const long long o = ((long long)t * a.B + b) * H + unit;
    if (own) {
      const float* gp = a.gates + ((long long)t * a.B + b) * (4 * H);
      dho = __ldg(a.dh_out + o); si = __ldg(gp + unit); sf = __ldg(gp + H + unit); tg = __ldg(gp + 2 * H + unit); so = __ldg(gp + 3 * H + unit);
      ccur = __ldg(a.cs + o); cprev = t == 0 ? __ldg(a.c0 + unit) : __ldg(a.cs + o - (long long)a.B * H);
    }
    if (t < a.T - 1) cluster_wait_();                                               // dgates_{t+1} of every slice has landed
    float s[LSTM_RB];
#pragma unroll
    for (int r = 0; r < LSTM_RB; ++r) s[r] = 0.f;
    if (t < a.T - 1) {
      const float (*d)[4 * H] = dbuf[(t + 1) & 1];
#pragma unroll
      for (int i = 0; i < KQ / 4; ++i) {
#pragma unroll
        for (int r = 0; r < LSTM_RB; ++r) {
          const float4 dv = *reinterpret_cast<const float4*>(&d[r][4 * (16 * i + p)]);
          s[r] = fmaf(dv.x, w[4 * i], s[r]); s[r] = fmaf(dv.y, w[4 * i + 1], s[r]); s[r] = fmaf(dv.z, w[4 * i + 2], s[r]); s[r] = fmaf(dv.w, w[4 * i + 3], s[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < LSTM_RB; ++r) {
        s[r] += __shfl_xor_sync(0xffffffffu, s[r], 1); s[r] += __shfl_xor_sync(0xffffffffu, s[r], 2);
        s[r] += __shfl_xor_sync(0xffffffffu, s[r], 4); s[r] += __shfl_xor_sync(0xffffffffu, s[r], 8);
      }
    }
    float di = 0.f, df = 0.f, dg = 0.f, dO = 0.f;
    if (p < LSTM_RB) {
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < LSTM_RB; ++r) if (p == r) acc = s[r];
      if (own) {
        const float dh = dho + acc;
        const float tc = tanhf(ccur);
        const float dc = dc_reg + dh * so * (1.f - tc * tc);
        di = dc * tg * si * (1.f - si); df = dc * cprev * sf * (1.f - sf); dg = dc * si * (1.f - tg * tg); dO = dh * tc * so * (1.f - so);
        dc_reg = dc * sf;
      }
      const uint32_t off = db_addr + (uint32_t)(((t & 1) * LSTM_RB + p) * (4 * H) + unit) * 4u;
#pragma unroll
      for (uint32_t rr = 0; rr < LSTM_CL; ++rr) {
        const uint32_t ra = mapa_u32(off, rr);
        st_cluster_f32(ra, di); st_cluster_f32(ra + (uint32_t)H * 4u, df); st_cluster_f32(ra + (uint32_t)(2 * H) * 4u, dg); st_cluster_f32(ra + (uint32_t)(3 * H) * 4u, dO);
      }
    }
    __syncwarp();
    cluster_arrive_();                                                              // (.aligned: the whole warp, converged) the global stores ride behind it
    if (own) {
      float* op = a.dgates + ((long long)t * a.B + b) * (4 * H);
      op[unit] = di; op[H + unit] = df; op[2 * H + unit] = dg; op[3 * H + unit] = dO;
    }
  }
  cluster_wait_();
}

inline bool lstm_seq_supported(int H) { return H == 64 || H == 128 || H == 256; }
inline size_t lstm_seq_bwd_smem(int H) { return (size_t)2 * LSTM_RB * 4 * H * sizeof(float); }

template <int H> inline cudaError_t lstm_seq_fwd_launch_t(const LstmSeqFwdArgs& a, int nch, cudaStream_t st) {
  lstm_seq_fwd_kernel<H><<<dim3(LSTM_CL, (a.B + LSTM_RB - 1) / LSTM_RB, nch), 2 * H, 0, st>>>(a);
  return cudaGetLastError();
}
template <int H> inline cudaError_t lstm_seq_bwd_launch_t(const LstmSeqBwdArgs& a, cudaStream_t st) {
  static bool attr[64] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 64 && !attr[dev]) {
    cudaError_t r = cudaFuncSetAttribute(lstm_seq_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_seq_bwd_smem(H));
    if (r != cudaSuccess) return r;
    attr[dev] = true;
  }
  lstm_seq_bwd_kernel<H><<<dim3(LSTM_CL, (a.B + LSTM_RB - 1) / LSTM_RB), 2 * H, lstm_seq_bwd_smem(H), st>>>(a);
  return cudaGetLastError();
}
inline cudaError_t lstm_seq_fwd_launch(const LstmSeqFwdArgs& a, int nch, cudaStream_t st) {
  return a.H == 64 ? lstm_seq_fwd_launch_t<64>(a, nch, st) : a.H == 128 ? lstm_seq_fwd_launch_t<128>(a, nch, st) : lstm_seq_fwd_launch_t<256>(a, nch, st);
}
inline cudaError_t lstm_seq_bwd_launch(const LstmSeqBwdArgs& a, cudaStream_t st) {
  return a.H == 64 ? lstm_seq_bwd_launch_t<64>(a, st) : a.H == 128 ? lstm_seq_bwd_launch_t<128>(a, st) : lstm_seq_bwd_launch_t<256>(a, st);
}

}  // namespace dqn
