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

}  // namespace dqn
