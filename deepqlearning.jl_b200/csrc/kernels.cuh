// kernels.cuh - the HBM-bound / latency-bound kernels of the step: sum-tree sampling and refresh,
// transition gather, fused dueling + Double-Q + Bellman target + IS-Huber head, fused Adam.
// Reference lines are cited per kernel (PER = src/prioritized_experience_replay.jl, SOLVER = src/solver.jl).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "igemm.cuh"

namespace dqn {

// Device-resident scalar state (everything a captured graph must read or advance without new arguments).
struct DevState {
  long long curr_size;            // r._curr_size                      PER:26
  unsigned long long sample_call; // Philox counter word: number of sampling calls so far
  double b1p, b2p;                // Adam running beta powers, Float64
  unsigned int gradmax_bits;      // max |g| of this step as fp32 bits (globalnorm, HELPERS:38-46)
  float loss;                     // loss_val SOLVER:223-224
  int error;                      // sticky error flags raised by kernels (bit 0: sampler did not converge,
                                  //   bit 1: non-positive priority PER:78, bit 2: td0+eps <= 0 PER:66, bit 3: bad action)
  int pad;
  unsigned long long step;        // gradient steps finished so far: step s publishes its scalars to host slot s & 1  (64-bit: a 32-bit
                                  //   counter wraps after 18 days at 2700 steps/s, and the epochs below must never decrease)
  unsigned long long tree_epoch;  // s once step s has gathered its rows and refreshed its priorities: from then on it touches neither the
                                  //   replay rows nor the tree, and new transitions may be ingested beside the rest of it (epoch_wait_kernel)
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11); oracle/philox.py restates it and checks the Random123 vectors.
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__host__ __device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t call, uint32_t slot, uint32_t attempt) {
  uint32_t c[4] = {slot, attempt, (uint32_t)call, (uint32_t)(call >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (float)(c[0] >> 8) * 5.9604644775390625e-08f;   // 2^-24, exact
}

// Float32 ^ Float32 through Float64, rounded once (the reference's Float32 pow; oracle pow_f32)
__device__ __forceinline__ float pow_f32(float x, float y) { return (float)pow((double)x, (double)y); }

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Sum-tree.  tree[1] root, children of k at 2k, 2k+1, leaves at P + i.  Internal nodes are always
// recomputed as fl(left + right) - the tree is a pure function of its leaves (bit-exact vs oracle/sumtree.py).
constexpr int TREE_TOP = 1024;            // nodes 1..1023 (the first 10 levels) are staged in shared memory by the sampler
// One binary step of the descent: children (l, r) of the current node.
__device__ __forceinline__ void tree_step(float& v, int& node, float l, float r) {
  const bool left = (v < l) || (r == 0.f);
  if (!left) v = __fsub_rn(v, l);
  node = 2 * node + (left ? 0 : 1);
}
// Root-to-leaf descent.  Below the staged top every level would be one dependent L2 round trip (11 of them at P = 2^20); internal nodes
// are always fl(left + right) of their children (tree_update / tree_level kernels), so the children AND grandchildren sums of a node
// can be rebuilt bit-exactly from its eight great-grandchildren - one 32-byte load serves three levels.
__device__ __forceinline__ int tree_descend(const float* __restrict__ tree, const float* __restrict__ top, int P, float u) {
  float v = __fmul_rn(u, top ? top[1] : tree[1]);
  int node = 1;
  while (node < P) {
    if (top && 2 * node + 1 < TREE_TOP) {
      const float2 ch = *reinterpret_cast<const float2*>(top + 2 * node);
      tree_step(v, node, ch.x, ch.y);
    } else if (8LL * node < 2LL * P) {                          // three levels from the eight great-grandchildren
      const float4 a = *reinterpret_cast<const float4*>(tree + 8LL * node), b = *reinterpret_cast<const float4*>(tree + 8LL * node + 4);
      const float g0 = __fadd_rn(a.x, a.y), g1 = __fadd_rn(a.z, a.w), g2 = __fadd_rn(b.x, b.y), g3 = __fadd_rn(b.z, b.w);
      const int n0 = node;
      tree_step(v, node, __fadd_rn(g0, g1), __fadd_rn(g2, g3));
      const bool hi = node != 2 * n0;
      tree_step(v, node, hi ? g2 : g0, hi ? g3 : g1);
      const int q = node - 4 * n0;                              // grandchild 0..3
      const float l = q == 0 ? a.x : q == 1 ? a.z : q == 2 ? b.x : b.z, r = q == 0 ? a.y : q == 1 ? a.w : q == 2 ? b.y : b.w;
      tree_step(v, node, l, r);
    } else {
      const float2 ch = *reinterpret_cast<const float2*>(tree + 2 * node);
      tree_step(v, node, ch.x, ch.y);
    }
  }
  return node - P;
}

// Mass of `node` once the leaves held by earlier slots are taken out: fl(tree[node] - sum of the held priorities below it), the
// held leaves added in slot order (fixed order: bit-exact against oracle/sumtree.py::descend_excl).  shift = levels below `node`.
__device__ __forceinline__ float tree_mass_excl(const float* __restrict__ tree, int P, int node, int shift, const int* held, int B) {
  float ex = 0.f;
  for (int i = 0; i < B; ++i) { const int l = held[i]; if (l >= 0 && ((P + l) >> shift) == node) ex = __fadd_rn(ex, tree[P + l]); }
  const float m = __fsub_rn(tree[node], ex);
  return m > 0.f ? m : 0.f;
}
// Descent over the tree minus the held leaves: one exact draw of successive sampling without replacement.
__device__ int tree_descend_excl(const float* __restrict__ tree, int P, float u, const int* held, int B) {
  int depth = 0; while ((1 << depth) < P) ++depth;
  float v = __fmul_rn(u, tree_mass_excl(tree, P, 1, depth, held, B));
  int node = 1;
  for (int shift = depth - 1; shift >= 0; --shift) {
    const float l = tree_mass_excl(tree, P, 2 * node, shift, held, B), r = tree_mass_excl(tree, P, 2 * node + 1, shift, held, B);
    const bool left = (v < l) || (r == 0.f);
    if (!left) v = __fsub_rn(v, l);
    node = 2 * node + (left ? 0 : 1);
  }
  return node - P;
}

// B draws without replacement (StatsBase.sample(...; replace=false), PER:85): slot j redraws while a slot
// i < j holds the same leaf; one check per round over a shared hash table.  One CTA, B <= 1024 threads.
// Rejection is exactly successive sampling without replacement, but it only terminates quickly while the held leaves carry a small
// share of the priority mass.  The reference's sampler always succeeds once curr_size >= batch_size, so slots that are still rejected
// after SAMPLE_MAX_ROUNDS redraws (curr_size barely above B, or a few leaves holding nearly all the mass) are resolved exactly: in slot
// order, each by a descent over the tree with the held leaves' mass taken out (attempt numbers 0x80000000 + t name those uniforms).
// The same CTA then gathers the per-sample metadata and importance weights (get_batch, PER:94-102).
constexpr int SAMPLE_MAX_ROUNDS = 64;
constexpr int SAMPLE_EXACT_TRIES = 4096;
__global__ void sample_kernel(const float* __restrict__ tree, int P, int B, uint64_t seed, DevState* st,
                              int use_call, uint64_t call_in, long long* __restrict__ idx_out,
                              const int* __restrict__ act, const float* __restrict__ rew, const uint8_t* __restrict__ done, float beta,
                              int* __restrict__ a_b, float* __restrict__ r_b, float* __restrict__ d_b, float* __restrict__ w_b) {
  extern __shared__ int sh[];
  int HT = 1; while (HT < 4 * B) HT <<= 1;      // hash table: HT keys + HT owners, then the staged tree top
  int* keys = sh;
  int* owner = sh + HT;
  float* top = reinterpret_cast<float*>(sh + 2 * HT);
  const int ntop = min(TREE_TOP, 2 * P);
  for (int t = threadIdx.x; t < ntop; t += blockDim.x) top[t] = tree[t];
  __syncthreads();
  const int j = threadIdx.x;
  const uint64_t call = use_call ? call_in : st->sample_call;
  uint32_t attempt = 0;
  int leaf = (j < B) ? tree_descend(tree, top, P, philox_uniform(seed, call, j, 0)) : -1;
  for (int round = 0; ; ++round) {
    for (int t = threadIdx.x; t < HT; t += blockDim.x) { keys[t] = -1; owner[t] = 0x7fffffff; }
    __syncthreads();
    int slot = -1;
    if (j < B) {
      unsigned h = ((unsigned)leaf * 2654435761u) & (HT - 1);
      while (true) {
        int prev = atomicCAS(&keys[h], -1, leaf);
        if (prev == -1 || prev == leaf) { slot = h; break; }
        h = (h + 1) & (HT - 1);
      }
      atomicMin(&owner[slot], j);
    }
    __syncthreads();
    const int rej = (j < B) && (owner[slot] < j);
    const int any = __syncthreads_or(rej);
    if (!any) break;
    if (round < SAMPLE_MAX_ROUNDS) {
      if (rej) { ++attempt; leaf = tree_descend(tree, top, P, philox_uniform(seed, call, j, attempt)); }
      continue;
    }
    // exact resolution of the slots that are still rejected (see above); `owner` is reused as the list of held leaves
    int* held = owner;
    __syncthreads();
    if (j < B) held[j] = rej ? -1 : leaf;
    __syncthreads();
    if (j == 0) {
      for (int s = 0; s < B; ++s) {
        if (held[s] >= 0) continue;
        int pick = -1;
        for (int t = 0; t < SAMPLE_EXACT_TRIES; ++t) {
          pick = tree_descend_excl(tree, P, philox_uniform(seed, call, (uint32_t)s, 0x80000000u + (uint32_t)t), held, B);
          bool dup = false;                       // rounding residue can leave a held leaf reachable: draw again
          for (int i = 0; i < B; ++i) dup = dup || (held[i] == pick);
          if (!dup) break;
          pick = -1;
        }
        if (pick < 0) { atomicOr(&st->error, 1); pick = 0; }
        held[s] = pick;
      }
    }
    __syncthreads();
    if (j < B) leaf = held[j];
    break;
  }
  if (j < B) {
    idx_out[j] = leaf;
    if (a_b) {
      a_b[j] = act[leaf]; r_b[j] = rew[leaf]; d_b[j] = done[leaf] ? 1.f : 0.f;
      const float p = __fdiv_rn(tree[P + leaf], top[1]);
      w_b[j] = pow_f32(__fmul_rn((float)st->curr_size, p), -beta);
    }
  }
}

// Per-sample metadata gather + importance weights (get_batch, PER:91-102):
//   w_i = (n * prio_i / sum prio)^(-beta), sum prio = tree root, n = curr_size.
__global__ void batch_meta_kernel(const long long* __restrict__ idx, int B, const float* __restrict__ tree, int P,
                                  const int* __restrict__ act, const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                  const DevState* __restrict__ st, float beta,
                                  int* __restrict__ a_b, float* __restrict__ r_b, float* __restrict__ d_b, float* __restrict__ w_b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= B) return;
  const long long i = idx[j];
  a_b[j] = act[i];
  r_b[j] = rew[i];
  d_b[j] = done[i] ? 1.f : 0.f;
  const float p = __fdiv_rn(tree[P + i], tree[1]);
  w_b[j] = pow_f32(__fmul_rn((float)st->curr_size, p), -beta);
}

// Observation gather: rows idx[j] of the s / s' stores -> contiguous batch (s rows 0..B-1, s' rows B..2B-1).
// 16-byte loads, 4 in flight per thread; grid = (chunks, 2B).
// Tensor-core path with byte observations: the same pass also writes the batch as fp32 raw values k (exact in TF32, so this
// operand needs no lo term; the 1/255 is folded into a scaled copy of the first layer's weights).
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* __restrict__ store_s, const uint8_t* __restrict__ store_sp,
                                                           const long long* __restrict__ idx, int B, long long row_bytes,
                                                           uint8_t* __restrict__ out, float* __restrict__ out_f, long long lo_delta, int obs_u8) {
  const int row = blockIdx.y;
  const long long i = idx[row < B ? row : row - B];
  const uint8_t* src = (row < B ? store_s : store_sp) + i * row_bytes;
  uint8_t* dst = out + (long long)row * row_bytes;
  if ((row_bytes & 15) == 0) {
    const long long nvec = row_bytes >> 4;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    long long v0 = ((long long)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
    int4 tmp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { long long v = v0 + (long long)u * blockDim.x; if (v < nvec) tmp[u] = __ldg(s4 + v); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      long long v = v0 + (long long)u * blockDim.x;
      if (v >= nvec) continue;
      d4[v] = tmp[u];
      if (out_f) {
        if (obs_u8) {          // 16 bytes -> 16 floats
          float* o = out_f + (long long)row * row_bytes + v * 16;
          const uint32_t w[4] = {(uint32_t)tmp[u].x, (uint32_t)tmp[u].y, (uint32_t)tmp[u].z, (uint32_t)tmp[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(o + 4 * q) = make4((float)(w[q] & 255u), (float)((w[q] >> 8) & 255u), (float)((w[q] >> 16) & 255u), (float)(w[q] >> 24));
        }
      }
    }
  } else {
    for (long long b = (long long)blockIdx.x * blockDim.x * 4 + threadIdx.x; b < row_bytes && b < ((long long)blockIdx.x + 1) * blockDim.x * 4; b += blockDim.x)
      dst[b] = src[b];
  }
}

// [s0, s1) of theta is the first conv layer's weight block; when the observations are raw bytes the tensor-core path reads
// a copy scaled by 1/255 (the byte values k are exact TF32 operands, Float32(k)/255f0 is not)
__global__ void scale_block_kernel(const float* __restrict__ w, float* __restrict__ out, long long s0, long long s1, float scale) {
  for (long long e = s0 + (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; e < s1; e += (long long)gridDim.x * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(w + e);
    *reinterpret_cast<float4*>(out + (e - s0)) = make4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

// Refresh of the touched leaf-to-root paths, one CTA: every level is recomputed from its children after a
// barrier, so the result is independent of thread order.  Also the end-of-step bookkeeping (beta powers,
// sampling-call counter) so that a captured graph advances its own state.
__global__ void tree_update_kernel(float* __restrict__ tree, int P, const long long* __restrict__ idx, const float* __restrict__ newp,
                                   int n, int write_leaves, DevState* st, int end_of_step, double beta1, double beta2,
                                   int advance_sampler, float* __restrict__ publish, int publish_epoch = 0) {
  const int j = threadIdx.x;
  if (write_leaves) {
    for (int t = j; t < n; t += blockDim.x) {
      const float p = newp[t];
      if (!(p > 0.f)) atomicOr(&st->error, 2);
      tree[P + idx[t]] = p;
    }
  }
  __syncthreads();
  if (n > 0) {
    for (int shift = 1; (P >> shift) >= 1; ++shift) {
      for (int t = j; t < n; t += blockDim.x) {
        const long long node = (P + idx[t]) >> shift;
        tree[node] = __fadd_rn(tree[2 * node], tree[2 * node + 1]);
      }
      __syncthreads();
    }
  }
  if (publish_epoch && j == 0) {          // (all levels are written: the barrier above closed the last one)
    __threadfence();
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&st->tree_epoch), "l"(st->step + 1ULL) : "memory");
  }
  if (end_of_step && j == 0) {
    st->b1p *= beta1; st->b2p *= beta2;
    if (advance_sampler) st->sample_call += 1;
    const unsigned long long s = ++st->step;
    if (publish) {      // scalars of the finished step -> mapped pinned host words (loss, grad_norm, error flags, step), two slots:
      publish += 4 * (s & 1);           // the host may read step s while step s+1 runs (dqn_step_result)
      publish[0] = st->loss; publish[1] = __uint_as_float(st->gradmax_bits); reinterpret_cast<int*>(publish)[2] = st->error;
      reinterpret_cast<unsigned int*>(publish)[3] = (unsigned int)s;
    }
  }
}

// First kernel of an ingest on the ingest lane: returns once step `expected` (the one in flight when dqn_replay_add was called) has
// published its tree epoch.  One warp, sleeping between polls: it shares an SM with a persistent GEMM CTA without taking anything from it.
__global__ void epoch_wait_kernel(DevState* st, unsigned long long expected) {
  if (threadIdx.x != 0) return;
  const long long t0 = clock64();
  while (true) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&st->tree_epoch) : "memory");
    if (v >= expected) break;
    if (clock64() - t0 > 240000000000LL) { atomicOr(&st->error, 32); break; }   // ~2 min: the step in flight may itself be waiting for a peer rank
    __nanosleep(200);
  }
}

// Full rebuild, level by level (used after bulk fills): parents[k] = fl(tree[2k] + tree[2k+1]) for k in [lo, 2lo)
__global__ void tree_level_kernel(float* __restrict__ tree, long long lo) {
  const long long k = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < 2 * lo) tree[k] = __fadd_rn(tree[2 * k], tree[2 * k + 1]);
}

// add_exp! for n transitions already on the device in Flux layout (C,H,W per sample): transposes to the
// store's HWC layout, writes a/r/done and the new leaf priority (td0+eps)^alpha (PER:65-74).
__global__ void ingest_kernel(const uint8_t* __restrict__ s_in, const uint8_t* __restrict__ sp_in, const int* __restrict__ a_in,
                              const float* __restrict__ r_in, const uint8_t* __restrict__ d_in, const float* __restrict__ td0,
                              long long t0, long long cursor, long long cap, int C, int HW, int elem_bytes,
                              uint8_t* __restrict__ store_s, uint8_t* __restrict__ store_sp, int* __restrict__ act, float* __restrict__ rew,
                              uint8_t* __restrict__ done, float* __restrict__ tree, int P, long long* __restrict__ slot_idx,
                              float alpha, float eps, int n_actions, long long new_size, DevState* st) {
  const long long t = t0 + blockIdx.y;      // transition
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) st->curr_size = new_size;
  const long long slot = (cursor + t) % cap;
  // the reference's asserts (PER:66 td_err + eps > 0; a valid action index): an offending transition raises the sticky flag and is
  // NOT stored - its ring slot keeps its previous contents and priority (the host path rejects the whole batch before any copy)
  const int a_t = a_in[t];
  const float base_t = __fadd_rn(td0[t], eps);
  const bool bad_a = a_t < 1 || a_t > n_actions, bad_p = !(base_t > 0.f);
  if (bad_a || bad_p) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(&st->error, (bad_a ? 8 : 0) | (bad_p ? 4 : 0)); slot_idx[t] = slot; }
    return;
  }
  const long long elems = (long long)C * HW;
  const uint8_t* si = s_in + t * elems * elem_bytes;
  const uint8_t* pi = sp_in + t * elems * elem_bytes;
  uint8_t* so = store_s + slot * elems * elem_bytes;
  uint8_t* po = store_sp + slot * elems * elem_bytes;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < elems; e += (long long)gridDim.x * blockDim.x) {
    const long long hw = e / C, c = e % C;            // e indexes the HWC store
    const long long src = c * HW + hw;                // CHW input
    if (elem_bytes == 1) { so[e] = si[src]; po[e] = pi[src]; }
    else { ((float*)so)[e] = ((const float*)si)[src]; ((float*)po)[e] = ((const float*)pi)[src]; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    act[slot] = a_t; rew[slot] = r_in[t]; done[slot] = d_in[t] ? 1 : 0;
    tree[P + slot] = pow_f32(base_t, alpha);
    slot_idx[t] = slot;
  }
}

// Synthetic replay fill, transition i a pure function of (seed, i): obs bytes / floats from
// Philox(key = seed, counter = (word, i_lo, i_hi, stream)); a ~ U{1..|A|}; r ~ U(-1,1); done ~ Bern(0.01);
// priority (|r|+eps)^alpha.  oracle/synthetic.py regenerates any transition on the CPU.
__global__ void fill_synthetic_kernel(uint8_t* __restrict__ store_s, uint8_t* __restrict__ store_sp, int* __restrict__ act,
                                      float* __restrict__ rew, uint8_t* __restrict__ done, float* __restrict__ tree, int P,
                                      long long i0, long long elems, int obs_u8, int n_actions, float alpha, float eps, uint64_t seed) {
  const long long i = i0 + blockIdx.y;
  const long long words = obs_u8 ? (elems + 15) / 16 : (elems + 3) / 4;     // one Philox block = 16 bytes or 4 floats
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int stream = 0; stream < 2; ++stream) {
      uint32_t c[4] = {(uint32_t)w, (uint32_t)i, (uint32_t)(i >> 32), (uint32_t)stream};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      uint8_t* base = (stream == 0 ? store_s : store_sp);
      if (obs_u8) {
        uint8_t* o = base + i * elems + w * 16;
        if ((elems & 15) == 0) { *reinterpret_cast<uint4*>(o) = make_uint4(c[0], c[1], c[2], c[3]); continue; }   // little-endian bytes
        for (int q = 0; q < 4; ++q)
          for (int b = 0; b < 4; ++b) { long long e = w * 16 + q * 4 + b; if (e < elems) o[q * 4 + b] = (uint8_t)(c[q] >> (8 * b)); }
      } else {
        float* o = (float*)base + i * elems + w * 4;
        for (int q = 0; q < 4; ++q) { long long e = w * 4 + q; if (e < elems) o[q] = ((float)(c[q] >> 8) * 5.9604644775390625e-08f) * 2.f - 1.f; }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t c[4] = {0xFFFFFFFFu, (uint32_t)i, (uint32_t)(i >> 32), 2u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    act[i] = 1 + (int)(c[0] % (uint32_t)n_actions);
    const float r = ((float)(c[1] >> 8) * 5.9604644775390625e-08f) * 2.f - 1.f;
    rew[i] = r;
    done[i] = ((c[2] >> 8) < 167772u) ? 1 : 0;          // 167772 / 2^24 ~ 0.01
    tree[P + i] = pow_f32(__fadd_rn(fabsf(r), eps), alpha);
  }
}

// Store (HWC, u8 or f32) rows -> Float32 Flux layout (CHW) on the way out (get_batch / replay_read), or
// host CHW observations -> HWC compute layout on the way in (q_values).  dir 0: HWC->CHW, 1: CHW->HWC.
__global__ void relayout_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long rows, int C, int HW,
                                int in_u8, int out_f32_from_u8, int dir) {
  const long long elems = (long long)C * HW;
  const long long total = rows * elems;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / elems, e = t % elems;
    long long src, dst;
    if (dir == 0) { const long long c = e / HW, hw = e % HW; src = hw * C + c; dst = e; }      // e indexes CHW output
    else          { const long long hw = e / C, c = e % C; src = c * HW + hw; dst = e; }       // e indexes HWC output
    if (in_u8) {
      const uint8_t v = in[row * elems + src];
      if (out_f32_from_u8) ((float*)out)[row * elems + dst] = u8_to_f32(v); else out[row * elems + dst] = v;
    } else ((float*)out)[row * elems + dst] = ((const float*)in)[row * elems + src];
  }
}

// ------------------------------------------------------------------------------------------------
// Fused head: dueling combine, Double-Q argmax, Bellman target, IS-weighted Huber loss, dQ seed, TD errors
// and new priorities.  One thread per sample, one CTA (B <= 1024).   SOLVER:209-225, DUEL:8-11, HELPERS:14-19,
// PER:76-80; evaluation order as SURVEY App. A steps 4-9, 12 (no FMA contraction on the target).
struct HeadArgs {
  // tower outputs; dueling: V [rows][1], A [rows][nA]; else Q in A_* and V_* unused
  const float* V_on; const float* A_on;      // online net, rows 0..B-1 = s, B..2B-1 = s'
  const float* V_tg; const float* A_tg;      // target net, rows 0..B-1 = s'
  const int* a_b; const float* r_b; const float* d_b; const float* w_b;
  float* dV; float* dA;                      // dL/d(last-layer outputs), already times act'(out)
  float* q_s; float* q_sp_on; float* q_sp_tg;// (B, nA) diagnostics
  float* y; int* best_a; float* td; float* newp;
  int B, nA, dueling, double_q, act_v, act_a;
  float gamma, alpha, eps, inv_world_B;      // inv_world_B = 1/(B*world): data-parallel ranks form one batch of B*world
  DevState* st;
  float* part; unsigned int* ticket;         // several CTAs (recurrent step: trace_length * batch rows): per-CTA partial sums + arrival counter
};
constexpr int HEAD_MAX_ACTIONS = 64;

__device__ __forceinline__ void dueling_q(const float* V, const float* A, int row, int nA, int dueling, float* q) {
  if (dueling) {
    float m = A[(long long)row * nA];
    for (int k = 1; k < nA; ++k) m = __fadd_rn(m, A[(long long)row * nA + k]);
    m = __fdiv_rn(m, (float)nA);
    const float v = V[row];
    for (int k = 0; k < nA; ++k) q[k] = __fsub_rn(__fadd_rn(v, A[(long long)row * nA + k]), m);
  } else {
    for (int k = 0; k < nA; ++k) q[k] = A[(long long)row * nA + k];
  }
}

// per-sample part of the head: returns huber(w td); q_* are this sample's three Q rows after the dueling combine
__device__ __forceinline__ float head_sample(const HeadArgs& h, int i, const float* q_s_row, const float* q_on_row, const float* q_tg_row,
                                             float v_on_s, const float* a_on_s, float* dv_out, float* da_out) {
  const int nA = h.nA;
  for (int k = 0; k < nA; ++k) h.q_sp_on[(long long)i * nA + k] = q_on_row[k];
  int best = 0;
  for (int k = 1; k < nA; ++k) if (q_on_row[k] > q_on_row[best]) best = k;          // first maximal index (Julia argmax)
  for (int k = 0; k < nA; ++k) h.q_sp_tg[(long long)i * nA + k] = q_tg_row[k];
  float qsp;
  if (h.double_q) qsp = q_tg_row[best];
  else { best = 0; for (int k = 1; k < nA; ++k) if (q_tg_row[k] > q_tg_row[best]) best = k; qsp = q_tg_row[best]; }
  // y = r + ((1 - d) * gamma) * q'
  const float y = __fadd_rn(h.r_b[i], __fmul_rn(__fmul_rn(__fsub_rn(1.f, h.d_b[i]), h.gamma), qsp));
  for (int k = 0; k < nA; ++k) h.q_s[(long long)i * nA + k] = q_s_row[k];
  int a = h.a_b[i] - 1;
  if (a < 0 || a >= nA) { atomicOr(&h.st->error, 8); a = min(max(a, 0), nA - 1); }   // never index q[] out of range
  const float td = __fsub_rn(q_s_row[a], y);
  const float w = h.w_b[i];
  const float x = __fmul_rn(w, td);
  const float ax = fabsf(x), quad = fminf(ax, 1.f), lin = __fsub_rn(ax, quad);
  const float hub = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, quad), quad), lin);
  const float g = __fmul_rn(__fmul_rn(w, fminf(fmaxf(x, -1.f), 1.f)), h.inv_world_B);
  if (h.dueling) {
    dv_out[0] = g * act_deriv(v_on_s, h.act_v);
    const float gm = __fdiv_rn(g, (float)nA);
    for (int k = 0; k < nA; ++k) da_out[k] = ((k == a ? g : 0.f) - gm) * act_deriv(a_on_s[k], h.act_a);
  } else {
    for (int k = 0; k < nA; ++k) da_out[k] = (k == a ? g : 0.f) * act_deriv(a_on_s[k], h.act_a);
  }
  h.y[i] = y; h.best_a[i] = best + 1; h.td[i] = td;
  h.newp[i] = pow_f32(__fadd_rn(fabsf(td), h.eps), h.alpha);
  return hub;
}

// The same head for small action sets (|A| <= NA: Atari's 6 / 18 fit NA = 8 / 32): every array lives in registers (loops unrolled and
// predicated, runtime indices turned into selects) and all of a sample's inputs are loaded before anything is computed - the generic
// kernel walks 64-entry local-memory arrays with dependent loads in between and took 12 us for 256 samples, alone on the GPU between the
// forward and the reverse pass.  Same operations in the same order: bit-identical results.
template <int NA>
__global__ void __launch_bounds__(256) head_loss_small_kernel(HeadArgs h) {
  __shared__ float red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x, nA = h.nA;
  float hub = 0.f;
  if (i < h.B) {
    float as[NA], ao[NA], at[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      const bool in = k < nA;
      as[k] = in ? __ldg(h.A_on + (long long)i * nA + k) : 0.f;
      ao[k] = in ? __ldg(h.A_on + (long long)(h.B + i) * nA + k) : 0.f;
      at[k] = in ? __ldg(h.A_tg + (long long)i * nA + k) : 0.f;
    }
    const float vs = h.dueling ? __ldg(h.V_on + i) : 0.f, vo = h.dueling ? __ldg(h.V_on + h.B + i) : 0.f, vt = h.dueling ? __ldg(h.V_tg + i) : 0.f;
    const float r = __ldg(h.r_b + i), d = __ldg(h.d_b + i), w = __ldg(h.w_b + i);
    int a = __ldg(h.a_b + i) - 1;
    float q[NA], qo[NA], qt[NA];
    if (h.dueling) {                                             // Q = (V + A) - mean(A), the mean summed in index order (DUEL:8-11)
      float ms = as[0], mo = ao[0], mt = at[0];
#pragma unroll
      for (int k = 1; k < NA; ++k) if (k < nA) { ms = __fadd_rn(ms, as[k]); mo = __fadd_rn(mo, ao[k]); mt = __fadd_rn(mt, at[k]); }
      ms = __fdiv_rn(ms, (float)nA); mo = __fdiv_rn(mo, (float)nA); mt = __fdiv_rn(mt, (float)nA);
#pragma unroll
      for (int k = 0; k < NA; ++k) { q[k] = __fsub_rn(__fadd_rn(vs, as[k]), ms); qo[k] = __fsub_rn(__fadd_rn(vo, ao[k]), mo); qt[k] = __fsub_rn(__fadd_rn(vt, at[k]), mt); }
    } else {
#pragma unroll
      for (int k = 0; k < NA; ++k) { q[k] = as[k]; qo[k] = ao[k]; qt[k] = at[k]; }
    }
    int best = 0; float qbest = qo[0];
#pragma unroll
    for (int k = 1; k < NA; ++k) if (k < nA && qo[k] > qbest) { best = k; qbest = qo[k]; }     // first maximal index (Julia argmax)
    float qsp = qt[0];
    if (h.double_q) {
#pragma unroll
      for (int k = 1; k < NA; ++k) if (k == best) qsp = qt[k];
    } else {
      best = 0;
#pragma unroll
      for (int k = 1; k < NA; ++k) if (k < nA && qt[k] > qsp) { best = k; qsp = qt[k]; }
    }
    const float y = __fadd_rn(r, __fmul_rn(__fmul_rn(__fsub_rn(1.f, d), h.gamma), qsp));
    if (a < 0 || a >= nA) { atomicOr(&h.st->error, 8); a = min(max(a, 0), nA - 1); }
    float qa = q[0];
#pragma unroll
    for (int k = 1; k < NA; ++k) if (k == a) qa = q[k];
    const float td = __fsub_rn(qa, y);
    const float x = __fmul_rn(w, td);
    const float ax = fabsf(x), quad = fminf(ax, 1.f), lin = __fsub_rn(ax, quad);
    hub = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, quad), quad), lin);
    const float g = __fmul_rn(__fmul_rn(w, fminf(fmaxf(x, -1.f), 1.f)), h.inv_world_B);
    const float gm = __fdiv_rn(g, (float)nA);
    if (h.dueling) h.dV[i] = g * act_deriv(vs, h.act_v);
#pragma unroll
    for (int k = 0; k < NA; ++k) if (k < nA) {
      h.dA[(long long)i * nA + k] = h.dueling ? ((k == a ? g : 0.f) - gm) * act_deriv(as[k], h.act_a) : (k == a ? g : 0.f) * act_deriv(as[k], h.act_a);
      h.q_s[(long long)i * nA + k] = q[k]; h.q_sp_on[(long long)i * nA + k] = qo[k]; h.q_sp_tg[(long long)i * nA + k] = qt[k];
    }
    h.y[i] = y; h.best_a[i] = best + 1; h.td[i] = td;
    h.newp[i] = pow_f32(__fadd_rn(fabsf(td), h.eps), h.alpha);
  }
  // loss = sum(huber) / B: warp sums, then warp 0 adds them in warp order (one CTA: B <= 256 here)
  float sacc = hub;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sacc;
  __syncthreads();
  if (threadIdx.x < 32) {
    sacc = (threadIdx.x < (blockDim.x + 31) / 32) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    if (threadIdx.x == 0) { h.st->loss = __fdiv_rn(sacc, (float)h.B); h.st->gradmax_bits = 0u; }
  }
}

__global__ void head_loss_kernel(HeadArgs h) {
  __shared__ float red[32];
  __shared__ bool last;
  float hub = 0.f;
  // one thread per sample; grid-stride (one CTA for the feed-forward step; the recurrent step has trace_length * batch_size rows and
  // runs several CTAs, whose partial sums the last CTA to arrive adds in CTA order: deterministic)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h.B; i += gridDim.x * blockDim.x) {
    float q[HEAD_MAX_ACTIONS], qo[HEAD_MAX_ACTIONS], qt[HEAD_MAX_ACTIONS], da[HEAD_MAX_ACTIONS], dv = 0.f;
    const int nA = h.nA;
    dueling_q(h.V_on, h.A_on, h.B + i, nA, h.dueling, qo);
    dueling_q(h.V_tg, h.A_tg, i, nA, h.dueling, qt);
    dueling_q(h.V_on, h.A_on, i, nA, h.dueling, q);
    hub += head_sample(h, i, q, qo, qt, h.dueling ? h.V_on[i] : 0.f, h.A_on + (long long)i * nA, &dv, da);
    if (h.dueling) h.dV[i] = dv;
    for (int k = 0; k < nA; ++k) h.dA[(long long)i * nA + k] = da[k];
  }
  // loss = sum(huber) / B  (block reduction; fp32)
  float s = hub;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x + 31) / 32) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (gridDim.x == 1) {
      if (threadIdx.x == 0) { h.st->loss = __fdiv_rn(s, (float)h.B); h.st->gradmax_bits = 0u; }
      return;
    }
    if (threadIdx.x == 0) {
      h.part[blockIdx.x] = s;
      __threadfence();
      last = atomicAdd(h.ticket, 1u) == gridDim.x - 1;
    }
  }
  if (gridDim.x == 1) return;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float tot = 0.f;
    for (unsigned int k = 0; k < gridDim.x; ++k) tot += reinterpret_cast<volatile float*>(h.part)[k];
    h.st->loss = __fdiv_rn(tot, (float)h.B); h.st->gradmax_bits = 0u;
    *h.ticket = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused Adam + max|g| (Flux.Optimise.Adam with Float64 scalars; globalnorm HELPERS:38-46).
// One pass: read w,m,v,g - write w,m,v (+ optionally the target copy).  Per element, in Float64:
//   m = b1 m + (1-b1) g ; v = b2 v + ((1-b2) g) g ; d = m/(1-b1p) / (sqrt(v/(1-b2p)) + eps) * eta ; w -= d
// with m, v, d each rounded to Float32 on store exactly as the reference's broadcasts do.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ g, long long n4, double eta, double beta1, double beta2,
                                                    double eps, float gscale, DevState* st,
                                                    float* __restrict__ w_hi, long long lo_delta, long long s0, long long s1, float scale) {
  const double c1 = 1.0 - st->b1p, c2 = 1.0 - st->b2p;
  const double omb1 = 1.0 - beta1, omb2 = 1.0 - beta2;
  // m/c1 and v/c2 as multiplications by the reciprocals: <= 1 ulp(double) away from the reference's divisions, invisible once
  // the update is rounded to Float32 (the parity test allows 2 ulp of the parameter); one fp64 division per element remains
  const double rc1 = 1.0 / c1, rc2 = 1.0 / c2;
  float gmax = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 W = reinterpret_cast<float4*>(w)[i], Mv = reinterpret_cast<float4*>(m)[i], Vv = reinterpret_cast<float4*>(v)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float* wp = &W.x; float* mp = &Mv.x; float* vp = &Vv.x; const float* gp = &G.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gf = gp[k] * gscale;
      gmax = fmaxf(gmax, fabsf(gf));
      const double gd = (double)gf;
      // explicit _rn intrinsics: no FMA contraction, every Float64 op rounds separately as the reference's broadcast does
      const float mt = (float)__dadd_rn(__dmul_rn(beta1, (double)mp[k]), __dmul_rn(omb1, gd));
      const float vt = (float)__dadd_rn(__dmul_rn(beta2, (double)vp[k]), __dmul_rn(__dmul_rn(omb2, gd), gd));
      const float d = (float)__dmul_rn(__ddiv_rn(__dmul_rn((double)mt, rc1), __dadd_rn(__dsqrt_rn(__dmul_rn((double)vt, rc2)), eps)), eta);
      mp[k] = mt; vp[k] = vt; wp[k] = wp[k] - d;
    }
    reinterpret_cast<float4*>(w)[i] = W; reinterpret_cast<float4*>(m)[i] = Mv; reinterpret_cast<float4*>(v)[i] = Vv;
    const long long e = i * 4;
    if (w_hi && e >= s0 && e < s1)                // first conv layer on raw bytes: refresh its 1/255-scaled copy in the same pass
      *reinterpret_cast<float4*>(w_hi + (e - s0)) = make4(W.x * scale, W.y * scale, W.z * scale, W.w * scale);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) gmax = fmaxf(gmax, red[k]);
    atomicMax(&st->gradmax_bits, __float_as_uint(gmax));     // non-negative floats order like their bits
  }
}


// update_priorities! (PER:76-80) standalone: td -> (|td|+eps)^alpha in place
__global__ void td_to_priority_kernel(float* __restrict__ p, long long n, float alpha, float eps) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = pow_f32(__fadd_rn(fabsf(p[i]), eps), alpha);
}
__global__ void scatter_leaves_kernel(float* __restrict__ tree, int P, const long long* __restrict__ idx, const float* __restrict__ p, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) tree[P + idx[i]] = p[i];
}
// Q = (V + A) - mean(A)  for the acting path (DUEL:8-11)
__global__ void dueling_combine_kernel(const float* __restrict__ V, const float* __restrict__ A, int rows, int nA, float* __restrict__ q) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float tmp[HEAD_MAX_ACTIONS];
  dueling_q(V, A, r, nA, 1, tmp);
  for (int k = 0; k < nA; ++k) q[(long long)r * nA + k] = tmp[k];
}

// Vectorised acting (SOLVER:83 action(exploration_policy, policy, k, obs) over many lanes; POLICY:38-46 argmax; POMDPTools EpsGreedyPolicy):
// per row Q = dueling combine, greedy action = FIRST maximal index (Julia argmax); with probability eps a uniform random action instead.
// Uniforms are counter based - Philox(seed; lane, 0 | 1, call) - so a host restatement reproduces the actions.  Actions are 1-based.
__global__ void act_select_kernel(const float* __restrict__ V, const float* __restrict__ A, int rows, int nA, int dueling, float eps, uint64_t seed,
                                  uint64_t call, int lane0, int* __restrict__ actions, float* __restrict__ q_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float q[HEAD_MAX_ACTIONS];
  dueling_q(V, A, r, nA, dueling, q);
  int best = 0;
  for (int k = 1; k < nA; ++k) if (q[k] > q[best]) best = k;
  if (q_out) for (int k = 0; k < nA; ++k) q_out[(long long)r * nA + k] = q[k];
  int a = best;
  if (eps > 0.f && philox_uniform(seed, call, (uint32_t)(lane0 + r), 0u) < eps) {
    a = (int)__fmul_rn(philox_uniform(seed, call, (uint32_t)(lane0 + r), 1u), (float)nA);
    if (a >= nA) a = nA - 1;
  }
  actions[r] = a + 1;
}

// Synthetic vectorised environment (bench, config 5 of BASELINE.json): lane i's next observation bytes, reward and done flag are a
// pure function of (seed, step, lane) - the same generator as the synthetic replay fill, so the lanes feed dqn_replay_add_device
// without touching the host.  obs_next doubles as the lanes' current observation of the next step.
__global__ void synth_env_step_kernel(uint8_t* __restrict__ obs_next, float* __restrict__ rew, uint8_t* __restrict__ done, float* __restrict__ td0,
                                      long long elems, int lanes, uint64_t seed, uint64_t step) {
  const int lane = blockIdx.y;
  const long long words = (elems + 15) / 16;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)w, (uint32_t)lane, (uint32_t)step, (uint32_t)(step >> 32) ^ 0x5EEDu};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint8_t* o = obs_next + (long long)lane * elems + w * 16;
    if ((elems & 15) == 0) *reinterpret_cast<uint4*>(o) = make_uint4(c[0], c[1], c[2], c[3]);
    else for (int q = 0; q < 16; ++q) { const long long e = w * 16 + q; if (e < elems) o[q] = (uint8_t)(c[q >> 2] >> (8 * (q & 3))); }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t c[4] = {0xFFFFFFFEu, (uint32_t)lane, (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float r = ((float)(c[1] >> 8) * 5.9604644775390625e-08f) * 2.f - 1.f;
    rew[lane] = r; done[lane] = ((c[2] >> 8) < 167772u) ? 1 : 0; td0[lane] = fabsf(r);
  }
}

__global__ void copy_f4_kernel(float4* __restrict__ dst, const float4* __restrict__ src, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void fill_u32_kernel(uint32_t* __restrict__ p, long long n, uint32_t v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
// scalars of the finished step -> mapped/pinned host words (loss, grad_norm, error flags)
__global__ void publish_kernel(const DevState* __restrict__ st, float* __restrict__ out) {
  out[0] = st->loss; out[1] = __uint_as_float(st->gradmax_bits); reinterpret_cast<int*>(out)[2] = st->error;
}

// Column sums of a row-major matrix D[P][N] (N % 4 == 0): the bias gradient of a conv layer, db[n] = sum_pixels delta[pix][n].
// Deterministic: every CTA sums a fixed slab of rows in a fixed order into part[cta][N]; the last CTA to finish (ticket) adds the
// partials in CTA order.  grid = G CTAs of 256 threads, part holds G*N floats, ticket is a zeroed counter that the kernel resets.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ D, long long P, int N, float* __restrict__ out,
                                                     float* __restrict__ part, unsigned int* __restrict__ ticket) {
  __shared__ float4 sh[256];
  __shared__ int last;
  const int nc = N >> 2;                                      // float4 columns
  const int tx = threadIdx.x % nc, ty = threadIdx.x / nc, rows_per = 256 / nc;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(P, r0 + per);
  float4 acc = make4(0.f, 0.f, 0.f, 0.f);
  if (ty < rows_per) {
    long long r = r0 + ty;
    for (; r + 7LL * rows_per < r1; r += 8LL * rows_per) {   // eight independent loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(D + (r + (long long)j * rows_per) * N) + tx);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    for (; r < r1; r += rows_per) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(D + r * N) + tx);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < nc) {
    float4 s4 = make4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < rows_per; ++j) { const float4 v = sh[j * nc + threadIdx.x]; s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w; }
    reinterpret_cast<float4*>(part + (long long)blockIdx.x * N)[threadIdx.x] = s4;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // final pass, same fixed shape: thread (tx, ty) adds the partials of CTAs ty, ty + rows_per, ... in order, then the ty sums are added in order
  acc = make4(0.f, 0.f, 0.f, 0.f);
  if (ty < rows_per) {
    for (unsigned c = ty; c < gridDim.x; c += rows_per) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part + (long long)c * N) + tx);   // L2: written by other CTAs
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  __syncthreads();
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < nc) {
    float4 s4 = make4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < rows_per; ++j) { const float4 v = sh[j * nc + threadIdx.x]; s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w; }
    reinterpret_cast<float4*>(out)[threadIdx.x] = s4;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// The last Dense layer of each tower (Dense(k, 1) and Dense(k, |A|), DUEL:36-58) is far too thin for a tiled contraction:
// 3.6 kFLOP per row.  One warp per (tower, row): lanes stride over k, N <= HEADS_MAXN running sums, a butterfly per output,
// lane 0 adds the bias row and applies the activation.  Fixed summation order: deterministic.
constexpr int HEADS_MAXN = 8;
struct HeadJob {
  const float* X; long long ldx;       // input rows (previous layer's output)
  const float* W;                      // augmented matrix [(K+1)][N]
  float* C;                            // out [rows][N]
  int rows, N, K, act;
};
struct HeadJobs { HeadJob j[4]; int n; };
__global__ void __launch_bounds__(256) heads_fwd_kernel(const HeadJobs jobs) {
  int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  int ji = 0;
  while (ji < jobs.n && w >= jobs.j[ji].rows) { w -= jobs.j[ji].rows; ++ji; }
  if (ji >= jobs.n) return;
  const HeadJob& jb = jobs.j[ji];
  const float* x = jb.X + (long long)w * jb.ldx;
  float acc[HEADS_MAXN];
#pragma unroll
  for (int n = 0; n < HEADS_MAXN; ++n) acc[n] = 0.f;
#pragma unroll 4
  for (int k = lane; k < jb.K; k += 32) {
    const float xv = __ldg(x + k);
    const float* wr = jb.W + (long long)k * jb.N;
#pragma unroll
    for (int n = 0; n < HEADS_MAXN; ++n) if (n < jb.N) acc[n] = fmaf(xv, __ldg(wr + n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < HEADS_MAXN; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int n = 0; n < HEADS_MAXN; ++n)
      if (n < jb.N) jb.C[(long long)w * jb.N + n] = act_apply(acc[n] + jb.W[(long long)jb.K * jb.N + n], jb.act);
  }
}

// reverse of the same layer: dX[b][j] = (sum_n D[b][n] W[j][n]) * act'(Y[b][j]); one thread per (b, j)
struct HeadGradJob { const float* D; const float* W; const float* Y; float* dX; int rows, N, K, act; };   // K = width of the previous layer
struct HeadGradJobs { HeadGradJob j[2]; int n; };
__global__ void __launch_bounds__(256) heads_dgrad_kernel(const HeadGradJobs jobs) {
  const HeadGradJob& jb = jobs.j[blockIdx.y];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (blockIdx.y >= jobs.n || i >= (long long)jb.rows * jb.K) return;
  const int b = (int)(i / jb.K), j = (int)(i - (long long)b * jb.K);
  const float* d = jb.D + (long long)b * jb.N;
  const float* wr = jb.W + (long long)j * jb.N;
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < HEADS_MAXN; ++n) if (n < jb.N) s = fmaf(__ldg(d + n), __ldg(wr + n), s);
  jb.dX[i] = s * act_deriv(jb.Y[i], jb.act);
}

// ------------------------------------------------------------------------------------------------
// The whole head in ONE launch (feed-forward step, thin output layers): for sample i one warp computes the output layers of both towers
// on the three rows that matter (online s, online s', target s') - lanes stride over k, a butterfly per output, exactly the sums of
// heads_fwd_kernel - then the per-sample head (dueling combine, Double-Q argmax, target, IS-Huber, dQ seed, priority) and the gradient
// into the last hidden layer (heads_dgrad_kernel's formula).  Replaces four launches of 7-13 us each that sat one behind the other in
// the middle of the step (heads_fwd x2, head_loss, heads_dgrad).  The loss is reduced deterministically: per-sample values to a buffer,
// the last CTA to finish (ticket) adds them in index order.
struct FusedHeadArgs {
  HeadArgs h;
  const float* H_on[2]; const float* H_tg[2];   // last hidden layer outputs per tower: online [2B][K], target [B][K]
  const float* W_on[2]; const float* W_tg[2];   // output layers, augmented [K+1][N]
  float* out_on[2]; float* out_tg[2];           // V / A outputs kept for the diagnostics ([2B][N], [B][N])
  float* dH[2];                                 // gradient into the last hidden layer [B][K], times act'(H)
  int N[2], act_out[2], act_hidden[2], K, ntow;
  float* hub; unsigned int* ticket;
  int fwd_done;                                 // the output layers were already computed (heads_fwd_kernel on each pass's own lane): out_* are inputs
};
__device__ __forceinline__ void head_rowdot(const float* __restrict__ x, const float* __restrict__ W, int K, int N, int act, int lane, float* out) {
  float acc[HEADS_MAXN];
#pragma unroll
  for (int n = 0; n < HEADS_MAXN; ++n) acc[n] = 0.f;
#pragma unroll 4
  for (int k = lane; k < K; k += 32) {
    const float xv = __ldg(x + k);
    const float* wr = W + (long long)k * N;
#pragma unroll
    for (int n = 0; n < HEADS_MAXN; ++n) if (n < N) acc[n] = fmaf(xv, __ldg(wr + n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < HEADS_MAXN; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    out[n] = n < N ? act_apply(acc[n] + __ldg(W + (long long)K * N + n), act) : 0.f;
  }
}
__global__ void __launch_bounds__(256) head_fused_kernel(const FusedHeadArgs a) {
  const HeadArgs& h = a.h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  __shared__ int last;
  if (i < h.B) {
    float o_s[2][HEADS_MAXN], o_sp[2][HEADS_MAXN], o_tg[2][HEADS_MAXN];
    if (a.fwd_done) {
      for (int t = 0; t < a.ntow; ++t) {
#pragma unroll
        for (int n = 0; n < HEADS_MAXN; ++n) {
          const bool in = n < a.N[t];
          o_s[t][n] = in ? a.out_on[t][(long long)i * a.N[t] + n] : 0.f;
          o_sp[t][n] = in ? a.out_on[t][(long long)(h.B + i) * a.N[t] + n] : 0.f;
          o_tg[t][n] = in ? a.out_tg[t][(long long)i * a.N[t] + n] : 0.f;
        }
      }
    } else
    for (int t = 0; t < a.ntow; ++t) {
      head_rowdot(a.H_on[t] + (long long)i * a.K, a.W_on[t], a.K, a.N[t], a.act_out[t], lane, o_s[t]);
      head_rowdot(a.H_on[t] + (long long)(h.B + i) * a.K, a.W_on[t], a.K, a.N[t], a.act_out[t], lane, o_sp[t]);
      head_rowdot(a.H_tg[t] + (long long)i * a.K, a.W_tg[t], a.K, a.N[t], a.act_out[t], lane, o_tg[t]);
      if (lane < a.N[t]) {
        a.out_on[t][(long long)i * a.N[t] + lane] = o_s[t][lane];
        a.out_on[t][(long long)(h.B + i) * a.N[t] + lane] = o_sp[t][lane];
        a.out_tg[t][(long long)i * a.N[t] + lane] = o_tg[t][lane];
      }
    }
    const int nA = h.nA, ta = a.ntow - 1;                        // advantage (or only) tower
    float q[HEADS_MAXN], qo[HEADS_MAXN], qt[HEADS_MAXN], da[HEADS_MAXN], dv = 0.f;
    if (h.dueling) {                                             // Q = (V + A) - mean(A), the mean summed in index order (DUEL:8-11)
      float ms = o_s[1][0], mo = o_sp[1][0], mt = o_tg[1][0];
      for (int k = 1; k < nA; ++k) { ms = __fadd_rn(ms, o_s[1][k]); mo = __fadd_rn(mo, o_sp[1][k]); mt = __fadd_rn(mt, o_tg[1][k]); }
      ms = __fdiv_rn(ms, (float)nA); mo = __fdiv_rn(mo, (float)nA); mt = __fdiv_rn(mt, (float)nA);
      for (int k = 0; k < nA; ++k) {
        q[k] = __fsub_rn(__fadd_rn(o_s[0][0], o_s[1][k]), ms); qo[k] = __fsub_rn(__fadd_rn(o_sp[0][0], o_sp[1][k]), mo); qt[k] = __fsub_rn(__fadd_rn(o_tg[0][0], o_tg[1][k]), mt);
      }
    } else {
      for (int k = 0; k < nA; ++k) { q[k] = o_s[0][k]; qo[k] = o_sp[0][k]; qt[k] = o_tg[0][k]; }
    }
    float hub = 0.f;
    if (lane == 0) {
      hub = head_sample(h, i, q, qo, qt, h.dueling ? o_s[0][0] : 0.f, o_s[ta], &dv, da);
      a.hub[i] = hub;
      if (h.dueling) h.dV[i] = dv;
      for (int k = 0; k < nA; ++k) h.dA[(long long)i * nA + k] = da[k];
    }
    dv = __shfl_sync(0xffffffffu, dv, 0);
#pragma unroll
    for (int k = 0; k < HEADS_MAXN; ++k) da[k] = __shfl_sync(0xffffffffu, k < nA ? da[k] : 0.f, 0);
    // gradient into the last hidden layer: dH[j] = (sum_n d[n] W[j][n]) act'(H[j])
    for (int t = 0; t < a.ntow; ++t) {
      const float* W = a.W_on[t]; const float* Hr = a.H_on[t] + (long long)i * a.K; float* o = a.dH[t] + (long long)i * a.K;
      const int N = a.N[t];
      for (int j = lane; j < a.K; j += 32) {
        float sacc = 0.f;
        if (h.dueling && t == 0) sacc = fmaf(dv, __ldg(W + j), 0.f);
        else {
#pragma unroll
          for (int n = 0; n < HEADS_MAXN; ++n) if (n < N) sacc = fmaf(da[n], __ldg(W + (long long)j * N + n), sacc);
        }
        o[j] = sacc * act_deriv(Hr[j], a.act_hidden[t]);
      }
    }
  }
  // deterministic loss: the last CTA sums the per-sample values in index order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int k = threadIdx.x; k < h.B; k += 32) s += __ldcg(a.hub + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) { h.st->loss = __fdiv_rn(s, (float)h.B); h.st->gradmax_bits = 0u; *a.ticket = 0u; }
  }
}
#endif  // __CUDACC__

}  // namespace dqn
