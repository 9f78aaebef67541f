// peer_ar.cuh - the gradient all-reduce of the data-parallel step over NVLink peer memory (the one exchange step of the path).
//
// The bucket is 12.9 MB (Dense towers) + 0.3 MB (conv trunk) per step; an NCCL ring on 16 CTAs moves it at 155-183 GB/s bus bandwidth
// (117-148 us at 4-8 GPUs, bench r02) and its small-message latency sits on the step's tail.  All ranks of one box see each other's
// memory through NVSwitch, so the exchange is written directly:
//   two-shot, owner computes - rank r owns slice r of the bucket.  It PULLS slice r of every rank's gradient (P2P loads), adds them in
//   rank order (fixed order: every rank ends up with bit-identical sums, and the same sums whichever rank computes them), and PUSHES
//   the result into slice r of every rank's gradient vector (P2P stores).  Nobody else reads or writes slice r, so the gradients are
//   reduced in place.
//   flags - per rank a block of 64-bit epochs in its own memory, written by its peers: ready[p] (peer p's gradients are final: its
//   kernel has started, stream order put every weight-gradient kernel before it) gates the pulls, done[p][c] (CTA c of peer p has
//   pushed its part, fenced at system scope) gates the kernel's exit - what follows in the stream (Adam) sees the complete sum.
//   Epochs come from the device-side step counter, so a captured CUDA graph replays the kernel unchanged.
// Spins are bounded (~2 min): a peer that never arrives raises the sticky error bit 16 (DQN_ERR_NCCL at the next scalar fetch) instead of
// hanging for ever.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace dqn {

constexpr int PEER_MAX = 8;            // ranks (one NVSwitch box)
constexpr int PEER_MAXG = 64;          // CTAs of one reduction
constexpr int PEER_READY = 0;          // flag block: ready[PEER_MAX], then done[PEER_MAX][PEER_MAXG]
constexpr int PEER_DONE = PEER_MAX;
constexpr int PEER_FLAGS = PEER_MAX + PEER_MAX * PEER_MAXG;
constexpr long long PEER_SPIN_CLOCKS = 240000000000LL; // ~2 min at 1.9 GHz: a last resort against a dead peer, far beyond any host-side skew between ranks

struct PeerArArgs {
  float* grad[PEER_MAX];                 // every rank's gradient vector as mapped in this process (own entry: the local pointer)
  unsigned long long* flags[PEER_MAX];   // every rank's flag block
  int world, rank, bucket;               // bucket 0 / 1: the two reductions of one step (distinct epochs)
  long long off, n;                      // elements [off, off + n) of the vector; both multiples of 4
  DevState* st;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool peer_wait(const unsigned long long* p, unsigned long long epoch, DevState* st) {
  const long long t0 = clock64();
  while (ld_acquire_sys_u64(p) < epoch) {
    if (clock64() - t0 > PEER_SPIN_CLOCKS) { atomicOr(&st->error, 16); return false; }
    __nanosleep(64);
  }
  return true;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  // .cg: no L1 (every address is read once, after its owner's ready flag was acquired at system scope); a plain weak load, so a warp's
  // 512 contiguous bytes go out as whole-line requests, and an unrolled group is issued back to back (a peer load is a ~2-3 us round trip)
  return __ldcg(reinterpret_cast<const float4*>(p));
}

// pull-add-push of the float4 elements [lo4, hi4) of the slice, U elements per thread in flight (U x world loads outstanding per thread:
// NVLink needs ~2.5 MB in flight per direction to run at full rate, 25 dependent round trips per thread ran it at 170 GB/s)
template <int U, int WM>
__device__ __forceinline__ void peer_reduce_range(const PeerArArgs& a, long long base, long long lo4, long long hi4) {
  const int W = a.world;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi4; i0 += stride * U) {
    float4 v[U][WM];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int p = 0; p < WM; ++p) if (p < W && i0 + u * stride < hi4) v[u][p] = ld_peer4(a.grad[p] + 4 * (base + i0 + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= hi4) break;
      float4 s = v[u][0];
#pragma unroll
      for (int p = 1; p < WM; ++p) if (p < W) { s.x = __fadd_rn(s.x, v[u][p].x); s.y = __fadd_rn(s.y, v[u][p].y); s.z = __fadd_rn(s.z, v[u][p].z); s.w = __fadd_rn(s.w, v[u][p].w); }
#pragma unroll
      for (int p = 0; p < WM; ++p) if (p < W) *reinterpret_cast<float4*>(a.grad[p] + 4 * (base + i0 + u * stride)) = s;
    }
  }
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerArArgs a) {
  const int tid = threadIdx.x, W = a.world, G = gridDim.x;
  const unsigned long long epoch = 2ULL * a.st->step + 1ULL + (unsigned long long)a.bucket;
  unsigned long long* mine = a.flags[a.rank];
  // 1. my gradients are final: tell every peer (one CTA does it), then wait until every peer has said the same
  if (blockIdx.x == 0 && tid < W) st_release_sys_u64(a.flags[tid] + PEER_READY + a.rank, epoch);
  __shared__ int ok;
  if (tid == 0) ok = 1;
  __syncthreads();
  if (tid < W && !peer_wait(mine + PEER_READY + tid, epoch, a.st)) ok = 0;
  __syncthreads();
  if (ok) {
    // 2. my slice: pull from everyone, add in rank order, push to everyone
    const long long n4 = a.n >> 2;
    const long long per4 = (n4 + W - 1) / W;
    const long long lo4 = (long long)a.rank * per4, hi4 = lo4 + per4 < n4 ? lo4 + per4 : n4;
    if (W <= 2) peer_reduce_range<4, 2>(a, a.off >> 2, lo4, hi4);
    else if (W <= 4) peer_reduce_range<2, 4>(a, a.off >> 2, lo4, hi4);
    else peer_reduce_range<1, 8>(a, a.off >> 2, lo4, hi4);
  }
  // 3. my pushes are visible everywhere, then say so to every peer
  __threadfence_system();
  __syncthreads();
  if (tid < W) st_release_sys_u64(a.flags[tid] + PEER_DONE + a.rank * PEER_MAXG + blockIdx.x, epoch);
  // 4. the kernel ends when every CTA of every peer has pushed (one CTA waits; the others leave their SMs)
  if (blockIdx.x == 0) {
    for (int f = tid; f < W * G; f += blockDim.x) peer_wait(mine + PEER_DONE + (f / G) * PEER_MAXG + (f % G), epoch, a.st);
  }
}

}  // namespace dqn
