// conv1_tc.cuh - the first conv layer on raw byte observations (Nature-DQN geometry: 8x8 taps, stride 4, 4 byte channels -> 32 maps),
// forward, as its own tcgen05 kernel.  The generic contraction kernel (tc_gemm_impl.cuh) spends 8 ring stages per 128-pixel tile on
// this layer (K = 256 is short, one stage is mostly barrier hops) and re-reads every input byte four times as im2col rows; here
//   * the layer's whole weight matrix (K = 256 x 32, hi and lo TF32 planes side by side = ONE N = 64 operand) stays resident in
//     shared memory for the life of the persistent CTA: no B traffic at all in the main loop;
//   * a tile's input is ONE TMA box per image it touches - the 36 input rows x 84 pixels (12 KB, each byte fetched once) under its
//     128 output pixels - instead of 32 KB of im2col chunks;
//   * eight converter warps expand the im2col rows straight out of that patch: thread = output pixel, one filter row (8 taps x 4
//     channels = 32 contiguous bytes) per step -> 32 fp32 words (byte values are exact TF32 operands: no lo plane) -> tcgen05.st into
//     a ring of eight TMEM A stages;
//   * two MMA warps take alternate stages, each into its own accumulator (the interleave that bounds the truncation error of the
//     tensor core's accumulate), 4 x tcgen05.mma 128x64x8 per stage with B descriptors that just walk the resident planes;
//   * four epilogue warps add the accumulators (round-to-nearest), bias, ReLU, 128-byte row stores; two accumulator sets, so the
//     epilogue of tile i overlaps the main loop of tile i+1.
// Same arithmetic as the generic byte path: operand = the byte value k, weights pre-scaled by 1/255 (w1s), two products
// k*w_hi + k*w_lo, fp32 accumulation in TMEM.  Checked against the same CPU functor executor (tests/csrc/tc_selftest.cu).
#pragma once
#include "tc_gemm_impl.cuh"

namespace c1 {
using namespace tc;

constexpr int C1_THREADS = 512;
constexpr int KH = 8, KW = 8, CIN = 4, COUT = 32, S = 4;
constexpr int K = KH * KW * CIN;                 // 256
constexpr int PATCH_ROWS = 36;                   // input rows under 8 output rows: 7 * 4 + 8
constexpr int B_PLANE = K * COUT * 4;            // 32 KB per TF32 plane
constexpr int WARP_TMA = 0, WARP_MMA0 = 1, WARP_CONV0 = 4, WARP_EPI0 = 12;   // 1-2: MMA issue, 3: idle, 4-11: converters, 12-15: epilogue
constexpr int NSTG = KH;                         // TMEM A stages per tile = filter rows
constexpr uint32_t ACC_COLS = 256, ACOL0 = 256;  // 2 buffers x 2 interleaved accumulators x 64 columns | 8 A stages x 32 columns

struct Params {
  const float* w1s;        // [K][32] weights pre-scaled by 1/255
  const float* bias;       // [32]
  float* Y;                // [M][32]
  int act;
  int nimg, IH, IW, OH, OW, M;
  int patch_pitch;         // bytes per patch row = IW * 4
  int patch_bytes;         // PATCH_ROWS * patch_pitch
  int patch_slot;          // patch_bytes rounded up to 1024
  // direct mode: the patches come straight from the replay store (no gathered batch on the critical path).  Image n of the pass is
  // transition idx[n % half_rows]; the first half_rows images read the s store (map 1) and the rest the s' store (map 2) - or the s'
  // store from the start (sp_first: the target pass).  idx == nullptr: the images are the rows of the gathered batch (map 1).
  const long long* idx;
  int half_rows, sp_first;
};

__host__ __device__ inline int smem_bytes(int patch_slot) { return 2 * B_PLANE + 2 * 2 * patch_slot + 1024 + 512; }

__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmx2, const Params p, int ntiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t b_base = sbase;                                   // hi plane, lo plane right behind it (n atom 1 of the N = 64 operand)
  const uint32_t patch0 = sbase + 2 * B_PLANE;                     // [2 tile buffers][2 image segments][patch_slot]
  const uint32_t bars = patch0 + 4 * p.patch_slot;
  const uint32_t bar_pfull = bars, bar_pempty = bars + 16, bar_afull = bars + 32, bar_aempty = bars + 96, bar_accf = bars + 160, bar_acce = bars + 176;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * B_PLANE + 4 * p.patch_slot + 192);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(bar_pfull + 8 * b, 1); mbar_init(bar_pempty + 8 * b, 8); mbar_init(bar_accf + 8 * b, 2); mbar_init(bar_acce + 8 * b, 4); }
    for (int s = 0; s < NSTG; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA0) tmem_alloc<512>(smem_u32(tmem_slot));
  // resident weights: row k, 16-byte chunk g of the 128-byte row (32 output maps) -> SWIZZLE_128B_BASE32B atoms of 4 k rows
  for (int e = tid; e < K * 8; e += C1_THREADS) {
    const int k = e >> 3, g = e & 7;
    const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1s + k * COUT + g * 4));
    const uint32_t off = (uint32_t)((k >> 2) * 512 + (k & 3) * 128 + ((g ^ ((k & 3) << 1)) * 16));
    sts128(b_base + off, w);
    sts128(b_base + B_PLANE + off, lo_of_trunc4(w));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int PPI = p.OH * p.OW;                                     // output pixels per image

  if (warp == WARP_TMA) {
    // ===== producer: the input rows under the tile, one box per image the tile touches =====
    int buf = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int m0 = t * BM, n0 = m0 / PPI, q0 = m0 - n0 * PPI, oh0 = q0 / p.OW;
      const bool two = (q0 + BM > PPI) && (n0 + 1 < p.nimg);
      mbar_wait(bar_pempty + 8 * buf, ph ^ 1);
      if (lane == 0) {
        const uint32_t dst = patch0 + buf * 2 * p.patch_slot, bar = bar_pfull + 8 * buf;
        mbar_expect_tx(bar, (uint32_t)(p.patch_bytes * (two ? 2 : 1)));
        const CUtensorMap* ma = &tmx; const CUtensorMap* mb = &tmx;
        int ia = n0, ib = n0 + 1;
        if (p.idx) {
          const int ha = n0 / p.half_rows, hb = (n0 + 1) / p.half_rows;
          ia = (int)__ldg(p.idx + (n0 - ha * p.half_rows));
          if ((ha ^ p.sp_first) & 1) ma = &tmx2;
          if (two) { ib = (int)__ldg(p.idx + (n0 + 1 - hb * p.half_rows)); if ((hb ^ p.sp_first) & 1) mb = &tmx2; }
        }
        tma_load_3d(dst, ma, 0, oh0 * S, ia, bar);
        if (two) tma_load_3d(dst + p.patch_slot, mb, 0, 0, ib, bar);
      }
      __syncwarp();
      if (++buf == 2) { buf = 0; ph ^= 1; }
    }
  } else if (warp >= WARP_CONV0 && warp < WARP_EPI0) {
    // ===== converters: two groups of four warps take alternate filter rows; thread = output pixel of the tile =====
    const int cw = warp - WARP_CONV0, grp = cw >> 2, q4 = warp & 3;
    const int row = q4 * 32 + lane;
    int buf = 0; uint32_t pph = 0, aph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int m0 = t * BM, n0 = m0 / PPI, q0 = m0 - n0 * PPI, oh0 = q0 / p.OW;
      const int m = m0 + row;
      const bool live = m < p.M;
      uint32_t src = 0;
      if (live) {
        const int n = m / PPI, q = m - n * PPI, oh = q / p.OW, ow = q - oh * p.OW;
        const int seg = n - n0;                                    // 0: first image of the tile, 1: the next one (its patch starts at input row 0)
        src = patch0 + (uint32_t)(buf * 2 * p.patch_slot + seg * p.patch_slot + (seg ? oh * S : (oh - oh0) * S) * p.patch_pitch + ow * (S * CIN));
      }
      mbar_wait(bar_pfull + 8 * buf, pph);
#pragma unroll 1
      for (int j = grp; j < NSTG; j += 2) {
        uint32_t v[32];
        if (live) {
          const float4 f0 = lds128(src + (uint32_t)(j * p.patch_pitch)), f1 = lds128(src + (uint32_t)(j * p.patch_pitch) + 16);
          const uint32_t w[8] = {__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.z), __float_as_uint(f0.w),
                                 __float_as_uint(f1.x), __float_as_uint(f1.y), __float_as_uint(f1.z), __float_as_uint(f1.w)};
#pragma unroll
          for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int b = 0; b < 4; ++b) v[4 * q + b] = byte_to_f32(w[q], b);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        mbar_wait(bar_aempty + 8 * j, aph ^ 1);                    // the MMAs that read this TMEM stage in the previous tile have retired
        tc_fence_after();
        const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16) + ACOL0 + (uint32_t)(j * 32);
        tmem_st16(ta, *reinterpret_cast<const uint32_t(*)[16]>(&v[0]));
        tmem_st16(ta + 16, *reinterpret_cast<const uint32_t(*)[16]>(&v[16]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull + 8 * j);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pempty + 8 * buf);            // this warp is done reading the patch
      aph ^= 1;
      if (++buf == 2) { buf = 0; pph ^= 1; }
    }
  } else if (warp >= WARP_EPI0) {
    // ===== epilogue: (hi + hi') + (lo + lo') + bias -> activation -> 128-byte rows =====
    const int q4 = warp & 3;
    const bool st8 = dqn::al32(p.Y);
    int buf = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int m = t * BM + q4 * 32 + lane;
      mbar_wait(bar_accf + 8 * buf, ph);
      tc_fence_after();
      const uint32_t tb = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * 128);
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t hh0[16], hl0[16], hh1[16], hl1[16];
        tmem_ld16_nowait(tb + c0, hh0); tmem_ld16_nowait(tb + 32 + c0, hl0);
        tmem_ld16_nowait(tb + 64 + c0, hh1); tmem_ld16_nowait(tb + 96 + c0, hl1);
        float4 bq[4];                                              // bias of the chunk, fetched in the shadow of the TMEM loads
#pragma unroll
        for (int j = 0; j < 4; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * j));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < p.M) {
          float* out = p.Y + (long long)m * COUT + c0;
          float4 o4[4];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = bq[j >> 2];
            float y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float hi = __fadd_rn(__uint_as_float(hh0[j + u]), __uint_as_float(hh1[j + u]));
              const float lo = __fadd_rn(__uint_as_float(hl0[j + u]), __uint_as_float(hl1[j + u]));
              y[u] = __fadd_rn(hi, lo);
            }
            o4[j >> 2] = make4(dqn::act_apply(y[0] + b.x, p.act), dqn::act_apply(y[1] + b.y, p.act), dqn::act_apply(y[2] + b.z, p.act), dqn::act_apply(y[3] + b.w, p.act));
          }
          if (st8) { dqn::st_global_v8(out, o4[0], o4[1]); dqn::st_global_v8(out + 8, o4[2], o4[3]); }     // whole sectors (see igemm.cuh)
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(out)[j] = o4[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
      if (++buf == 2) { buf = 0; ph ^= 1; }
    }
  } else if (warp == WARP_MMA0 || warp == WARP_MMA0 + 1) {
    // ===== MMA issue: warp r takes the filter rows j = r, r + 2, ... into accumulator r of the tile's buffer =====
    constexpr uint32_t idesc = make_idesc2(BM, 2 * COUT, false, true);
    const int role = warp - WARP_MMA0;
    if (lane == 0) {                                               // one lane runs the whole loop: nothing but waits, MMAs and commits
      const uint64_t dbase = make_desc(b_base, B_PLANE, 512, 1u);
      int buf = 0; uint32_t ph = 0, aph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(bar_acce + 8 * buf, ph ^ 1);                     // the epilogue has drained this buffer (first two tiles: immediate)
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(buf * 128 + role * 64);
#pragma unroll 1
        for (int j = role; j < NSTG; j += 2) {
          mbar_wait(bar_afull + 8 * j, aph);
          tc_fence_after();
          const uint32_t a_t = tmem + ACOL0 + (uint32_t)(j * 32);
          const uint64_t d0 = dbase + (uint64_t)(j * 4 * (1024 >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_tf32_ts(acc, a_t + ks * 8, d0 + (uint64_t)(ks * (1024 >> 4)), idesc, (ks > 0 || j > role) ? 1u : 0u);
          umma_commit(bar_aempty + 8 * j);
          if (j + 2 >= NSTG) umma_commit(bar_accf + 8 * buf);
        }
        aph ^= 1;
        if (++buf == 2) { buf = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == WARP_MMA0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// host: geometry this kernel covers, and the 3-D tensor map of the byte batch viewed as u32 pixels {IW, IH, images}
inline bool geometry_ok(const dqn::ConvGeom& g) {
  return g.KH == KH && g.KW == KW && g.Cin == CIN && g.Cout == COUT && g.S == S && g.IW <= 256 && g.IH >= KH && (g.IW * CIN) % 16 == 0 &&
         ((BM - 1) / g.OW + 1) * S + KH - S <= PATCH_ROWS;     // 128 consecutive output pixels of one image span at most 8 output rows
}
inline bool make_map(CUtensorMap* m, const void* xb, int nimg, const dqn::ConvGeom& g) {
  if (!tma_api().load() || (reinterpret_cast<uintptr_t>(xb) & 15)) return false;
  cuuint64_t gd[3] = {(cuuint64_t)g.IW, (cuuint64_t)g.IH, (cuuint64_t)nimg}, gs[2] = {(cuuint64_t)g.IW * 4, (cuuint64_t)g.IH * g.IW * 4};
  cuuint32_t bx[3] = {(cuuint32_t)g.IW, (cuuint32_t)PATCH_ROWS, 1}, es[3] = {1, 1, 1};
  return tma_api().tiled(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(xb), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline Params make_params(const dqn::ConvFwdOp& op) {
  Params p{};
  p.w1s = op.Ws; p.bias = op.W + (long long)op.K * op.N; p.Y = op.Y; p.act = op.act;
  p.nimg = op.nimg; p.IH = op.g.IH; p.IW = op.g.IW; p.OH = op.g.OH; p.OW = op.g.OW; p.M = op.M;
  p.patch_pitch = op.g.IW * CIN; p.patch_bytes = PATCH_ROWS * p.patch_pitch; p.patch_slot = (p.patch_bytes + 1023) / 1024 * 1024;
  return p;
}
}  // namespace c1

#ifndef TC_KERNEL_ONLY
namespace {
// first conv layer, forward, on the raw byte batch (DQN_MATH_3XTF32 with byte observations and the Nature-DQN first-layer geometry)
bool tc_conv1_eligible(dqn_engine* e, const dqn::ConvGeom& g) {
  return e->cfg.math_mode == DQN_MATH_3XTF32 && e->tc_c1 && e->arena && e->a8 && e->elem_bytes == 1 && g.Cout == c1::COUT && c1::geometry_ok(g);
}
// tensor maps of the two observation stores (direct mode: one image = one stored transition)
void tc_conv1_init(dqn_engine* e) {
  e->c1_maps_ok = false;
  if (e->convs.empty() || e->lstm || !tc_conv1_eligible(e, e->convs[0].g) || e->cap > 0x7fffffffLL) return;
  e->c1_maps_ok = c1::make_map(&e->c1_map_s, e->store_s, (int)e->cap, e->convs[0].g) && c1::make_map(&e->c1_map_sp, e->store_sp, (int)e->cap, e->convs[0].g);
}
bool tc_conv1_fwd(dqn_engine* e, const char* name, const dqn::ConvFwdOp& op, double flops, double bytes) {
  if (e->cfg.math_mode != DQN_MATH_3XTF32 || !e->tc_c1 || !e->arena) return false;
  if (!op.a8 || !op.x_u8 || !op.Ws || op.N != c1::COUT || op.K != c1::K || !c1::geometry_ok(op.g)) return false;
  CUtensorMap tm, tm2;
  c1::Params p = c1::make_params(op);
  if (e->c1_direct_now) {                      // this step's passes read the replay store through the sampled indices (engine.cu: enqueue_step)
    if (!e->c1_maps_ok) return false;
    tm = e->c1_map_s; tm2 = e->c1_map_sp;
    p.idx = e->idx_d; p.half_rows = e->B; p.sp_first = (op.X != (const void*)e->xb) ? 1 : 0;
  } else {
    if (!c1::make_map(&tm, op.X, op.nimg, op.g)) return false;
    tm2 = tm;
  }
  const int smem = c1::smem_bytes(p.patch_slot);
  if (smem > 227 * 1024) return false;
  static unsigned long long attr_set[4] = {0, 0, 0, 0};
  const int dev = e->cfg.device & 255;
  if (!((attr_set[dev >> 6] >> (dev & 63)) & 1ull)) {
    CK(cudaFuncSetAttribute(c1::conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev >> 6] |= 1ull << (dev & 63);
  }
  const int ntiles = (op.M + tc::BM - 1) / tc::BM;
  Scope sc(e, name, flops, bytes);
  c1::conv1_fwd_kernel<<<std::min(ntiles, e->nsm), c1::C1_THREADS, smem, e->ls>>>(tm, tm2, p, ntiles);
  CK(cudaGetLastError());
  return true;
}
}  // namespace
#endif
