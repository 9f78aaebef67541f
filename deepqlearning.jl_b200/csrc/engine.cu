// engine.cu - host side of libdqn_b200.so: handle lifecycle, HBM layout, the step schedule (eager or one
// CUDA graph), parameter import/export in Flux order, and the C-ABI of include/dqn_b200.h.
//
// The step (dqn_train_step) is the reference's batch_train! (src/solver.jl:191-236):
//   sample -> gather -> online forward on [s ; s'] -> target forward on s' -> fused head (dueling, Double-Q,
//   Bellman target, IS-Huber, dQ, td, new priorities) -> reverse pass -> [NCCL all-reduce] -> fused Adam +
//   max|g| -> sum-tree refresh -> publish (loss, grad_norm).
// There is no CPU fallback anywhere in this file: every numerical result is produced by a kernel launch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dqn_b200.h"
#include "igemm.cuh"
#include "kernels.cuh"
#include <unistd.h>
#include "lstm.cuh"
#include "peer_ar.cuh"
#include "tc_gemm.cuh"

using namespace dqn;

namespace {

thread_local std::string g_create_error;

struct Err {
  int code; std::string msg;
};

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess) throw Err{DQN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)}; \
  } while (0)

[[noreturn]] void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  throw Err{code, buf};
}

// ---- NCCL through dlopen: the library resolves the NCCL already loaded in the process (torch's), else the system one.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& why) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { why = "NCCL symbols missing"; return false; }
    return true;
  }
};
NcclApi g_nccl;

struct Mat { long long off; int K, N, act; };        // augmented matrix [(K+1)][N] at theta + off
struct ConvL { ConvGeom g; Mat w; };

struct ProfRec { cudaEvent_t a, b; std::string name; double flops, bytes; };

constexpr int MAXD = DQN_MAX_LAYERS;
constexpr int COLSUM_CTAS = 296;

struct ActBufs {
  std::vector<float*> conv_out;
  float* tow_out[2][MAXD] = {};
};

}  // namespace

struct dqn_engine {
  dqn_config_t cfg{};
  std::string err;
  cudaStream_t stream = nullptr;
  // second lane: the target-network forward and the weight gradients run beside the online forward / the dgrad chain
  cudaStream_t stream2 = nullptr; float* ws2 = nullptr;
  cudaStream_t stream3 = nullptr;                       // NCCL lane: the fc bucket is reduced while the conv backward still runs
  long long tower_off = 0;                              // internal offset where the tower (Dense) parameters start
  cudaStream_t ls = nullptr; float* lws = nullptr;     // lane the contraction launchers currently target
  std::vector<cudaEvent_t> evs; size_t ev_next = 0; int use_streams = 1;
  int nsm = 148;
  // topology
  std::vector<ConvL> convs;
  int hwc = 0;                     // 1: observations stored H,W,C (conv trunk), 0: flat in Flux order
  long long obs_elems = 0, obs_row_bytes = 0; int elem_bytes = 4;
  int feat = 0;                    // trunk output features (== obs_elems when there is no trunk)
  int ntow = 1, depth = 0;
  Mat tow[2][MAXD];
  long long nflux = 0, nint = 0;   // Flux parameter count, internal (padded) float count
  std::vector<long long> perm;     // Flux flat index -> internal index
  // parameters / optimiser state
  float *theta = nullptr, *theta_t = nullptr, *adam_m = nullptr, *adam_v = nullptr, *grad = nullptr;
  // replay
  long long cap = 0, cursor = 0, curr_size = 0; int P = 0;
  uint8_t *store_s = nullptr, *store_sp = nullptr, *done = nullptr;
  int* act = nullptr; float* rew = nullptr; float* tree = nullptr;
  DevState* st = nullptr;
  // batch
  int B = 0, rows_on = 0;
  long long* idx_d = nullptr; uint8_t* xb = nullptr;
  int* a_b = nullptr; float *r_b = nullptr, *d_b = nullptr, *w_b = nullptr;
  ActBufs on, tg;
  std::vector<float*> conv_delta; float* tow_delta[2][MAXD] = {};
  float *q_s = nullptr, *q_sp_on = nullptr, *q_sp_tg = nullptr, *y = nullptr, *td = nullptr, *newp = nullptr; int* best_a = nullptr;
  float* ws = nullptr; long long ws_floats = 0;
  // staging for host I/O
  uint8_t* stage = nullptr; long long stage_bytes = 0;
  float* host_out = nullptr; float* host_out_dev = nullptr;   // mapped pinned, two slots of (loss, grad_norm, error, step): step s -> slot s & 1
  unsigned long long n_launched = 0;                           // gradient steps launched so far (== DevState.step once they finish)
  cudaEvent_t step_ev[2] = {nullptr, nullptr};                 // recorded behind step s in slot s & 1 (dqn_step_result)
  cudaStream_t copy_stream = nullptr;                          // dqn_replay_add: host -> staging copies run beside the step in flight
  uint8_t* add_stage[2] = {nullptr, nullptr}; long long add_stage_bytes[2] = {0, 0};
  cudaEvent_t add_free[2] = {nullptr, nullptr};                // staging slot consumed by its ingest kernels
  unsigned long long add_seq = 0;
  // dqn_replay_add's ingest kernels run on their own lane: behind the priority update of the step in flight (which publishes a tree
  // epoch in device memory; an external event node in the captured graph cost 8 us per step) instead of behind the whole step, so a
  // caller that runs one step ahead never leaves the GPU idle between steps.  Every other call joins the lane first (guard).
  cudaStream_t ingest_stream = nullptr; cudaEvent_t ingest_done = nullptr, main_mark = nullptr;
  int ingest_lane = 1;                                         // DQN_INGEST_LANE=0: ingest on the main stream, behind the whole step in flight
  bool ingest_pending = false;                                 // ingest work the main stream has not been ordered behind yet
  bool main_dirty = true;                                      // main-stream work since the last step that the ingest lane is not ordered behind
  long long* idx_h = nullptr;                                  // pinned
  // graph
  cudaGraphExec_t graph_sample = nullptr, graph_idx = nullptr;
  // nccl
  ncclComm_t comm = nullptr;
  int nccl_ctas = 0;       // CTAs NCCL may use (NCCL_MAX_CTAS) = SMs the persistent grids leave free while a reduction is in flight
  int ar_ctas = 0;         // CTAs of the reduction launched last
  int sm_reserve = 0;      // SMs currently held back from the persistent tensor-core grids
  // measurement
  cudaEvent_t t0 = nullptr, t1 = nullptr, copy_done = nullptr;
  int counting = 0, launches = 0, launches_per_step = 0;
  int profiling = 0; std::vector<ProfRec> prof; std::vector<cudaEvent_t> ev_pool;
  uint32_t* flush_buf = nullptr; long long flush_n = 0;
  bool capturing = false;
  // tensor-core side buffers: the sampled byte batch as fp32 (raw values k, exact in TF32), the first conv layer's weights
  // pre-scaled by 1/255 for both networks, and the {1,0,0,0} chunk that realises the bias column of [x 1]
  float* arena = nullptr; long long lo_delta = 0;
  float* xb_f = nullptr;
  float *w_on_s = nullptr, *w_tg_s = nullptr, *ones = nullptr;
  long long w_scale_lo = 0, w_scale_hi = 0;
  int tc_split = 0; int tc_tail = 0; int tc_tma = 1; int tc_c1 = 1; int tc_tma_wgrad = 0; int dgrad_merge = 1;
  float* wm[DQN_MAX_LAYERS] = {};   // per strided conv layer: weights rearranged for the class-merged dgrad (ConvDgradMergedOp)
  float* colsum_part = nullptr; unsigned int* colsum_ticket = nullptr;
  bool towers_updated = false;
  // recurrent engines (one LSTM as the trunk, SOLVER:239-287): the network batch is trace_length * batch_size rows, time-major
  int lstm = 0, lstm_in = 0, Hh = 0, T = 0, Bep = 0, Lmax = 0;
  Mat wi{}, wh{}; long long h0_off = 0, c0_off = 0;   // internal parameters: W_i^T with the bias row, W_h^T, state0
  float *ep_s = nullptr, *ep_sp = nullptr, *ep_r = nullptr; int *ep_a = nullptr, *ep_len = nullptr; uint8_t* ep_done = nullptr;
  float *xproj_on = nullptr, *xproj_tg = nullptr;      // [2TB][4H], [TB][4H]: x W_i^T + b of every step
  float *hs_on = nullptr, *hs_tg = nullptr;            // [(2T+1) Bep][H]: h0 | h of the s pass | h of the s' pass;  [(T+1) Bep][H]
  float *cs_on = nullptr, *cs_sp = nullptr, *cs_tg = nullptr;   // cell states: [(T+1) Bep][H] of the s pass (kept for BPTT), ping-pong pairs of the others
  float *gates_s = nullptr, *dgates = nullptr, *dcell = nullptr, *whT = nullptr;
  float *h_act = nullptr, *c_act = nullptr; int act_rows = 0;   // acting hidden state (POLICY:32-34), one row per lane
  int* ep_start_d = nullptr;
  // trunk hand-over to the towers (conv trunk: the last conv layer; LSTM trunk: the hidden states)
  float *trunk_on = nullptr, *trunk_tg = nullptr, *trunk_delta = nullptr; int trunk_act = 0;
  // gradient all-reduce over NVLink peer memory (peer_ar.cuh); NCCL stays the fallback (no peer access, world > 8, DQN_PEER_AR=0)
  int peer_ctas = 16;      // CTAs of a large reduction (DQN_PEER_CTAS, <= PEER_MAXG): bytes in flight over NVLink vs SMs taken from the GEMMs
  bool peer_ar = false; PeerArArgs peer{}; unsigned long long* peer_flags = nullptr; void* peer_opened[2 * PEER_MAX] = {}; int peer_nopened = 0;
  // first conv layer straight from the replay store: its TMA patch loads take the sampled indices, so the row gather leaves the critical
  // path (it still runs, on the third lane, for the weight-gradient operand).  DQN_C1_DIRECT=0: gather first, as before.
  int c1_direct = 1; bool c1_maps_ok = false, c1_direct_now = false; CUtensorMap c1_map_s, c1_map_sp;
  int head_small = 1;      // register-resident head kernel for small action sets (DQN_HEAD_SMALL=0: the generic one)
  int lstm_seq = 1;        // recurrent engines: the whole recurrence of a pass in one cluster launch (lstm_seq_*_kernel); 0 = one launch per time step
  int fuse_head_all = 1;   // output layers of all three passes + head + their input gradient in one launch (head_fused_kernel)
  float* hub = nullptr;    // per-sample Huber values of that kernel (deterministic loss reduction)
  int fuse_heads = 1;      // thin output layers (N <= 8) by heads_fwd_kernel / heads_dgrad_kernel instead of the tiled contraction
  int a8 = 0;              // 1: the first conv layer (forward and weight gradient) reads the byte batch directly, no fp32 copy of it exists
  int merge_fwd = 0;       // 1: online and target forward share launches layer by layer; measured slower than two lanes on B200 (0.571 vs 0.539 ms/step)
};

namespace {

using E = dqn_engine;

template <class T> T* dalloc(long long n) {
  T* p = nullptr;
  CK(cudaMalloc(&p, std::max<long long>(n, 1) * sizeof(T)));
  CK(cudaMemset(p, 0, std::max<long long>(n, 1) * sizeof(T)));
  return p;
}

// ---- launch bookkeeping -------------------------------------------------------------------------
struct Scope {
  E* e; bool on; size_t slot = 0;
  Scope(E* e_, const char* name, double flops, double bytes) : e(e_), on(e_->profiling && !e_->capturing) {
    if (e->counting) e->launches++;
    if (on) {
      ProfRec r; r.name = name; r.flops = flops; r.bytes = bytes;
      CK(cudaEventCreate(&r.a)); CK(cudaEventCreate(&r.b));
      CK(cudaEventRecord(r.a, e->ls));
      slot = e->prof.size();
      e->prof.push_back(r);
    }
  }
  ~Scope() { if (on) cudaEventRecord(e->prof[slot].b, e->ls); }
};

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- fp32 implicit-GEMM launch ------------------------------------------------------------------
template <class Op>
void launch_igemm(E* e, const char* name, Op a, Op b, int nz, bool allow_split, double flops, double bytes) {
  Op a0 = a;
  if (Op::Z_IS_CLASS) a0.set_class(0);
  int M = a0.M, N = a0.N, K = a0.K;
  if (!Op::Z_IS_CLASS && nz == 2) { M = std::max(M, b.M); N = std::max(N, b.N); K = std::max(K, b.K); }
  if (M <= 0 || N <= 0) return;
  int cfgid;   // 0: 128x64, 1: 64x64, 2: 128x32
  auto ctas = [&](int bm, int bn) { return (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn) * nz; };
  if (N <= 32) cfgid = 2;
  else if (ctas(128, 64) >= e->nsm) cfgid = 0;
  else cfgid = 1;
  const int bm = cfgid == 1 ? 64 : 128, bn = cfgid == 2 ? 32 : 64;
  int nsplit = 1;
  if (allow_split || (!Op::Z_IS_CLASS && 2 * ctas(bm, bn) <= e->nsm)) {     // also split tiny grids (the 512->1|6 heads): latency, not flops
    const int ktiles = (K + IGEMM_BK - 1) / IGEMM_BK;
    long long c = ctas(bm, bn);
    nsplit = (int)std::max<long long>(1, std::min<long long>({(2LL * e->nsm + c - 1) / c, (long long)ktiles / 4, 64LL}));
    const long long stride = (long long)M * N;
    if (nsplit > 1 && (long long)nz * nsplit * stride > e->ws_floats) nsplit = (int)std::max<long long>(1, e->ws_floats / (nz * stride));
  }
  dim3 grid((M + bm - 1) / bm, (N + bn - 1) / bn, nz * nsplit);
  const long long ws_stride = (long long)M * N;
  {
    Scope sc(e, name, flops, bytes);
    if (cfgid == 0) igemm_kernel<128, 64, 8, 4, Op><<<grid, IGEMM_THREADS, 0, e->ls>>>(a, b, nsplit, e->lws, ws_stride);
    else if (cfgid == 1) igemm_kernel<64, 64, 4, 4, Op><<<grid, IGEMM_THREADS, 0, e->ls>>>(a, b, nsplit, e->lws, ws_stride);
    else igemm_kernel<128, 32, 4, 4, Op><<<grid, IGEMM_THREADS, 0, e->ls>>>(a, b, nsplit, e->lws, ws_stride);
    CK(cudaGetLastError());
  }
  if (nsplit > 1) {
    Scope sc(e, "splitk_reduce", 0, (double)(nsplit + 1) * ws_stride * nz * 4);
    dim3 g2((unsigned)std::min<long long>((ws_stride + 255) / 256, 4 * e->nsm), nz);
    splitk_reduce_kernel<Op><<<g2, 256, 0, e->ls>>>(a, b, nsplit, e->lws, ws_stride);
    CK(cudaGetLastError());
  }
}

// lane switch + fork/join helpers (work identically in eager mode and under stream capture)
struct Lane {
  E* e; cudaStream_t s0; float* w0;
  Lane(E* e_, bool second) : e(e_), s0(e_->ls), w0(e_->lws) { if (second) { e->ls = e->stream2; e->lws = e->ws2; } }
  ~Lane() { e->ls = s0; e->lws = w0; }
};
cudaEvent_t next_event(E* e) {
  if (e->ev_next == e->evs.size()) { cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e->evs.push_back(ev); }
  return e->evs[e->ev_next++];
}
void order_after(E* e, cudaStream_t later, cudaStream_t earlier) {     // everything enqueued on `later` from now on waits for `earlier` up to now
  cudaEvent_t ev = next_event(e);
  CK(cudaEventRecord(ev, earlier));
  CK(cudaStreamWaitEvent(later, ev, 0));
}

// ---- network schedule ---------------------------------------------------------------------------
// One forward pass: parameters P applied to the rows of X.  Xs: the input batch as fp32 for the tensor-core path (null => fp32
// CUDA-core kernels only); w1s: the first conv layer's weights pre-scaled by 1/255 when Xs holds raw byte values.
struct Pass { const float* P; const void* X; int x_u8; int rows; ActBufs* bufs; const char* tag; const float* Xs; const float* w1s; bool tc; const float* trunk_out = nullptr; bool skip_last = false; };   // skip_last: the output layers are computed by head_fused_kernel

// Layer by layer over all passes.  On the tensor-core path the passes of a layer (online network on [s ; s'], target network on s')
// and the two towers of a Dense layer share ONE launch: the persistent kernels serialise on the machine anyway, and one launch
// rounds the tile count to SM multiples once instead of once per pass.
void forward(E* e, const Pass* ps, int np) {
  const void* cur[2]; int cur_u8[2]; const float* cur_s[2];
  for (int p = 0; p < np; ++p) { cur[p] = ps[p].X; cur_u8[p] = ps[p].x_u8; cur_s[p] = ps[p].Xs; }
  for (int p = 0; p < np; ++p) if (ps[p].trunk_out) { cur[p] = ps[p].trunk_out; cur_u8[p] = 0; cur_s[p] = ps[p].trunk_out; }   // recurrent trunk: the towers read the hidden states
  char nm[64];
  for (size_t l = 0; l < e->convs.size(); ++l) {
    const ConvL& c = e->convs[l];
    ConvFwdOp ops[2]; double fl[2], by[2], fls = 0, bys = 0;
    for (int p = 0; p < np; ++p) {
      ConvFwdOp& op = ops[p]; op = ConvFwdOp{};
      const int rows = ps[p].rows;
      op.X = cur[p]; op.x_u8 = cur_u8[p]; op.W = ps[p].P + c.w.off; op.Y = ps[p].bufs->conv_out[l]; op.act = c.w.act; op.nimg = rows; op.g = c.g;
      op.M = rows * c.g.OH * c.g.OW; op.N = c.g.Cout; op.K = c.w.K;
      op.vecA = (c.g.Cin % 4 == 0); op.vecB = (c.g.Cout % 4 == 0);
      if (ps[p].tc) {
        const bool raw = (l == 0 && cur_u8[p]);                 // raw byte observations: exact single-plane operand, weights pre-scaled by 1/255
        op.Xs = cur_s[p]; op.Ws = raw ? ps[p].w1s : ps[p].P + c.w.off; op.a_single = raw;
        if (raw && e->a8) { op.a8 = 1; op.Xs = nullptr; }       // ... read straight from the byte batch
      }
      fl[p] = 2.0 * op.M * op.N * op.K;
      by[p] = (double)rows * c.g.IH * c.g.IW * c.g.Cin * (cur_u8[p] ? 1 : 4) + (double)(op.K + 1) * op.N * 4 + (double)op.M * op.N * 4;
      fls += fl[p]; bys += by[p];
    }
    snprintf(nm, sizeof nm, "conv%zu_fwd", l + 1);
    if (!(np > 1 && tc_conv_fwd(e, nm, ops, np, fls, bys))) {
      for (int p = 0; p < np; ++p) {
        snprintf(nm, sizeof nm, "conv%zu_fwd_%s", l + 1, ps[p].tag);
        if (l == 0 && tc_conv1_fwd(e, nm, ops[p], fl[p], by[p])) continue;      // raw byte observations: the dedicated first-layer kernel
        if (!tc_conv_fwd(e, nm, &ops[p], 1, fl[p], by[p])) launch_igemm(e, nm, ops[p], ops[p], 1, false, fl[p], by[p]);
      }
    }
    for (int p = 0; p < np; ++p) { cur[p] = ps[p].bufs->conv_out[l]; cur_u8[p] = 0; cur_s[p] = ps[p].bufs->conv_out[l]; }
  }
  for (int l = 0; l < e->depth; ++l) {
    DenseFwdOp ops[4]; double fl[2] = {0, 0}, by[2] = {0, 0};
    for (int p = 0; p < np; ++p)
      for (int t = 0; t < e->ntow; ++t) {
        const Mat& w = e->tow[t][l];
        const int rows = ps[p].rows;
        DenseFwdOp& op = ops[p * e->ntow + t]; op = DenseFwdOp{};
        op.X = l == 0 ? cur[p] : (const void*)ps[p].bufs->tow_out[t][l - 1]; op.ldx = w.K; op.x_u8 = l == 0 ? cur_u8[p] : 0;
        op.W = ps[p].P + w.off; op.C = ps[p].bufs->tow_out[t][l]; op.ldc = w.N; op.act = w.act; op.M = rows; op.N = w.N; op.K = w.K;
        op.vecA = (w.K % 4 == 0) && al16(op.X); op.vecB = (w.N % 4 == 0);
        if (ps[p].tc) { op.Xs = l == 0 ? (cur_u8[p] ? nullptr : cur_s[p]) : ps[p].bufs->tow_out[t][l - 1]; op.Ws = ps[p].P + w.off; op.a_single = 0; }
        fl[p] += 2.0 * rows * op.N * op.K;
        by[p] += 4.0 * ((double)rows * op.K / (l == 0 ? e->ntow : 1) + (double)(op.K + 1) * op.N + (double)rows * op.N);
      }
    snprintf(nm, sizeof nm, "dense%d_fwd", l + 1);
    if (l == e->depth - 1 && ps[0].skip_last) continue;
    if (l == e->depth - 1 && l > 0 && e->fuse_heads && e->tow[e->ntow - 1][l].N <= HEADS_MAXN && e->tow[0][l].N <= HEADS_MAXN) {
      // the thin output layers of all passes and towers: one warp per row
      HeadJobs jobs{}; int total = 0;
      for (int p = 0; p < np; ++p)
        for (int t = 0; t < e->ntow; ++t) {
          const DenseFwdOp& op = ops[p * e->ntow + t];
          jobs.j[jobs.n++] = HeadJob{(const float*)op.X, op.ldx, op.W, op.C, op.M, op.N, op.K, op.act};
          total += op.M;
        }
      snprintf(nm, sizeof nm, "heads_fwd_%s", np > 1 ? "both" : ps[0].tag);
      Scope sc(e, nm, fl[0] + fl[1], by[0] + by[1]);
      heads_fwd_kernel<<<(total + 7) / 8, 256, 0, e->ls>>>(jobs);
      CK(cudaGetLastError());
      continue;
    }
    if (!(np > 1 && tc_dense_fwd(e, nm, ops, np * e->ntow, fl[0] + fl[1], by[0] + by[1]))) {
      for (int p = 0; p < np; ++p) {
        DenseFwdOp* o = ops + p * e->ntow;
        snprintf(nm, sizeof nm, "dense%d_fwd_%s", l + 1, ps[p].tag);
        if (!tc_dense_fwd(e, nm, o, e->ntow, fl[p], by[p])) launch_igemm(e, nm, o[0], o[e->ntow - 1], e->ntow, false, fl[p], by[p]);
      }
    }
  }
}

void enqueue_adam(E* e, long long lo, long long hi, cudaStream_t s);

// rearranged weight copies of the strided conv layers for the class-merged input gradient: the online weights do not change inside a
// step, so this runs early on the target lane, off the critical path
void prepare_dgrad_weights(E* e) {
  if (!e->arena || !e->dgrad_merge || e->cfg.math_mode != DQN_MATH_3XTF32) return;
  for (size_t l = 1; l < e->convs.size(); ++l) {
    const ConvL& c = e->convs[l];
    if (!e->wm[l] || !ConvDgradMergedOp::geometry_ok(c.g)) continue;
    Scope sc(e, "dgrad_merge_weights", 0, 8.0 * c.w.K * c.g.Cout);
    dgrad_merge_weights_kernel<<<64, 256, 0, e->ls>>>(e->theta + c.w.off, e->wm[l], c.g.KH, c.g.KW, c.g.S, c.g.Cin, c.g.Cout);
    CK(cudaGetLastError());
  }
}

// sum of the gradient elements [off, off + n) over the ranks, in place, on stream s: peer-memory kernel or NCCL
void allreduce_grad(E* e, long long off, long long n, int bucket, cudaStream_t s) {
  if (e->peer_ar && (off % 4) == 0 && (n % 4) == 0 && n > 0) {
    PeerArArgs a = e->peer; a.off = off; a.n = n; a.bucket = bucket;
    const long long per4 = (n / 4 + a.world - 1) / a.world;
    const int G = (int)std::max<long long>(1, std::min<long long>(e->peer_ctas, (per4 + 1023) / 1024));     // >= 2 float4 per thread, else fewer CTAs
    peer_allreduce_kernel<<<G, 512, 0, s>>>(a);
    CK(cudaGetLastError());
    e->ar_ctas = G;
    return;
  }
  ncclResult_t r = g_nccl.AllReduce(e->grad + off, e->grad + off, (size_t)n, ncclFloat, ncclSum, e->comm, s);
  if (r != ncclSuccess) fail(DQN_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  e->ar_ctas = e->nccl_ctas;
}

void backward(E* e, bool conc, bool head_fused = false) {
  const int B = e->B;
  char nm[64];
  const bool trunk = !e->convs.empty() || e->lstm;
  float* dfeat = trunk ? e->trunk_delta : nullptr;
  for (int l = e->depth - 1; l >= 0; --l) {
    // weight + bias gradients of both towers in one launch
    DenseWgradOp wg[2];
    for (int t = 0; t < e->ntow; ++t) {
      const Mat& w = e->tow[t][l];
      DenseWgradOp& op = wg[t]; op = DenseWgradOp{};
      if (l == 0) { op.X = trunk ? (const void*)e->trunk_on : (const void*)e->xb; op.x_u8 = trunk ? 0 : (e->elem_bytes == 1); }
      else { op.X = e->on.tow_out[t][l - 1]; op.x_u8 = 0; }
      op.ldx = w.K; op.D = e->tow_delta[t][l]; op.ldd = w.N; op.dW = e->grad + w.off; op.M = w.K + 1; op.N = w.N; op.K = B;
      op.vecA = (w.K % 4 == 0) && al16(op.X); op.vecB = (w.N % 4 == 0);
      if (e->arena) {
        const bool raw_bytes = (l == 0 && !trunk && e->elem_bytes == 1);
        op.Xs = l == 0 ? (trunk ? e->trunk_on : (raw_bytes ? nullptr : (const float*)e->xb)) : e->on.tow_out[t][l - 1];
        op.Ds = e->tow_delta[t][l]; op.ones = e->ones; op.a_single = 0; op.out_scale = 0.f;
      }
    }
    // TMA feed: the ones row of [x 1] is not a box of the activation matrix - the bias gradient (column sums of delta) goes to colsum_kernel
    // (measured: the extra column-sum launch costs more than the TMA feed gains on this operand, 27.1 + 12.7 us against 27.0 us for fc1 -
    //  the weight gradients keep the cp.async feed with the ones row unless DQN_TC_TMA_WGRAD=1)
    bool split_bias = e->arena && e->tc_tma && e->tc_tma_wgrad && e->cfg.math_mode == DQN_MATH_3XTF32 && B >= 32;
    for (int t = 0; t < e->ntow; ++t) { const Mat& w = e->tow[t][l]; split_bias = split_bias && wg[t].Xs && (w.K % 4 == 0) && w.K >= 64 && (w.N % 32 == 0) && w.N <= 1024; }
    if (split_bias) for (int t = 0; t < e->ntow; ++t) { wg[t].no_bias = 1; wg[t].M = e->tow[t][l].K; }
    if (e->ntow == 1) wg[1] = wg[0];
    snprintf(nm, sizeof nm, "dense%d_wgrad", l + 1);
    double fl = 0, by = 0;
    for (int t = 0; t < e->ntow; ++t) { fl += 2.0 * wg[t].M * wg[t].N * B; by += 4.0 * ((double)B * wg[t].M + (double)B * wg[t].N + (double)wg[t].M * wg[t].N); }
    {
      if (conc) order_after(e, e->stream2, e->stream);          // delta of this layer is complete on the main lane
      Lane lane(e, conc);
      if (!tc_dense_wgrad(e, nm, wg, e->ntow, fl, by)) launch_igemm(e, nm, wg[0], wg[1], e->ntow, true, fl, by);
      if (split_bias) {
        snprintf(nm, sizeof nm, "dense%d_bgrad", l + 1);
        for (int t = 0; t < e->ntow; ++t) {
          const Mat& w = e->tow[t][l];
          Scope sc(e, nm, 0, 4.0 * B * w.N);
          colsum_kernel<<<std::min(COLSUM_CTAS, (B + 7) / 8), 256, 0, e->ls>>>(e->tow_delta[t][l], (long long)B, w.N, e->grad + w.off + (long long)w.K * w.N,
                                                                             e->colsum_part + (e->ls == e->stream ? 0 : COLSUM_CTAS * 1024), e->colsum_ticket + (e->ls == e->stream ? 0 : 1));
          CK(cudaGetLastError());
        }
      }
    }
    // input gradients
    if (l > 0 && l == e->depth - 1 && head_fused) {
      // head_fused_kernel already wrote the gradient into the last hidden layer
    } else if (l > 0) {
      DenseDgradOp dg[2];
      for (int t = 0; t < e->ntow; ++t) {
        const Mat& w = e->tow[t][l]; const Mat& wp = e->tow[t][l - 1];
        DenseDgradOp& op = dg[t]; op = DenseDgradOp{};
        op.D = e->tow_delta[t][l]; op.ldd = w.N; op.W = e->theta + w.off; op.dX = e->tow_delta[t][l - 1]; op.ldx = w.K;
        op.Y = e->on.tow_out[t][l - 1]; op.ldy = w.K; op.act = wp.act; op.accumulate = 0; op.apply_act = 1;
        op.M = B; op.N = w.K; op.K = w.N; op.vecA = (w.N % 4 == 0); op.vecB = (w.N % 4 == 0);
        if (e->arena) { op.Ds = e->tow_delta[t][l]; op.Ws = e->theta + w.off; op.a_single = 0; }
      }
      if (e->ntow == 1) dg[1] = dg[0];
      snprintf(nm, sizeof nm, "dense%d_dgrad", l + 1);
      double fl = 0, by = 0;
      for (int t = 0; t < e->ntow; ++t) { fl += 2.0 * B * dg[t].N * dg[t].K; by += 4.0 * ((double)B * dg[t].K + (double)dg[t].N * dg[t].K + 2.0 * B * dg[t].N); }
      if (l == e->depth - 1 && e->fuse_heads && dg[0].K <= HEADS_MAXN && dg[e->ntow - 1].K <= HEADS_MAXN && dg[0].N == dg[e->ntow - 1].N) {
        HeadGradJobs jobs{};
        for (int t = 0; t < e->ntow; ++t) jobs.j[jobs.n++] = HeadGradJob{dg[t].D, dg[t].W, dg[t].Y, dg[t].dX, B, dg[t].K, dg[t].N, dg[t].act};
        Scope sc(e, "heads_dgrad", fl, by);
        heads_dgrad_kernel<<<dim3((unsigned)(((long long)B * dg[0].N + 255) / 256), e->ntow), 256, 0, e->ls>>>(jobs);
        CK(cudaGetLastError());
      } else if (!tc_dense_dgrad2(e, nm, dg, e->ntow, fl, by)) launch_igemm(e, nm, dg[0], dg[1], e->ntow, false, fl, by);
    } else if (trunk) {
      // gradient into the trunk: both towers in one contraction (K = N_val + N_adv) - no second launch, no read-modify-write
      const Mat& w0 = e->tow[0][0]; const Mat& w1 = e->tow[e->ntow - 1][0];
      DenseDgradOp op{};
      op.D = e->tow_delta[0][0]; op.ldd = w0.N; op.W = e->theta + w0.off; op.dX = dfeat; op.ldx = w0.K;
      op.Y = e->trunk_on; op.ldy = w0.K; op.act = e->trunk_act; op.accumulate = 0; op.apply_act = e->lstm ? 0 : 1;
      op.M = B; op.N = w0.K; op.K = w0.N; op.K1 = 0;
      if (e->ntow == 2) { op.K1 = w0.N; op.K = w0.N + w1.N; op.D2 = e->tow_delta[1][0]; op.ldd2 = w1.N; op.W2 = e->theta + w1.off; }
      op.vecA = (w0.N % 4 == 0) && (w1.N % 4 == 0); op.vecB = op.vecA;
      if (e->arena) {
        op.Ds = e->tow_delta[0][0]; op.Ws = e->theta + w0.off; op.a_single = 0;
        if (e->ntow == 2) { op.Ds2 = e->tow_delta[1][0]; op.Ws2 = e->theta + w1.off; }
      }
      snprintf(nm, sizeof nm, "dense1_dgrad");
      const double fl = 2.0 * B * op.N * op.K, by = 4.0 * ((double)B * op.K + (double)op.N * op.K + 2.0 * B * op.N);
      if (e->ntow == 1 || w0.N % 4 == 0) {
        if (!tc_dense_dgrad(e, nm, op, fl, by)) launch_igemm(e, nm, op, op, 1, false, fl, by);
      } else {                                 // tower widths not 16-byte granular: one launch per tower, accumulating in a fixed order
        for (int t = 0; t < e->ntow; ++t) {
          const Mat& w = e->tow[t][0];
          DenseDgradOp o2{};
          o2.D = e->tow_delta[t][0]; o2.ldd = w.N; o2.W = e->theta + w.off; o2.dX = dfeat; o2.ldx = w.K;
          o2.Y = e->trunk_on; o2.ldy = w.K; o2.act = e->trunk_act; o2.accumulate = t > 0; o2.apply_act = (t == e->ntow - 1) && !e->lstm;
          o2.M = B; o2.N = w.K; o2.K = w.N; o2.vecA = o2.vecB = 0;
          snprintf(nm, sizeof nm, "dense1_dgrad_t%d", t);
          launch_igemm(e, nm, o2, o2, 1, false, 2.0 * B * o2.N * o2.K, 4.0 * ((double)B * o2.K + (double)o2.N * o2.K + 2.0 * B * o2.N));
        }
      }
    }
  }
  e->towers_updated = false;
  if (conc && trunk && !e->lstm) {
    // Every Dense gradient is enqueued (weight gradients on lane 2) and so is the last reader of the Dense weights (the gradient into the
    // trunk, on the main lane): reduce that bucket and run its Adam update now, on the third lane, behind the conv backward.  The Dense
    // layers hold 98 % of the parameters (12.9 of 13.2 MB in config 3), so almost all of the optimizer's HBM traffic leaves the critical path.
    order_after(e, e->stream3, e->stream2);
    if (e->cfg.world > 1) {
      cudaStream_t keep = e->ls; e->ls = e->stream3;
      {
        Scope sc(e, "allreduce_dense", 0, 2.0 * (e->nint - e->tower_off) * 4);
        allreduce_grad(e, e->tower_off, e->nint - e->tower_off, 0, e->stream3);
      }
      e->ls = keep;
      e->sm_reserve = e->ar_ctas;                            // the conv backward below runs beside this reduction: leave its CTAs room
    }
    order_after(e, e->stream3, e->stream);
    enqueue_adam(e, e->tower_off, e->nint, e->stream3);
    e->towers_updated = true;
  }
  for (int l = (int)e->convs.size() - 1; l >= 0; --l) {
    const ConvL& c = e->convs[l];
    ConvWgradOp wg{};
    wg.X = l == 0 ? (const void*)e->xb : (const void*)e->on.conv_out[l - 1]; wg.x_u8 = l == 0 ? (e->elem_bytes == 1) : 0;
    wg.D = e->conv_delta[l]; wg.dW = e->grad + c.w.off; wg.nimg = B; wg.g = c.g;
    wg.M = c.w.K + 1; wg.N = c.g.Cout; wg.K = B * c.g.OH * c.g.OW; wg.vecA = (c.g.Cin % 4 == 0); wg.vecB = (c.g.Cout % 4 == 0);
    if (e->arena) {
      wg.a_single = (l == 0 && e->elem_bytes == 1);
      wg.Xs = l == 0 ? (wg.a_single ? e->xb_f : (const float*)e->xb) : e->on.conv_out[l - 1]; wg.Ds = e->conv_delta[l]; wg.ones = e->ones;
    }
    snprintf(nm, sizeof nm, "conv%d_wgrad", l + 1);
    double fl = 2.0 * wg.M * wg.N * wg.K;
    double by = (double)B * c.g.IH * c.g.IW * c.g.Cin * (wg.x_u8 ? 1 : 4) + 4.0 * wg.K * wg.N + 4.0 * wg.M * wg.N;
    ConvWgradOp wg_tc = wg;                    // the tensor-core operand holds raw byte values: fold the 1/255 into its epilogue only
    wg_tc.out_scale = wg.a_single ? 1.0f / 255.0f : 0.f;
    // a bias row that would open a 128-row tile of its own (257 = 2 x 128 + 1 rows in the first layer) is summed by colsum_kernel instead
    // ... and whenever the TMA feed takes the launch (the ones row of [x 1] is not an im2col box)
    const bool tma_wgrad = e->arena && e->tc_tma && e->tc_tma_wgrad && e->cfg.math_mode == DQN_MATH_3XTF32 && l > 0 && (c.g.Cin % 32 == 0) && (c.g.Cout % 32 == 0) && c.g.Cout <= 1024;
    const bool split_bias = tma_wgrad || (e->arena && (c.w.K % 128 == 0) && c.w.K <= 256 && (c.g.Cout % 4 == 0) && c.g.Cout <= 1024);   // worth a launch only when it saves >= 1/3 of the tiles
    if (split_bias) { wg_tc.M = c.w.K; wg_tc.no_bias = 1; }
    if (wg.a_single && e->a8) { wg_tc.a8 = 1; wg_tc.Xs = nullptr; }     // first layer: the byte batch itself is the operand
    bool bias_by_colsum = false;
    {
      if (conc) order_after(e, e->stream2, e->stream);
      Lane lane(e, conc);
      if (!tc_conv_wgrad(e, nm, wg_tc, fl, by)) launch_igemm(e, nm, wg, wg, 1, true, fl, by);
      else bias_by_colsum = split_bias;
      // the first layer has no input gradient: its column sums go to the (idle) main lane and run beside the weight-gradient contraction
      // instead of behind it and its reduction on the tail of the step
      if (bias_by_colsum && !(conc && l == 0)) {
        snprintf(nm, sizeof nm, "conv%d_bgrad", l + 1);
        Scope sc(e, nm, 0, 4.0 * wg.K * wg.N);
        colsum_kernel<<<COLSUM_CTAS, 256, 0, e->ls>>>(e->conv_delta[l], (long long)wg.K, wg.N, e->grad + c.w.off + (long long)c.w.K * c.g.Cout,
                                                     e->colsum_part + (e->ls == e->stream ? 0 : COLSUM_CTAS * 1024), e->colsum_ticket + (e->ls == e->stream ? 0 : 1));
        CK(cudaGetLastError());
        bias_by_colsum = false;
      }
    }
    if (bias_by_colsum) {
      snprintf(nm, sizeof nm, "conv%d_bgrad", l + 1);
      Scope sc(e, nm, 0, 4.0 * wg.K * wg.N);
      colsum_kernel<<<COLSUM_CTAS, 256, 0, e->stream>>>(e->conv_delta[l], (long long)wg.K, wg.N, e->grad + c.w.off + (long long)c.w.K * c.g.Cout,
                                                       e->colsum_part, e->colsum_ticket);
      CK(cudaGetLastError());
    }
    if (l > 0) {
      ConvDgradOp dg{};
      dg.D = e->conv_delta[l]; dg.W = e->theta + c.w.off; dg.dX = e->conv_delta[l - 1]; dg.Yprev = e->on.conv_out[l - 1];
      dg.act = e->convs[l - 1].w.act; dg.apply_act = 1; dg.nimg = B; dg.g = c.g;
      dg.vecA = (c.g.Cout % 4 == 0); dg.vecB = (c.g.Cout % 4 == 0);
      if (e->arena) { dg.Ds = e->conv_delta[l]; dg.Ws = e->theta + c.w.off; dg.a_single = 0; }
      snprintf(nm, sizeof nm, "conv%d_dgrad", l + 1);
      fl = 2.0 * B * c.g.OH * c.g.OW * c.g.Cout * c.w.K;
      by = 4.0 * ((double)B * c.g.OH * c.g.OW * c.g.Cout + (double)c.w.K * c.g.Cout + 2.0 * B * c.g.IH * c.g.IW * c.g.Cin);
      // strided layers: the S*S parity classes share their source pixels - one contraction over a rearranged copy of the weights
      if (e->arena && e->dgrad_merge && e->cfg.math_mode == DQN_MATH_3XTF32 && ConvDgradMergedOp::geometry_ok(c.g) && e->wm[l]) {
        ConvDgradMergedOp mg{};                                  // (its weight copy wm[l] was rebuilt on the target lane, prepare_dgrad_weights)
        mg.D = e->conv_delta[l]; mg.Wm = e->wm[l]; mg.dX = e->conv_delta[l - 1]; mg.Yprev = e->on.conv_out[l - 1]; mg.act = e->convs[l - 1].w.act; mg.apply_act = 1;
        mg.nimg = B; mg.g = c.g; mg.vecA = mg.vecB = 1; mg.Ds = mg.D; mg.Ws = mg.Wm; mg.a_single = 0;
        mg.init();
        if (tc_conv_dgrad_merged(e, nm, mg, fl, by)) continue;
      }
      if (!tc_conv_dgrad(e, nm, dg, fl, by)) launch_igemm(e, nm, dg, dg, c.g.S * c.g.S, false, fl, by);
    }
  }
}

// Adam over the parameter range [lo, hi) (both multiples of 4 floats), plus its share of max|g|
void enqueue_adam(E* e, long long lo, long long hi, cudaStream_t s) {
  if (hi <= lo) return;
  const long long n4 = (hi - lo) / 4;
  const int grid = (int)std::min<long long>(4LL * e->nsm, (n4 + 255) / 256);
  cudaStream_t keep = e->ls; e->ls = s;                       // profiling events go to the lane the kernel runs on
  {
    Scope sc(e, "adam", 0, 7.0 * (hi - lo) * 4);
    adam_kernel<<<grid, 256, 0, s>>>(e->theta + lo, e->adam_m + lo, e->adam_v + lo, e->grad + lo, n4,
                                     (double)e->cfg.learning_rate, e->cfg.adam_beta1, e->cfg.adam_beta2, e->cfg.adam_eps, 1.0f, e->st,
                                     e->w_on_s, 0, e->w_scale_lo - lo, e->w_scale_hi - lo, 1.0f / 255.0f);
    CK(cudaGetLastError());
  }
  e->ls = keep;
}

void enqueue_gather(E* e, cudaStream_t gs = nullptr) {       // observation rows of the sampled transitions -> batch (and its tensor-core operand planes)
  if (!gs) gs = e->stream;
  cudaStream_t keep_ls = e->ls; e->ls = gs;
  const long long rb = e->obs_row_bytes;
  const long long per = (rb % 16 == 0) ? 256LL * 4 * 16 : 256LL * 4;
  dim3 grid((unsigned)((rb + per - 1) / per), 2 * e->B);
  {
    Scope sc(e, "gather_rows", 0, 4.0 * e->B * rb);
    gather_rows_kernel<<<grid, 256, 0, gs>>>(e->store_s, e->store_sp, e->idx_d, e->B, rb, e->xb, (rb % 16 == 0 && e->elem_bytes == 1) ? e->xb_f : nullptr, 0, e->elem_bytes == 1);
    CK(cudaGetLastError());
  }
  e->ls = keep_ls;
}
void enqueue_batch_prep(E* e, cudaStream_t gs = nullptr) {   // get_batch (PER:89-104) for indices given by the caller
  {
    Scope sc(e, "batch_meta", 0, e->B * 40.0);
    batch_meta_kernel<<<(e->B + 255) / 256, 256, 0, e->stream>>>(e->idx_d, e->B, e->tree, e->P, e->act, e->rew, e->done, e->st,
                                                                e->cfg.beta, e->a_b, e->r_b, e->d_b, e->w_b);
    CK(cudaGetLastError());
  }
  if (gs) order_after(e, gs, e->stream);
  enqueue_gather(e, gs);
}
size_t sample_smem(int B) { int HT = 1; while (HT < 4 * B) HT <<= 1; return 2 * HT * sizeof(int) + TREE_TOP * sizeof(float); }

// ---- recurrent batch_train! (SOLVER:239-287) ---------------------------------------------------------------------------------------
// x W_i^T + b for all time steps of a pass at once: rows x [in] -> rows x [4H]
void lstm_xproj(E* e, const float* P, const float* X, int rows, float* out, const char* nm) {
  DenseFwdOp op{};
  op.X = X; op.ldx = e->lstm_in; op.x_u8 = 0; op.W = P + e->wi.off; op.C = out; op.ldc = 4 * e->Hh; op.act = DQN_ACT_IDENTITY;
  op.M = rows; op.N = 4 * e->Hh; op.K = e->lstm_in; op.vecA = (e->lstm_in % 4 == 0) && al16(X); op.vecB = 1;
  if (e->arena) { op.Xs = X; op.Ws = op.W; op.a_single = 0; }
  const double fl = 2.0 * rows * op.N * op.K, by = 4.0 * ((double)rows * op.K + (double)(op.K + 1) * op.N + (double)rows * op.N);
  if (!tc_dense_fwd(e, nm, &op, 1, fl, by)) launch_igemm(e, nm, op, op, 1, false, fl, by);
}
void lstm_step_launch(E* e, const LstmFwdArgs& a, int nch, const char* nm) {
  Scope sc(e, nm, 2.0 * nch * a.B * 4.0 * a.H * a.H, 4.0 * nch * a.B * 10.0 * a.H);
  dim3 grid((a.H + 31) / 32, (a.B + LSTM_TB - 1) / LSTM_TB, nch), block(32, LSTM_TB);
  lstm_fwd_step_kernel<<<grid, block, LSTM_TB * a.H * sizeof(float), e->ls>>>(a);
  CK(cudaGetLastError());
}
void lstm_broadcast(E* e, float* dst, const float* src, int rows) {
  const long long n = (long long)rows * e->Hh;
  lstm_broadcast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->ls>>>(dst, src, rows, e->Hh);
  CK(cudaGetLastError());
}

void enqueue_step_recurrent(E* e, bool sample) {
  const int TB = e->B, T = e->T, Bep = e->Bep, H = e->Hh, N4 = 4 * e->Hh;
  const long long d = e->obs_elems, blk = (long long)Bep * H;
  float* xs = reinterpret_cast<float*>(e->xb);                 // rows 0..TB-1: s (time-major), TB..2TB-1: s'
  if (sample) {
    Scope sc(e, "episode_sample", 0, Bep * 8.0 * 22);
    sample_kernel<<<1, (Bep + 31) / 32 * 32, sample_smem(Bep), e->stream>>>(e->tree, e->P, Bep, e->cfg.seed, e->st, 0, 0, e->idx_d, nullptr, nullptr, nullptr, 0.f,
                                                                         nullptr, nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
  }
  {
    Scope sc(e, "episode_gather", 0, 2.0 * TB * d * 4 * 2);
    episode_gather_kernel<<<dim3(T, Bep), 128, 0, e->stream>>>(e->idx_d, Bep, T, e->Lmax, d, e->cfg.seed, e->st, 0, 0, e->ep_len, e->ep_s, e->ep_sp, e->ep_a, e->ep_r,
                                                               e->ep_done, xs, xs + (long long)TB * d, e->a_b, e->r_b, e->d_b, e->w_b, e->ep_start_d);
    CK(cudaGetLastError());
  }
  const bool conc = e->use_streams && !e->profiling;
  const bool seq = e->lstm_seq && lstm_seq_supported(H);      // the whole recurrence in one cluster launch per pass (lstm.cuh)
  e->ev_next = 0;
  // ---- target network over s' (lane 2) beside the online network over s and s' (two chains per launch)
  if (conc) order_after(e, e->stream2, e->stream);
  {
    Lane lane(e, conc);
    lstm_xproj(e, e->theta_t, xs + (long long)TB * d, TB, e->xproj_tg, "lstm_xproj_target");
    lstm_broadcast(e, e->hs_tg, e->theta_t + e->h0_off, Bep);
    lstm_broadcast(e, e->cs_tg + 2 * blk, e->theta_t + e->c0_off, Bep);
    if (seq) {
      LstmSeqFwdArgs a{};
      a.xproj[0] = e->xproj_tg; a.h0[0] = e->theta_t + e->h0_off; a.c0[0] = e->theta_t + e->c0_off; a.hs[0] = e->hs_tg + blk; a.cs[0] = nullptr; a.gates[0] = nullptr;
      a.Wh = e->theta_t + e->wh.off; a.B = Bep; a.H = H; a.T = T;
      Scope sc(e, "lstm_seq_target", 2.0 * TB * 4.0 * H * H, 4.0 * TB * 10.0 * H);
      CK(lstm_seq_fwd_launch(a, 1, e->ls));
    } else
    for (int t = 0; t < T; ++t) {
      LstmFwdArgs a{};
      a.xproj[0] = e->xproj_tg + (long long)t * Bep * N4; a.h_prev[0] = e->hs_tg + t * blk; a.h_out[0] = e->hs_tg + (t + 1) * blk;
      a.c_prev[0] = t == 0 ? e->cs_tg + 2 * blk : e->cs_tg + ((t - 1) & 1) * blk; a.c_out[0] = e->cs_tg + (t & 1) * blk; a.gates[0] = nullptr;
      a.Wh = e->theta_t + e->wh.off; a.B = Bep; a.H = H;
      lstm_step_launch(e, a, 1, "lstm_step_target");
    }
    const Pass p_tg{e->theta_t, nullptr, 0, TB, &e->tg, "target", nullptr, nullptr, e->arena != nullptr, e->trunk_tg};
    forward(e, &p_tg, 1);
  }
  lstm_xproj(e, e->theta, xs, 2 * TB, e->xproj_on, "lstm_xproj_online");
  lstm_broadcast(e, e->hs_on, e->theta + e->h0_off, Bep);
  lstm_broadcast(e, e->cs_on, e->theta + e->c0_off, Bep);
  if (seq) {
    // chain 0: the s pass (states and gates kept for BPTT); chain 1: the s' pass
    LstmSeqFwdArgs a{};
    a.xproj[0] = e->xproj_on; a.xproj[1] = e->xproj_on + (long long)TB * N4;
    a.h0[0] = a.h0[1] = e->theta + e->h0_off; a.c0[0] = a.c0[1] = e->theta + e->c0_off;
    a.hs[0] = e->hs_on + blk; a.hs[1] = e->hs_on + (T + 1) * blk; a.cs[0] = e->cs_on + blk; a.cs[1] = nullptr; a.gates[0] = e->gates_s; a.gates[1] = nullptr;
    a.Wh = e->theta + e->wh.off; a.B = Bep; a.H = H; a.T = T;
    Scope sc(e, "lstm_seq_online", 2.0 * 2 * TB * 4.0 * H * H, 4.0 * 2 * TB * 10.0 * H);
    CK(lstm_seq_fwd_launch(a, 2, e->ls));
  } else
  for (int t = 0; t < T; ++t) {
    LstmFwdArgs a{};
    // chain 0: the s pass (states kept for BPTT); chain 1: the s' pass.  Block 0 of hs_on / cs_on is state0, block t+1 the s pass after step t,
    // block T+1+t the s' pass after step t
    a.xproj[0] = e->xproj_on + (long long)t * Bep * N4; a.h_prev[0] = e->hs_on + t * blk; a.h_out[0] = e->hs_on + (t + 1) * blk;
    a.c_prev[0] = e->cs_on + t * blk; a.c_out[0] = e->cs_on + (t + 1) * blk; a.gates[0] = e->gates_s + (long long)t * Bep * N4;
    a.xproj[1] = e->xproj_on + ((long long)TB + (long long)t * Bep) * N4; a.h_prev[1] = t == 0 ? e->hs_on : e->hs_on + (T + t) * blk; a.h_out[1] = e->hs_on + (T + 1 + t) * blk;
    a.c_prev[1] = t == 0 ? e->cs_on : e->cs_sp + ((t - 1) & 1) * blk; a.c_out[1] = e->cs_sp + (t & 1) * blk; a.gates[1] = nullptr;
    a.Wh = e->theta + e->wh.off; a.B = Bep; a.H = H;
    lstm_step_launch(e, a, 2, "lstm_step_online");
  }
  {
    const Pass p_on{e->theta, nullptr, 0, 2 * TB, &e->on, "online", nullptr, nullptr, e->arena != nullptr, e->trunk_on};
    forward(e, &p_on, 1);
  }
  if (conc) order_after(e, e->stream, e->stream2);
  {
    HeadArgs h{};
    const int L = e->depth - 1;
    if (e->cfg.dueling) { h.V_on = e->on.tow_out[0][L]; h.A_on = e->on.tow_out[1][L]; h.V_tg = e->tg.tow_out[0][L]; h.A_tg = e->tg.tow_out[1][L];
                          h.dV = e->tow_delta[0][L]; h.dA = e->tow_delta[1][L]; h.act_v = e->tow[0][L].act; h.act_a = e->tow[1][L].act; }
    else { h.A_on = e->on.tow_out[0][L]; h.A_tg = e->tg.tow_out[0][L]; h.dA = e->tow_delta[0][L]; h.act_a = e->tow[0][L].act; }
    h.a_b = e->a_b; h.r_b = e->r_b; h.d_b = e->d_b; h.w_b = e->w_b;      // w = the trace mask: loss = (1/T) sum_t sum_i huber(m td) / B  (SOLVER:279-281)
    h.q_s = e->q_s; h.q_sp_on = e->q_sp_on; h.q_sp_tg = e->q_sp_tg; h.y = e->y; h.best_a = e->best_a; h.td = e->td; h.newp = e->newp;
    h.B = TB; h.nA = e->cfg.n_actions; h.dueling = e->cfg.dueling; h.double_q = e->cfg.double_q;
    h.gamma = e->cfg.discount; h.alpha = e->cfg.alpha; h.eps = e->cfg.eps;
    h.inv_world_B = 1.0f / ((float)TB * (float)e->cfg.world);
    h.st = e->st;
    h.part = e->hub; h.ticket = e->colsum_ticket + 3;
    Scope sc(e, "head_loss", 0, TB * (double)(3 * (e->cfg.n_actions + 1) + 12) * 4);
    head_loss_kernel<<<(TB + 127) / 128, 128, 0, e->stream>>>(h);
    CK(cudaGetLastError());
  }
  backward(e, conc);                                           // the towers: weight gradients + the gradient into the hidden states (trunk_delta)
  // ---- BPTT through the cell
  if (seq) {
    LstmSeqBwdArgs a{};
    a.dh_out = e->trunk_delta; a.Wh = e->theta + e->wh.off; a.gates = e->gates_s; a.cs = e->cs_on + blk; a.c0 = e->theta + e->c0_off; a.dgates = e->dgates;
    a.B = Bep; a.H = H; a.T = T;
    Scope sc(e, "lstm_seq_bptt", 2.0 * TB * 4.0 * H * H, 4.0 * TB * 14.0 * H);
    CK(lstm_seq_bwd_launch(a, e->stream));
  } else {
  {
    Scope sc(e, "lstm_transpose", 0, 8.0 * H * N4);
    transpose_kernel<<<dim3((N4 + 31) / 32, (H + 31) / 32), dim3(32, 8), 0, e->stream>>>(e->theta + e->wh.off, e->whT, H, N4);
    CK(cudaGetLastError());
  }
  for (int t = T - 1; t >= 0; --t) {
    LstmBwdArgs a{};
    a.dh_out = e->trunk_delta + t * blk; a.dg_next = t == T - 1 ? nullptr : e->dgates + (long long)(t + 1) * Bep * N4; a.WhT = e->whT;
    a.gates = e->gates_s + (long long)t * Bep * N4; a.c_prev = e->cs_on + t * blk; a.c_cur = e->cs_on + (t + 1) * blk;
    a.dc = e->dcell; a.dgates = e->dgates + (long long)t * Bep * N4; a.B = Bep; a.H = H; a.first = t == T - 1;
    Scope sc(e, "lstm_bptt_step", 2.0 * Bep * 4.0 * H * H, 4.0 * Bep * 14.0 * H);
    lstm_bwd_step_kernel<<<dim3((H + 31) / 32, (Bep + LSTM_TB - 1) / LSTM_TB), dim3(32, LSTM_TB), LSTM_TB * N4 * sizeof(float), e->stream>>>(a);
    CK(cudaGetLastError());
  }
  }
  if (conc) order_after(e, e->stream2, e->stream);
  {
    // dW_h = h_{t-1}^T dgates over all (t, b) rows; [dW_i; db] = [x 1]^T dgates: the same weight-gradient contraction the Dense layers use
    Lane lane(e, conc);
    DenseWgradOp wh{};
    wh.X = e->hs_on; wh.ldx = H; wh.x_u8 = 0; wh.D = e->dgates; wh.ldd = N4; wh.dW = e->grad + e->wh.off; wh.M = H; wh.N = N4; wh.K = TB; wh.no_bias = 1;
    wh.vecA = (H % 4 == 0); wh.vecB = 1;
    if (e->arena) { wh.Xs = e->hs_on; wh.Ds = e->dgates; wh.ones = e->ones; wh.a_single = 0; wh.out_scale = 0.f; }
    double fl = 2.0 * wh.M * wh.N * TB, by = 4.0 * ((double)TB * wh.M + (double)TB * wh.N + (double)wh.M * wh.N);
    if (!tc_dense_wgrad(e, "lstm_wh_wgrad", &wh, 1, fl, by)) launch_igemm(e, "lstm_wh_wgrad", wh, wh, 1, true, fl, by);
    DenseWgradOp wi{};
    wi.X = xs; wi.ldx = e->lstm_in; wi.x_u8 = 0; wi.D = e->dgates; wi.ldd = N4; wi.dW = e->grad + e->wi.off; wi.M = e->lstm_in + 1; wi.N = N4; wi.K = TB;
    wi.vecA = (e->lstm_in % 4 == 0) && al16(xs); wi.vecB = 1;
    if (e->arena) { wi.Xs = xs; wi.Ds = e->dgates; wi.ones = e->ones; wi.a_single = 0; wi.out_scale = 0.f; }
    fl = 2.0 * wi.M * wi.N * TB; by = 4.0 * ((double)TB * wi.M + (double)TB * wi.N + (double)wi.M * wi.N);
    if (!tc_dense_wgrad(e, "lstm_wi_wgrad", &wi, 1, fl, by)) launch_igemm(e, "lstm_wi_wgrad", wi, wi, 1, true, fl, by);
  }
  if (conc) order_after(e, e->stream, e->stream2);
  if (e->cfg.world > 1) {
    Scope sc(e, "grad_allreduce", 0, 2.0 * e->nint * 4);
    allreduce_grad(e, 0, e->nint, 0, e->stream);
  }
  enqueue_adam(e, 0, e->nint, e->stream);                      // state0 (h0, c0) has a zero gradient: Adam leaves it where it is
  {
    Scope sc(e, "end_of_step", 0, 12);
    tree_update_kernel<<<1, 32, 0, e->stream>>>(e->tree, e->P, e->idx_d, e->newp, 0, 0, e->st, 1, e->cfg.adam_beta1, e->cfg.adam_beta2, sample ? 1 : 0, e->host_out_dev, 1);
    CK(cudaGetLastError());
  }
}

void enqueue_step(E* e, bool sample) {
  if (e->lstm) { enqueue_step_recurrent(e, sample); return; }
  const int B = e->B;
  // first conv layer straight from the replay store (conv1_tc.cuh, direct mode): the gather leaves the critical path for the third lane
  const bool direct = e->c1_direct && e->c1_maps_ok && e->use_streams && !e->profiling && !e->merge_fwd && !e->convs.empty() && tc_conv1_eligible(e, e->convs[0].g);
  e->ev_next = 0;
  if (sample) {
    {
      Scope sc(e, "sumtree_sample", 0, B * 8.0 * 22);
      sample_kernel<<<1, (B + 31) / 32 * 32, sample_smem(B), e->stream>>>(e->tree, e->P, B, e->cfg.seed, e->st, 0, 0, e->idx_d,
                                                                            e->act, e->rew, e->done, e->cfg.beta, e->a_b, e->r_b, e->d_b, e->w_b);
      CK(cudaGetLastError());
    }
    if (direct) order_after(e, e->stream3, e->stream);
    enqueue_gather(e, direct ? e->stream3 : nullptr);
  } else enqueue_batch_prep(e, direct ? e->stream3 : nullptr);
  const float* xs = (e->arena && e->obs_row_bytes % 16 == 0) ? (e->elem_bytes == 1 ? e->xb_f : (const float*)e->xb) : nullptr;
  const bool conc = e->use_streams && !e->profiling;          // profiling wants clean per-kernel times: one lane
  const bool tcp = e->arena && e->obs_row_bytes % 16 == 0 && (xs != nullptr || e->a8);
  // thin output layers (N <= 8) behind at least one hidden Dense layer: the whole head is one launch
  bool head_fused = e->fuse_heads && e->fuse_head_all && e->depth >= 2 && e->cfg.n_actions <= HEADS_MAXN && B <= 65536;
  for (int t = 0; t < e->ntow; ++t) head_fused = head_fused && e->tow[t][e->depth - 1].N <= HEADS_MAXN && e->tow[t][e->depth - 1].K == e->tow[0][e->depth - 1].K;
  Pass p_on{e->theta, e->xb, e->elem_bytes == 1, 2 * B, &e->on, "online", xs, e->w_on_s, tcp};
  Pass p_tg{e->theta_t, e->xb + (long long)B * e->obs_row_bytes, e->elem_bytes == 1, B, &e->tg, "target", xs ? xs + (long long)B * e->obs_elems : nullptr, e->w_tg_s, tcp};
  // fuse_head_all: 0 = separate kernels (default); 2 = output layers on their passes' own lanes (heads_fwd_kernel), then loss + gradient
  // into the last hidden layer in one launch; 1 = the output layers in that launch as well (joins the three passes early)
  const bool head_fwd_in_pass = head_fused && e->fuse_head_all == 2;
  p_on.skip_last = p_tg.skip_last = head_fused && !head_fwd_in_pass;
  e->c1_direct_now = direct;
  if (tcp && e->cfg.math_mode == DQN_MATH_3XTF32 && e->merge_fwd) {   // tensor-core path: both networks layer by layer in shared launches
    const Pass both[2] = {p_on, p_tg};
    prepare_dgrad_weights(e);
    forward(e, both, 2);
  } else {
    if (conc) order_after(e, e->stream2, e->stream);          // fork: the gathered batch is ready
    {
      Lane lane(e, conc);
      prepare_dgrad_weights(e);
      forward(e, &p_tg, 1);
    }
    forward(e, &p_on, 1);
    if (conc) order_after(e, e->stream, e->stream2);          // join before the head needs Q_target(s')
  }
  e->c1_direct_now = false;
  if (direct) order_after(e, e->stream, e->stream3);          // the gathered batch (first layer's weight-gradient operand) is complete
  {
    HeadArgs h{};
    const int L = e->depth - 1;
    if (e->cfg.dueling) { h.V_on = e->on.tow_out[0][L]; h.A_on = e->on.tow_out[1][L]; h.V_tg = e->tg.tow_out[0][L]; h.A_tg = e->tg.tow_out[1][L];
                          h.dV = e->tow_delta[0][L]; h.dA = e->tow_delta[1][L]; h.act_v = e->tow[0][L].act; h.act_a = e->tow[1][L].act; }
    else { h.A_on = e->on.tow_out[0][L]; h.A_tg = e->tg.tow_out[0][L]; h.dA = e->tow_delta[0][L]; h.act_a = e->tow[0][L].act; }
    h.a_b = e->a_b; h.r_b = e->r_b; h.d_b = e->d_b; h.w_b = e->w_b;
    h.q_s = e->q_s; h.q_sp_on = e->q_sp_on; h.q_sp_tg = e->q_sp_tg; h.y = e->y; h.best_a = e->best_a; h.td = e->td; h.newp = e->newp;
    h.B = B; h.nA = e->cfg.n_actions; h.dueling = e->cfg.dueling; h.double_q = e->cfg.double_q;
    h.gamma = e->cfg.discount; h.alpha = e->cfg.alpha; h.eps = e->cfg.eps;
    h.inv_world_B = 1.0f / ((float)B * (float)e->cfg.world);
    h.st = e->st;
    if (head_fused) {
      FusedHeadArgs fa{}; fa.h = h; fa.ntow = e->ntow; fa.K = e->tow[0][L].K; fa.hub = e->hub; fa.ticket = e->colsum_ticket + 2; fa.fwd_done = head_fwd_in_pass ? 1 : 0;
      for (int t = 0; t < e->ntow; ++t) {
        const Mat& w = e->tow[t][L];
        fa.H_on[t] = e->on.tow_out[t][L - 1]; fa.H_tg[t] = e->tg.tow_out[t][L - 1]; fa.W_on[t] = e->theta + w.off; fa.W_tg[t] = e->theta_t + w.off;
        fa.out_on[t] = e->on.tow_out[t][L]; fa.out_tg[t] = e->tg.tow_out[t][L]; fa.dH[t] = e->tow_delta[t][L - 1];
        fa.N[t] = w.N; fa.act_out[t] = w.act; fa.act_hidden[t] = e->tow[t][L - 1].act;
      }
      Scope sc(e, head_fwd_in_pass ? "head_loss_dgrad" : "head_fused", 6.0 * B * fa.K * (e->cfg.n_actions + 1), 4.0 * B * fa.K * 5);
      head_fused_kernel<<<(B + 7) / 8, 256, 0, e->stream>>>(fa);
      CK(cudaGetLastError());
    } else {
      Scope sc(e, "head_loss", 0, B * (double)(3 * (e->cfg.n_actions + 1) + 12) * 4);
      if (e->head_small && B <= 256 && e->cfg.n_actions <= 8) head_loss_small_kernel<8><<<1, (B + 31) / 32 * 32, 0, e->stream>>>(h);
      else if (e->head_small && B <= 256 && e->cfg.n_actions <= 32) head_loss_small_kernel<32><<<1, (B + 31) / 32 * 32, 0, e->stream>>>(h);
      else head_loss_kernel<<<1, std::min(1024, (B + 31) / 32 * 32), 0, e->stream>>>(h);
      CK(cudaGetLastError());
    }
  }
  // update_priorities! (PER:76-80) needs nothing but the TD errors: on the third lane now, beside the whole reverse pass, instead of
  // 13 us alone at the end of the step (one CTA, twenty dependent tree levels)
  const bool early_tree = conc;
  if (early_tree) {
    order_after(e, e->stream3, e->stream);
    cudaStream_t keep = e->ls; e->ls = e->stream3;
    {
      Scope sc(e, "sumtree_update", 0, B * 12.0 * 21);
      // (publishes the tree epoch: from here on the step touches neither the replay rows nor the tree - new transitions may be ingested
      //  beside its reverse pass, dqn_replay_add)
      tree_update_kernel<<<1, std::min(1024, (B + 31) / 32 * 32), 0, e->stream3>>>(e->tree, e->P, e->idx_d, e->newp, e->cfg.prioritized_replay ? B : 0,
                                                                                  1, e->st, 0, e->cfg.adam_beta1, e->cfg.adam_beta2, 0, nullptr, 1);
      CK(cudaGetLastError());
    }
    e->ls = keep;
  }
  backward(e, conc, head_fused);
  e->sm_reserve = 0;
  if (conc) order_after(e, e->stream, e->stream2);            // all weight gradients are in
  if (e->cfg.world > 1) {
    const bool split = conc && !e->convs.empty();          // the Dense bucket is already in flight on the NCCL lane
    Scope sc(e, "grad_allreduce", 0, 2.0 * e->nint * 4);
    const long long n = split ? e->tower_off : e->nint;
    cudaStream_t cs = split ? e->stream3 : e->stream;
    if (split) order_after(e, e->stream3, e->stream);
    cudaStream_t keep = e->ls; e->ls = cs;
    allreduce_grad(e, 0, n, split ? 1 : 0, cs);
    e->ls = keep;
    if (split) order_after(e, e->stream, e->stream3);
  }
  if (e->towers_updated) {
    order_after(e, e->stream, e->stream3);                  // the Dense bucket's update (and its max|g|) is complete
    enqueue_adam(e, 0, e->tower_off, e->stream);
  } else enqueue_adam(e, 0, e->nint, e->stream);
  if (early_tree) {
    if (!e->towers_updated) order_after(e, e->stream, e->stream3);
    Scope sc(e, "end_of_step", 0, 12);                        // beta powers, sampler counter, (loss, grad_norm, flags) -> host words
    tree_update_kernel<<<1, 32, 0, e->stream>>>(e->tree, e->P, e->idx_d, e->newp, 0, 0, e->st, 1, e->cfg.adam_beta1, e->cfg.adam_beta2, sample ? 1 : 0, e->host_out_dev);
    CK(cudaGetLastError());
  } else {
    Scope sc(e, "sumtree_update", 0, B * 12.0 * 21);
    tree_update_kernel<<<1, std::min(1024, (B + 31) / 32 * 32), 0, e->stream>>>(e->tree, e->P, e->idx_d, e->newp, e->cfg.prioritized_replay ? B : 0,
                                                                               1, e->st, 1, e->cfg.adam_beta1, e->cfg.adam_beta2, sample ? 1 : 0, e->host_out_dev, 1);
    CK(cudaGetLastError());
  }
}

void run_step(E* e, bool sample) {
  const int need = e->lstm ? e->Bep : e->B;
  if (e->curr_size < need) fail(DQN_ERR_STATE, "replay holds %lld %s, batch_size is %d (PER:83 / episode_replay.jl:73 @assert r._curr_size >= r.batch_size)", e->curr_size, e->lstm ? "episodes" : "transitions", need);
  cudaGraphExec_t& gx = sample ? e->graph_sample : e->graph_idx;
  if (e->cfg.use_graph && !e->profiling) {
    if (!gx) {
      cudaGraph_t g = nullptr;
      e->capturing = true;
      e->counting = 1; e->launches = 0;
      CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
      try { enqueue_step(e, sample); } catch (...) { cudaStreamEndCapture(e->stream, &g); if (g) cudaGraphDestroy(g); e->capturing = false; e->counting = 0; throw; }
      CK(cudaStreamEndCapture(e->stream, &g));
      e->capturing = false; e->counting = 0; e->launches_per_step = e->launches;
      CK(cudaGraphInstantiate(&gx, g, 0));
      CK(cudaGraphDestroy(g));
    }
    CK(cudaGraphLaunch(gx, e->stream));
  } else {
    e->counting = 1; e->launches = 0;
    enqueue_step(e, sample);
    e->counting = 0; e->launches_per_step = e->launches;
  }
  e->n_launched += 1;
  CK(cudaEventRecord(e->step_ev[e->n_launched & 1], e->stream));
  e->main_dirty = e->lstm;      // the ingest lane is ordered behind this step through tree_free (feed-forward steps record it)
}

// Scalars of the step launched `back` steps before the latest one (0: the latest; 1: the one before it, which can be read while the
// latest is still running - the two host slots alternate).
void fetch_scalars(E* e, float* loss, float* gn, int back = 0) {
  if (back < 0 || back > 1) fail(DQN_ERR_INVALID, "back must be 0 or 1");
  if (e->n_launched < (unsigned long long)back + 1) fail(DQN_ERR_STATE, "no such step: %llu launched so far", e->n_launched);
  const unsigned long long s = e->n_launched - back;
  if (back == 0) CK(cudaStreamSynchronize(e->stream)); else CK(cudaEventSynchronize(e->step_ev[s & 1]));
  volatile float* slot = e->host_out + 4 * (s & 1);
  const int err = reinterpret_cast<volatile int*>(slot)[2];
  if (loss) *loss = slot[0];
  if (gn) *gn = slot[1];
  if (err) {
    CK(cudaStreamSynchronize(e->stream));
    int zero = 0;
    CK(cudaMemcpy(&e->st->error, &zero, sizeof(int), cudaMemcpyHostToDevice));
    reinterpret_cast<int*>(e->host_out)[2] = 0; reinterpret_cast<int*>(e->host_out)[6] = 0;
    if (err & 1) fail(DQN_ERR_STATE, "sum-tree sampler did not reach %d distinct indices", e->B);
    if (err & 2) fail(DQN_ERR_STATE, "non-positive priority (PER:78 @assert all(new_priorities .> 0f0))");
    if (err & 4) fail(DQN_ERR_STATE, "td_err + eps <= 0 (PER:66 @assert)");
    if (err & 8) fail(DQN_ERR_INVALID, "action index outside 1..n_actions");
    if (err & 32) fail(DQN_ERR_STATE, "ingest lane: the step in flight never published its tree epoch");
    if (err & 16) fail(DQN_ERR_NCCL, "gradient all-reduce over peer memory: a rank did not arrive within the spin limit");
  }
}

void check_dev_errors(E* e) {
  int err = 0;
  CK(cudaMemcpyAsync(&err, &e->st->error, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (err) {
    int zero = 0;
    CK(cudaMemcpy(&e->st->error, &zero, sizeof(int), cudaMemcpyHostToDevice));
    if (err & 2) fail(DQN_ERR_STATE, "non-positive priority (PER:78 @assert all(new_priorities .> 0f0))");
    if (err & 4) fail(DQN_ERR_STATE, "td_err + eps <= 0 (PER:66 @assert)");
    if (err & 8) fail(DQN_ERR_INVALID, "action index outside 1..n_actions");
    fail(DQN_ERR_STATE, "device error flags %d", err);
  }
}

void ensure_stage(E* e, long long bytes) {
  if (bytes <= e->stage_bytes) return;
  if (e->stage) CK(cudaFree(e->stage));
  e->stage = nullptr; e->stage_bytes = 0;
  CK(cudaMalloc(&e->stage, bytes));
  e->stage_bytes = bytes;
}

void rebuild_tree(E* e) {
  for (long long lo = e->P / 2; lo >= 1; lo /= 2) {
    tree_level_kernel<<<(unsigned)((lo + 255) / 256), 256, 0, e->stream>>>(e->tree, lo);
    CK(cudaGetLastError());
  }
}

void set_curr_size(E* e) {
  CK(cudaMemcpyAsync(&e->st->curr_size, &e->curr_size, sizeof(long long), cudaMemcpyHostToDevice, e->stream));
}

// n transitions already on the device (Flux layout) -> ring
void ingest_device(E* e, const uint8_t* s, const int* a, const float* r, const uint8_t* sp, const uint8_t* d, const float* td0, long long n, long long* slots,
                   cudaStream_t st = nullptr) {
  if (!st || n > 4096) st = e->stream;                       // (a bulk add rebuilds the whole tree: main stream only)
  const int C = e->hwc ? e->cfg.obs_c : 1;
  const int HW = e->hwc ? e->cfg.obs_h * e->cfg.obs_w : (int)e->obs_elems;
  if (n > e->cap) fail(DQN_ERR_INVALID, "adding %lld transitions to a buffer of %lld", n, e->cap);
  const long long new_size = std::min(e->cap, e->curr_size + n);
  for (long long t0 = 0; t0 < n; t0 += 32768) {
    const long long cnt = std::min<long long>(32768, n - t0);
    dim3 grid((unsigned)std::min<long long>((e->obs_elems + 255) / 256, 64), (unsigned)cnt);
    ingest_kernel<<<grid, 256, 0, st>>>(s, sp, a, r, d, td0, t0, e->cursor, e->cap, C, HW, e->elem_bytes, e->store_s, e->store_sp,
                                               e->act, e->rew, e->done, e->tree, e->P, slots, e->cfg.alpha, e->cfg.eps, e->cfg.n_actions, new_size, e->st);
    CK(cudaGetLastError());
  }
  if (n <= 4096) {
    tree_update_kernel<<<1, (int)std::min<long long>(1024, (n + 31) / 32 * 32), 0, st>>>(e->tree, e->P, slots, nullptr, (int)n, 0, e->st, 0, 1.0, 1.0, 0, nullptr);
    CK(cudaGetLastError());
  } else rebuild_tree(e);
  e->cursor = (e->cursor + n) % e->cap;
  e->curr_size = new_size;
}

void relayout(E* e, const void* in, void* out, long long rows, int in_u8, int out_f32_from_u8, int dir) {
  const int C = e->hwc ? e->cfg.obs_c : 1;
  const int HW = e->hwc ? e->cfg.obs_h * e->cfg.obs_w : (int)e->obs_elems;
  const long long total = rows * e->obs_elems;
  relayout_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 8LL * e->nsm), 256, 0, e->stream>>>((const uint8_t*)in, (uint8_t*)out, rows, C, HW, in_u8, out_f32_from_u8, dir);
  CK(cudaGetLastError());
}

// ---- topology -----------------------------------------------------------------------------------
void build_topology(E* e) {
  const dqn_config_t& c = e->cfg;
  if (c.n_layers < 1 || c.n_layers > DQN_MAX_LAYERS) fail(DQN_ERR_INVALID, "n_layers must be in 1..%d", DQN_MAX_LAYERS);
  int i = 0, H = c.obs_h, W = c.obs_w, C = c.obs_c;
  std::vector<dqn_layer_t> dense;
  while (i < c.n_layers && c.layers[i].kind == DQN_LAYER_CONV) {
    const dqn_layer_t& l = c.layers[i];
    if (l.in != C) fail(DQN_ERR_INVALID, "conv layer %d expects %d input channels, gets %d", i + 1, l.in, C);
    if (l.kh < 1 || l.kw < 1 || l.stride < 1 || l.kh > H || l.kw > W) fail(DQN_ERR_INVALID, "conv layer %d geometry", i + 1);
    ConvL cl{};
    cl.g.IH = H; cl.g.IW = W; cl.g.Cin = C; cl.g.OH = (H - l.kh) / l.stride + 1; cl.g.OW = (W - l.kw) / l.stride + 1;
    cl.g.Cout = l.out; cl.g.KH = l.kh; cl.g.KW = l.kw; cl.g.S = l.stride;
    if (l.stride > 4) fail(DQN_ERR_UNSUPPORTED, "conv stride above 4");
    cl.g.init();
    cl.w = Mat{0, l.kh * l.kw * C, l.out, l.act};
    e->convs.push_back(cl);
    H = cl.g.OH; W = cl.g.OW; C = l.out; ++i;
  }
  e->hwc = !e->convs.empty();
  e->feat = H * W * C;
  for (; i < c.n_layers; ++i) {
    const dqn_layer_t& l = c.layers[i];
    if (l.kind == DQN_LAYER_FLATTEN) { if (!dense.empty()) fail(DQN_ERR_UNSUPPORTED, "flattenbatch after a Dense layer"); continue; }
    if (l.kind == DQN_LAYER_LSTM) {
      // Chain([flattenbatch,] LSTM(in, out), Dense...) - the recurrent models of the reference's tests (test/runtests.jl:116,133,150)
      if (e->lstm || !dense.empty() || !e->convs.empty()) fail(DQN_ERR_UNSUPPORTED, "layer %d: one LSTM, directly on the (flattened) observation, is supported", i + 1);
      if (l.in != e->feat) fail(DQN_ERR_INVALID, "LSTM expects %d inputs, gets %d", l.in, e->feat);
      if (l.out < 1 || l.out % 4 != 0) fail(DQN_ERR_UNSUPPORTED, "LSTM width must be a multiple of 4");
      if (c.obs_dtype != DQN_OBS_F32) fail(DQN_ERR_UNSUPPORTED, "recurrent engines store Float32 observations");
      e->lstm = 1; e->lstm_in = l.in; e->Hh = l.out; e->feat = l.out;
      continue;
    }
    if (l.kind != DQN_LAYER_DENSE) fail(DQN_ERR_UNSUPPORTED, "layer %d: only Conv* [flattenbatch] [LSTM] Dense+ chains are supported", i + 1);
    dense.push_back(l);
  }
  if (dense.empty()) fail(DQN_ERR_INVALID, "DeepQLearningError: the qnetwork provided is incompatible with dueling (no trailing Dense layer, DUEL:47-50)");
  int in = e->feat;
  for (auto& l : dense) { if (l.in != in) fail(DQN_ERR_INVALID, "Dense layer expects %d inputs, gets %d", l.in, in); in = l.out; }
  if (dense.back().out != c.n_actions) fail(DQN_ERR_INVALID, "last Dense has %d outputs, n_actions is %d", dense.back().out, c.n_actions);
  if (c.n_actions < 1 || c.n_actions > HEAD_MAX_ACTIONS) fail(DQN_ERR_UNSUPPORTED, "n_actions must be in 1..%d", HEAD_MAX_ACTIONS);
  e->depth = (int)dense.size();
  e->ntow = c.dueling ? 2 : 1;
  // internal flat layout: convs, then towers (val, adv), every matrix start aligned to 4 floats
  long long off = 0;
  auto place = [&](Mat& m) { m.off = off; off += ((long long)(m.K + 1) * m.N + 3) / 4 * 4; };
  for (auto& cl : e->convs) place(cl.w);
  if (e->lstm) {
    // LSTMCell: W_i (4H, in) and W_h (4H, H) column-major = row-major [in][4H], [H][4H]; b rides as the bias row of W_i; state0 = (h0, c0)
    e->wi = Mat{0, e->lstm_in, 4 * e->Hh, DQN_ACT_IDENTITY}; place(e->wi);
    e->wh = Mat{0, e->Hh, 4 * e->Hh, DQN_ACT_IDENTITY}; place(e->wh);        // its bias row stays zero
    e->h0_off = off; off += (e->Hh + 3) / 4 * 4;
    e->c0_off = off; off += (e->Hh + 3) / 4 * 4;
  }
  e->tower_off = off;
  for (int t = 0; t < e->ntow; ++t)
    for (int l = 0; l < e->depth; ++l) {
      Mat m{0, dense[l].in, dense[l].out, dense[l].act};
      if (c.dueling && t == 0 && l == e->depth - 1) { m.N = 1; m.act = DQN_ACT_IDENTITY; }   // fresh Dense(k, 1), DUEL:53-54
      place(m);
      e->tow[t][l] = m;
    }
  e->nint = off;
  // Flux.params order -> internal index (DUEL:2-6,13: base, val, adv; weight then bias)
  e->perm.clear();
  for (auto& cl : e->convs) {
    const ConvGeom& g = cl.g;
    for (int co = 0; co < g.Cout; ++co) for (int ci = 0; ci < g.Cin; ++ci) for (int kh = 0; kh < g.KH; ++kh) for (int kw = 0; kw < g.KW; ++kw) {
      const int j = g.KH - 1 - kh, ii = g.KW - 1 - kw;      // true convolution -> cross-correlation taps
      e->perm.push_back(cl.w.off + ((long long)(j * g.KW + ii) * g.Cin + ci) * g.Cout + co);
    }
    for (int co = 0; co < g.Cout; ++co) e->perm.push_back(cl.w.off + (long long)cl.w.K * g.Cout + co);
  }
  if (e->lstm) {                               // Flux.params(LSTM): Wi, Wh, b, state0[1], state0[2]
    const int N4 = 4 * e->Hh;
    for (int k = 0; k < e->lstm_in; ++k) for (int n = 0; n < N4; ++n) e->perm.push_back(e->wi.off + (long long)k * N4 + n);
    for (int k = 0; k < e->Hh; ++k) for (int n = 0; n < N4; ++n) e->perm.push_back(e->wh.off + (long long)k * N4 + n);
    for (int n = 0; n < N4; ++n) e->perm.push_back(e->wi.off + (long long)e->lstm_in * N4 + n);
    for (int j = 0; j < e->Hh; ++j) e->perm.push_back(e->h0_off + j);
    for (int j = 0; j < e->Hh; ++j) e->perm.push_back(e->c0_off + j);
  }
  const int HWt = H * W;
  for (int t = 0; t < e->ntow; ++t)
    for (int l = 0; l < e->depth; ++l) {
      const Mat& m = e->tow[t][l];
      for (int ki = 0; ki < m.K; ++ki) {
        int row = ki;
        if (l == 0 && e->hwc) { const int ch = ki / HWt, hw = ki % HWt; row = hw * C + ch; }   // Flux flatten (w + W h + W H c) -> HWC
        for (int o = 0; o < m.N; ++o) e->perm.push_back(m.off + (long long)row * m.N + o);
      }
      for (int o = 0; o < m.N; ++o) e->perm.push_back(m.off + (long long)m.K * m.N + o);
    }
  e->nflux = (long long)e->perm.size();
}

void allocate(E* e) {
  const dqn_config_t& c = e->cfg;
  if (e->lstm) {                               // the network sees trace_length * batch_size rows per pass
    e->T = c.trace_length > 0 ? c.trace_length : 40;            // SOLVER:15
    e->Lmax = c.max_episode_length > 0 ? c.max_episode_length : 100;   // SOLVER:21
    e->Bep = c.batch_size;
    if (e->T > e->Lmax) e->Lmax = e->T;
    if ((long long)e->T * e->Bep > 65536) fail(DQN_ERR_UNSUPPORTED, "trace_length * batch_size above 65536");
  }
  const int B = e->B = e->lstm ? e->T * c.batch_size : c.batch_size;
  e->rows_on = std::max(2 * B, c.max_act_rows > 0 ? c.max_act_rows : 2 * B);
  e->elem_bytes = c.obs_dtype == DQN_OBS_U8 ? 1 : 4;
  e->obs_elems = (long long)c.obs_c * c.obs_h * c.obs_w;
  e->obs_row_bytes = e->obs_elems * e->elem_bytes;
  e->theta = dalloc<float>(e->nint); e->theta_t = dalloc<float>(e->nint);
  e->adam_m = dalloc<float>(e->nint); e->adam_v = dalloc<float>(e->nint); e->grad = dalloc<float>(e->nint);
  e->cap = c.buffer_size;
  int P = 2; while (P < e->cap) P <<= 1;
  e->P = P;
  if (e->lstm) {
    // EpisodeReplayBuffer: cap episodes of up to Lmax steps; the transition stores of the feed-forward path stay empty
    const long long steps = e->cap * e->Lmax;
    e->ep_s = dalloc<float>(steps * e->obs_elems); e->ep_sp = dalloc<float>(steps * e->obs_elems);
    e->ep_a = dalloc<int>(steps); e->ep_r = dalloc<float>(steps); e->ep_done = dalloc<uint8_t>(steps); e->ep_len = dalloc<int>(e->cap);
    e->ep_start_d = dalloc<int>(e->Bep);
    const long long TB = B, H = e->Hh;
    e->xproj_on = dalloc<float>(2 * TB * 4 * H); e->xproj_tg = dalloc<float>(TB * 4 * H);
    e->hs_on = dalloc<float>((2LL * e->T + 1) * e->Bep * H); e->hs_tg = dalloc<float>((e->T + 1LL) * e->Bep * H);
    e->cs_on = dalloc<float>((e->T + 1LL) * e->Bep * H); e->cs_sp = dalloc<float>(2LL * e->Bep * H); e->cs_tg = dalloc<float>(2LL * e->Bep * H);
    e->gates_s = dalloc<float>(TB * 4 * H); e->dgates = dalloc<float>(TB * 4 * H); e->dcell = dalloc<float>((long long)e->Bep * H);
    e->whT = dalloc<float>(4 * H * H);
    e->act_rows = e->rows_on;
    e->h_act = dalloc<float>((long long)e->act_rows * H); e->c_act = dalloc<float>((long long)e->act_rows * H);
    e->trunk_on = e->hs_on + (long long)e->Bep * H; e->trunk_tg = e->hs_tg + (long long)e->Bep * H;
    e->trunk_delta = dalloc<float>(TB * H); e->trunk_act = DQN_ACT_IDENTITY;
  }
  e->store_s = dalloc<uint8_t>(e->lstm ? 1 : e->cap * e->obs_row_bytes);
  e->store_sp = dalloc<uint8_t>(e->lstm ? 1 : e->cap * e->obs_row_bytes);
  e->act = dalloc<int>(e->cap); e->rew = dalloc<float>(e->cap); e->done = dalloc<uint8_t>(e->cap);
  e->tree = dalloc<float>(2LL * P);
  e->st = dalloc<DevState>(1);
  DevState init{}; init.b1p = c.adam_beta1; init.b2p = c.adam_beta2;
  CK(cudaMemcpy(e->st, &init, sizeof init, cudaMemcpyHostToDevice));
  e->idx_d = dalloc<long long>(B);                            // (recurrent: the first Bep entries are the sampled episodes)
  e->xb = dalloc<uint8_t>((long long)e->rows_on * e->obs_row_bytes);
  e->a_b = dalloc<int>(B); e->r_b = dalloc<float>(B); e->d_b = dalloc<float>(B); e->w_b = dalloc<float>(B);
  for (auto& cl : e->convs) {
    const long long per = (long long)cl.g.OH * cl.g.OW * cl.g.Cout;
    e->on.conv_out.push_back(dalloc<float>(e->rows_on * per));
    e->tg.conv_out.push_back(dalloc<float>(B * per));
    e->conv_delta.push_back(dalloc<float>(B * per));
  }
  if (!e->convs.empty()) { e->trunk_on = e->on.conv_out.back(); e->trunk_tg = e->tg.conv_out.back(); e->trunk_delta = e->conv_delta.back(); e->trunk_act = e->convs.back().w.act; }
  for (int t = 0; t < e->ntow; ++t)
    for (int l = 0; l < e->depth; ++l) {
      e->on.tow_out[t][l] = dalloc<float>((long long)e->rows_on * e->tow[t][l].N);
      e->tg.tow_out[t][l] = dalloc<float>((long long)B * e->tow[t][l].N);
      e->tow_delta[t][l] = dalloc<float>((long long)B * e->tow[t][l].N);
    }
  const int nA = c.n_actions;
  e->q_s = dalloc<float>((long long)B * nA); e->q_sp_on = dalloc<float>((long long)B * nA); e->q_sp_tg = dalloc<float>((long long)B * nA);
  e->y = dalloc<float>(B); e->td = dalloc<float>(B); e->newp = dalloc<float>(B); e->best_a = dalloc<int>(B);
  e->colsum_part = dalloc<float>(2LL * COLSUM_CTAS * 1024);
  e->colsum_ticket = dalloc<unsigned int>(4);                 // [0], [1]: column sums per lane; [2]: head_fused_kernel
  e->hub = dalloc<float>(B);
  e->ws_floats = 16LL << 20;      // 64 MB split-K workspace
  e->ws = dalloc<float>(e->ws_floats);
  e->ws2 = dalloc<float>(e->ws_floats);
  e->lws = e->ws;
  CK(cudaHostAlloc(&e->host_out, 32, cudaHostAllocMapped));
  memset(e->host_out, 0, 32);
  CK(cudaHostGetDevicePointer(&e->host_out_dev, e->host_out, 0));
  CK(cudaHostAlloc(&e->idx_h, sizeof(long long) * B, cudaHostAllocDefault));
  CK(cudaEventCreate(&e->t0)); CK(cudaEventCreate(&e->t1)); CK(cudaEventCreateWithFlags(&e->copy_done, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->ingest_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->ingest_done, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->main_mark, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&e->step_ev[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->add_free[i], cudaEventDisableTiming));
  }
}

void destroy(E* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->graph_sample) cudaGraphExecDestroy(e->graph_sample);
  if (e->graph_idx) cudaGraphExecDestroy(e->graph_idx);
  for (int i = 0; i < e->peer_nopened; ++i) if (e->peer_opened[i]) cudaIpcCloseMemHandle(e->peer_opened[i]);
  if (e->peer_flags) cudaFree(e->peer_flags);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  tc_destroy(e);
  void* ptrs[] = {e->theta, e->theta_t, e->adam_m, e->adam_v, e->grad, e->store_s, e->store_sp, e->done, e->act, e->rew, e->tree, e->st,
                  e->idx_d, e->xb, e->a_b, e->r_b, e->d_b, e->w_b, e->q_s, e->q_sp_on, e->q_sp_tg, e->y, e->td, e->newp, e->best_a, e->ws, e->ws2,
                  e->stage, e->flush_buf, e->colsum_part, e->colsum_ticket, e->hub, e->ep_s, e->ep_sp, e->ep_r, e->ep_a, e->ep_len, e->ep_done, e->ep_start_d,
                  e->xproj_on, e->xproj_tg, e->hs_on, e->hs_tg, e->cs_on, e->cs_sp, e->cs_tg, e->gates_s, e->dgates, e->dcell, e->whT, e->h_act, e->c_act,
                  e->lstm ? e->trunk_delta : nullptr};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto p : e->on.conv_out) cudaFree(p);
  for (auto p : e->tg.conv_out) cudaFree(p);
  for (auto p : e->conv_delta) cudaFree(p);
  for (int t = 0; t < 2; ++t) for (int l = 0; l < MAXD; ++l) { if (e->on.tow_out[t][l]) cudaFree(e->on.tow_out[t][l]); if (e->tg.tow_out[t][l]) cudaFree(e->tg.tow_out[t][l]); if (e->tow_delta[t][l]) cudaFree(e->tow_delta[t][l]); }
  if (e->host_out) cudaFreeHost(e->host_out);
  if (e->idx_h) cudaFreeHost(e->idx_h);
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (e->t0) cudaEventDestroy(e->t0);
  if (e->t1) cudaEventDestroy(e->t1);
  if (e->copy_done) cudaEventDestroy(e->copy_done);
  for (int i = 0; i < 2; ++i) {
    if (e->step_ev[i]) cudaEventDestroy(e->step_ev[i]);
    if (e->add_free[i]) cudaEventDestroy(e->add_free[i]);
    if (e->add_stage[i]) cudaFree(e->add_stage[i]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->ingest_stream) { cudaStreamSynchronize(e->ingest_stream); cudaStreamDestroy(e->ingest_stream); }
  for (cudaEvent_t ev : {e->ingest_done, e->main_mark}) if (ev) cudaEventDestroy(ev);
  for (auto ev : e->evs) cudaEventDestroy(ev);
  if (e->stream3) cudaStreamDestroy(e->stream3);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

// host <-> internal parameter images
void scatter_params(E* e, float* dev, const float* flat) {
  std::vector<float> h((size_t)e->nint, 0.f);
  for (long long i = 0; i < e->nflux; ++i) h[(size_t)e->perm[(size_t)i]] = flat[i];
  CK(cudaMemcpyAsync(dev, h.data(), sizeof(float) * e->nint, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
}
void gather_params(E* e, const float* dev, float* flat) {
  std::vector<float> h((size_t)e->nint);
  CK(cudaMemcpyAsync(h.data(), dev, sizeof(float) * e->nint, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (long long i = 0; i < e->nflux; ++i) flat[i] = h[(size_t)e->perm[(size_t)i]];
}

// ro: the call enqueues nothing on the main stream that touches the replay state (scalar fetches; dqn_replay_add, which has its own lane)
template <class F> int guard(E* e, F&& f, bool ro = false) {
  try {
    if (!e) return DQN_ERR_INVALID;
    CK(cudaSetDevice(e->cfg.device));
    if (!ro) {
      if (e->ingest_pending) { CK(cudaStreamWaitEvent(e->stream, e->ingest_done, 0)); e->ingest_pending = false; }
      e->main_dirty = true;
    }
    f();
    return DQN_OK;
  } catch (const Err& x) {
    if (e) e->err = x.msg;
    return x.code;
  } catch (const std::exception& x) {
    if (e) e->err = x.what();
    return DQN_ERR_INVALID;
  }
}

template <class T> void d2h(E* e, T* dst, const T* src, long long n) {
  CK(cudaMemcpyAsync(dst, src, sizeof(T) * n, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
}

}  // namespace

// tensor-core path (tc_gemm.cuh) needs the engine definition
#include "tc_gemm_impl.cuh"
#include "conv1_tc.cuh"

// =================================================================================================
extern "C" {

int dqn_config_default(dqn_config_t* c) {
  if (!c) return DQN_ERR_INVALID;
  memset(c, 0, sizeof *c);
  c->abi_version = DQN_ABI_VERSION;
  c->obs_h = c->obs_w = 1; c->obs_dtype = DQN_OBS_F32;
  c->dueling = c->double_q = c->prioritized_replay = 1;          // SOLVER:10,11,16
  c->batch_size = 32; c->buffer_size = 1000;                     // SOLVER:5,20
  c->alpha = 0.6f; c->beta = 0.4f; c->eps = 1e-3f;               // PER:43-45 (the solver's own PER fields are dead)
  c->learning_rate = 1e-4f; c->discount = 1.0f;                  // SOLVER:3; default_discount HELPERS:83
  c->adam_beta1 = 0.9; c->adam_beta2 = 0.999; c->adam_eps = 1e-8;
  c->seed = 0; c->math_mode = DQN_MATH_FP32; c->use_graph = 1; c->rank = 0; c->world = 1;
  return DQN_OK;
}

const char* dqn_last_error(const dqn_engine_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dqn_device_count(int* n) { return cudaGetDeviceCount(n) == cudaSuccess ? DQN_OK : DQN_ERR_CUDA; }

int dqn_nccl_unique_id(uint8_t id_out[DQN_NCCL_ID_BYTES]) {
  std::string why;
  if (!g_nccl.load(why)) { g_create_error = why; return DQN_ERR_NCCL; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return DQN_ERR_NCCL; }
  static_assert(sizeof(ncclUniqueId) == DQN_NCCL_ID_BYTES, "ncclUniqueId size");
  memcpy(id_out, &id, DQN_NCCL_ID_BYTES);
  return DQN_OK;
}

// Peer-memory all-reduce set-up (peer_ar.cuh): every rank publishes its gradient vector and its flag block - as CUDA IPC handles between
// processes, as plain pointers between the engines of one process (dqn_group_create) - through one ncclAllGather on the communicator
// that exists anyway, and maps its peers'.  Any failure leaves NCCL in charge.
struct PeerInfo { long long pid; unsigned long long grad, flags; int dev, pad; cudaIpcMemHandle_t hgrad, hflags; };
void peer_setup(E* e) {
  const char* v = getenv("DQN_PEER_AR");
  const int W = e->cfg.world;
  if ((v && atoi(v) == 0) || W > PEER_MAX || !g_nccl.AllGather) return;
  { const char* c = getenv("DQN_PEER_CTAS"); if (c) e->peer_ctas = std::max(1, std::min(PEER_MAXG, atoi(c))); }
  e->peer_flags = dalloc<unsigned long long>(PEER_FLAGS);
  PeerInfo mine{};
  mine.pid = (long long)getpid(); mine.grad = (unsigned long long)e->grad; mine.flags = (unsigned long long)e->peer_flags; mine.dev = e->cfg.device;
  bool ok = cudaIpcGetMemHandle(&mine.hgrad, e->grad) == cudaSuccess && cudaIpcGetMemHandle(&mine.hflags, e->peer_flags) == cudaSuccess;
  cudaGetLastError();
  mine.pad = ok ? 1 : 0;
  PeerInfo* d_all = dalloc<PeerInfo>(W + 1);
  CK(cudaMemcpy(d_all + W, &mine, sizeof mine, cudaMemcpyHostToDevice));
  ncclResult_t r = g_nccl.AllGather(d_all + W, d_all, sizeof(PeerInfo), ncclChar, e->comm, e->stream);
  if (r != ncclSuccess) { cudaFree(d_all); return; }
  CK(cudaStreamSynchronize(e->stream));
  std::vector<PeerInfo> all(W);
  CK(cudaMemcpy(all.data(), d_all, sizeof(PeerInfo) * W, cudaMemcpyDeviceToHost));
  cudaFree(d_all);
  PeerArArgs a{};
  a.world = W; a.rank = e->cfg.rank; a.st = e->st;
  for (int p = 0; p < W && ok; ++p) {
    if (!all[p].pad) { ok = false; break; }
    if (p == e->cfg.rank) { a.grad[p] = e->grad; a.flags[p] = e->peer_flags; continue; }
    if (all[p].pid == mine.pid) {                              // same process: the pointers are valid here once peer access is on
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, e->cfg.device, all[p].dev) != cudaSuccess || !can) { ok = false; break; }
      cudaError_t pe = cudaDeviceEnablePeerAccess(all[p].dev, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { ok = false; break; }
      cudaGetLastError();
      a.grad[p] = reinterpret_cast<float*>(all[p].grad); a.flags[p] = reinterpret_cast<unsigned long long*>(all[p].flags);
    } else {
      void* g = nullptr; void* f = nullptr;
      if (cudaIpcOpenMemHandle(&g, all[p].hgrad, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      e->peer_opened[e->peer_nopened++] = g;
      if (cudaIpcOpenMemHandle(&f, all[p].hflags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      e->peer_opened[e->peer_nopened++] = f;
      a.grad[p] = reinterpret_cast<float*>(g); a.flags[p] = reinterpret_cast<unsigned long long*>(f);
    }
  }
  cudaGetLastError();
  // every rank must take the same path: agree through one more reduction (sum of the ok flags == world)
  float* d_ok = dalloc<float>(1);
  const float mine_ok = ok ? 1.f : 0.f;
  CK(cudaMemcpy(d_ok, &mine_ok, sizeof(float), cudaMemcpyHostToDevice));
  r = g_nccl.AllReduce(d_ok, d_ok, 1, ncclFloat, ncclSum, e->comm, e->stream);
  float tot = 0.f;
  if (r == ncclSuccess) { CK(cudaStreamSynchronize(e->stream)); CK(cudaMemcpy(&tot, d_ok, sizeof(float), cudaMemcpyDeviceToHost)); }
  cudaFree(d_ok);
  if ((int)(tot + 0.5f) == W) { e->peer = a; e->peer_ar = true; }
}

int dqn_engine_create(const dqn_config_t* cfg, dqn_engine_t** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return DQN_ERR_INVALID; }
  *out = nullptr;
  E* e = new E();
  try {
    if (cfg->abi_version != DQN_ABI_VERSION) fail(DQN_ERR_INVALID, "abi_version %d, library is %d", cfg->abi_version, DQN_ABI_VERSION);
    e->cfg = *cfg;
    if (cfg->batch_size < 1 || cfg->batch_size > 1024) fail(DQN_ERR_INVALID, "batch_size must be in 1..1024");
    if (cfg->buffer_size < cfg->batch_size) fail(DQN_ERR_STATE, "max_size < batch_size (PER:84 @assert)");
    if (cfg->buffer_size > (1LL << 30)) fail(DQN_ERR_UNSUPPORTED, "buffer_size above 2^30");
    if (cfg->obs_c < 1 || cfg->obs_h < 1 || cfg->obs_w < 1) fail(DQN_ERR_INVALID, "observation shape");
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) fail(DQN_ERR_INVALID, "rank/world");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 1) fail(DQN_ERR_CUDA, "no CUDA device: libdqn_b200 has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) fail(DQN_ERR_INVALID, "device %d of %d", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) fail(DQN_ERR_UNSUPPORTED, "compute capability %d.%d: this library is built for sm_100a only", prop.major, prop.minor);
    e->nsm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->stream3, cudaStreamNonBlocking));
    e->ls = e->stream;
    { const char* v = getenv("DQN_STREAMS"); e->use_streams = v ? atoi(v) : 1; }
    { const char* v = getenv("DQN_MERGE_FWD"); e->merge_fwd = v ? atoi(v) : 0; }
    { const char* v = getenv("DQN_FUSE_HEADS"); e->fuse_heads = v ? atoi(v) : 1; }
    { const char* v = getenv("DQN_LSTM_SEQ"); e->lstm_seq = v ? atoi(v) : 1; }
    { const char* v = getenv("DQN_HEAD_SMALL"); e->head_small = v ? atoi(v) : 1; }
    { const char* v = getenv("DQN_C1_DIRECT"); e->c1_direct = v ? atoi(v) : 1; }
    { const char* v = getenv("DQN_INGEST_LANE"); e->ingest_lane = v ? atoi(v) : 1; }
    // (measured, ms/step: 0 separate kernels 0.401; 2 loss + dgrad in one launch 0.413; 1 output layers as well 0.430 - a warp per sample
    //  is too little parallelism for the 256 x 1024 gradient that heads_dgrad_kernel spreads over 262 144 threads)
    { const char* v = getenv("DQN_FUSE_HEAD_ALL"); e->fuse_head_all = v ? atoi(v) : 0; }   // one launch for output layers + loss + dH: measured slower (0.452 vs 0.415 ms/step) - it joins the three passes early
    build_topology(e);
    allocate(e);
    tc_init(e);
    tc_conv1_init(e);
    if (cfg->world > 1) {
      std::string why;
      if (!g_nccl.load(why)) fail(DQN_ERR_NCCL, "%s", why.c_str());
      ncclUniqueId id; memcpy(&id, cfg->nccl_id, DQN_NCCL_ID_BYTES);
      // the collective's CTAs must find SMs beside the persistent tensor-core grids: bound NCCL's CTA count and keep that many SMs free
      // while a reduction is in flight (tc_launch_v); DQN_NCCL_CTAS overrides, an NCCL_MAX_CTAS already in the environment wins
      { const char* v = getenv("DQN_NCCL_CTAS"); e->nccl_ctas = v ? atoi(v) : 16; }
      if (e->nccl_ctas > 0) { char buf[16]; snprintf(buf, sizeof buf, "%d", e->nccl_ctas); setenv("NCCL_MAX_CTAS", buf, 0); }
      ncclResult_t r = g_nccl.CommInitRank(&e->comm, cfg->world, id, cfg->rank);
      if (r != ncclSuccess) fail(DQN_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
      // NCCL sets its channels up lazily at the first collectives: do that here, not inside the first timed steps
      for (int i = 0; i < 4; ++i) {
        cudaStream_t cs = (i & 1) ? e->stream3 : e->stream;
        r = g_nccl.AllReduce(e->grad, e->grad, (size_t)((i < 2) ? e->nint : std::max<long long>(e->tower_off, 4)), ncclFloat, ncclSum, e->comm, cs);
        if (r != ncclSuccess) fail(DQN_ERR_NCCL, "ncclAllReduce (warm-up): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        CK(cudaStreamSynchronize(cs));
      }
    }
    if (cfg->world > 1) peer_setup(e);
    CK(cudaStreamSynchronize(e->stream));
    *out = e;
    return DQN_OK;
  } catch (const Err& x) {
    g_create_error = x.msg;
    destroy(e);
    return x.code;
  }
}

void dqn_engine_destroy(dqn_engine_t* h) { destroy(h); }

int64_t dqn_num_params(const dqn_engine_t* h) { return h ? h->nflux : 0; }
int64_t dqn_tree_nodes(const dqn_engine_t* h) { return h ? 2LL * h->P : 0; }

int dqn_set_params(dqn_engine_t* h, int which, const float* flat, int64_t n) {
  return guard(h, [&] {
    if (n != h->nflux || !flat) fail(DQN_ERR_INVALID, "expected %lld parameters, got %lld", h->nflux, (long long)n);
    scatter_params(h, which == DQN_NET_TARGET ? h->theta_t : h->theta, flat);
    tc_params_changed(h);
  });
}
int dqn_get_params(dqn_engine_t* h, int which, float* flat, int64_t n) {
  return guard(h, [&] {
    if (n != h->nflux || !flat) fail(DQN_ERR_INVALID, "expected %lld parameters, got %lld", h->nflux, (long long)n);
    gather_params(h, which == DQN_NET_TARGET ? h->theta_t : h->theta, flat);
  });
}
int dqn_sync_target(dqn_engine_t* h) {
  return guard(h, [&] {
    copy_f4_kernel<<<2 * h->nsm, 256, 0, h->stream>>>((float4*)h->theta_t, (const float4*)h->theta, h->nint / 4);
    CK(cudaGetLastError());
    tc_params_changed(h);
  });
}
int dqn_get_adam_state(dqn_engine_t* h, float* m, float* v, double bp[2], int64_t n) {
  return guard(h, [&] {
    if (n != h->nflux) fail(DQN_ERR_INVALID, "expected %lld parameters", h->nflux);
    if (m) gather_params(h, h->adam_m, m);
    if (v) gather_params(h, h->adam_v, v);
    if (bp) { DevState s; d2h(h, &s, h->st, 1); bp[0] = s.b1p; bp[1] = s.b2p; }
  });
}

int dqn_replay_add(dqn_engine_t* h, const void* s, const int32_t* a, const float* r, const void* sp, const uint8_t* done,
                   const float* td0, int64_t n) {
  return guard(h, [&] {
    if (n < 0 || (n > 0 && (!s || !a || !r || !sp || !done || !td0))) fail(DQN_ERR_INVALID, "null argument");
    if (h->lstm) fail(DQN_ERR_UNSUPPORTED, "recurrent engine: transitions are added as whole episodes (dqn_episode_add)");
    // the reference's asserts (PER:66 td_err + eps > 0; a valid action index) are checked on the host arguments before anything is
    // enqueued: the call needs no device round trip, and a rejected batch leaves the buffer untouched.  (The ingest kernel raises the
    // same sticky flags for device-resident input, dqn_replay_add_device; they surface at the next dqn_train_step.)
    for (long long i = 0; i < n; ++i) {
      const float base = td0[i] + h->cfg.eps;
      if (!(base > 0.f)) fail(DQN_ERR_STATE, "td_err + eps <= 0 (PER:66 @assert)");
      if (a[i] < 1 || a[i] > h->cfg.n_actions) fail(DQN_ERR_INVALID, "action index outside 1..n_actions");
    }
    const long long rb = h->obs_row_bytes;
    const long long chunk = std::max<long long>(1, std::min<long long>(n, (256LL << 20) / std::max<long long>(1, 2 * rb)));
    const long long per = 2 * rb + 4 + 4 + 1 + 4;
    const long long need = chunk * per + 64 * 6 + chunk * 8;
    // The host -> device copies go to one of two staging areas on their own stream, so they run beside whatever step is still in
    // flight on the engine's stream; only the ingest kernels (and the sum-tree refresh) are ordered behind that step.  The call
    // returns when the copies are done - the caller's buffers are free again - with the ingest still in flight.
    // Small adds (the per-env-step case) are ingested on their own lane, ordered behind the priority update of the step in flight only;
    // a bulk add goes to the main stream like every other call.
    const bool lane = h->ingest_lane && chunk <= 4096 && n <= 4096;
    if (!lane) {
      if (h->ingest_pending) { CK(cudaStreamWaitEvent(h->stream, h->ingest_done, 0)); h->ingest_pending = false; }
      h->main_dirty = true;
    }
    cudaStream_t is = lane ? h->ingest_stream : h->stream;
    for (long long t0 = 0; t0 < n; t0 += chunk) {
      const long long c = std::min(chunk, n - t0);
      const int slot = (int)(h->add_seq++ & 1);
      if (h->add_stage_bytes[slot] < need) {
        if (h->add_stage[slot]) { CK(cudaStreamSynchronize(h->stream)); CK(cudaStreamSynchronize(h->ingest_stream)); CK(cudaFree(h->add_stage[slot])); h->add_stage[slot] = nullptr; h->add_stage_bytes[slot] = 0; }
        CK(cudaMalloc(&h->add_stage[slot], need)); h->add_stage_bytes[slot] = need;
      }
      uint8_t* p = h->add_stage[slot];
      auto carve = [&](long long bytes) { uint8_t* q = p; p += (bytes + 63) / 64 * 64; return q; };
      uint8_t* ds = carve(c * rb); uint8_t* dsp = carve(c * rb);
      int* da = (int*)carve(c * 4); float* dr = (float*)carve(c * 4); float* dtd = (float*)carve(c * 4); uint8_t* dd = carve(c);
      long long* slots = (long long*)carve(c * 8);
      CK(cudaStreamWaitEvent(h->copy_stream, h->add_free[slot], 0));      // the ingest that last read this area has finished
      CK(cudaMemcpyAsync(ds, (const uint8_t*)s + t0 * rb, c * rb, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaMemcpyAsync(dsp, (const uint8_t*)sp + t0 * rb, c * rb, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaMemcpyAsync(da, a + t0, c * 4, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaMemcpyAsync(dr, r + t0, c * 4, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaMemcpyAsync(dtd, td0 + t0, c * 4, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaMemcpyAsync(dd, done + t0, c, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaEventRecord(h->copy_done, h->copy_stream));
      if (lane) {
        if (h->main_dirty) {                                  // main-stream work of earlier calls (fills, priority updates, reads): all of it first
          CK(cudaEventRecord(h->main_mark, h->stream)); CK(cudaStreamWaitEvent(is, h->main_mark, 0)); h->main_dirty = false;
        }
        // the step in flight (if any) must have gathered its rows and refreshed its priorities: it says so in device memory (tree epoch)
        epoch_wait_kernel<<<1, 32, 0, is>>>(h->st, h->n_launched);
        CK(cudaGetLastError());
      }
      CK(cudaStreamWaitEvent(is, h->copy_done, 0));
      ingest_device(h, ds, da, dr, dsp, dd, dtd, c, slots, is);
      CK(cudaEventRecord(h->add_free[slot], is));
      if (lane) { CK(cudaEventRecord(h->ingest_done, is)); h->ingest_pending = true; }
    }
    if (n > 0) CK(cudaEventSynchronize(h->copy_done));
  }, /*ro=*/true);
}

int dqn_replay_add_device(dqn_engine_t* h, const void* s, const int32_t* a, const float* r, const void* sp, const uint8_t* done,
                          const float* td0, int64_t n) {
  return guard(h, [&] {
    if (n <= 0) return;
    if (h->lstm) fail(DQN_ERR_UNSUPPORTED, "recurrent engine: transitions are added as whole episodes (dqn_episode_add)");
    ensure_stage(h, n * 8);
    ingest_device(h, (const uint8_t*)s, a, r, (const uint8_t*)sp, done, td0, n, (long long*)h->stage);
    check_dev_errors(h);
  });
}

int dqn_replay_size(const dqn_engine_t* h, int64_t* curr_size, int64_t* cursor) {
  if (!h) return DQN_ERR_INVALID;
  if (curr_size) *curr_size = h->curr_size;
  if (cursor) *cursor = h->cursor;
  return DQN_OK;
}

int dqn_replay_fill_synthetic(dqn_engine_t* h, int64_t n, uint64_t seed) {
  return guard(h, [&] {
    if (n < 0 || n > h->cap) fail(DQN_ERR_INVALID, "n outside 0..buffer_size");
    const long long words = h->elem_bytes == 1 ? (h->obs_elems + 15) / 16 : (h->obs_elems + 3) / 4;
    for (long long i0 = 0; i0 < n; i0 += 32768) {
      const long long cnt = std::min<long long>(32768, n - i0);
      dim3 grid((unsigned)std::min<long long>((words + 255) / 256, 16), (unsigned)cnt);
      fill_synthetic_kernel<<<grid, 256, 0, h->stream>>>(h->store_s, h->store_sp, h->act, h->rew, h->done, h->tree, h->P, i0, h->obs_elems,
                                                         h->elem_bytes == 1, h->cfg.n_actions, h->cfg.alpha, h->cfg.eps, seed);
      CK(cudaGetLastError());
    }
    if (n < h->cap) { fill_u32_kernel<<<2 * h->nsm, 256, 0, h->stream>>>((uint32_t*)(h->tree + h->P + n), h->P - n, 0u); CK(cudaGetLastError()); }
    rebuild_tree(h);
    h->cursor = n % h->cap; h->curr_size = n;
    set_curr_size(h);
  });
}

int dqn_replay_read(dqn_engine_t* h, const int64_t* idx, int64_t n, void* s, int32_t* a, float* r, void* sp, uint8_t* done) {
  return guard(h, [&] {
    const long long rb = h->obs_row_bytes;
    for (long long t = 0; t < n; ++t) {
      const long long i = idx[t];
      if (i < 0 || i >= h->cap) fail(DQN_ERR_INVALID, "index %lld outside the buffer", i);
      ensure_stage(h, 2 * rb);
      if (s) { relayout(h, h->store_s + i * rb, h->stage, 1, h->elem_bytes == 1, 0, 0); CK(cudaMemcpyAsync((uint8_t*)s + t * rb, h->stage, rb, cudaMemcpyDeviceToHost, h->stream)); }
      if (sp) { relayout(h, h->store_sp + i * rb, h->stage + rb, 1, h->elem_bytes == 1, 0, 0); CK(cudaMemcpyAsync((uint8_t*)sp + t * rb, h->stage + rb, rb, cudaMemcpyDeviceToHost, h->stream)); }
      if (a) CK(cudaMemcpyAsync(a + t, h->act + i, 4, cudaMemcpyDeviceToHost, h->stream));
      if (r) CK(cudaMemcpyAsync(r + t, h->rew + i, 4, cudaMemcpyDeviceToHost, h->stream));
      if (done) CK(cudaMemcpyAsync(done + t, h->done + i, 1, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
  });
}

static void upload_idx(dqn_engine_t* h, const int64_t* idx, long long n, long long* dst, long long limit) {
  for (long long t = 0; t < n; ++t) if (idx[t] < 0 || idx[t] >= limit) fail(DQN_ERR_INVALID, "index %lld outside 0..%lld", (long long)idx[t], limit - 1);
  CK(cudaMemcpyAsync(dst, idx, sizeof(long long) * n, cudaMemcpyHostToDevice, h->stream));
}

int dqn_update_priorities(dqn_engine_t* h, const int64_t* idx, const float* td, int64_t n) {
  return guard(h, [&] {
    if (n <= 0) return;
    ensure_stage(h, n * 12 + 128);
    long long* di = (long long*)h->stage; float* dp = (float*)(h->stage + (n * 8 + 63) / 64 * 64);
    upload_idx(h, idx, n, di, h->cap);
    CK(cudaMemcpyAsync(dp, td, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    td_to_priority_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dp, n, h->cfg.alpha, h->cfg.eps);
    CK(cudaGetLastError());
    tree_update_kernel<<<1, 1024, 0, h->stream>>>(h->tree, h->P, di, dp, (int)n, 1, h->st, 0, 1.0, 1.0, 0, nullptr);
    CK(cudaGetLastError());
    check_dev_errors(h);
  });
}

int dqn_set_priorities(dqn_engine_t* h, const int64_t* idx, const float* prio, int64_t n) {
  return guard(h, [&] {
    if (n <= 0) return;
    ensure_stage(h, n * 12 + 128);
    long long* di = (long long*)h->stage; float* dp = (float*)(h->stage + (n * 8 + 63) / 64 * 64);
    upload_idx(h, idx, n, di, h->cap);
    CK(cudaMemcpyAsync(dp, prio, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    scatter_leaves_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->tree, h->P, di, dp, n);
    CK(cudaGetLastError());
    rebuild_tree(h);
    CK(cudaStreamSynchronize(h->stream));
  });
}

int dqn_get_priorities(dqn_engine_t* h, float* out, int64_t n) {
  return guard(h, [&] { if (n < 0 || n > h->cap) fail(DQN_ERR_INVALID, "n"); d2h(h, out, h->tree + h->P, n); });
}
int dqn_get_tree(dqn_engine_t* h, float* out, int64_t n_nodes) {
  return guard(h, [&] { if (n_nodes < 0 || n_nodes > 2LL * h->P) fail(DQN_ERR_INVALID, "n_nodes"); d2h(h, out, h->tree, n_nodes); });
}

int dqn_sample_indices(dqn_engine_t* h, uint64_t call, int64_t* idx_out) {
  return guard(h, [&] {
    if (h->curr_size < h->B) fail(DQN_ERR_STATE, "replay holds %lld transitions, batch_size is %d (PER:83)", h->curr_size, h->B);
    ensure_stage(h, h->B * 8);
    sample_kernel<<<1, (h->B + 31) / 32 * 32, sample_smem(h->B), h->stream>>>(h->tree, h->P, h->B, h->cfg.seed, h->st, 1, call, (long long*)h->stage,
                                                                               nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
    d2h(h, (long long*)idx_out, (const long long*)h->stage, h->B);
    check_dev_errors(h);
  });
}

int dqn_get_batch(dqn_engine_t* h, const int64_t* idx, float* s, int32_t* a, float* r, float* sp, float* done, float* weights) {
  return guard(h, [&] {
    const int B = h->B;
    upload_idx(h, idx, B, h->idx_d, h->curr_size > 0 ? h->curr_size : 1);
    enqueue_batch_prep(h);
    const long long fbytes = (long long)B * h->obs_elems * 4;
    ensure_stage(h, 2 * fbytes);
    relayout(h, h->xb, h->stage, 2LL * B, h->elem_bytes == 1, 1, 0);
    if (s) CK(cudaMemcpyAsync(s, h->stage, fbytes, cudaMemcpyDeviceToHost, h->stream));
    if (sp) CK(cudaMemcpyAsync(sp, h->stage + fbytes, fbytes, cudaMemcpyDeviceToHost, h->stream));
    if (a) CK(cudaMemcpyAsync(a, h->a_b, 4LL * B, cudaMemcpyDeviceToHost, h->stream));
    if (r) CK(cudaMemcpyAsync(r, h->r_b, 4LL * B, cudaMemcpyDeviceToHost, h->stream));
    if (done) CK(cudaMemcpyAsync(done, h->d_b, 4LL * B, cudaMemcpyDeviceToHost, h->stream));
    if (weights) CK(cudaMemcpyAsync(weights, h->w_b, 4LL * B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  });
}

int dqn_train_step(dqn_engine_t* h, float* loss, float* grad_norm) {
  return guard(h, [&] { run_step(h, true); fetch_scalars(h, loss, grad_norm); });
}
int dqn_train_step_with_indices(dqn_engine_t* h, const int64_t* idx, float* loss, float* grad_norm) {
  return guard(h, [&] {
    upload_idx(h, idx, h->lstm ? h->Bep : h->B, h->idx_d, h->curr_size > 0 ? h->curr_size : 1);   // recurrent: batch_size episode indices
    run_step(h, false);
    fetch_scalars(h, loss, grad_norm);
  });
}
int dqn_train_step_async(dqn_engine_t* h) { return guard(h, [&] { run_step(h, true); }); }
int dqn_sync(dqn_engine_t* h, float* loss, float* grad_norm) { return guard(h, [&] { fetch_scalars(h, loss, grad_norm); }, /*ro=*/true); }
int dqn_step_result(dqn_engine_t* h, int back, float* loss, float* grad_norm) { return guard(h, [&] { fetch_scalars(h, loss, grad_norm, back); }, /*ro=*/true); }

int dqn_q_values(dqn_engine_t* h, int which, const void* obs, int64_t n, float* q_out) {
  return guard(h, [&] {
    if (n < 0 || (n > 0 && (!obs || !q_out))) fail(DQN_ERR_INVALID, "null argument");
    const long long rb = h->obs_row_bytes; const int nA = h->cfg.n_actions; const int L = h->depth - 1;
    if (h->lstm) {
      // policy.qnetwork(obatch) with a Recur layer (POLICY:38-64): row i is lane i, its hidden state is carried from call to call until
      // dqn_policy_reset (resetstate!).  Online network only: the target network never acts.
      if (n > h->act_rows) fail(DQN_ERR_INVALID, "%lld rows, the acting state holds %d lanes (max_act_rows)", (long long)n, h->act_rows);
      if (which != DQN_NET_ONLINE) fail(DQN_ERR_UNSUPPORTED, "recurrent engine: acting uses the online network");
      if (n == 0) return;
      const int H = h->Hh;
      float* xs = reinterpret_cast<float*>(h->xb);
      CK(cudaMemcpyAsync(xs, obs, n * rb, cudaMemcpyHostToDevice, h->stream));
      lstm_xproj(h, h->theta, xs, (int)n, h->xproj_on, "lstm_xproj_act");
      LstmFwdArgs a{};
      a.xproj[0] = h->xproj_on; a.h_prev[0] = h->h_act; a.h_out[0] = h->hs_on; a.c_prev[0] = h->c_act; a.c_out[0] = h->c_act; a.gates[0] = nullptr;
      a.Wh = h->theta + h->wh.off; a.B = (int)n; a.H = H;
      lstm_step_launch(h, a, 1, "lstm_step_act");
      CK(cudaMemcpyAsync(h->h_act, h->hs_on, sizeof(float) * n * H, cudaMemcpyDeviceToDevice, h->stream));
      const Pass pa{h->theta, nullptr, 0, (int)n, &h->on, "act", nullptr, nullptr, false, h->hs_on};
      forward(h, &pa, 1);
      float* q = h->on.tow_out[h->ntow - 1][L];
      if (h->cfg.dueling) {
        q = (float*)h->ws;
        dueling_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->on.tow_out[0][L], h->on.tow_out[1][L], (int)n, nA, q);
        CK(cudaGetLastError());
      }
      CK(cudaMemcpyAsync(q_out, q, sizeof(float) * n * nA, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      return;
    }
    const int chunk = h->rows_on;
    ensure_stage(h, (long long)chunk * rb);
    for (long long t0 = 0; t0 < n; t0 += chunk) {
      const int c = (int)std::min<long long>(chunk, n - t0);
      CK(cudaMemcpyAsync(h->stage, (const uint8_t*)obs + t0 * rb, c * rb, cudaMemcpyHostToDevice, h->stream));
      relayout(h, h->stage, h->xb, c, h->elem_bytes == 1, 0, 1);
      const Pass pa{which == DQN_NET_TARGET ? h->theta_t : h->theta, h->xb, h->elem_bytes == 1, (int)c, &h->on, "act", nullptr, nullptr, false};
      forward(h, &pa, 1);
      float* q = h->on.tow_out[h->ntow - 1][L];
      if (h->cfg.dueling) {
        q = (float*)h->ws;
        dueling_combine_kernel<<<(c + 255) / 256, 256, 0, h->stream>>>(h->on.tow_out[0][L], h->on.tow_out[1][L], c, nA, q);
        CK(cudaGetLastError());
      }
      CK(cudaMemcpyAsync(q_out + t0 * nA, q, sizeof(float) * c * nA, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
  });
}

int dqn_get_last_indices(dqn_engine_t* h, int64_t* out) { return guard(h, [&] { d2h(h, (long long*)out, (const long long*)h->idx_d, h->B); }); }
int dqn_get_td(dqn_engine_t* h, float* out) { return guard(h, [&] { d2h(h, out, (const float*)h->td, h->B); }); }
int dqn_get_is_weights(dqn_engine_t* h, float* out) { return guard(h, [&] { d2h(h, out, (const float*)h->w_b, h->B); }); }
int dqn_get_q(dqn_engine_t* h, int which, float* out) {
  return guard(h, [&] {
    const float* src = which == DQN_Q_S_ONLINE ? h->q_s : which == DQN_Q_SP_ONLINE ? h->q_sp_on : h->q_sp_tg;
    d2h(h, out, src, (long long)h->B * h->cfg.n_actions);
  });
}
int dqn_get_targets(dqn_engine_t* h, float* y, int32_t* best_a) {
  return guard(h, [&] { if (y) d2h(h, y, (const float*)h->y, h->B); if (best_a) d2h(h, best_a, (const int*)h->best_a, h->B); });
}
int dqn_get_grads(dqn_engine_t* h, float* flat, int64_t n) {
  return guard(h, [&] { if (n != h->nflux) fail(DQN_ERR_INVALID, "expected %lld", h->nflux); gather_params(h, h->grad, flat); });
}

// Output of one layer of the online network on the s rows of the last step (rows 0..B-1), in the reference's memory image:
// conv stage -> (B, C, OH, OW) row-major == Flux (OW, OH, C, B); tower stage -> (B, N).  Parity tests take the ReLU masks from here.
int dqn_get_activation(dqn_engine_t* h, int stage, int tower, float* out, int64_t n) {
  return guard(h, [&] {
    const int nc = (int)h->convs.size();
    if (stage < 0 || stage >= nc + h->depth || tower < 0 || tower >= h->ntow || !out) fail(DQN_ERR_INVALID, "stage/tower");
    if (stage < nc) {
      const ConvGeom& g = h->convs[stage].g;
      const long long per = (long long)g.OH * g.OW * g.Cout;
      if (n != h->B * per) fail(DQN_ERR_INVALID, "expected %lld floats", h->B * per);
      ensure_stage(h, n * 4);
      const long long total = n;
      relayout_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 8LL * h->nsm), 256, 0, h->stream>>>((const uint8_t*)h->on.conv_out[stage], h->stage, h->B, g.Cout, g.OH * g.OW, 0, 0, 0);
      CK(cudaGetLastError());
      d2h(h, out, (const float*)h->stage, n);
    } else {
      const Mat& w = h->tow[tower][stage - nc];
      if (n != (long long)h->B * w.N) fail(DQN_ERR_INVALID, "expected %lld floats", (long long)h->B * w.N);
      d2h(h, out, (const float*)h->on.tow_out[tower][stage - nc], n);
    }
  });
}

// ---- EpisodeReplayBuffer (src/episode_replay.jl) -----------------------------------------------------------------------------------
int dqn_episode_add(dqn_engine_t* h, const float* s, const int32_t* a, const float* r, const float* sp, const uint8_t* done, int64_t len) {
  return guard(h, [&] {
    if (!h->lstm) fail(DQN_ERR_UNSUPPORTED, "not a recurrent engine (no LSTM layer in the chain)");
    if (len < 1 || len > h->Lmax || !s || !a || !r || !sp || !done) fail(DQN_ERR_INVALID, "episode length %lld outside 1..%d (max_episode_length) or null argument", (long long)len, h->Lmax);
    for (long long t = 0; t < len; ++t) if (a[t] < 1 || a[t] > h->cfg.n_actions) fail(DQN_ERR_INVALID, "action index outside 1..n_actions");
    const long long slot = h->cursor, d = h->obs_elems, base = slot * h->Lmax;
    CK(cudaMemcpyAsync(h->ep_s + base * d, s, sizeof(float) * len * d, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->ep_sp + base * d, sp, sizeof(float) * len * d, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->ep_a + base, a, sizeof(int) * len, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->ep_r + base, r, sizeof(float) * len, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->ep_done + base, done, len, cudaMemcpyHostToDevice, h->stream));
    const int len32 = (int)len; const float one = 1.f;
    CK(cudaMemcpyAsync(h->ep_len + slot, &len32, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tree + h->P + slot, &one, sizeof(float), cudaMemcpyHostToDevice, h->stream));   // uniform sampling = unit priorities
    ensure_stage(h, 64);
    CK(cudaMemcpyAsync(h->stage, &slot, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    tree_update_kernel<<<1, 32, 0, h->stream>>>(h->tree, h->P, (const long long*)h->stage, nullptr, 1, 0, h->st, 0, 1.0, 1.0, 0, nullptr);
    CK(cudaGetLastError());
    h->cursor = (h->cursor + 1) % h->cap;                      // r._idx = mod1(r._idx + 1, r.max_size)
    h->curr_size = std::min(h->cap, h->curr_size + 1);
    set_curr_size(h);
    CK(cudaStreamSynchronize(h->stream));                      // the caller's arrays (and the stack words above) are free again
  });
}
int dqn_episode_count(const dqn_engine_t* h, int64_t* curr_size, int64_t* cursor) {
  if (!h || !h->lstm) return DQN_ERR_INVALID;
  if (curr_size) *curr_size = h->curr_size;
  if (cursor) *cursor = h->cursor;
  return DQN_OK;
}
int dqn_episode_sample(dqn_engine_t* h, uint64_t call, int64_t* idx_out, int32_t* start_out) {
  return guard(h, [&] {
    if (!h->lstm) fail(DQN_ERR_UNSUPPORTED, "not a recurrent engine");
    if (h->curr_size < h->Bep) fail(DQN_ERR_STATE, "replay holds %lld episodes, batch_size is %d (episode_replay.jl:73)", h->curr_size, h->Bep);
    const int TB = h->B;
    ensure_stage(h, h->Bep * 8);
    sample_kernel<<<1, (h->Bep + 31) / 32 * 32, sample_smem(h->Bep), h->stream>>>(h->tree, h->P, h->Bep, h->cfg.seed, h->st, 1, call, (long long*)h->stage,
                                                                                 nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
    float* xs = reinterpret_cast<float*>(h->xb);
    episode_gather_kernel<<<dim3(h->T, h->Bep), 128, 0, h->stream>>>((const long long*)h->stage, h->Bep, h->T, h->Lmax, h->obs_elems, h->cfg.seed, h->st, 1, call, h->ep_len,
                                                                     h->ep_s, h->ep_sp, h->ep_a, h->ep_r, h->ep_done, xs, xs + (long long)TB * h->obs_elems, h->a_b, h->r_b,
                                                                     h->d_b, h->w_b, h->ep_start_d);
    CK(cudaGetLastError());
    if (idx_out) d2h(h, (long long*)idx_out, (const long long*)h->stage, h->Bep);
    if (start_out) d2h(h, start_out, (const int*)h->ep_start_d, h->Bep);
    check_dev_errors(h);
  });
}
int dqn_policy_reset(dqn_engine_t* h) {
  return guard(h, [&] {
    if (!h->lstm) return;                                      // Flux.reset! is a no-op without recurrent layers
    cudaStream_t keep = h->ls; h->ls = h->stream;
    lstm_broadcast(h, h->h_act, h->theta + h->h0_off, h->act_rows);
    lstm_broadcast(h, h->c_act, h->theta + h->c0_off, h->act_rows);
    h->ls = keep;
  });
}

// ---- vectorised acting on the device (the caller side of the path) ------------------------------------------------
// One forward of the online network over n lanes (the tensor-core path when the engine runs DQN_MATH_3XTF32), dueling combine, first-max
// argmax and the epsilon-greedy draw in one kernel: actions never bounce through the host.  obs_layout 0: Flux layout (C,H,W per lane,
// what dqn_q_values takes), 1: the engine's own H,W,C layout (what a device-side environment writes / dqn_replay_add_device reads back).
static void act_rows(dqn_engine_t* h, const void* obs_dev, long long n, int obs_layout, float eps, uint64_t call, int32_t* actions_dev, float* q_dev) {
  const long long rb = h->obs_row_bytes; const int nA = h->cfg.n_actions; const int L = h->depth - 1;
  if (h->lstm) fail(DQN_ERR_UNSUPPORTED, "recurrent engine: act through dqn_q_values (carried hidden state)");
  const int chunk = h->rows_on;
  for (long long t0 = 0; t0 < n; t0 += chunk) {
    const int c = (int)std::min<long long>(chunk, n - t0);
    const uint8_t* src = (const uint8_t*)obs_dev + t0 * rb;
    if (obs_layout == 1 || !h->hwc) CK(cudaMemcpyAsync(h->xb, src, c * rb, cudaMemcpyDeviceToDevice, h->stream));
    else relayout(h, src, h->xb, c, h->elem_bytes == 1, 0, 1);
    const bool tcp = h->arena && h->cfg.math_mode == DQN_MATH_3XTF32 && rb % 16 == 0 && (h->elem_bytes == 4 || h->a8);
    const float* xs = (h->arena && rb % 16 == 0 && h->elem_bytes == 4) ? (const float*)h->xb : nullptr;
    const Pass pa{h->theta, h->xb, h->elem_bytes == 1, c, &h->on, "act", xs, h->w_on_s, tcp};
    forward(h, &pa, 1);
    Scope sc(h, "act_select", 0, (double)c * (nA + 1) * 4);
    act_select_kernel<<<(c + 127) / 128, 128, 0, h->stream>>>(h->cfg.dueling ? h->on.tow_out[0][L] : nullptr, h->on.tow_out[h->ntow - 1][L], c, nA, h->cfg.dueling, eps,
                                                            h->cfg.seed ^ 0xAC7105EEDull, call, (int)t0, actions_dev + t0, q_dev ? q_dev + t0 * nA : nullptr);
    CK(cudaGetLastError());
  }
}
int dqn_act_device(dqn_engine_t* h, const void* obs_dev, int64_t n, int obs_layout, float eps, uint64_t call, int32_t* actions_dev, float* q_dev) {
  return guard(h, [&] {
    if (n < 0 || (n > 0 && (!obs_dev || !actions_dev))) fail(DQN_ERR_INVALID, "null argument");
    cudaStream_t keep = h->ls; h->ls = h->stream;
    act_rows(h, obs_dev, n, obs_layout, eps, call, actions_dev, q_dev);
    h->ls = keep;
  });
}
// the same with host buffers (batched evaluation rollouts, SURVEY 8f row 4: one call acts for every evaluation environment)
int dqn_act(dqn_engine_t* h, const void* obs, int64_t n, float eps, uint64_t call, int32_t* actions_out, float* q_out) {
  return guard(h, [&] {
    if (n < 0 || (n > 0 && (!obs || !actions_out))) fail(DQN_ERR_INVALID, "null argument");
    if (n == 0) return;
    const long long rb = h->obs_row_bytes; const int nA = h->cfg.n_actions;
    ensure_stage(h, n * rb + n * 4 + n * nA * 4 + 256);
    uint8_t* d_obs = h->stage; int* d_act = (int*)(h->stage + (n * rb + 63) / 64 * 64); float* d_q = (float*)((uint8_t*)d_act + (n * 4 + 63) / 64 * 64);
    CK(cudaMemcpyAsync(d_obs, obs, n * rb, cudaMemcpyHostToDevice, h->stream));
    cudaStream_t keep = h->ls; h->ls = h->stream;
    act_rows(h, d_obs, n, 0, eps, call, d_act, q_out ? d_q : nullptr);
    h->ls = keep;
    CK(cudaMemcpyAsync(actions_out, d_act, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (q_out) CK(cudaMemcpyAsync(q_out, d_q, sizeof(float) * n * nA, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  });
}
// synthetic vectorised environment for the bench (config 5): lanes' next observations (engine layout), rewards, done flags, |r| - all on the device
int dqn_synth_env_step(dqn_engine_t* h, void* obs_next_dev, float* rew_dev, uint8_t* done_dev, float* td0_dev, int64_t lanes, uint64_t seed, uint64_t step) {
  return guard(h, [&] {
    if (h->elem_bytes != 1 || lanes < 1) fail(DQN_ERR_UNSUPPORTED, "synthetic lanes produce byte observations");
    const long long words = (h->obs_elems + 15) / 16;
    synth_env_step_kernel<<<dim3((unsigned)std::min<long long>((words + 255) / 256, 8), (unsigned)lanes), 256, 0, h->stream>>>((uint8_t*)obs_next_dev, rew_dev, done_dev, td0_dev,
                                                                                                                    h->obs_elems, (int)lanes, seed, step);
    CK(cudaGetLastError());
  });
}

int dqn_timer_start(dqn_engine_t* h) { return guard(h, [&] { CK(cudaEventRecord(h->t0, h->stream)); }); }
int dqn_timer_stop(dqn_engine_t* h, float* ms) {
  return guard(h, [&] { CK(cudaEventRecord(h->t1, h->stream)); CK(cudaEventSynchronize(h->t1)); CK(cudaEventElapsedTime(ms, h->t0, h->t1)); });
}
int dqn_launches_per_step(const dqn_engine_t* h) { return h ? h->launches_per_step : 0; }
int dqn_collective_kind(const dqn_engine_t* h) { return !h || h->cfg.world <= 1 ? 0 : (h->peer_ar ? 2 : 1); }
int dqn_set_profiling(dqn_engine_t* h, int on) {
  return guard(h, [&] {
    CK(cudaStreamSynchronize(h->stream));
    for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->prof.clear();
    h->profiling = on;
  });
}
int dqn_get_profile(dqn_engine_t* h, char* buf, int64_t buflen) {
  return guard(h, [&] {
    CK(cudaStreamSynchronize(h->stream));
    struct Agg { double ms = 0, flops = 0, bytes = 0; int n = 0; int order = 0; };
    std::map<std::string, Agg> agg; int order = 0;
    for (auto& r : h->prof) {
      float ms = 0; CK(cudaEventElapsedTime(&ms, r.a, r.b));
      Agg& a = agg[r.name]; if (a.n == 0) a.order = order++;
      a.ms += ms; a.flops = r.flops; a.bytes = r.bytes; a.n++;
    }
    std::vector<std::pair<std::string, Agg>> v(agg.begin(), agg.end());
    std::sort(v.begin(), v.end(), [](auto& x, auto& y) { return x.second.order < y.second.order; });
    std::string s;
    char line[256];
    for (auto& kv : v) {
      snprintf(line, sizeof line, "%s %.6f %d %.0f %.0f\n", kv.first.c_str(), kv.second.ms / kv.second.n, kv.second.n, kv.second.bytes, kv.second.flops);
      s += line;
    }
    if ((int64_t)s.size() + 1 > buflen) fail(DQN_ERR_INVALID, "profile needs %zu bytes", s.size() + 1);
    memcpy(buf, s.c_str(), s.size() + 1);
  });
}
int dqn_flush_l2(dqn_engine_t* h) {
  return guard(h, [&] {
    if (!h->flush_buf) { h->flush_n = (256LL << 20) / 4; h->flush_buf = dalloc<uint32_t>(h->flush_n); }
    fill_u32_kernel<<<4 * h->nsm, 256, 0, h->stream>>>(h->flush_buf, h->flush_n, 1u);
    CK(cudaGetLastError());
  });
}
void* dqn_stream(dqn_engine_t* h) { return h ? (void*)h->stream : nullptr; }

int dqn_host_alloc(void** p, int64_t bytes) { return cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? DQN_OK : DQN_ERR_CUDA; }
int dqn_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? DQN_OK : DQN_ERR_CUDA; }

}  // extern "C"

// =================================================================================================
// Single-process data-parallel group: the reference's host is ONE Julia process (SURVEY 8b "Multi-GPU" row), so the boundary offers
// ndev engines behind one handle.  Every engine keeps its own worker thread (its CUDA context, its NCCL rank): ncclCommInitRank must
// be entered by all ranks at once, and a step's launches are issued concurrently, exactly as one process per GPU would.
struct dqn_group {
  struct Worker {
    std::thread th; std::mutex mu; std::condition_variable cv;
    std::function<int()> job; bool has_job = false, done = false, quit = false; int rc = 0;
  };
  std::vector<dqn_engine*> eng;
  std::vector<Worker*> wk;
  std::string err;
  static void loop(Worker* w) {
    for (;;) {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->has_job || w->quit; });
      if (w->quit) return;
      std::function<int()> j = w->job; w->has_job = false;
      lk.unlock();
      const int rc = j();
      lk.lock();
      w->rc = rc; w->done = true;
      w->cv.notify_all();
    }
  }
  // run f(rank) on every worker, wait for all; returns the first non-zero status
  int run(const std::function<int(int)>& f) {
    for (size_t r = 0; r < wk.size(); ++r) {
      Worker* w = wk[r];
      std::lock_guard<std::mutex> lk(w->mu);
      w->job = [f, r] { return f((int)r); }; w->has_job = true; w->done = false;
      w->cv.notify_all();
    }
    int rc = 0;
    for (size_t r = 0; r < wk.size(); ++r) {
      Worker* w = wk[r];
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->done; });
      if (w->rc != 0 && rc == 0) { rc = w->rc; err = "rank " + std::to_string(r) + ": " + (eng[r] ? eng[r]->err : g_create_error); }
    }
    return rc;
  }
};

extern "C" {

int dqn_group_create(const dqn_config_t* cfg, int ndev, const int* devices, dqn_group_t** out) {
  if (!cfg || !out || ndev < 1) { g_create_error = "null argument / ndev < 1"; return DQN_ERR_INVALID; }
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) { g_create_error = "no CUDA device: libdqn_b200 has no CPU path"; return DQN_ERR_CUDA; }
  uint8_t id[DQN_NCCL_ID_BYTES] = {0};
  if (ndev > 1) { const int rc = dqn_nccl_unique_id(id); if (rc != DQN_OK) return rc; }
  dqn_group* g = new dqn_group();
  g->eng.assign(ndev, nullptr);
  for (int r = 0; r < ndev; ++r) { auto* w = new dqn_group::Worker(); g->wk.push_back(w); w->th = std::thread(dqn_group::loop, w); }
  std::vector<std::string> errs(ndev);
  const int rc = g->run([&](int r) {
    dqn_config_t c = *cfg;
    c.device = devices ? devices[r] : r; c.rank = r; c.world = ndev;
    c.seed = cfg->seed + (uint64_t)r;                           // every shard samples its own stream (bench.py: shard_seeds)
    memcpy(c.nccl_id, id, DQN_NCCL_ID_BYTES);
    const int rr = dqn_engine_create(&c, &g->eng[r]);
    if (rr != DQN_OK) errs[r] = g_create_error;                // thread-local on the worker: carry it out
    return rr;
  });
  if (rc != DQN_OK) {
    for (int r = 0; r < ndev; ++r) if (!errs[r].empty()) { g_create_error = "rank " + std::to_string(r) + ": " + errs[r]; break; }
    dqn_group_destroy(g);
    return rc;
  }
  *out = g;
  return DQN_OK;
}

void dqn_group_destroy(dqn_group_t* g) {
  if (!g) return;
  g->run([&](int r) { if (g->eng[r]) { dqn_engine_destroy(g->eng[r]); g->eng[r] = nullptr; } return 0; });
  for (auto* w : g->wk) {
    { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
    w->th.join();
    delete w;
  }
  delete g;
}

int dqn_group_size(const dqn_group_t* g) { return g ? (int)g->eng.size() : 0; }
dqn_engine_t* dqn_group_engine(dqn_group_t* g, int rank) { return (g && rank >= 0 && rank < (int)g->eng.size()) ? g->eng[rank] : nullptr; }
const char* dqn_group_last_error(const dqn_group_t* g) { return g ? g->err.c_str() : g_create_error.c_str(); }

int dqn_group_set_params(dqn_group_t* g, int which, const float* flat, int64_t n) {
  if (!g) return DQN_ERR_INVALID;
  return g->run([&](int r) { return dqn_set_params(g->eng[r], which, flat, n); });
}
int dqn_group_sync_target(dqn_group_t* g) {
  if (!g) return DQN_ERR_INVALID;
  return g->run([&](int r) { return dqn_sync_target(g->eng[r]); });
}
// one data-parallel batch_train!: every shard samples its own batch, the gradient bucket is all-reduced, identical Adam everywhere.
// loss = mean of the shards' losses (the loss of the combined batch of B * ndev samples), grad_norm = max|g| of the reduced gradient.
int dqn_group_train_step(dqn_group_t* g, float* loss, float* grad_norm) {
  if (!g) return DQN_ERR_INVALID;
  std::vector<float> l(g->eng.size(), 0.f), gn(g->eng.size(), 0.f);
  const int rc = g->run([&](int r) { return dqn_train_step(g->eng[r], &l[r], &gn[r]); });
  if (rc != DQN_OK) return rc;
  double s = 0; for (float v : l) s += v;
  if (loss) *loss = (float)(s / l.size());
  if (grad_norm) *grad_norm = gn[0];
  return DQN_OK;
}

}  // extern "C"
