/* dqn_b200.h - C-ABI of libdqn_b200.so, the B200-native engine behind the `batch_train!` hot path of
 * JuliaPOMDP/DeepQLearning.jl.  Plain pointers and sizes only (no torch / C++ types); this is what the
 * reference's Julia host binds with `ccall` (see INTEGRATION.md and julia/DeepQLearningB200.jl), and what
 * the Python host mirror in deepqlearning.jl_b200/ binds with ctypes.
 *
 * Each entry point cites the reference interface it replaces (paths into the reference tree):
 *   PER  = src/prioritized_experience_replay.jl, SOLVER = src/solver.jl, DUEL = src/dueling.jl,
 *   POLICY = src/policy.jl, HELPERS = src/helpers.jl.
 *
 * Conventions
 *   - every function returns 0 (DQN_OK) or a negative dqn_status; dqn_last_error() gives the message.
 *     Nothing throws or aborts across the boundary.  The reference's @assert / throw sites map to
 *     DQN_ERR_STATE / DQN_ERR_INVALID (PER:66,78,83,84,90; SOLVER:46; DUEL:47-50; POLICY:44).
 *   - all pointers are HOST pointers owned by the caller and only touched during the call, unless the
 *     name ends in _device.  The engine owns all device memory.
 *   - arrays use the reference's memory images: observations as Flux stores them (W,H,C,N column-major
 *     == N,C,H,W row-major), parameters as the concatenation of Flux.params(active_q) in Flux order
 *     (DuelingNetwork fields base, val, adv - DUEL:2-6,13; per layer weight then bias; Dense weight
 *     (out,in) column-major, Conv weight (kw,kh,cin,cout) column-major, true convolution).
 *   - actions are 1-based Int32 as in DQExperience (PER:3-9); sampled indices are 0-based int64
 *     (the reference's Vector{Int64} minus one).
 *   - a handle is not thread-safe; one handle per GPU, calls on it serialised by the caller.
 */
#ifndef DQN_B200_H
#define DQN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQN_ABI_VERSION 1
#define DQN_MAX_LAYERS 16
#define DQN_NCCL_ID_BYTES 128

typedef struct dqn_engine dqn_engine_t;
typedef struct dqn_group dqn_group_t;

typedef enum {
  DQN_OK = 0,
  DQN_ERR_INVALID = -1,     /* bad argument / unsupported topology (DUEL:47-50) */
  DQN_ERR_CUDA = -2,        /* CUDA runtime or driver error */
  DQN_ERR_STATE = -3,       /* reference @assert would fire (e.g. curr_size < batch_size, PER:83) */
  DQN_ERR_NCCL = -4,
  DQN_ERR_UNSUPPORTED = -5
} dqn_status;

enum { DQN_ACT_IDENTITY = 0, DQN_ACT_RELU = 1, DQN_ACT_TANH = 2, DQN_ACT_SIGMOID = 3 };
enum { DQN_LAYER_DENSE = 0, DQN_LAYER_CONV = 1, DQN_LAYER_FLATTEN = 2,
       DQN_LAYER_LSTM = 3 };   /* Flux.LSTM(in, out): makes the engine recurrent (SOLVER:239-287, EpisodeReplayBuffer) */
enum { DQN_OBS_F32 = 0, DQN_OBS_U8 = 1 };   /* U8: value k stands for Float32(k)/255f0 */
enum { DQN_NET_ONLINE = 0, DQN_NET_TARGET = 1 };
enum { DQN_Q_S_ONLINE = 0, DQN_Q_SP_ONLINE = 1, DQN_Q_SP_TARGET = 2 };
enum { DQN_MATH_FP32 = 0,      /* CUDA-core fp32 FMA contractions */
       DQN_MATH_3XTF32 = 1 };  /* tcgen05 kind::tf32, error-compensated 3-pass split */

/* One layer of the Chain handed to DeepQLearningSolver(qnetwork = ...) (SOLVER:2), before the dueling split. */
typedef struct {
  int32_t kind;          /* DQN_LAYER_* ; FLATTEN stands for flattenbatch (HELPERS:6-8) */
  int32_t act;           /* DQN_ACT_* */
  int32_t in, out;       /* Dense: in/out features.  Conv: cin/cout */
  int32_t kh, kw, stride;/* Conv only (pad = 0) */
} dqn_layer_t;

/* Mirrors the DeepQLearningSolver fields that reach the hot path (SOLVER:1-28) plus the buffer
 * constructor's keywords (PER:39-45).  Defaults of the reference: alpha=0.6, beta=0.4, eps=1e-3,
 * learning_rate=1e-4, batch_size=32, buffer_size=1000, double_q=dueling=prioritized_replay=true. */
typedef struct {
  int32_t abi_version;        /* DQN_ABI_VERSION */
  int32_t device;             /* CUDA ordinal */
  int32_t obs_c, obs_h, obs_w;/* observation (W,H,C) in Flux terms; a flat observation of d features is c=d,h=w=1 */
  int32_t obs_dtype;          /* DQN_OBS_* : storage type of the replay store and of `s`/`sp` passed in */
  int32_t n_actions;
  int32_t n_layers;
  dqn_layer_t layers[DQN_MAX_LAYERS];
  int32_t dueling;            /* SOLVER:11, split rule DUEL:36-58 */
  int32_t double_q;           /* SOLVER:10 */
  int32_t prioritized_replay; /* SOLVER:16 : 0 => priorities are never updated and new ones use td0 as given */
  int32_t batch_size;         /* SOLVER:5, <= 1024 */
  int64_t buffer_size;        /* SOLVER:20 */
  float alpha, beta, eps;     /* PER:43-45 */
  float learning_rate;        /* SOLVER:3 (Float32, widened to Float64 inside Adam) */
  float discount;             /* gamma = Float32(discount) SOLVER:208 */
  double adam_beta1, adam_beta2, adam_eps;   /* Flux.Optimise.Adam defaults 0.9, 0.999, 1e-8 */
  uint64_t seed;              /* Philox key of the sum-tree sampler */
  int32_t math_mode;          /* DQN_MATH_* */
  int32_t use_graph;          /* 1: replay the whole step as one CUDA graph */
  int32_t rank, world;        /* data-parallel group; world=1 => no collective */
  uint8_t nccl_id[DQN_NCCL_ID_BYTES];   /* ncclUniqueId from dqn_nccl_unique_id (rank 0), ignored if world==1 */
  int32_t max_act_rows;       /* rows per dqn_q_values chunk (0 => 2*batch_size) */
  int32_t trace_length;       /* recurrent engines: SOLVER:15 trace_length (default 40) */
  int32_t max_episode_length; /* recurrent engines: steps stored per episode (SOLVER:21 max_episode_length, default 100) */
  int32_t reserved[5];
} dqn_config_t;

/* ---- lifecycle ------------------------------------------------------------------------------------ */
int dqn_config_default(dqn_config_t* cfg);                       /* fills the reference defaults */
int dqn_engine_create(const dqn_config_t* cfg, dqn_engine_t** out);   /* replaces SOLVER:40-66 setup: buffer, dueling split, target copy, Adam */
void dqn_engine_destroy(dqn_engine_t* h);
const char* dqn_last_error(const dqn_engine_t* h);               /* h may be NULL: error of the last failed create on this thread */
int dqn_nccl_unique_id(uint8_t id_out[DQN_NCCL_ID_BYTES]);
int dqn_device_count(int* n);

/* ---- single-process data-parallel group (the Julia host is one process, SURVEY 8b "Multi-GPU") -----------------------------------
 * ndev engines behind one handle: rank r on devices[r] (NULL: 0..ndev-1), world = ndev, sampler seed cfg->seed + r, one NCCL
 * communicator created inside.  Per-shard calls (dqn_replay_add, dqn_replay_fill_synthetic, dqn_get_params, ...) go through
 * dqn_group_engine(g, r); the calls below act on all ranks at once (each on its own worker thread). */
int dqn_group_create(const dqn_config_t* cfg, int ndev, const int* devices, dqn_group_t** out);
void dqn_group_destroy(dqn_group_t* g);
int dqn_group_size(const dqn_group_t* g);
dqn_engine_t* dqn_group_engine(dqn_group_t* g, int rank);
const char* dqn_group_last_error(const dqn_group_t* g);
int dqn_group_set_params(dqn_group_t* g, int which, const float* flat, int64_t n);   /* the same parameters on every rank */
int dqn_group_sync_target(dqn_group_t* g);
int dqn_group_train_step(dqn_group_t* g, float* loss, float* grad_norm);            /* one data-parallel batch_train! (SOLVER:191-236 over B * ndev samples) */

/* ---- parameters: Flux.params(active_q) / loadparams! (SOLVER:143-144, 173-174, 292, 314-316) ------- */
int64_t dqn_num_params(const dqn_engine_t* h);
int dqn_set_params(dqn_engine_t* h, int which, const float* flat, int64_t n);
int dqn_get_params(dqn_engine_t* h, int which, float* flat, int64_t n);
int dqn_sync_target(dqn_engine_t* h);                            /* Flux.loadparams!(target_q, params(active_q)) SOLVER:142-145 */
int dqn_get_adam_state(dqn_engine_t* h, float* m_flat, float* v_flat, double beta_pow[2], int64_t n);

/* ---- replay writes: add_exp! (PER:65-74; callers SOLVER:88-95, PER:121-122) ------------------------ */
/* n transitions appended at the ring cursor; priority (td0+eps)^alpha; DQN_ERR_STATE if td0+eps <= 0.
 * The call returns when the caller's buffers have been copied (its own copy stream); the ring / sum-tree writes run asynchronously, in
 * program order with every other call on the handle - if a step is in flight (dqn_train_step_async) they are ordered behind that step's
 * priority update, not behind the whole step, so add + async step + dqn_step_result(back = 1) keeps the GPU busy between steps. */
int dqn_replay_add(dqn_engine_t* h, const void* s, const int32_t* a, const float* r, const void* sp,
                   const uint8_t* done, const float* td0, int64_t n);
int dqn_replay_add_device(dqn_engine_t* h, const void* s, const int32_t* a, const float* r, const void* sp,
                          const uint8_t* done, const float* td0, int64_t n);
int dqn_replay_size(const dqn_engine_t* h, int64_t* curr_size, int64_t* cursor);    /* _curr_size, _idx-1 */
/* Synthetic fill on the device (bench / full-size tests): transition i is a pure function of (seed, i);
 * oracle/synthetic.py regenerates any subset on the CPU. */
int dqn_replay_fill_synthetic(dqn_engine_t* h, int64_t n, uint64_t seed);
int dqn_replay_read(dqn_engine_t* h, const int64_t* idx, int64_t n, void* s, int32_t* a, float* r, void* sp, uint8_t* done);

/* ---- recurrent engines (a DQN_LAYER_LSTM in the chain): EpisodeReplayBuffer (src/episode_replay.jl) ----------------------------
 * buffer_size counts EPISODES here (SOLVER:184 EpisodeReplayBuffer(env, solver.buffer_size, batch_size, trace_length)); observations
 * are Float32; dqn_train_step is then the recurrent batch_train! (SOLVER:239-287): batch_size episodes, trace_length steps, the
 * start-offset behaviour of src/episode_replay.jl:81-92 preserved.  The step's diagnostics (dqn_get_q, dqn_get_td, dqn_get_targets,
 * dqn_get_is_weights = the Int32 trace mask as floats) have trace_length * batch_size rows, time-major (row t * batch_size + i). */
int dqn_episode_add(dqn_engine_t* h, const float* s, const int32_t* a, const float* r, const float* sp, const uint8_t* done, int64_t len);  /* add_episode! :60-66 */
int dqn_episode_count(const dqn_engine_t* h, int64_t* curr_size, int64_t* cursor);
int dqn_episode_sample(dqn_engine_t* h, uint64_t call, int64_t* idx_out, int32_t* start_out);   /* indices / ep_start sampling call `call` draws; no state change */
int dqn_policy_reset(dqn_engine_t* h);                           /* resetstate!(policy) POLICY:32-34: acting hidden state <- state0 */

/* ---- priorities: update_priorities! (PER:76-80) --------------------------------------------------- */
int dqn_update_priorities(dqn_engine_t* h, const int64_t* idx, const float* td, int64_t n);
int dqn_set_priorities(dqn_engine_t* h, const int64_t* idx, const float* prio, int64_t n);  /* raw leaves (tests) */
int dqn_get_priorities(dqn_engine_t* h, float* out, int64_t n);          /* r._priorities[1:n] */
int dqn_get_tree(dqn_engine_t* h, float* out, int64_t n_nodes);          /* heap order, node 1 = root */
int64_t dqn_tree_nodes(const dqn_engine_t* h);

/* ---- sampling: StatsBase.sample(r) / get_batch (PER:82-104) --------------------------------------- */
int dqn_sample_indices(dqn_engine_t* h, uint64_t call, int64_t* idx_out);  /* indices sampling call number `call` would draw; no state change */
int dqn_get_batch(dqn_engine_t* h, const int64_t* idx, float* s, int32_t* a, float* r, float* sp,
                  float* done, float* weights);                            /* get_batch(r, idx): Float32 arrays in Flux layout, a 1-based */

/* ---- the step: batch_train! (SOLVER:191-236) ------------------------------------------------------ */
int dqn_train_step(dqn_engine_t* h, float* loss, float* grad_norm);        /* sample + step; scalars as SOLVER:235 */
int dqn_train_step_with_indices(dqn_engine_t* h, const int64_t* idx, float* loss, float* grad_norm);
int dqn_train_step_async(dqn_engine_t* h);                                 /* enqueue only */
int dqn_sync(dqn_engine_t* h, float* loss, float* grad_norm);              /* wait, fetch the last step's scalars */
/* (loss_val, grad_norm) of the step launched `back` steps before the latest one: back = 0 is dqn_sync; back = 1 waits only for the
 * step before the latest, so a caller that logs the scalars (SOLVER:147-166 is their only use) can add the next transitions and
 * launch the next step while this one runs:  add_k; async_k; result(back=1) -> step k-1. */
int dqn_step_result(dqn_engine_t* h, int back, float* loss, float* grad_norm);

/* ---- acting: policy.qnetwork(obatch) (POLICY:38-64, SOLVER:83) ------------------------------------ */
int dqn_q_values(dqn_engine_t* h, int which, const void* obs, int64_t n, float* q_out);   /* q_out (n, n_actions) row-major == (|A|, n) column-major */

/* vectorised acting (SOLVER:83 action(exploration_policy, policy, k, obs) for n lanes; POLICY:38-46): forward + dueling combine + first-max
 * argmax + epsilon-greedy draw on the device.  Uniforms: Philox(seed ^ 0xAC7105EED; lane, {0, 1}, call): lane i explores iff u0 < eps and
 * then takes action 1 + floor(u1 * n_actions).  Actions are 1-based Int32.  _device: obs / actions / q are DEVICE pointers; obs_layout
 * 0 = Flux layout per lane (C,H,W), 1 = the engine's H,W,C layout (what dqn_replay_add_device stores).  q may be NULL. */
int dqn_act(dqn_engine_t* h, const void* obs, int64_t n, float eps, uint64_t call, int32_t* actions_out, float* q_out);
int dqn_act_device(dqn_engine_t* h, const void* obs_dev, int64_t n, int obs_layout, float eps, uint64_t call, int32_t* actions_dev, float* q_dev);
/* synthetic vectorised environment step for the bench (BASELINE.json configs[4]): next observations (uint8, engine layout), rewards,
 * done flags and |r| of `lanes` lanes as a pure function of (seed, step, lane), written to device buffers */
int dqn_synth_env_step(dqn_engine_t* h, void* obs_next_dev, float* rew_dev, uint8_t* done_dev, float* td0_dev, int64_t lanes, uint64_t seed, uint64_t step);

/* ---- diagnostics of the last step (parity tests) -------------------------------------------------- */
int dqn_get_last_indices(dqn_engine_t* h, int64_t* idx_out);
int dqn_get_td(dqn_engine_t* h, float* td_out);                  /* td_vals SOLVER:222 */
int dqn_get_is_weights(dqn_engine_t* h, float* w_out);           /* importance_weights PER:101-102 */
int dqn_get_q(dqn_engine_t* h, int which_q, float* q_out);       /* (B, n_actions) */
int dqn_get_targets(dqn_engine_t* h, float* y_out, int32_t* best_a_out);   /* q_targets SOLVER:217, best_a 1-based SOLVER:212 */
int dqn_get_grads(dqn_engine_t* h, float* flat, int64_t n);      /* gs in Flux.params order/layout */
/* output of layer `stage` (conv layers first, then the Dense layers of tower `tower`: 0 = val or the only tower, 1 = adv) of the
 * online network on the s rows of the last step, in the reference's memory image ((B,C,OH,OW) row-major / (B,N)); the parity
 * tests read the ReLU masks of the engine's own forward pass from here (Zygote's relu pullback, SOLVER:219-225) */
int dqn_get_activation(dqn_engine_t* h, int stage, int tower, float* out, int64_t n);

/* ---- measurement ---------------------------------------------------------------------------------- */
int dqn_timer_start(dqn_engine_t* h);                            /* CUDA event on the engine's stream */
int dqn_timer_stop(dqn_engine_t* h, float* ms);                  /* second event, synchronise, elapsed */
int dqn_launches_per_step(const dqn_engine_t* h);                /* kernels of ours launched by one dqn_train_step */
int dqn_collective_kind(const dqn_engine_t* h);                  /* gradient all-reduce of this engine: 0 none (world 1), 1 NCCL, 2 the engine's own
                                                                  * kernel over NVLink peer memory (peer_ar.cuh) */
int dqn_set_profiling(dqn_engine_t* h, int on);                  /* eager launches bracketed by events */
int dqn_get_profile(dqn_engine_t* h, char* buf, int64_t buflen); /* "name ms count bytes flops\n" per kernel */
int dqn_flush_l2(dqn_engine_t* h);                               /* writes a buffer larger than L2 */
void* dqn_stream(dqn_engine_t* h);                               /* cudaStream_t of the engine */
int dqn_host_alloc(void** p, int64_t bytes);                     /* pinned host memory for dqn_replay_add / dqn_q_values callers */
int dqn_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* DQN_B200_H */
