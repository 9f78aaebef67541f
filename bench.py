#!/usr/bin/env python
"""bench.py - DQN gradient-steps/s of the batch_train! hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          our engine (one process per GPU under torchrun)
    python bench.py --impl reference ...                   the reference's CPU path (torch-CPU restatement, oracle/)

Workload (N=1): BASELINE.json configs[2] - synthetic Atari-shaped uint8 observations 84x84x4, Nature-DQN conv trunk
+ dueling heads (|A|=6), batch 256, 1M-transition prioritized replay shard resident in HBM, double-Q + PER.
A "step" is one full batch_train!: sum-tree sample, gather, 3 forwards, fused head, reverse pass, Adam, priority
write-back.  Under N>1 every rank owns a private 1M shard and the gradient bucket is all-reduced (weak scaling).
One JSON line is printed by rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "dqn_gradient_steps_per_sec"
UNIT = "steps/s"
TRAIN_FREQ = 4          # env steps per gradient step (src/solver.jl:6) => transitions ingested per e2e step


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sus=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.rows, self.proc, self.dev = [], None, dev

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.dev)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def load_traffic():
    """per-launch DRAM bytes of the step's kernels from the committed ncu capture (profiles/*_traffic_*.json), newest file wins"""
    import glob
    best = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic_*.json"))):
        try:
            best = json.load(open(f))
            best["_file"] = os.path.basename(f)
        except Exception:
            pass
    return best


def run_reference(args):
    """The reference's CPU path on this box's host cores (Flux-equivalent restatement, torch-CPU, all threads)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import util
    from oracle.cpu_baseline import CpuBaseline
    spec = util.SPECS["c3_conv"]
    cb = CpuBaseline(spec["layers"], spec["obs"], spec["nA"], 256, args.buffer, True, store_rows=4096)
    ts = cb.time_steps(args.steps, args.warmup)
    total = float(np.sum(ts))
    v = args.steps / total
    sample = f"{args.steps} full steps (B=256, O(N) sampling over N={args.buffer} priorities; observation store bounded to 4096 rows, index mod 4096)"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": workload_config(args, 1),
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cb.threads, "kind": "port", "sample": sample,
                                       "label": "Flux-equivalent CPU restatement (torch-CPU); Julia is not installed in this image"},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


MATH_LABEL = {"fp32": "fp32 (CUDA-core FMA)", "3xtf32": "3xtf32 (tcgen05 kind::tf32, 3-pass split)"}


def resolve_math(args):
    """both arms print the same `config`: 'auto' is the engine's benchmarked mode"""
    key = "3xtf32" if args.math in ("auto", "3xtf32") or args.math.startswith("3xtf32") else "fp32"
    args.math_key = key
    args.math = MATH_LABEL[key]


def drqn_config(args):
    return {"workload": "BASELINE.json configs[3]: recurrent DRQN, Chain(flattenbatch, LSTM(128,128), Dense(128,16)) + dueling, double-Q, "
                        "EpisodeReplayBuffer of 1000 synthetic episodes (32..100 steps of dim-128 Float32 observations), trace_length 32, batch 64",
            "batch_per_gpu": 64, "trace_length": 32, "parallelism": "dp1",
            "l2": "latency-bound by nature (32 sequential LSTM steps per pass); working set far below L2", "math": args.math}


def run_drqn(args):
    """configs[3]: the recurrent batch_train! (src/solver.jl:239-287).  Device-resident steps/s, end-to-end steps/s (one whole episode added
    from pinned host memory per 8 steps + the scalars read back), and the CPU restatement (oracle/recurrent.py, numpy) beside it."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    d, H, nA, T, B, cap, L = 128, 128, 16, 32, 64, 1000, 100
    rng = np.random.default_rng(11)
    import oracle as O
    if args.impl == "reference":
        net = O.make_recurrent_q(d, H, [(H, nA, 0)], True, rng); tgt = O.make_recurrent_q(d, H, [(H, nA, 0)], True, rng)
        s = rng.normal(size=(T, B, d)).astype(np.float32); sp = rng.normal(size=(T, B, d)).astype(np.float32)
        a = rng.integers(1, nA + 1, (T, B)).astype(np.int32); r = rng.uniform(-1, 1, (T, B)).astype(np.float32)
        done = np.zeros((T, B), np.float32); mask = np.ones((T, B), np.int32)
        opt = O.Adam(1e-4)
        ts = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter(); O.batch_train_recurrent(net, tgt, opt, (s, a, r, sp, done, mask), 0.99, True); ts.append(time.perf_counter() - t0)
        total = float(np.sum(ts[args.warmup:])); v = args.steps / total
        print(json.dumps({"impl": "reference", "metric": "drqn_gradient_steps_per_sec", "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": drqn_config(args), "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                                                        "sample": f"{args.steps} full recurrent steps, numpy restatement (oracle/recurrent.py)"},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import dqn_b200 as lib
    layers = [dict(kind=2, act=0, in_=0, out=0), dict(kind=3, act=0, in_=d, out=H), dict(kind=0, act=0, in_=H, out=nA)]
    math_mode = lib.MATH_3XTF32 if args.math_key == "3xtf32" else lib.MATH_FP32
    cfg = lib.make_config(layers, (d,), nA, obs_dtype="f32", dueling=True, double_q=True, prioritized_replay=False, batch_size=B, buffer_size=cap,
                          learning_rate=1e-4, discount=0.99, seed=2, math_mode=math_mode, use_graph=not args.no_graph, trace_length=T, max_episode_length=L)
    eng = lib.Engine(cfg)
    net = O.make_recurrent_q(d, H, [(H, nA, 0)], True, rng)
    eng.set_params(np.concatenate([p.ravel() for p in net.params()]), 0)
    eng.sync_target()
    eps = []
    for _ in range(cap):
        n = int(rng.integers(32, L + 1))
        ep = (rng.normal(size=(n, d)).astype(np.float32), rng.integers(1, nA + 1, n).astype(np.int32), rng.uniform(-1, 1, n).astype(np.float32),
              rng.normal(size=(n, d)).astype(np.float32), np.r_[np.zeros(n - 1), 1].astype(np.uint8))
        eng.episode_add(*ep)
        if len(eps) < 16:
            eps.append(ep)
    for _ in range(args.warmup):
        eng.train_step_async()
    eng.sync()
    clocks = ClockSampler(0); clocks.start()
    eng.timer_start()
    for _ in range(args.steps):
        eng.train_step_async()
    ms = eng.timer_stop()
    loss, gn = eng.sync()
    value = args.steps / (ms * 1e-3)
    h2d = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        if i % 8 == 0:
            ep = eps[(i // 8) % len(eps)]
            eng.episode_add(*ep)
            h2d += sum(x.nbytes for x in ep)
        loss, gn = eng.train_step()
    e2e = args.steps / (time.perf_counter() - t0)
    clk = clocks.stop()
    eng.set_profiling(1)
    for _ in range(5):
        eng.train_step_async()
    eng.sync()
    prof = eng.get_profile()
    eng.set_profiling(0)
    tot = sum(k["ms"] * k["count"] / 5 for k in prof)
    kernels = [{"name": k["name"], "ms": round(k["ms"], 5), "launches_per_step": k["count"] // 5, "share": round(k["ms"] * k["count"] / 5 / tot, 4)}
               for k in sorted(prof, key=lambda k: -k["ms"] * k["count"])][:10]
    cpu = None
    if not args.no_cpu:
        tgt = O.make_recurrent_q(d, H, [(H, nA, 0)], True, rng)
        s = rng.normal(size=(T, B, d)).astype(np.float32); sp = rng.normal(size=(T, B, d)).astype(np.float32)
        a = rng.integers(1, nA + 1, (T, B)).astype(np.int32); r = rng.uniform(-1, 1, (T, B)).astype(np.float32)
        opt = O.Adam(1e-4)
        ts = []
        for i in range(2 + min(args.cpu_steps, 20)):
            t1 = time.perf_counter(); O.batch_train_recurrent(net, tgt, opt, (s, a, r, sp, np.zeros((T, B), np.float32), np.ones((T, B), np.int32)), 0.99, True)
            ts.append(time.perf_counter() - t1)
        cpu = {"value": 1.0 / float(np.median(ts[2:])), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"median of {len(ts) - 2} full recurrent steps, numpy restatement (oracle/recurrent.py), BLAS threads as available"}
    lstm = [k for k in prof if k["name"].startswith("lstm_step") or k["name"] == "lstm_bptt_step"]
    seq_ms = sum(k["ms"] * k["count"] / 5 for k in lstm)
    print(json.dumps({"metric": "drqn_gradient_steps_per_sec", "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": drqn_config(args),
                      "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": 8,
                              "what": "per step: batch_train! + (loss, grad_norm) read back; one whole episode added from host memory every 8 steps; wall clock"},
                      "gpu_launches": eng.launches_per_step() * args.steps, "clocks": clk,
                      "roofline": {"kernel": "lstm_step_* (sequential recurrence)", "bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                                   "note": f"{seq_ms:.3f} ms of the eager step are the 3 x 32 dependent recurrence launches; no HBM or tensor roofline applies (SURVEY 8d: C4 is latency-bound)"},
                      "cpu_baseline": cpu, "kernels": kernels, "last_loss": loss, "last_grad_norm": gn}))
    eng.close()


def mlp_config(args):
    return {"workload": "BASELINE.json configs[1]: synthetic vector obs (dim 128, Float32), |A|=16, 3x256 Dense MLP + dueling (two towers), batch 256, "
                        f"{args.buffer}-transition PER buffer, double-Q, Adam lr 1e-4",
            "batch_per_gpu": 256, "buffer_per_gpu": args.buffer, "parallelism": "dp1",
            "l2": "latency-bound by nature (0.82 GFLOP, 9.6 MB per step: SURVEY 8d); the gathered rows come from a 1 GB store, everything else lives in L2",
            "math": args.math}


def run_mlp(args):
    """configs[1]: the PER step on the MLP.  Device-resident steps/s, end-to-end steps/s (4 transitions added from pinned host memory and the
    scalars read back every step, one step ahead like the conv workload), CPU restatement beside it."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import util
    spec = util.SPECS["c2_mlp"]
    if args.impl == "reference":
        from oracle.cpu_baseline import CpuBaseline
        cb = CpuBaseline(spec["layers"], spec["obs"], spec["nA"], 256, args.buffer, False)
        ts = cb.time_steps(args.steps, args.warmup)
        total = float(np.sum(ts)); v = args.steps / total
        print(json.dumps({"impl": "reference", "metric": "mlp_gradient_steps_per_sec", "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": mlp_config(args), "cpu_baseline": {"value": v, "unit": UNIT, "cores": cb.threads, "kind": "port", "sample": f"{args.steps} full steps"},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import dqn_b200 as lib
    import oracle as O
    math_mode = lib.MATH_3XTF32 if args.math_key == "3xtf32" else lib.MATH_FP32
    cfg = lib.make_config(util.layer_descs(spec), (128,), 16, obs_dtype="f32", batch_size=256, buffer_size=args.buffer, learning_rate=1e-4, discount=0.99,
                          seed=2, math_mode=math_mode, use_graph=not args.no_graph)
    eng = lib.Engine(cfg)
    eng.set_params(O.flat_params(util.make_oracle_net(spec, True, seed=1)), 0)
    eng.sync_target()
    eng.replay_fill_synthetic(args.buffer, seed=1000)
    for _ in range(args.warmup):
        eng.train_step_async()
    eng.sync()
    clocks = ClockSampler(0); clocks.start()
    eng.timer_start()
    for _ in range(args.steps):
        eng.train_step_async()
    ms = eng.timer_stop()
    loss, gn = eng.sync()
    value = args.steps / (ms * 1e-3)
    rng = np.random.default_rng(5)
    s_h = lib._capi.pinned_empty((TRAIN_FREQ, 128), np.float32); sp_h = lib._capi.pinned_empty((TRAIN_FREQ, 128), np.float32)
    s_h[...] = rng.normal(size=s_h.shape); sp_h[...] = rng.normal(size=sp_h.shape)
    a_h = rng.integers(1, 17, TRAIN_FREQ).astype(np.int32); r_h = rng.uniform(-1, 1, TRAIN_FREQ).astype(np.float32)
    d_h = np.zeros(TRAIN_FREQ, np.uint8); td_h = np.abs(r_h)
    h2d = int(s_h.nbytes + sp_h.nbytes + a_h.nbytes + r_h.nbytes + d_h.nbytes + td_h.nbytes)
    for _ in range(args.warmup):
        eng.replay_add(s_h, a_h, r_h, sp_h, d_h, td_h); eng.train_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.replay_add(s_h, a_h, r_h, sp_h, d_h, td_h)
        loss, gn = eng.train_step()
    e2e_sync = args.steps / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for k in range(args.steps):
        eng.replay_add(s_h, a_h, r_h, sp_h, d_h, td_h)
        eng.train_step_async()
        if k:
            loss, gn = eng.step_result(1)
    loss, gn = eng.step_result(0)
    e2e = args.steps / (time.perf_counter() - t0)
    clk = clocks.stop()
    eng.set_profiling(1)
    for _ in range(5):
        eng.train_step_async()
    eng.sync()
    prof = eng.get_profile()
    eng.set_profiling(0)
    tot = sum(k["ms"] * k["count"] / 5 for k in prof)
    kernels = [{"name": k["name"], "ms": round(k["ms"], 5), "share": round(k["ms"] * k["count"] / 5 / tot, 4)} for k in sorted(prof, key=lambda k: -k["ms"] * k["count"])][:10]
    adam = next((k for k in prof if k["name"].startswith("adam")), None)
    peaks = load_peaks()
    roof = None
    if adam:
        ach = adam["bytes"] / (adam["ms"] * 1e-3) / 1e9
        roof = {"kernel": "adam", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None,
                "note": "9.3 MB of optimizer state per launch: L2-resident and launch-latency-bound at this size; the step as a whole is latency-bound (SURVEY 8d), "
                        f"eager per-kernel sum {tot:.3f} ms"}
    cpu = None
    if not args.no_cpu:
        from oracle.cpu_baseline import CpuBaseline
        cb = CpuBaseline(spec["layers"], spec["obs"], spec["nA"], 256, args.buffer, False)
        ts = cb.time_steps(args.cpu_steps, 3)
        cpu = {"value": 1.0 / float(np.median(ts)), "unit": UNIT, "cores": cb.threads, "kind": "port",
               "sample": f"median of {args.cpu_steps} full steps after 3 warm-ups (B=256, O(N) sampling over N={args.buffer})",
               "label": "Flux-equivalent CPU restatement (torch-CPU); Julia is not installed in this image"}
    print(json.dumps({"metric": "mlp_gradient_steps_per_sec", "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": mlp_config(args),
                      "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "sync_value": e2e_sync,
                              "what": f"per step: add_exp! x{TRAIN_FREQ} from pinned host memory + batch_train! + (loss, grad_norm) read back; wall clock; one step ahead (sync_value: strictly serial)"},
                      "gpu_launches": eng.launches_per_step() * args.steps, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
                      "last_loss": loss, "last_grad_norm": gn}))
    eng.close()


def workload_config(args, world):
    return {"workload": "BASELINE.json configs[2]: synthetic Atari-shaped obs 84x84x4 (u8), Nature-DQN conv + dueling, |A|=6, batch 256/GPU, "
                        f"{args.buffer}-transition PER shard/GPU, double-Q, Adam lr 1e-4",
            "batch_per_gpu": 256, "buffer_per_gpu": args.buffer, "parallelism": f"dp{world}",
            "l2": "inputs larger than L2: every step gathers 14.5 MB of fresh rows from a 56 GB store and streams 92 MB of optimizer state",
            "math": args.math}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--buffer", type=int, default=1_000_000)
    ap.add_argument("--math", default=os.environ.get("DQN_MATH", "auto"), choices=["auto", "fp32", "3xtf32"])
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (tuning runs)")
    ap.add_argument("--workload", default="conv", choices=["conv", "drqn", "mlp"],
                    help="conv: BASELINE.json configs[2] (the headline, default); drqn: configs[3] (LSTM-128, seq 32, batch 64); mlp: configs[1]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    resolve_math(args)
    if args.workload == "drqn":
        return run_drqn(args)
    if args.workload == "mlp":
        return run_mlp(args)
    if args.impl == "reference":
        return run_reference(args)

    import dqn_b200 as lib
    import util
    cp = lib.ControlPlane(args.gpus)
    rank, world, local = cp.rank, cp.world, cp.local_rank
    peaks = load_peaks()
    spec = util.SPECS["c3_conv"]
    nccl_id = None
    if world > 1:
        nccl_id = cp.broadcast_bytes(lib.nccl_unique_id() if rank == 0 else None, 128)
    seeds = lib.shard_seeds(0, rank)
    math_mode = lib.MATH_3XTF32 if args.math_key == "3xtf32" else lib.MATH_FP32
    cfg = lib.make_config(util.layer_descs(spec), (84, 84, 4), 6, obs_dtype="u8", batch_size=256, buffer_size=args.buffer, learning_rate=1e-4,
                          discount=0.99, seed=seeds["sampler"], device=local, math_mode=math_mode, use_graph=not args.no_graph, rank=rank, world=world, nccl_id=nccl_id)
    eng = lib.Engine(cfg)
    net = util.make_oracle_net(spec, True, seed=seeds["weights"])   # glorot-uniform weights from seed 1 (same on every rank)
    import oracle as O
    theta = O.flat_params(net)
    eng.set_params(theta, 0)
    eng.sync_target()
    eng.replay_fill_synthetic(args.buffer, seed=seeds["replay"])    # per-GPU shard

    barrier = cp.barrier

    # ---- device-resident throughput (value) ------------------------------------------------------
    # (the clock sampler - an nvidia-smi child process - starts BEFORE the barrier: spawning it takes tens of milliseconds on rank 0, and
    #  ranks that entered the timed loop meanwhile would count that wait inside their first all-reduce)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        eng.train_step_async()
    eng.sync()
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        eng.train_step_async()
    ms = eng.timer_stop()
    loss, gn = eng.sync()
    barrier()
    ms_by_rank = cp.gather_over_ranks(ms / args.steps)
    ms = cp.max_over_ranks(ms)
    launches = eng.launches_per_step() * args.steps
    value = world * args.steps / (ms * 1e-3)
    if args.quick:
        if rank == 0:
            clocks.stop()
            print(json.dumps({"quick": True, "value": value, "ms_per_step": ms / args.steps, "n_gpus": world, "env": {k: v for k, v in os.environ.items() if k.startswith("DQN_")}, "loss": loss}))
        eng.close()
        cp.close()
        return

    # ---- end to end through the public call with host buffers (e2e) -------------------------------
    #   per gradient step: add_exp! of train_freq=4 fresh transitions from pinned host memory (H2D), batch_train!, read (loss, grad_norm) (D2H)
    s_h = lib._capi.pinned_empty((TRAIN_FREQ, 4, 84, 84), np.uint8)
    sp_h = lib._capi.pinned_empty((TRAIN_FREQ, 4, 84, 84), np.uint8)
    rng = np.random.default_rng(5 + rank)
    s_h[...] = rng.integers(0, 256, s_h.shape, dtype=np.uint8)
    sp_h[...] = rng.integers(0, 256, sp_h.shape, dtype=np.uint8)
    a_h = rng.integers(1, 7, TRAIN_FREQ).astype(np.int32)
    r_h = rng.uniform(-1, 1, TRAIN_FREQ).astype(np.float32)
    d_h = np.zeros(TRAIN_FREQ, np.uint8)
    td_h = np.abs(r_h)
    h2d = int(s_h.nbytes + sp_h.nbytes + a_h.nbytes + r_h.nbytes + d_h.nbytes + td_h.nbytes)
    #   (a) strictly synchronous, the reference's own loop shape: add_exp!, batch_train!, its scalars, next env steps ...
    for _ in range(args.warmup):
        eng.replay_add(s_h, a_h, r_h, sp_h, d_h, td_h)
        eng.train_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.replay_add(s_h, a_h, r_h, sp_h, d_h, td_h)
        loss, gn = eng.train_step()
    e2e_sync_s = cp.max_over_ranks(time.perf_counter() - t0)
    e2e_sync = world * args.steps / e2e_sync_s
    #   (b) the same calls one step ahead: the host adds the next transitions and launches step k+1 while step k runs, and reads
    #       step k's (loss, grad_norm) then (dqn_step_result back=1).  Same work, same order on the device, every result read.
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        eng.replay_add(s_h, a_h, r_h, sp_h, d_h, td_h)
        eng.train_step_async()
        if k:
            loss, gn = eng.step_result(1)
    loss, gn = eng.step_result(0)
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None                  # sampled every 10 ms over the device-timed region and the end-to-end regions
    e2e_s = cp.max_over_ranks(e2e_s)
    e2e = world * args.steps / e2e_s

    # ---- per-kernel timing (eager launches bracketed by CUDA events on the engine's stream) -> roofline ---
    roof, kernels = None, None
    eng.set_profiling(1)                 # every rank runs the profiled steps (they contain the collective); rank 0 reports
    for _ in range(10):
        eng.train_step_async()
    eng.sync()
    prof = eng.get_profile()
    eng.set_profiling(0)
    barrier()
    if rank == 0:
        tot = sum(k["ms"] * 1 for k in prof)
        kernels = [{"name": k["name"], "ms": round(k["ms"], 5), "share": round(k["ms"] / tot, 4)} for k in sorted(prof, key=lambda k: -k["ms"])]
        top = max(prof, key=lambda k: k["ms"])
        traffic = load_traffic()
        tr = traffic.get(top["name"], {}).get("dram_bytes")
        if top["flops"] > 0:
            tf32_peak = 0.5 * peaks["bf16_sus"]
            ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
            roof = {"kernel": top["name"], "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak, "traffic": tr,
                    "peak_note": f"dense TF32 = 1/2 of the {peaks['src']} sustained bf16 cuBLAS peak ({peaks['bf16_sus']} TFLOP/s); algorithmic FLOPs of the launch",
                    # 3xTF32 executes three tensor-core products per algorithmic product (two where A is raw bytes: the first conv layer)
                    "achieved_executed": ach * (2.0 if top["name"].startswith("conv1_fwd") or top["name"] == "conv1_wgrad" else 3.0)}
        else:
            ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            roof = {"kernel": top["name"], "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": tr,
                    "peak_note": f"{peaks['src']} HBM copy bandwidth"}
        step_flops = 26.36e9
        roof["step_tflops_algorithmic"] = step_flops * (args.steps / (ms * 1e-3)) / 1e12
        roof["traffic_source"] = (traffic.get("_file") or "") + ("@" + str(traffic.get("_commit")) if traffic.get("_commit") else "")
        coll = [k for k in prof if k["name"] in ("nccl_allreduce", "grad_allreduce")]
        if coll:                                                    # a collective is measured against NVLink, not HBM: bus bandwidth of a ring all-reduce
            b = sum(k["bytes"] for k in coll) / 2.0                 # payload bytes (the Scope records 2 x payload)
            t = sum(k["ms"] for k in coll) * 1e-3
            busbw = b * 2.0 * (world - 1) / world / t / 1e9
            kind = {1: "ncclAllReduce", 2: "peer_allreduce_kernel (own kernel over NVLink peer memory, two-shot, owner computes)"}.get(eng.collective_kind(), "?")
            roof["collective"] = {"kernel": kind, "payload_bytes": b, "busbw": busbw, "peak": 770.0, "unit": "GB/s", "frac": busbw / 770.0,
                                  "peak_note": "measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)"}
        # the HBM-bound kernels of the step, against the measured copy bandwidth (algorithmic bytes of DESIGN.md section 2)
        roof["hbm_kernels"] = [{"kernel": k["name"], "achieved": k["bytes"] / (k["ms"] * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                "frac": k["bytes"] / (k["ms"] * 1e-3) / 1e9 / peaks["hbm"], "traffic": traffic.get(k["name"], {}).get("dram_bytes")}
                               for k in prof if k["name"] in ("gather_rows", "adam")]

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.cpu_baseline import CpuBaseline
        cb = CpuBaseline(spec["layers"], spec["obs"], spec["nA"], 256, args.buffer, True, store_rows=4096)
        ts = cb.time_steps(args.cpu_steps, 3)
        cpu = {"value": 1.0 / float(np.median(ts)), "unit": UNIT, "cores": cb.threads, "kind": "port",
               "sample": f"median of {args.cpu_steps} full steps after 3 warm-ups (B=256, O(N) sampling over N={args.buffer}; obs store bounded to 4096 rows)",
               "label": "Flux-equivalent CPU restatement (torch-CPU); Julia is not installed in this image"}

    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "ms_per_step_by_rank": [round(v, 5) for v in ms_by_rank],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": workload_config(args, world),
                          "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                                  "what": f"per step: add_exp! x{TRAIN_FREQ} from pinned host memory + batch_train! + (loss, grad_norm) read back; wall clock",
                                  "mode": "one step ahead: the host adds step k+1's transitions and launches it while step k runs, then reads step k's scalars (dqn_step_result)",
                                  "sync_value": e2e_sync, "sync_mode": "strictly serial host loop: add, step, wait, read"},
                          "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
                          "last_loss": loss, "last_grad_norm": gn}))
    eng.close()
    cp.close()


if __name__ == "__main__":
    main()
