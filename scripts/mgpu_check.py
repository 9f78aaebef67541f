"""Run under torchrun on >= 2 GPUs: data-parallel step parity.  Every rank holds different transitions; after ONE
step the all-reduced gradient must equal the oracle's gradient of the combined batch of B*world samples, and the
updated parameters must be identical on all ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dqn_b200 as lib       # noqa: E402
import oracle as O           # noqa: E402
import util                  # noqa: E402

cp = lib.ControlPlane(2)
rank, world = cp.rank, cp.world
ok = True
finals = {}
for name, mode, ar in (("testmdp", 0, "peer"), ("conv_small", 0, "peer"), ("conv_small", 1, "peer"), ("conv_small", 1, "nccl")):
    # ar: the engine's own all-reduce over NVLink peer memory (default when the ranks can map each other's memory) or NCCL (DQN_PEER_AR=0)
    os.environ["DQN_PEER_AR"] = "1" if ar == "peer" else "0"
    spec = util.SPECS[name]
    B = spec["B"]
    nccl_id = cp.broadcast_bytes(lib.nccl_unique_id() if rank == 0 else None, 128)     # a ncclUniqueId is single-use: one per communicator
    net = util.make_oracle_net(spec, True, seed=21)
    tgt = util.perturbed_copy(net, seed=22)
    cfg = lib.make_config(util.layer_descs(spec), tuple(reversed(spec["obs"])), spec["nA"], obs_dtype="u8" if spec["u8"] else "f32",
                          batch_size=B, buffer_size=spec["N"], learning_rate=spec["lr"], discount=0.99, seed=2, device=cp.local_rank,
                          math_mode=mode, rank=rank, world=world, nccl_id=nccl_id)
    eng = lib.Engine(cfg)
    eng.set_params(O.flat_params(net), 0)
    eng.set_params(O.flat_params(tgt), 1)
    data = [util.random_transitions(spec, 100, seed=50 + r) for r in range(world)]
    s, a, r, sp, done = data[rank]
    eng.replay_add(s, a, r, sp, done, np.abs(r))
    idx = np.arange(B, dtype=np.int64)
    loss, gn = eng.train_step_with_indices(idx)
    g = eng.grads()
    # oracle: combined batch; IS weights per shard (each shard's own n and sum p)
    S, A, R, SP, D, W = [], [], [], [], [], []
    for rr in range(world):
        buf = util.make_oracle_replay(spec)
        s_, a_, r_, sp_, d_ = data[rr]
        buf.add_batch(s_, a_, r_, sp_, d_, np.abs(r_))
        sb, ab, rb, spb, db, _, w = buf.get_batch(idx, total="tree", dequant=util.dequant)
        S.append(sb); A.append(ab); R.append(rb); SP.append(spb); D.append(db); W.append(w)
    out = O.forward_backward(net, tgt, np.concatenate(S), np.concatenate(A) - 1, np.concatenate(R), np.concatenate(SP), np.concatenate(D),
                             np.concatenate(W), 0.99, True, np.float64)
    gref = np.concatenate([x.ravel() for x in out["grads"]])
    err = float(np.abs(g - gref).max() / np.abs(gref).max())
    theta = eng.get_params(0)
    tsum = cp.sum_over_ranks(float(np.abs(theta).sum()))
    same = abs(tsum / world - float(np.abs(theta).sum())) <= 1e-9 * tsum
    gmax = cp.max_over_ranks(gn)
    kind = eng.collective_kind()
    ok = ok and kind == (2 if ar == "peer" else 1)
    for _ in range(6):                                   # more steps through the captured graph (two buckets per step, fresh epochs every step)
        eng.train_step()
    th2 = eng.get_params(0)
    t2 = cp.sum_over_ranks(float(np.abs(th2).sum()))
    same = same and abs(t2 / world - float(np.abs(th2).sum())) <= 1e-9 * t2 and bool(np.isfinite(th2).all())
    finals[(name, mode, ar)] = th2
    if ar == "nccl" and world == 2:                      # a sum of two terms has one rounding whichever way it is ordered: both reductions agree exactly
        same = same and np.array_equal(th2, finals[(name, mode, "peer")])
    print(f"[rank {rank}] {name} mode {mode} all-reduce {ar} (kind {kind}): grad rel err vs combined-batch oracle {err:.2e}, params identical across ranks {same}, grad_norm {gn:.6e} (max {gmax:.6e})", flush=True)
    ok = ok and err < 2e-4 and same and gmax == gn
    eng.close()
cp.barrier()
print(f"[rank {rank}] MGPU {'OK' if ok else 'FAILED'}", flush=True)
cp.close()
sys.exit(0 if ok else 1)
