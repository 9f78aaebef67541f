"""world_size-2 gloo tests (CPU) of the host-side logic of the N>1 path: id broadcast, barrier, max over ranks, and the
rule that makes data parallelism equal one big batch (mean of per-rank gradients scaled by 1/(B*world))."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import importlib.util
    spec = importlib.util.spec_from_file_location("dqn_dist", os.path.join(ROOT, "deepqlearning.jl_b200", "dist.py"))
    d = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(d)
    cp = d.ControlPlane(world)
    assert (cp.rank, cp.world) == (rank, world)
    ident = bytes(range(128)) if rank == 0 else None
    got = cp.broadcast_bytes(ident, 128)
    assert got == bytes(range(128))
    cp.barrier()
    assert cp.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    assert cp.gather_over_ranks(3.0 * rank) == [3.0 * r for r in range(world)]
    seeds = d.shard_seeds(0, rank)
    assert seeds["weights"] == 1 and seeds["replay"] == 1000 + rank

    # data-parallel equivalence on the oracle: sum over ranks of grads of loss/(B*world) == grad of one batch of B*world
    import oracle as O
    import util
    spec_ = util.SPECS["testmdp"]
    net = util.make_oracle_net(spec_, True, seed=21)
    tgt = util.perturbed_copy(net, seed=22)
    B = 8
    s, a, r, sp, done = util.random_transitions(spec_, B * world, seed=5)
    w = np.random.default_rng(6).uniform(0.5, 2.0, B * world).astype(np.float32)
    full = O.forward_backward(net, tgt, s, a - 1, r, sp, done.astype(np.float32), w, 0.99, True, np.float64)
    sl = slice(rank * B, (rank + 1) * B)
    part = O.forward_backward(net, tgt, s[sl], a[sl] - 1, r[sl], sp[sl], done[sl].astype(np.float32), w[sl], 0.99, True, np.float64)
    import torch
    g = torch.tensor(np.concatenate([x.ravel() for x in part["grads"]]) / world)     # each rank scales by 1/world (loss/(B*world))
    cp.dist.all_reduce(g)
    gfull = np.concatenate([x.ravel() for x in full["grads"]])
    np.testing.assert_allclose(g.numpy(), gfull, rtol=1e-10, atol=1e-14)
    out.put((rank, True))
    cp.close()


def test_control_plane_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)]


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference --gpus 2` launched as the driver launches it (torchrun, 2 ranks, no GPU needed): rank 0 alone times the
    CPU restatement and prints ONE JSON line with the contract's keys; the other rank exits 0 without work."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "3", "--buffer", "4096"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "dqn_gradient_steps_per_sec" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "configs[2]" in d["config"]["workload"]
