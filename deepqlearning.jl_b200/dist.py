"""Control plane for the data-parallel group (one process per GPU).  Only three things cross it: the 128-byte
ncclUniqueId (rank 0 -> all), barriers, and max-over-ranks of timings.  The data path (ONE gradient all-reduce per
step) is NCCL inside the engine; this module uses torch.distributed's gloo backend so that it also runs on CPU-only
hosts (tests/test_dist_cpu.py, world_size 2)."""
import os


class ControlPlane:
    def __init__(self, world_hint=1):
        self.dist = None
        self.rank, self.world, self.local_rank = 0, 1, 0
        if world_hint > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
            if "RANK" in os.environ:
                import torch.distributed as dist
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29533")
                if not dist.is_initialized():
                    dist.init_process_group("gloo")
                self.dist = dist
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
                self.local_rank = int(os.environ.get("LOCAL_RANK", self.rank))

    def broadcast_bytes(self, payload, nbytes, src=0):
        """rank `src` passes `payload` (bytes of length nbytes); every rank gets the bytes back."""
        if self.dist is None:
            return payload
        import torch
        t = torch.zeros(nbytes, dtype=torch.uint8)
        if self.rank == src:
            t = torch.tensor(list(payload), dtype=torch.uint8)
        self.dist.broadcast(t, src)
        return bytes(t.tolist())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])

    def gather_over_ranks(self, x):
        """list of every rank's value, in rank order, on every rank"""
        if self.dist is None:
            return [float(x)]
        import torch
        t = torch.zeros(self.world, dtype=torch.float64)
        t[self.rank] = float(x)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()


def shard_seeds(base_seed, rank):
    """Per-rank seeds: every rank owns a private replay shard and sampler stream; the initial weights are shared."""
    return dict(replay=base_seed + 1000 + rank, sampler=base_seed + 2 + rank, weights=base_seed + 1)
