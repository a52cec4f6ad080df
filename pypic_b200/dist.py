"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch; gloo for the CPU tests of the host logic).

Particles interact only through the grid, so the decomposition is by particles
(SURVEY.md 8e): contiguous index ranges per rank, ONE fp64 all-reduce of the grid
accumulators per deposit (the sheath packs [jh | j1 | absorbed counts] into a single
message per Picard iteration), and the field update replicated on every rank --
NCCL delivers identical bits to all ranks, so every rank takes the same Picard exit.
There is no other data-path collective.
"""
import torch
import torch.distributed as dist


def shard_range(N, rank, world):
    """Contiguous [start, stop) of rank's particles; sizes differ by at most one and
    concatenating the shards in rank order restores the global index order."""
    base, rem = divmod(int(N), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_split(n_split_global, start, stop):
    """Local index where species 2 begins inside the shard [start, stop)."""
    return int(min(max(n_split_global - start, 0), stop - start))


class Comm:
    """Thin wrapper so the simulation objects work with or without a process group."""

    def __init__(self, group=None, enabled=True):
        self.enabled = bool(enabled) and dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    def allreduce_sum(self, t):
        """In-place sum over ranks of the grid accumulators."""
        if self.enabled and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather_int(self, v, device="cpu"):
        if not (self.enabled and self.world > 1):
            return [int(v)]
        t = torch.tensor([int(v)], dtype=torch.int64, device=device)
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [int(o.item()) for o in out]

    def max_float(self, v, device="cpu"):
        if not (self.enabled and self.world > 1):
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        if self.enabled and self.world > 1:
            dist.barrier(group=self.group)
