"""Multi-GPU edge sharding: one process per GPU, edges split contiguously, worlds and robot
tables replicated, one small all-gather of the per-edge summary {first_hit: i32, cost: f32}.

The path has no other exchange step: units are independent and every reduction (over steps, over worlds)
stays inside one warp, so NCCL is used for exactly one collective per batch.
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_edges: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank; shards differ by at most one edge."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(n_edges, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_capacity(n_edges: int, world_size: int) -> int:
    """Rows every rank contributes to the all-gather (largest shard)."""
    return (n_edges + world_size - 1) // world_size


def pack_summary(first_hit, n_steps, cost, capacity: int):
    """[capacity, 3] int32: first_hit, n_steps, bit pattern of cost (float32); padded rows are -1."""
    import torch
    n = first_hit.shape[0]
    buf = torch.full((capacity, 3), -1, dtype=torch.int32, device=first_hit.device)
    buf[:n, 0] = first_hit.to(torch.int32)
    buf[:n, 1] = n_steps.to(torch.int32)
    buf[:n, 2] = cost.to(torch.float32).contiguous().view(torch.int32)
    return buf


def unpack_summary(gathered, n_edges: int, world_size: int):
    """Inverse of the gather: returns (first_hit[E], n_steps[E], cost[E]) in global edge order."""
    import torch
    cap = gathered.shape[0] // world_size
    parts = []
    for r in range(world_size):
        lo, hi = shard_bounds(n_edges, r, world_size)
        parts.append(gathered[r * cap: r * cap + (hi - lo)])
    allrows = torch.cat(parts, dim=0)
    return allrows[:, 0].contiguous(), allrows[:, 1].contiguous(), allrows[:, 2].contiguous().view(torch.float32)


def all_gather_summary(first_hit, n_steps, cost, n_edges: int, group=None, out=None):
    """All ranks receive every edge's (first_hit, n_steps, cost).  One collective."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    cap = shard_capacity(n_edges, ws)
    mine = pack_summary(first_hit, n_steps, cost, cap)
    if out is None:
        out = torch.empty((cap * ws, 3), dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return unpack_summary(out, n_edges, ws)


class ShardedEdgeChecker:
    """Wraps an EdgeChecker: validates this rank's shard of a replicated edge batch and gathers the summary."""

    def __init__(self, checker, group=None):
        import torch.distributed as dist
        self.checker = checker
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1

    def validate_edges(self, q_from, q_to, backward: bool = False):
        """q_from / q_to: the FULL batch [E, D] (replicated on every rank).  Returns
        (local EdgeBatchResult, (first_hit[E], n_steps[E], cost[E]) gathered over ranks)."""
        E = q_from.shape[0]
        lo, hi = shard_bounds(E, self.rank, self.world_size)
        # edge_base = lo: the on-device gate hashes the GLOBAL edge index, so sharding does not change its decisions
        local = self.checker.validate_edges(q_from[lo:hi], q_to[lo:hi], backward=backward, edge_base=lo)
        if self.world_size == 1:
            return local, (local.first_hit, local.n_steps, local.cost)
        return local, all_gather_summary(local.first_hit, local.n_steps, local.cost, E, self.group)
