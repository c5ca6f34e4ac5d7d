"""Multi-GPU sharding of a batch of independent MPC instances.

Every ``(p, u0)`` pair is independent, so the solve path has NO collective: each rank
(one process per GPU) solves a contiguous slice of scenarios with all of their multi-start
guesses (the best-of-starts argmin stays local), and only the results are gathered.
The gather uses ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple


def shard_range(n_scenarios: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous scenario slice ``[lo, hi)`` of ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_scenarios, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def best_of_starts(cost, u, exit_status, starts: int):
    """Per scenario, the start with the lowest cost (ties: lowest index); NaN costs lose.

    ``cost`` [n*starts], ``u`` [n*starts, 2N], ``exit_status`` [n*starts] (torch tensors).
    Returns ``(best_cost [n], best_u [n, 2N], best_status [n], best_index [n])``.
    """
    import torch
    n = cost.shape[0] // starts
    c = cost.reshape(n, starts)
    c = torch.where(torch.isnan(c), torch.full_like(c, float("inf")), c)
    idx = torch.argmin(c, dim=1)
    flat = torch.arange(n, device=cost.device) * starts + idx
    return cost[flat], u[flat], exit_status[flat], idx


def gather_results(local: Dict[str, "object"], n_scenarios: int, starts_kept: int = 1):
    """All-gather per-rank result tensors (first dim = local scenarios * starts_kept) into the
    global scenario order.  Ranks may hold different counts (uneven shards)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    out = {}
    for key, t in local.items():
        parts = []
        for r in range(world):
            lo, hi = shard_range(n_scenarios, r, world)
            shape = ((hi - lo) * starts_kept,) + tuple(t.shape[1:])
            parts.append(torch.empty(shape, dtype=t.dtype, device=t.device))
        lo, hi = shard_range(n_scenarios, rank, world)
        if t.shape[0] != (hi - lo) * starts_kept:
            raise ValueError(f"{key}: local rows {t.shape[0]} != shard size {(hi - lo) * starts_kept}")
        # all_gather needs equal shapes: pad to the largest shard, trim after
        mx = max(p.shape[0] for p in parts)
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        out[key] = torch.cat([b[: p.shape[0]] for b, p in zip(bufs, parts)], dim=0)
    return out
