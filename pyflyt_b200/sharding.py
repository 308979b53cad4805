"""Env-batch sharding over ranks: envs never interact, so rank r simply owns the global
env ids [r * per_rank, (r + 1) * per_rank) and there is NO collective on the step path.  The Philox
streams are keyed by GLOBAL env id (``env_offset``), so a trajectory does not depend on the world size.
Only end-of-run statistics cross ranks (one all-reduce of a few scalars)."""

from __future__ import annotations

from typing import Sequence


def shard_range(global_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first ``global_envs % world_size`` ranks get one extra env."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(int(global_envs), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_step_stats(values: Sequence[float], op: str = "max", device=None):
    """All-reduce a few scalars (timings, counters) over the default process group; identity if the
    process group is not initialised.  Returns a python list."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return t.tolist()
