"""Multi-GPU sharding of seeded sweeps: every trajectory is independent and
every stream is keyed by (base_seed, T, run), so ranks take contiguous blocks of runs with no
data-path communication; the only exchange is one all-reduce of the per-variant statistics
(MAX for the g(T) maximum -- order independent, bit-identical for any GPU count; SUM for
sum / sum of squares / count).  One process per GPU, torch.distributed (NCCL over NVLink on
GPUs; gloo on CPU tensors for the host-logic tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_runs(n_runs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, lo + n) of rank `rank` out of n_runs (remainder to the low ranks)."""
    if world < 1 or not (0 <= rank < world) or n_runs < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_runs, world)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def allreduce_stats(stats, group=None):
    """In-place all-reduce of a (..., 4) tensor of (max, sum, sumsq, count) rows."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    mx = stats[..., 0].contiguous()
    sm = stats[..., 1:].contiguous()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
    stats[..., 0] = mx
    stats[..., 1:] = sm
    return stats


def stats_to_curves(stats):
    """(max, sum, sumsq, count) -> dict(max, mean, sem) as in fast_driver._sem (ddof = 1)."""
    import torch
    mx, s1, s2, n = stats[..., 0], stats[..., 1], stats[..., 2], stats[..., 3]
    mean = s1 / torch.clamp(n, min=1)
    var = torch.clamp(s2 - n * mean * mean, min=0) / torch.clamp(n - 1, min=1)
    sem = torch.sqrt(var / torch.clamp(n, min=1))
    return {"max": mx, "mean": mean, "sem": sem}


def regret_sweep_sharded(family, n_runs: int, T: int, d: int, alg_flags, eta0s, *, key_a: int = 0, run_lo: int = 0,
                         want_regret: bool = False, group=None):
    """This rank's block of a seeded sweep + the all-reduced (A,4) statistics (CUDA tensors)."""
    import torch.distributed as dist
    from . import engine
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, n = shard_runs(n_runs, rank, world)
    reg, stats = engine.regret_sweep_device(family, n, T, d, alg_flags, eta0s, key_a=key_a, key_c0=run_lo + lo,
                                            want_regret=want_regret)
    if n == 0:
        stats[..., 0] = float("-inf")
    allreduce_stats(stats, group)
    return reg, stats
