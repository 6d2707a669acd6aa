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
    """In-place reduction over ranks of a (..., 4) tensor of (max, sum, sumsq, count) rows: MAX for column 0, SUM
    for the rest.  ONE collective: every rank gathers all ranks' rows (a few hundred bytes each) and folds them
    locally in rank order, so the sums are bit-identical on every rank and from run to run."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    world = dist.get_world_size(group)
    mine = stats.contiguous()
    if mine.is_cuda and dist.get_backend(group) == "gloo":
        mine = mine.cpu()           # several ranks sharing one GPU (tests on a 1-GPU box): gloo moves host tensors
    allst = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allst, mine, group=group)
    out = allst[0].clone()
    for r in range(1, world):                       # fixed order: deterministic sums
        out[..., 0] = torch.maximum(out[..., 0], allst[r][..., 0])
        out[..., 1:] = out[..., 1:] + allst[r][..., 1:]
    stats.copy_(out)             # (host -> device when the collective ran on host copies)
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


def _device_g_stats(T_grid, lo, n, base_seed, d, eta0, comparator):
    """This rank's (nT, 4) statistics of FTRL(eta0) regrets on runs [lo, lo + n): one fused launch per T, all
    enqueued on the current stream before anything is read back."""
    import torch
    from . import engine
    import numpy as np
    stats = torch.empty((len(T_grid), 4), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    for i, T in enumerate(T_grid):
        T = int(T)
        if comparator == "exact_max" and 2 <= d <= 8 and n > 0 and T > 0:
            # exact maximum at single-pass speed: the single-pass regret bounds the literal one from above for every run,
            # so only the runs whose bound reaches this rank's largest bound take the literal second pass (engine.g_sweep)
            reg, st = engine.regret_sweep_device("gpath", n, T, d, [0], [eta0], key_a=base_seed, key_c0=lo, comparator="linear")
            margin = 1e-9 * T + 2.3e-16 * T * T
            cand = torch.nonzero(reg[:, 0] >= st[0, 0] - margin).flatten()
            keys = np.stack([np.full(cand.numel(), base_seed, dtype=np.uint64), (lo + cand.cpu().numpy()).astype(np.uint64)], axis=1)
            z, y = engine.gen_sequences_device("gpath", keys.shape[0], T, d, keys=keys, dtype=torch.float64)
            exact = engine.simulate_batch_device(z, y, [0], [eta0])
            st = st.clone(); st[0, 0] = exact.max()
        else:
            _, st = engine.regret_sweep_device("gpath", n, T, d, [0], [eta0], key_a=base_seed, key_c0=lo, want_regret=False,
                                               comparator="two_pass" if comparator == "exact_max" else comparator)
        stats[i] = st[0]
    return stats


def g_sweep_sharded(T_grid, runs: int, *, base_seed: int = 0, d: int = 5, eta0: float = 2.0 ** 0.5,
                    comparator: str = "exact_max", group=None, stats_fn=None):
    """fast_algorithms.empirical_worst_case_thresholds (fast_algorithms.py:211-247) with the `runs` sharded over the
    ranks of `group` (contiguous blocks, every T of a run on the same rank) and ONE collective over the (nT, 4)
    per-T statistics.  Returns (g (nT,) = max(0, max_r regret), stats (nT, 4) = max, sum, sumsq, count) on every rank;
    g is bit-identical for any number of ranks (MAX is order independent) and, with the default comparator="exact_max"
    (or "two_pass"), bit-identical to the reference's value.  `stats_fn(T_grid, lo, n)` replaces the
    device sweep in the CPU (gloo) tests of this host logic."""
    import torch
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    lo, n = shard_runs(int(runs), rank, world)
    T_grid = [int(T) for T in T_grid]
    if stats_fn is not None:
        stats = stats_fn(T_grid, lo, n)
    else:
        stats = _device_g_stats(T_grid, lo, n, int(base_seed), d, float(eta0), comparator)
    if n == 0:
        stats[..., 0] = float("-inf")
    allreduce_stats(stats, group)
    g = torch.clamp(stats[..., 0], min=0.0)         # max_regret starts at 0.0 (fast_algorithms.py:228)
    return g, stats
