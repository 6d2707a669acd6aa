"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: run sharding and the single
all-reduce of per-variant statistics.  The per-rank partial statistics are
produced by the CPU oracle here (there is no GPU); on a B200 the same code path reduces the
kernels' outputs over NCCL."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from online_convex_optimization_b200 import dist as odist

R2 = math.sqrt(2.0)


def test_shard_runs_partition():
    for n in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            blocks = [odist.shard_runs(n, r, w) for r in range(w)]
            assert sum(c for _, c in blocks) == n
            pos = 0
            for lo, c in blocks:
                assert lo == pos
                pos += c
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        odist.shard_runs(5, 2, 2)


def _worker(rank, world, port, n_runs, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O   # checker stands in for the kernels on CPU
    lo, n = odist.shard_runs(n_runs, rank, world)
    flags, etas = [0, 1], [R2, R2]
    reg = O.regret_sweep(0, T, lo, n, 5, flags, etas, nthreads=1) if n else np.zeros((0, 2))
    st = torch.zeros((2, 4), dtype=torch.float64)
    st[:, 0] = torch.from_numpy(reg.max(axis=0)) if n else float("-inf")
    st[:, 1] = torch.from_numpy(reg.sum(axis=0))
    st[:, 2] = torch.from_numpy((reg ** 2).sum(axis=0))
    st[:, 3] = n
    odist.allreduce_stats(st)
    q.put((rank, st.numpy()))
    dist.destroy_process_group()


def test_allreduce_stats_world2_matches_single_process():
    from oracle import oracle as O
    n_runs, T, world = 37, 60, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_runs, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = O.regret_sweep(0, T, 0, n_runs, 5, [0, 1], [R2, R2], nthreads=1)
    for r in range(world):
        st = got[r]
        assert (st[:, 0] == full.max(axis=0)).all()          # MAX is order independent: bit-identical
        assert (st[:, 3] == n_runs).all()
        assert np.allclose(st[:, 1], full.sum(axis=0), rtol=1e-14)
        assert np.allclose(st[:, 2], (full ** 2).sum(axis=0), rtol=1e-14)
    curves = odist.stats_to_curves(torch.from_numpy(got[0]))
    assert np.allclose(curves["mean"].numpy(), full.mean(axis=0))
    assert np.allclose(curves["sem"].numpy(), full.std(axis=0, ddof=1) / math.sqrt(n_runs))


def _g_worker(rank, world, port, T_grid, runs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O   # checker stands in for the kernels on CPU

    def stats_fn(Tg, lo, n):
        st = torch.zeros((len(Tg), 4), dtype=torch.float64)
        for i, T in enumerate(Tg):
            reg = O.regret_sweep(0, T, lo, n, 5, [0], [R2], nthreads=1)[:, 0] if n else np.zeros(0)
            st[i] = torch.tensor([reg.max() if n else 0.0, reg.sum(), (reg ** 2).sum(), float(n)])
        return st
    g, st = odist.g_sweep_sharded(T_grid, runs, stats_fn=stats_fn)
    q.put((rank, g.numpy(), st.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("runs", [9, 1])          # runs = 1: rank 1 owns no run at all
def test_g_sweep_sharded_world2_matches_reference_semantics(runs):
    """T_grid-sharded g(T) sweep: one collective, g identical on both ranks and equal to the single-process
    empirical_worst_case_thresholds (max clamped at 0), also when a rank has an empty shard."""
    from oracle import oracle as O
    T_grid, world = [40, 75], 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + runs
    procs = [ctx.Process(target=_g_worker, args=(r, world, port, T_grid, runs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, g, st = q.get(timeout=120)
        got[r] = (g, st)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = O.empirical_worst_case_thresholds(np.array(T_grid), runs=runs, base_seed=0)
    for r in range(world):
        g, st = got[r]
        assert [float(v) for v in g] == [want[T] for T in T_grid]
        assert (st[:, 3] == runs).all()
    assert (got[0][1] == got[1][1]).all()           # bit-identical statistics on every rank
