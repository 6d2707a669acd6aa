"""Multi-rank tests (-m gpu): a g(T)-style sweep sharded over 2 ranks gives the single-GPU statistics -- MAX
bit-identical, SUM within 1e-14.  With >= 2 visible GPUs the ranks use one GPU each and NCCL; on a
1-GPU box the same sharding logic runs with both ranks on cuda:0 and the one collective over gloo (host copies of the
few hundred bytes of statistics), so the sharded path has a test on every box."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
R2 = math.sqrt(2.0)


def _worker(rank, world, port, n_runs, T, q, backend="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from online_convex_optimization_b200 import dist as odist
    reg, stats = odist.regret_sweep_sharded("gpath", n_runs, T, 5, [0, 1], [R2, R2], key_a=0, run_lo=0, want_regret=True)
    g = {}
    for mode in ("two_pass", "linear", "exact_max"):     # T_grid-sharded g(T) sweep: one collective over the (nT, 4) curve
        gg, gst = odist.g_sweep_sharded([100, 250], n_runs, base_seed=0, comparator=mode)
        g[mode] = (gg.cpu().numpy(), gst.cpu().numpy())
    q.put((rank, stats.cpu().numpy(), reg.cpu().numpy(), g))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("backend", ["nccl", "gloo_one_gpu"])
def test_sharded_sweep_equals_single_gpu(backend):
    if backend == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    from online_convex_optimization_b200 import engine
    n_runs, T, world = 1001, 300, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000) + (0 if backend == "nccl" else 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_runs, T, q, "nccl" if backend == "nccl" else "gloo")) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, st, reg, g = q.get(timeout=300)
        got[r] = (st, reg, g)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = engine.regret_sweep("gpath", n_runs, T, 5, [0, 1], [R2, R2], key_a=0, key_c0=0)
    full = np.concatenate([got[0][1], got[1][1]], axis=0)
    assert (full.view(np.uint64) == single["regret"].view(np.uint64)).all()       # same trajectories on any rank
    for r in range(world):
        st = got[r][0]
        assert (st[:, 0] == single["stats"][:, 0]).all()                              # MAX: bit-identical
        assert (st[:, 3] == n_runs).all()
        assert np.allclose(st[:, 1], single["stats"][:, 1], rtol=1e-14)
        assert np.allclose(st[:, 2], single["stats"][:, 2], rtol=1e-14)
    # g(T) sweep sharded over the two ranks == the single-GPU sweep (MAX bit-identical), both comparator modes
    for mode in ("two_pass", "linear", "exact_max"):
        g1, st1 = engine.g_sweep(np.array([100, 250]), runs=n_runs, base_seed=0, comparator=mode)
        for r in range(world):
            gg, gst = got[r][2][mode]
            assert (gg == g1).all() and (gst[:, 0] == st1[:, 0]).all() and (gst[:, 3] == n_runs).all()
            assert np.allclose(gst[:, 1], st1[:, 1], rtol=1e-13)
    assert (got[0][2]["exact_max"][0] == got[0][2]["two_pass"][0]).all()      # the pruned maximum IS the literal maximum
