"""Batched twin of the reference's ``fast_driver`` (fast_driver.py:48-127, 201-220).

Same ``ExperimentConfig`` / ``evaluate_stream_with_stats`` / ``main`` surface, but the reference's
nested Python loops (runs x T_grid x replicates x 4 algorithm calls, fast_driver.py:87-111)
become, per horizon T, one generator launch, one FTL/FTRL launch and one SMART launch over all
``runs * replicates`` sequences.  The statistics (mean over replicates, then mean and
1.96 * std(ddof=1)/sqrt(n) over runs, fast_driver.py:113-125) are formed on the host with the
reference's own numpy expressions on the per-sequence regrets, so FTRL / FTL columns are
bit-identical to the reference and the SMART columns agree to 1e-12 (warp-parallel prefix loss).

Plotting (matplotlib, fast_driver.py:134-194) is out of scope; ``main`` writes the numbers as JSON.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Mapping, Tuple

import numpy as np
from numpy.typing import NDArray

from . import engine
from ._lib import FAM_FLIP, FAM_IID, FAM_NOISY, FAM_SWITCH
from .fast_algorithms import empirical_worst_case_thresholds
from .sequence_generation import CASES, REPLICATES_BY_TITLE, RUNS_BY_TITLE

FloatArray = NDArray[np.float64]
IntArray = NDArray[np.int64]
Stats = Dict[str, Tuple[FloatArray, FloatArray]]

ALGO_KEYS = ("FTRL", "FTL", "SMART", "EMP")
CI_Z = 1.96
SQRT2 = math.sqrt(2)


@dataclass(frozen=True)
class ExperimentConfig:
    """fast_driver.py:48-52 (the ndarray default needs default_factory on Python >= 3.11)."""
    T_grid: IntArray = field(default_factory=lambda: np.arange(100, 1100, 100, dtype=int))
    base_seed: int = 0
    g_runs: int = 1000


def _sem(x: FloatArray) -> float:
    """fast_driver.py:60-64."""
    n = x.size
    if n <= 1:
        return 0.0
    return float(np.std(x, ddof=1) / math.sqrt(n))


def _regrets_for_T(stream_builder, T: int, theta_emp: float, runs: int, replicates: int, base_seed: int,
                   compute: str = "f64") -> np.ndarray:
    """Per-sequence regrets, shape (runs, replicates, 4) in ALGO_KEYS order, for one horizon."""
    samplers = [stream_builder(run_seed=base_seed + 2025 * (run + 1)) for run in range(runs)]
    if runs <= 0 or replicates <= 0:                    # the reference's loops simply do not execute: empty statistics
        return np.zeros((max(runs, 0), max(replicates, 0), 4))
    spec = getattr(samplers[0], "spec", None)
    B = runs * replicates
    thresholds = [math.sqrt(2 * T), float(theta_emp)]
    if spec is not None and engine._lib.device_count() > 0:
        # device-generated sequences stay in HBM; nothing but the regrets crosses PCIe
        import torch
        fam, d = spec["family"], spec["d"]
        keys = None
        if fam in (FAM_IID, FAM_NOISY):
            keys = np.array([[s.spec["run_seed"], s.spec["stream0"] + rep] for s in samplers for rep in range(replicates)],
                            dtype=np.uint64)
        z, y = engine.gen_sequences_device(fam, B, T, d, keys=keys, p_flip=spec["p"], block_len=spec.get("block_len", 20),
                                           dtype=torch.float32)
        alg = engine.simulate_batch_device(z, y, [0, 1], [SQRT2, SQRT2], compute=compute)
        smart, _ = engine.smart_batch_device(z, y, thresholds, [SQRT2, SQRT2])
        # stays on the device, nothing synchronises here: the caller gathers every horizon with one copy at the end,
        # so the kernels of all horizons run back to back
        return torch.cat([alg, smart], dim=1).reshape(runs, replicates, 4)
    else:
        # foreign stream builders: sample on the host exactly as the reference does, batch the rest
        zs, ys = [], []
        for s in samplers:
            for rep in range(replicates):
                z, y, _ = s(T, rep=rep)
                zs.append(z)
                ys.append(y)
        z, y = np.stack(zs), np.stack(ys)
        alg = engine.simulate_batch(z, y, [0, 1], [SQRT2, SQRT2], compute=compute)
        smart, _ = engine.smart_batch(z, y, thresholds, [SQRT2, SQRT2])
        out = np.concatenate([alg, smart], axis=1)
    return out.reshape(runs, replicates, 4)


def evaluate_stream_with_stats(
    stream_builder: Callable[..., Callable],
    T_grid: IntArray,
    g_emp: Mapping[int, float],
    *,
    runs: int = 1,
    replicates: int = 1,
    base_seed: int = 0,
    stream_name: str = "",
    compute: str = "f64",
) -> Stats:
    """fast_driver.py:71-127 -- mean regret and 95 % CI per algorithm at each T.
    compute="f32" plays FTRL / FTL with the float32-state semantics of algorithms.py (driver.py twin)."""
    return _finish_case(_enqueue_case(stream_builder, T_grid, g_emp, runs, replicates, base_seed, compute), runs)


def _enqueue_case(stream_builder, T_grid, g_emp, runs, replicates, base_seed, compute="f64") -> list:
    """Launch every horizon of one case; returns device tensors (enqueued, not synchronised) or host arrays."""
    return [_regrets_for_T(stream_builder, int(T), float(g_emp[int(T)]), runs, replicates, base_seed, compute)
            for T in T_grid]


def _finish_case(pending: list, runs: int) -> Stats:
    """Gather the per-sequence regrets of one case and form the reference's statistics (fast_driver.py:113-125)."""
    by_T = {k: [] for k in ALGO_KEYS}
    for reg in pending:
        if not isinstance(reg, np.ndarray):
            reg = reg.cpu().numpy()                 # first call waits for the stream; the rest are already there
        for ki, k in enumerate(ALGO_KEYS):
            # fast_driver.py:113-114: per run, float(np.mean(list of replicate regrets)) -- the row-wise mean of a
            # C-contiguous (runs, replicates) array runs the same pairwise add.reduce per row, bit for bit
            # (checked against the reference's own statistics by test_fast_driver_full_default_workload)
            by_T[k].append(np.ascontiguousarray(reg[:, :, ki]).mean(axis=1).tolist())
    stats: Stats = {}
    for k in ALGO_KEYS:
        means, cis = [], []
        for vals in by_T[k]:
            arr = np.asarray(vals, dtype=float)
            mu = float(np.mean(arr)) if arr.size else 0.0
            ci = CI_Z * _sem(arr) if arr.size > 1 else 0.0
            means.append(mu)
            cis.append(ci)
        stats[k] = (np.array(means, dtype=float), np.array(cis, dtype=float))
    return stats


def run_experiment(cfg: ExperimentConfig = None) -> dict:
    """fast_driver.main (fast_driver.py:201-220) without the plotting."""
    cfg = cfg or ExperimentConfig()
    g_emp = empirical_worst_case_thresholds(cfg.T_grid, runs=cfg.g_runs, base_seed=cfg.base_seed)
    # all four cases are enqueued before the first gather, so the GPU never waits for the host-side statistics
    pending = {title: _enqueue_case(builder, cfg.T_grid, g_emp, RUNS_BY_TITLE.get(title, 1),
                                    REPLICATES_BY_TITLE.get(title, 1), cfg.base_seed)
               for title, builder in CASES.items()}
    stats_by_case: Dict[str, Stats] = {title: _finish_case(p, RUNS_BY_TITLE.get(title, 1)) for title, p in pending.items()}
    return {"T_grid": [int(t) for t in cfg.T_grid], "g_emp": g_emp, "stats_by_case": stats_by_case}


def plot_empirical_g(T_grid, g_emp, *, out_path: str = "empirical_g_T_fast.png") -> None:
    """fast_driver.py:140-158 (needs matplotlib; see plotting.py)."""
    from . import plotting
    plotting.plot_empirical_g(T_grid, g_emp, out_path=out_path)


def plot_comparisons(T_grid, stats_by_case, *, out_path: str = "algorithm_comparison_fast.png") -> None:
    """fast_driver.py:161-194 (needs matplotlib; see plotting.py)."""
    from . import plotting
    plotting.plot_comparisons(T_grid, stats_by_case, out_path=out_path, runs_by_title=RUNS_BY_TITLE, reps_by_title=REPLICATES_BY_TITLE)


def main(out_path: str = "fast_driver_b200.json", *, plots: bool = True) -> None:
    """fast_driver.main (fast_driver.py:201-220): the experiment, its JSON, and -- when matplotlib is importable --
    the reference's two figures under the reference's file names."""
    res = run_experiment()
    if plots:
        try:
            plot_empirical_g(res["T_grid"], res["g_emp"], out_path="empirical_g_T_fast.png")
            plot_comparisons(res["T_grid"], res["stats_by_case"], out_path="algorithm_comparison_fast.png")
        except ImportError as e:
            print("figures skipped:", e)
    ser = {"T_grid": res["T_grid"], "g_emp": {str(k): v for k, v in res["g_emp"].items()},
           "stats_by_case": {t: {k: {"mean": m.tolist(), "ci": c.tolist()} for k, (m, c) in st.items()}
                             for t, st in res["stats_by_case"].items()}}
    with open(out_path, "w") as f:
        json.dump(ser, f, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
