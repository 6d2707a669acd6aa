"""Batched twin of the reference's ``exact_ftl_driver`` (exact_ftl_driver.py:41-206, 268-293):
g(T) from FTRL against the exact comparator, and FTRL vs "FTL (exact)" regret per sequence
family.  Per horizon T all ``runs * replicates`` sequences go through one exact-FTL prefix launch
(the reference performs T conic solves per sequence) and one FTRL launch with the exact
comparator action.  Plotting is out of scope; ``main`` writes JSON.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Mapping, Sequence, Tuple

import numpy as np
from numpy.typing import NDArray

from . import engine
from ._lib import FAM_IID, FAM_NOISY
from .exact_ftl import ExactFTLNoClip
from .fast_algorithms import _rng  # noqa: F401  (exact_ftl_driver.py:23 imports it)
from .sequence_generation import CASES, REPLICATES_BY_TITLE, RUNS_BY_TITLE

FloatArray = NDArray[np.float64]
IntArray = NDArray[np.int64]
Stats = Dict[str, Tuple[FloatArray, FloatArray]]
CI_Z = 1.96
SQRT2 = math.sqrt(2)


@dataclass(frozen=True)
class ExperimentConfig:
    """exact_ftl_driver.py:41-49."""
    T_grid: IntArray = field(default_factory=lambda: np.arange(100, 1100, 100, dtype=np.int64))
    base_seed: int = 0
    g_runs: int = 200
    norm: str = "l2"
    solver: str | None = None
    solver_opts: Mapping[str, object] | None = None
    reuse_grid: Sequence[int] = (1,)


def _sem(x: FloatArray) -> float:
    n = x.size
    if n <= 1:
        return 0.0
    return float(np.std(x, ddof=1) / math.sqrt(n))


def _exact_and_ftrl(z: np.ndarray, y: np.ndarray, *, norm: str = "l2", solver_opts: Mapping[str, object] | None = None,
                    only_comparator: bool = False):
    """(FTRL regret vs exact comparator, exact-FTL regret) for a batch z (B,T,d), y (B,T) on the unit `norm` ball.
    l2 in the linear regime: closed form for every prefix; anything else (linf / l1 balls, rows outside the ball,
    real-valued labels): the general interior-point solver with solver_opts {"tol", "max_iter"} -- every prefix, or
    only the full sequence when only the comparator is needed (exact_ftl_driver.py:104-112 solves just that one)."""
    z = np.ascontiguousarray(z, dtype=np.float64)      # exact_ftl_driver.py:158-159
    y = np.ascontiguousarray(y, dtype=np.float64)
    opts = dict(solver_opts or {})
    ex = None
    if norm == "l2" or z.shape[2] == 1:                # for d = 1 the three unit balls coincide
        ex = engine.exact_ftl_prefix_batch(z, y, want_actions=True)
        if ex["status"].any():
            ex = None
    if ex is None:
        if z.shape[2] > 32 or (norm == "l1" and z.shape[2] > 16):
            raise RuntimeError("Exact FTL (no clip) failed: status=outside_linear_regime (general solver covers d <= 32; l1: d <= 16)")
        ex = engine.exact_ftl_general_batch(z, y, norm=norm, tol=float(opts.get("tol", 1e-8)), max_iter=int(opts.get("max_iter", 300)),
                                            first_prefix=z.shape[1] if only_comparator else 0)
        if ex["status"].any():
            raise RuntimeError(f"Exact FTL (no clip) failed: status=max_iter_reached ({int(ex['status'].sum())} prefixes)")
    ftrl = engine.simulate_batch(z, y, [0], [SQRT2], comparator_action=ex["actions"][:, -1, :])
    return ftrl[:, 0], ex["regret"]


def empirical_worst_case_thresholds(T_grid: IntArray, *, runs: int, base_seed: int, norm: str = "l2",
                                    solver: str | None = None, solver_opts: Mapping[str, object] | None = None) -> Dict[int, float]:
    """exact_ftl_driver.py:64-117 -- g(T) = max(0, max_r FTRL regret vs the exact comparator), d = 5."""
    ExactFTLNoClip(d=5, T_max=1, norm=norm, solver=solver, solver_opts=None if solver_opts is None else dict(solver_opts))
    g_emp: Dict[int, float] = {}
    for T_val in T_grid:
        T = int(T_val)
        max_regret = 0.0
        if runs > 0:
            z, y = engine.gen_sequences("gpath", runs, T, 5, key_a=base_seed, key_c0=0)
            ftrl, _ = _exact_and_ftrl(z, y, norm=norm, solver_opts=solver_opts, only_comparator=True)
            for reg in ftrl:
                if reg > max_regret:
                    max_regret = float(reg)
        g_emp[T] = max_regret
    return g_emp


def evaluate_stream_with_stats(stream_builder: Callable[..., Callable], T_grid: IntArray, *, runs: int, replicates: int,
                               cfg: ExperimentConfig, stream_name: str) -> Stats:
    """exact_ftl_driver.py:120-206."""
    algo_labels = ["FTRL", "FTL (exact)"]
    ExactFTLNoClip(d=5, T_max=1, norm=cfg.norm, solver=cfg.solver)      # validates norm like the reference's constructor
    by_T = {label: [] for label in algo_labels}
    samplers = [stream_builder(run_seed=cfg.base_seed + 2025 * (run + 1)) for run in range(runs)]
    if runs <= 0 or replicates <= 0:                 # the reference returns empty statistics (means 0, CIs 0)
        z0 = np.zeros(len(T_grid), dtype=float)
        return {label: (z0.copy(), z0.copy()) for label in algo_labels}
    for T in T_grid:
        T_int = int(T)
        spec = getattr(samplers[0], "spec", None) if samplers else None
        if spec is not None:
            # the twin's own stream builders: one generator launch for all runs x replicates of this horizon
            keys = None
            if spec["family"] in (FAM_IID, FAM_NOISY):
                keys = np.array([[s.spec["run_seed"], s.spec["stream0"] + rep] for s in samplers for rep in range(replicates)],
                                dtype=np.uint64)
            zb, yb = engine.gen_sequences(spec["family"], runs * replicates, T_int, spec["d"], keys=keys, p_flip=spec["p"],
                                          block_len=spec.get("block_len", 20), dtype=np.float32)
        else:
            # foreign stream builders: sample on the host exactly as the reference does
            zs, ys = [], []
            for s in samplers:
                for rep in range(replicates):
                    z, y, _ = s(T_int, rep=rep)
                    zs.append(z)
                    ys.append(y)
            zb, yb = np.stack(zs), np.stack(ys)
        ftrl, ftl = _exact_and_ftrl(zb, yb, norm=cfg.norm, solver_opts=cfg.solver_opts)
        vals = {"FTRL": ftrl.reshape(runs, replicates), "FTL (exact)": ftl.reshape(runs, replicates)}
        for label in algo_labels:
            by_T[label].append([float(np.mean([float(v) for v in vals[label][run]])) for run in range(runs)])
    stats: Stats = {}
    for label in algo_labels:
        means, cis = [], []
        for v in by_T[label]:
            arr = np.asarray(v, dtype=float)
            means.append(float(np.mean(arr)) if arr.size else 0.0)
            cis.append(CI_Z * _sem(arr) if arr.size > 1 else 0.0)
        stats[label] = (np.array(means, dtype=float), np.array(cis, dtype=float))
    return stats


def run_experiment(cfg: ExperimentConfig = None) -> dict:
    cfg = cfg or ExperimentConfig()
    g_emp = empirical_worst_case_thresholds(cfg.T_grid, runs=cfg.g_runs, base_seed=cfg.base_seed, norm=cfg.norm,
                                            solver=cfg.solver, solver_opts=cfg.solver_opts)
    stats_by_case = {title: evaluate_stream_with_stats(builder, cfg.T_grid, runs=RUNS_BY_TITLE.get(title, 1),
                                                       replicates=REPLICATES_BY_TITLE.get(title, 1), cfg=cfg, stream_name=title)
                     for title, builder in CASES.items()}
    return {"T_grid": [int(t) for t in cfg.T_grid], "g_emp": g_emp, "stats_by_case": stats_by_case}


def main(out_path: str = "exact_ftl_driver_b200.json") -> None:
    res = run_experiment()
    ser = {"T_grid": res["T_grid"], "g_emp": {str(k): v for k, v in res["g_emp"].items()},
           "stats_by_case": {t: {k: {"mean": m.tolist(), "ci": c.tolist()} for k, (m, c) in st.items()}
                             for t, st in res["stats_by_case"].items()}}
    with open(out_path, "w") as f:
        json.dump(ser, f, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
