"""Batched twin of the reference's ``driver`` (driver.py:47-129, 204-226): identical to
``fast_driver`` except that it imports the float32-state ``algorithms`` module (driver.py:22) --
here: the FTRL / FTL columns use the engine's float32 compute mode, g(T) comes from
``algorithms.empirical_worst_case_thresholds``.  Plotting is out of scope."""
from __future__ import annotations

from typing import Dict

from . import fast_driver as _fd
from .algorithms import empirical_worst_case_thresholds
from .sequence_generation import CASES, REPLICATES_BY_TITLE, RUNS_BY_TITLE

ExperimentConfig = _fd.ExperimentConfig
ALGO_KEYS, CI_Z, Stats, _sem = _fd.ALGO_KEYS, _fd.CI_Z, _fd.Stats, _fd._sem


def evaluate_stream_with_stats(stream_builder, T_grid, g_emp, *, runs: int = 1, replicates: int = 1, base_seed: int = 0,
                               stream_name: str = "") -> Stats:
    """driver.py:70-129."""
    return _fd.evaluate_stream_with_stats(stream_builder, T_grid, g_emp, runs=runs, replicates=replicates,
                                          base_seed=base_seed, stream_name=stream_name, compute="f32")


def run_experiment(cfg: ExperimentConfig = None) -> dict:
    """driver.main (driver.py:204-226) without the plotting."""
    cfg = cfg or ExperimentConfig()
    g_emp = empirical_worst_case_thresholds(cfg.T_grid, runs=cfg.g_runs, base_seed=cfg.base_seed)
    stats_by_case: Dict[str, Stats] = {}
    for title, builder in CASES.items():
        stats_by_case[title] = evaluate_stream_with_stats(builder, cfg.T_grid, g_emp, runs=RUNS_BY_TITLE.get(title, 1),
                                                          replicates=REPLICATES_BY_TITLE.get(title, 1), base_seed=cfg.base_seed,
                                                          stream_name=title)
    return {"T_grid": [int(t) for t in cfg.T_grid], "g_emp": g_emp, "stats_by_case": stats_by_case}
