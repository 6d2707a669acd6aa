"""Drop-in twin of the reference's ``fast_algorithms`` module (same names, signatures and
return types; reference file:line cited per function).  ``fast_driver.py:23-28`` imports
exactly these names, so switching engines is that one import line:

    from online_convex_optimization_b200.fast_algorithms import (
        empirical_worst_case_thresholds, simulate_alg, simulate_SMART, simulate_empirical_g_SMART)

Single calls go through the batched C ABI with B = 1 (launch-latency bound; they exist for
signature compatibility and parity tests -- throughput comes from ``engine`` / ``fast_driver``).
No CPU fallback: a call without a CUDA device raises RuntimeError.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from . import engine


def simulate_alg(z: np.ndarray, y: np.ndarray, alg_flag: int, eta0: float) -> float:
    """fast_algorithms.py:171-177 -- regret of FTRL (alg_flag == 0) or FTL on one sequence."""
    return float(engine.simulate_batch(z, y, [int(alg_flag)], [float(eta0)])[0, 0])


def simulate_SMART_like(z: np.ndarray, y: np.ndarray, theta_thresh: float, eta0: float) -> float:
    """fast_algorithms.py:184-195 -- single-switch SMART; returns total_loss - comparator_loss."""
    reg, _ = engine.smart_batch(z, y, [float(theta_thresh)], [float(eta0)])
    return float(reg[0, 0])


def simulate_SMART(z: np.ndarray, y: np.ndarray, *, eta0: float = math.sqrt(2)) -> float:
    """fast_algorithms.py:198-200."""
    T = np.asarray(z).shape[0]
    return simulate_SMART_like(z, y, theta_thresh=math.sqrt(2 * T), eta0=eta0)


def simulate_empirical_g_SMART(z: np.ndarray, y: np.ndarray, theta_emp: float, *, eta0: float = math.sqrt(2)) -> float:
    """fast_algorithms.py:203-204."""
    return simulate_SMART_like(z, y, theta_thresh=theta_emp, eta0=eta0)


def empirical_worst_case_thresholds(T_grid: np.ndarray, *, runs: int = 5, base_seed: int = 0,
                                    comparator: str = "exact_max") -> Dict[int, float]:
    """fast_algorithms.py:211-247 -- for each T: max over `runs` seeded sequences of the FTRL
    (eta0 = sqrt 2) regret, clamped at 0.  One fused launch per T: sequences are generated on the
    device from the same numpy stream keys (base_seed, T, run) and never stored.
    `comparator` is this engine's only addition to the reference signature.  All three give max_r regret_r:
    "exact_max" (default): the reference's value BIT FOR BIT at single-pass speed -- the single-pass regret
    (cum - T/2 + ||theta_T||) is an upper bound of the literal one for every run, so only the runs whose bound reaches
    the largest bound (typically one) take the literal second pass of fast_algorithms.py:113-114;
    "two_pass": the literal second pass for every run (same bits, 1.25x slower at T = 1e4);
    "linear": the single-pass value itself (within 1e-12 * T/2 of the reference; see include/oco_b200.h)."""
    tg = np.asarray(T_grid)
    g, _ = engine.g_sweep(tg, runs=int(runs), base_seed=int(base_seed), comparator=comparator)
    return {int(T): float(v) for T, v in zip(tg, g)}


def _rng(base_seed: int, T: int, run: int) -> np.random.Generator:
    """fast_algorithms.py:254-257 (host-side generator with the reference's stream key)."""
    ss = np.random.SeedSequence([base_seed, T, run])
    return np.random.Generator(np.random.PCG64(ss))
