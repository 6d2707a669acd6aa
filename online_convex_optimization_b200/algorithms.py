"""Drop-in twin of the reference's ``algorithms`` module (the NumPy / float32-state implementation
``driver.py:22-27`` imports): same names and signatures as ``fast_algorithms``.

``simulate_alg`` runs the float32 compute mode of the engine (float32 theta / x, float64 loss
accumulator, algorithms.py:28-54); the contract of this mode is the north-star tolerance
1e-5 relative to the reference, not bit parity (the reference's comparator pass goes through a
float32 BLAS matvec whose summation order is not pinned).  The SMART variants and
``empirical_worst_case_thresholds`` use the float64 kernels: their float32 reference results differ
from float64 by ~1e-6 relative, inside the same tolerance.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from . import engine
from . import fast_algorithms as _fa

_rng = _fa._rng                                     # algorithms.py:177-180


def simulate_alg(z: np.ndarray, y: np.ndarray, alg_flag: int, eta0: float) -> float:
    """algorithms.py:28-54 (alg_flag 0 = FTRL, 1 = FTL)."""
    z32 = np.ascontiguousarray(z, dtype=np.float32)
    y32 = np.ascontiguousarray(y, dtype=np.float32)
    return float(engine.simulate_batch(z32, y32, [int(alg_flag)], [float(eta0)], compute="f32")[0, 0])


def simulate_SMART_like(z: np.ndarray, y: np.ndarray, theta_thresh: float, eta0: float) -> float:
    """algorithms.py:65-120."""
    return _fa.simulate_SMART_like(np.asarray(z, dtype=np.float32), np.asarray(y, dtype=np.float32), theta_thresh, eta0)


def simulate_SMART(z: np.ndarray, y: np.ndarray, *, eta0: float = math.sqrt(2)) -> float:
    """algorithms.py:123-125."""
    T = np.asarray(z).shape[0]
    return simulate_SMART_like(z, y, theta_thresh=math.sqrt(2 * T), eta0=eta0)


def simulate_empirical_g_SMART(z: np.ndarray, y: np.ndarray, theta_emp: float, *, eta0: float = math.sqrt(2)) -> float:
    """algorithms.py:127-128."""
    return simulate_SMART_like(z, y, theta_thresh=theta_emp, eta0=eta0)


def empirical_worst_case_thresholds(T_grid: np.ndarray, *, runs: int = 5, base_seed: int = 0) -> Dict[int, float]:
    """algorithms.py:135-171: g-path sequences cast to float32, FTRL(eta0 = sqrt 2), max regret per T.
    The float32-cast sequences come from the device generator; the play uses the float32 compute mode."""
    out: Dict[int, float] = {}
    for T_val in np.asarray(T_grid):
        T = int(T_val)
        mx = 0.0
        if runs > 0:
            z, y = engine.gen_sequences("gpath", int(runs), T, 5, key_a=int(base_seed), key_c0=0, dtype=np.float32)
            reg = engine.simulate_batch(z, y, [0], [math.sqrt(2)], compute="f32")[:, 0]
            mx = max(0.0, float(reg.max()))
        out[T] = mx
    return out
