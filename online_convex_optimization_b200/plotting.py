"""Plotting shim: the two figures of the reference's drivers (fast_driver.py:134-194,
driver.py / exact_ftl_driver.py alike) drawn from the batched twins' results.  matplotlib is imported lazily --
it is host-side presentation, absent from the build image, and nothing on the hot path depends on it; without it
the drivers write their JSON and these functions raise a clear ImportError."""
from __future__ import annotations

import math
from typing import Dict, Mapping, Tuple

import numpy as np


def _pyplot():
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception as e:   # pragma: no cover - depends on the host
        raise ImportError("matplotlib is needed for the figures (the numerical results are in the drivers' JSON output)") from e
    return plt


def plot_empirical_g(T_grid, g_emp: Mapping[int, float], *, out_path: str = "empirical_g_T_b200.png") -> None:
    """Empirical g(T) with the two reference curves sqrt(T/pi) and sqrt(2T) (fast_driver.py:140-158)."""
    plt = _pyplot()
    Ts = [int(T) for T in T_grid]
    fig, ax = plt.subplots(figsize=(7.5, 5.0))
    ax.plot(Ts, [float(g_emp[T]) for T in Ts], marker="o", label="Empirical g(T)")
    ax.plot(Ts, [math.sqrt(T / math.pi) for T in Ts], linestyle="--", label=r"$\sqrt{T/\pi}$")
    ax.plot(Ts, [math.sqrt(2 * T) for T in Ts], marker="x", label=r"$\sqrt{2T}$")
    ax.set_title("Empirical worst-case g(T) for SMART (ALG_WC = FTRL)", fontsize=18)
    ax.set_xlabel("T rounds", fontsize=16)
    ax.set_ylabel("g(T)", fontsize=16)
    ax.legend(prop={"size": 14})
    fig.tight_layout()
    fig.savefig(out_path, dpi=600, bbox_inches="tight")
    plt.close(fig)


def plot_comparisons(T_grid, stats_by_case: Dict[str, Dict[str, Tuple[np.ndarray, np.ndarray]]], *,
                     out_path: str = "algorithm_comparison_b200.png", labels: Mapping[str, str] | None = None,
                     runs_by_title: Mapping[str, int] | None = None, reps_by_title: Mapping[str, int] | None = None) -> None:
    """One panel per sequence family: mean regret per algorithm with its 95 % band (fast_driver.py:161-194)."""
    plt = _pyplot()
    labels = dict(labels or {"FTRL": "FTRL", "FTL": "FTL", "SMART": "SMART (√2T)", "EMP": "SMART (empirical g)",
                             "FTL (exact)": "FTL (exact)"})
    Ts = np.asarray([int(T) for T in T_grid])
    n = len(stats_by_case)
    rows = max(1, math.ceil(n / 2))
    fig, axes = plt.subplots(rows, 2, figsize=(12, 4.0 * rows), squeeze=False)
    flat = axes.ravel()
    for ax, (title, stats) in zip(flat, stats_by_case.items()):
        for key, (mean, ci) in stats.items():
            mean, ci = np.asarray(mean, dtype=float), np.asarray(ci, dtype=float)
            (line,) = ax.plot(Ts, mean, label=labels.get(key, key))
            if np.any(ci > 0.0):
                ax.fill_between(Ts, mean - ci, mean + ci, alpha=0.2, linewidth=0, color=line.get_color())
        suffix = ""
        if runs_by_title is not None and reps_by_title is not None:
            suffix = f" (runs={runs_by_title.get(title, 1)}, reps/T={reps_by_title.get(title, 1)})"
        ax.set_title(title + suffix, fontsize=16)
        ax.set_xlabel("T rounds", fontsize=14)
        ax.set_ylabel("Cumulative regret", fontsize=14)
        ax.legend(prop={"size": 12})
    for ax in flat[n:]:
        ax.axis("off")
    fig.suptitle("Online Linear Binary Classification", fontsize=20)
    fig.tight_layout()
    fig.savefig(out_path, dpi=600, bbox_inches="tight")
    plt.close(fig)
