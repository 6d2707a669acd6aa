"""CPU tests of the host-side logic of the driver twins (no GPU, no library call): the statistics of
fast_driver.py:113-125 as implemented by fast_driver._finish_case, the stateful buffer protocol of
exact_ftl.ExactFTLNoClip (exact_ftl.py:130-169) and the stream-builder surface of sequence_generation.py."""
import math

import numpy as np
import pytest

from online_convex_optimization_b200 import exact_ftl as ex
from online_convex_optimization_b200 import fast_driver as fd
from online_convex_optimization_b200 import sequence_generation as sg


def _reference_stats(per_T, runs):
    """Literal transcription of fast_driver.py:113-125 on per-sequence regrets (runs, replicates, 4)."""
    out = {}
    for ki, k in enumerate(fd.ALGO_KEYS):
        means, cis = [], []
        for reg in per_T:
            run_means = [float(np.mean([float(v) for v in reg[run, :, ki]])) for run in range(runs)]
            arr = np.asarray(run_means, dtype=float)
            mu = float(np.mean(arr)) if arr.size else 0.0
            n = arr.size
            sem = 0.0 if n <= 1 else float(np.std(arr, ddof=1) / math.sqrt(n))
            means.append(mu)
            cis.append(1.96 * sem if n > 1 else 0.0)
        out[k] = (np.array(means), np.array(cis))
    return out


@pytest.mark.parametrize("runs,reps", [(48, 16), (48, 20), (1, 1), (3, 7), (5, 129)])
def test_finish_case_statistics_are_the_references(runs, reps):
    rng = np.random.default_rng(runs * 1000 + reps)
    per_T = [rng.standard_normal((runs, reps, 4)) * 25.0 + 3.0 for _ in range(4)]
    got = fd._finish_case([a.copy() for a in per_T], runs)
    want = _reference_stats(per_T, runs)
    for k in fd.ALGO_KEYS:
        assert (got[k][0].view(np.uint64) == want[k][0].view(np.uint64)).all()      # means: bit-identical
        assert (got[k][1].view(np.uint64) == want[k][1].view(np.uint64)).all()      # CI half-widths: bit-identical


def test_experiment_config_and_cases_surface():
    cfg = fd.ExperimentConfig()
    assert list(cfg.T_grid) == list(range(100, 1100, 100)) and cfg.base_seed == 0 and cfg.g_runs == 1000
    assert list(sg.CASES) == ["Random i.i.d. (separable)", "Massart noise 10%", "Label flips", "Switching leaders"]
    assert sg.RUNS_BY_TITLE["Massart noise 10%"] == 48 and sg.REPLICATES_BY_TITLE["Massart noise 10%"] == 20
    for title, builder in sg.CASES.items():
        s = builder(run_seed=2025)                      # building a sampler needs no device
        assert callable(s) and s.spec["d"] == 5 and s.spec["run_seed"] == 2025
    # the task vector of the iid families is numpy's: unit norm in float32 (sequence_generation.py:55-59)
    u = sg._task_vector(2025, 11, 5)
    assert u.dtype == np.float32 and abs(float(np.linalg.norm(u)) - 1.0) < 1e-6


def test_exact_solver_buffer_protocol_without_device():
    """exact_ftl.py:130-169: argument checks and the row buffer bookkeeping do not touch the device."""
    with pytest.raises(ValueError):
        ex.ExactFTLNoClip(d=3, T_max=4, norm="l3")
    s = ex.ExactFTLNoClip(d=2, T_max=3)
    assert s._last_length == 0 and s._Z_buf.shape == (3, 2)
    z = np.arange(6, dtype=np.float64).reshape(3, 2) / 10.0
    y = np.array([1.0, -1.0, 1.0])
    s._set_prefix(z, y, 2)
    assert s._last_length == 2 and (s._Z_buf[:2] == z[:2]).all() and (s._Z_buf[2] == 0).all() and (s._w_buf == [1, 1, 0]).all()
    s._set_prefix(z, y, 1)                              # shrinking clears the tail (exact_ftl.py:163-168)
    assert s._last_length == 1 and (s._Z_buf[1:] == 0).all() and (s._y_buf[1:] == 0).all() and (s._w_buf == [1, 0, 0]).all()
    with pytest.raises(ValueError):
        s._set_prefix(z, y, 4)
    with pytest.raises(ValueError):
        s.solve_prefix(np.zeros((2, 3)), np.zeros(2))   # wrong dimension (exact_ftl.py:188-189)
    with pytest.raises(ValueError):
        s.solve_prefix(np.zeros((4, 2)), np.zeros(4))   # longer than T_max (exact_ftl.py:190-191)
    s.reset_buffers()
    assert s._last_length == 0 and (s._Z_buf == 0).all()
    assert (s._solve_current() == 0).all()              # empty prefix -> zero action, no solve
    with pytest.raises(ValueError):
        ex.replay_exact_ftl(z, y, np.zeros((3, 2)))     # actions must be (T+1, d) (exact_ftl.py:314-315)


def test_plotting_shim_draws_the_reference_figures(tmp_path, monkeypatch):
    """plotting.py with a stand-in matplotlib: g(T) figure = 3 curves, comparison figure = one panel per case with
    one line per algorithm and a CI band only where the CI is non-zero; without matplotlib a clear ImportError."""
    import sys
    import types
    from unittest.mock import MagicMock
    import numpy as np
    from online_convex_optimization_b200 import plotting
    fake = types.ModuleType("matplotlib"); fake.use = lambda *a, **k: None
    plt = MagicMock()
    axes_made = []

    def subplots(*a, **k):
        fig = MagicMock()
        if a and len(a) >= 2:
            arr = np.empty((a[0], a[1]), dtype=object)
            for i in range(a[0]):
                for j in range(a[1]):
                    ax = MagicMock(); ax.plot.return_value = (MagicMock(),); arr[i, j] = ax; axes_made.append(ax)
            return fig, arr
        ax = MagicMock(); ax.plot.return_value = (MagicMock(),); axes_made.append(ax)
        return fig, ax
    plt.subplots.side_effect = subplots
    fake.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", fake)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    Tg = np.array([100, 200, 300])
    plotting.plot_empirical_g(Tg, {100: 6.0, 200: 8.3, 300: 10.4}, out_path=str(tmp_path / "g.png"))
    assert axes_made[-1].plot.call_count == 3
    stats = {"A": {"FTRL": (np.ones(3), np.zeros(3)), "FTL": (np.ones(3), np.full(3, 0.1))},
             "B": {"FTRL": (np.ones(3), np.zeros(3))}, "C": {"FTRL": (np.ones(3), np.zeros(3))}}
    n0 = len(axes_made)
    plotting.plot_comparisons(Tg, stats, out_path=str(tmp_path / "c.png"))
    panels = axes_made[n0:]
    assert len(panels) == 4 and panels[0].plot.call_count == 2 and panels[0].fill_between.call_count == 1
    assert panels[3].axis.call_count == 1                      # unused fourth panel switched off
    monkeypatch.setitem(sys.modules, "matplotlib", None)
    import pytest
    with pytest.raises(ImportError):
        plotting.plot_empirical_g(Tg, {100: 6.0, 200: 8.3, 300: 10.4})
