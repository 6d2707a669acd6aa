"""GPU tests (-m gpu) of the driver-level drop-ins: the batched fast_driver twin against the
reference's own evaluate_stream_with_stats outputs (tests/golden/driver_stats.json, produced by
running the unmodified reference through the in-memory shim), and the exact_ftl twin against
the oracle's closed form and the published exact-FTL curves (parity unpinned: CVXPY absent)."""
import math

import numpy as np
import pytest

from conftest import hexf
from oracle import oracle as O

pytestmark = pytest.mark.gpu
R2 = math.sqrt(2.0)
REL = 1e-12


def _check_stats(got, want, exact_keys=("FTRL", "FTL")):
    for k, (wm, wc) in want.items():
        wm = np.array([hexf(v) for v in wm])
        wc = np.array([hexf(v) for v in wc])
        gm, gc = got[k]
        if k in exact_keys:
            assert (gm.view(np.uint64) == wm.view(np.uint64)).all(), k
            assert (gc.view(np.uint64) == wc.view(np.uint64)).all(), k
        else:   # SMART columns: warp-parallel prefix loss, 1e-12
            assert np.all(np.abs(gm - wm) <= REL * np.maximum(1.0, np.abs(wm))), (k, gm, wm)
            assert np.all(np.abs(gc - wc) <= 1e-10 * np.maximum(1.0, np.abs(wc))), (k, gc, wc)


def test_fast_driver_small_golden(golden_stats):
    from online_convex_optimization_b200 import fast_driver as fd
    from online_convex_optimization_b200 import sequence_generation as sg
    g_emp = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs1000"].items()}
    for title, rec in golden_stats["stats_small"].items():
        got = fd.evaluate_stream_with_stats(sg.CASES[title], np.array([100, 200, 300]), g_emp, runs=rec["runs"],
                                            replicates=rec["replicates"], base_seed=0, stream_name=title)
        _check_stats(got, rec["stats"])


def test_fast_driver_full_default_workload(golden_stats):
    """BASELINE config 1: the reference's complete default experiment (T = 100..1000, 1000 g-runs,
    48x16 + 48x20 + 1 + 1 sequences per T, 4 algorithms) reproduced from the same seeds."""
    from online_convex_optimization_b200 import fast_driver as fd
    res = fd.run_experiment()
    want_g = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs1000"].items()}
    assert res["g_emp"] == want_g
    for title, want in golden_stats["stats_full"].items():
        _check_stats(res["stats_by_case"][title], want)
    # SURVEY.md Appendix C spot values
    assert res["stats_by_case"]["Label flips"]["FTL"][0][-1] == 250.0
    assert res["stats_by_case"]["Switching leaders"]["FTL"][0][-1] == -426.0


def test_fast_driver_foreign_builder_host_path():
    """A stream builder that is not ours (no .spec): sampled on the host, simulated in batch."""
    from online_convex_optimization_b200 import fast_driver as fd

    def builder(*, run_seed):
        def sample(T, rep=0):
            z, y = O.gen_gpath(run_seed, T, rep, 3)
            return z.astype(np.float32), y.astype(np.float32), None
        return sample
    T_grid = np.array([50, 80])
    got = fd.evaluate_stream_with_stats(builder, T_grid, {50: 3.0, 80: 4.0}, runs=3, replicates=2)
    for ti, T in enumerate(T_grid):
        per_run = []
        for run in range(3):
            vals = []
            for rep in range(2):
                z, y = O.gen_gpath(2025 * (run + 1), int(T), rep, 3)
                vals.append(O.simulate_alg(z.astype(np.float32), y.astype(np.float32), 0, R2))
            per_run.append(float(np.mean(vals)))
        assert got["FTRL"][0][ti] == float(np.mean(np.asarray(per_run)))


# ------------------------------------------------------------------ exact FTL twin
def test_exact_ftl_matches_oracle_closed_form():
    from online_convex_optimization_b200 import engine
    zs, ys = zip(*[O.gen_gpath(4, 300, r, 5) for r in range(6)])
    z, y = np.stack(zs), np.stack(ys)
    res = engine.exact_ftl_prefix_batch(z, y)
    assert not res["status"].any()
    for b in range(6):
        act = O.exact_ftl_l2_prefix(z[b], y[b])
        assert (res["actions"][b].view(np.uint64) == act.view(np.uint64)).all()
        rep = O.replay_actions(z[b], y[b], act)
        assert res["cum_loss"][b] == rep["cum_loss"]
        assert abs(res["regret"][b] - rep["regret"]) <= REL * max(1.0, abs(rep["regret"]))
    bad = engine.exact_ftl_prefix_batch(z * 2.0, y)
    assert bad["status"].all()


def test_exact_ftl_twin_api_and_known_answers():
    from online_convex_optimization_b200 import exact_ftl as ex
    from online_convex_optimization_b200 import sequence_generation as sg
    for T, want in ((100, 25.0), (1000, 250.0)):                 # published curve: label flips T/4
        z, y, _ = sg.flip_sequence(T)
        solver = ex.ExactFTLNoClip(d=5, T_max=T)
        actions = ex.compute_prefix_actions(solver, z, y)
        assert actions.shape == (T + 1, 5) and (actions[0] == 0).all()
        r = ex.replay_exact_ftl(z, y, actions)
        assert isinstance(r, ex.RunResult) and r.regret == want and (r.x_last == actions[T]).all()
        assert solver._last_length == T
    z, y, _ = sg.switching_two_leaders_sequence(500)
    assert ex.run_ftl_exact(z, y).regret == 6.5                    # published: switching leaders 1.5 / 6.5 / 12.5
    # stateful protocol: append_row solutions equal the batched prefix actions
    z, y = O.gen_gpath(1, 40, 0, 3)
    solver = ex.ExactFTLNoClip(d=3, T_max=40)
    acts = ex.compute_prefix_actions(solver, z, y)
    solver.reset_buffers()
    for t in range(40):
        assert (solver.append_row(z[t], float(y[t])) == acts[t + 1]).all()
    with pytest.raises(ValueError):
        solver.append_row(z[0], 1.0)                                 # longer than T_max (exact_ftl.py:142-143)
    with pytest.raises(ValueError):
        ex.compute_prefix_actions(ex.ExactFTLNoClip(d=2, T_max=40), z, y)
    with pytest.raises(ValueError):
        ex.replay_exact_ftl(z, y, acts[:-1])
    with pytest.raises(ValueError):
        ex.ExactFTLNoClip(d=3, T_max=4, norm="l3")
    with pytest.raises(RuntimeError):
        ex.compute_prefix_actions(ex.ExactFTLNoClip(d=3, T_max=40), z * 5.0, y)   # outside the linear regime
    with pytest.raises(RuntimeError):
        ex.ExactFTLNoClip(d=3, T_max=4, norm="l1")
    # run_ftrl with the exact comparator == oracle FTRL play + comparator loss at the exact action
    r = ex.run_ftrl(z, y, eta0=R2, comparator_action=acts[-1])
    o = O.simulate_alg(z, y, 0, R2, trace=True)
    assert r.cum_loss == o["cum_loss"]
    comp = 0.5 * np.abs(z @ acts[-1] - y).sum()
    assert abs(r.comp_loss - comp) <= REL * comp and abs(r.regret - (o["cum_loss"] - comp)) <= 1e-10
    r2 = ex.run_ftrl(z, y, eta0=R2)                                   # comparator solved internally
    assert r2.regret == r.regret and r2.x_last.shape == (3,)
    assert (r2.x_last == o["x_trace"][-1]).all()


def test_exact_driver_twin_small():
    from online_convex_optimization_b200 import exact_ftl_driver as ed
    from online_convex_optimization_b200 import sequence_generation as sg
    cfg = ed.ExperimentConfig(T_grid=np.array([100, 200]), g_runs=20)
    g = ed.empirical_worst_case_thresholds(cfg.T_grid, runs=20, base_seed=0, norm="l2", solver=None, solver_opts=None)
    assert set(g) == {100, 200} and all(v >= 0 for v in g.values())
    st = ed.evaluate_stream_with_stats(sg.CASES["Label flips"], cfg.T_grid, runs=1, replicates=1, cfg=cfg, stream_name="x")
    assert list(st["FTL (exact)"][0]) == [25.0, 50.0]
    st = ed.evaluate_stream_with_stats(sg.CASES["Random i.i.d. (separable)"], cfg.T_grid, runs=3, replicates=2, cfg=cfg,
                                       stream_name="x")
    # in the linear regime exact FTL == linearised FTL, so its regret equals the fast path's FTL regret
    from online_convex_optimization_b200 import fast_driver as fd
    fast = fd.evaluate_stream_with_stats(sg.CASES["Random i.i.d. (separable)"], cfg.T_grid, {100: 1.0, 200: 1.0},
                                         runs=3, replicates=2)
    assert np.allclose(st["FTL (exact)"][0], fast["FTL"][0], rtol=1e-9, atol=1e-9)


def test_driver_twin_float32_columns(golden_stats):
    """driver.py twin (float32-state algorithms): within 1e-5 of the float64 golden statistics."""
    from online_convex_optimization_b200 import driver as dr
    from online_convex_optimization_b200 import sequence_generation as sg
    g_emp = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs1000"].items()}
    rec = golden_stats["stats_small"]["Massart noise 10%"]
    got = dr.evaluate_stream_with_stats(sg.CASES["Massart noise 10%"], np.array([100, 200, 300]), g_emp, runs=rec["runs"],
                                        replicates=rec["replicates"], base_seed=0)
    for k in ("FTRL", "FTL"):
        want = np.array([hexf(v) for v in rec["stats"][k][0]])
        assert np.all(np.abs(got[k][0] - want) <= 1e-5 * np.maximum(1.0, np.abs(want)))
