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
    # outside the linear regime and on the linf / l1 balls the general solver takes over (same API)
    big = ex.compute_prefix_actions(ex.ExactFTLNoClip(d=3, T_max=40), z * 5.0, y)
    assert big.shape == (41, 3) and (np.linalg.norm(big, axis=1) <= 1.0 + 1e-9).all()
    s1 = ex.ExactFTLNoClip(d=3, T_max=40, norm="l1")
    a1 = ex.compute_prefix_actions(s1, z, y)
    assert (np.abs(a1).sum(axis=1) <= 1.0 + 1e-9).all()
    last = s1.solve_prefix(z[:17], y[:17])                            # one prefix == the batched row (same objective value)
    f = lambda x: 0.5 * np.abs(z[:17] @ x - y[:17]).sum()
    assert abs(f(last) - f(a1[17])) <= 1e-6 * max(1.0, f(last))
    assert ex.run_ftl_exact(z, y, norm="linf").regret >= -1e-6
    with pytest.raises(RuntimeError):                                 # a solve that cannot converge reports failure
        ex.compute_prefix_actions(ex.ExactFTLNoClip(d=3, T_max=40, solver_opts={"max_iter": 2}), z * 5.0, y)
    with pytest.raises(RuntimeError):
        ex.ExactFTLNoClip(d=17, T_max=4, norm="l1")
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


def test_c2_exact_vs_fast_ftl_1e4_trials_bit_exact():
    """BASELINE configs[1]: exact FTL vs fast (linearised) FTL on the same 1e4 sequences.  In the linear regime both
    play -theta/||theta|| = v/||v||; the closed-form kernel uses the fast path's operation order, so the iterates
    (and the regrets) agree bit for bit, not just to solver tolerance."""
    from online_convex_optimization_b200 import engine
    from online_convex_optimization_b200._lib import FAM_IID, FAM_NOISY
    B, T, d = 10000, 100, 5
    keys = np.array([[2025 * (b // 16 + 1), 13 + b % 16] for b in range(B)], dtype=np.uint64)
    for fam in (FAM_IID, FAM_NOISY):
        z, y = engine.gen_sequences(fam, B, T, d, keys=keys, p_flip=0.1)
        ex = engine.exact_ftl_prefix_batch(z, y)
        sim = engine.simulate_batch(z, y, [1], [1.0], full=True, trace=True)
        assert (ex["status"] == 0).all()
        assert (ex["actions"][:, :T].view(np.uint64) == sim["x_trace"][:, 0].view(np.uint64)).all()
        assert (ex["regret"] == sim["regret"][:, 0]).all() and (ex["cum_loss"] == sim["cum_loss"][:, 0]).all()


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


# ------------------------------------------------------------------ general exact FTL (ADMM) vs scipy
def _obj(x, z, y):
    return 0.5 * np.abs(z @ x - y).sum()


def _scipy_opt(z, y, norm):
    """Optimal value from the scipy restatement of the reference's program (oracle/exact_ftl_scipy.py)."""
    from oracle import exact_ftl_scipy as XS
    return XS.solve(z, y, norm)[0]


@pytest.mark.parametrize("norm", ["l2", "linf", "l1"])
def test_exact_ftl_general_solver_vs_scipy(norm):
    """Outside the linear regime (rows up to norm ~2.5, real-valued labels) the batched interior-point solver reaches
    the optimum of an independent solver: never meaningfully worse (HiGHS is exact for linf / l1; SLSQP is a weaker
    reference for l2, hence the looser bound there), and strictly feasible."""
    from online_convex_optimization_b200 import engine
    rng = np.random.default_rng(7)
    B, T, d = 4, 40, 3
    z = rng.standard_normal((B, T, d)) * 0.9
    y = rng.standard_normal((B, T))
    res = engine.exact_ftl_general_batch(z, y, norm=norm)
    assert (res["status"] == 0).all(), res["status"]
    assert (res["actions"][:, 0] == 0).all()
    pnorm = {"l2": 2, "linf": np.inf, "l1": 1}[norm]
    for b in range(B):
        for k in (1, 2, 5, 17, 40):
            x = res["actions"][b, k]
            assert np.linalg.norm(x, pnorm) < 1.0
            f, f_ref = _obj(x, z[b, :k], y[b, :k]), _scipy_opt(z[b, :k], y[b, :k], norm)
            f0 = 0.5 * np.abs(y[b, :k]).sum()
            assert f <= f_ref + (2e-6 if norm == "l2" else 2e-8) * max(1.0, f0), (norm, b, k, f, f_ref)
        rep = O.replay_actions(z[b], y[b], res["actions"][b])
        assert abs(res["regret"][b] - rep["regret"]) <= 1e-10 * max(1.0, abs(rep["regret"]))
    # prefixes are solved independently: asking for the last one only gives the same bits
    last = engine.exact_ftl_general_batch(z, y, norm=norm, first_prefix=T)
    assert (last["actions"][:, T].view(np.uint64) == res["actions"][:, T].view(np.uint64)).all()
    assert (last["actions"][:, :T] == 0).all() and (last["status"] == 0).all()
    tight = engine.exact_ftl_general_batch(z, y, norm=norm, tol=1e-10)      # tolerance is honoured
    for b in range(B):
        d8, d10 = _obj(res["actions"][b, T], z[b], y[b]), _obj(tight["actions"][b, T], z[b], y[b])
        assert -1e-10 * T <= d8 - d10 <= 1e-8 * max(1.0, 0.5 * np.abs(y[b]).sum())
    starved = engine.exact_ftl_general_batch(z, y, norm=norm, max_iter=2)
    assert (starved["status"] > 0).all()
    if norm == "l2":     # sequences inside the linear regime keep the closed form, bit for bit
        zs, ys = O.gen_gpath(4, 60, 1, d)
        mixed = engine.exact_ftl_general_batch(np.stack([zs, zs * 3.0]), np.stack([ys, ys]))
        assert (mixed["status"] == 0).all()
        assert (mixed["actions"][0].view(np.uint64) == O.exact_ftl_l2_prefix(zs, ys).view(np.uint64)).all()


@pytest.mark.parametrize("norm,d", [("l2", 10), ("linf", 10), ("l2", 20), ("linf", 32), ("l2", 32), ("l1", 10), ("l1", 16)])
def test_exact_ftl_general_solver_wide_d(norm, d):
    """The general solver for 8 < d <= 32 (kernels_wide_ipm.cu: lane per variable, Hessian in shared memory) against
    the scipy restatement of the reference's program (l2, linf; l1 in split variables up to d = 16): rows up to norm ~3, real-valued labels, prefixes shorter than d
    (rank-deficient Hessians) and longer; includes the reference's own smoke shape (exact_ftl.py:460-475: d = 10)
    pushed out of the linear regime.  Same bounds as the d <= 8 test."""
    from online_convex_optimization_b200 import engine
    from online_convex_optimization_b200 import exact_ftl as X
    rng = np.random.default_rng(11 + d)
    B, T = 3, 48
    z = rng.standard_normal((B, T, d)) * (2.0 / np.sqrt(d))
    y = rng.standard_normal((B, T))
    res = engine.exact_ftl_general_batch(z, y, norm=norm)
    assert (res["status"] == 0).all(), res["status"]
    assert (res["actions"][:, 0] == 0).all()
    pnorm = {"l2": 2, "linf": np.inf, "l1": 1}[norm]
    for b in range(B):
        for k in (1, 3, d - 1, d + 1, 33, T):
            x = res["actions"][b, k]
            assert np.linalg.norm(x, pnorm) < 1.0
            f, f_ref = _obj(x, z[b, :k], y[b, :k]), _scipy_opt(z[b, :k], y[b, :k], norm)
            f0 = 0.5 * np.abs(y[b, :k]).sum()
            assert f <= f_ref + (2e-6 if norm == "l2" else 2e-8) * max(1.0, f0), (norm, d, b, k, f, f_ref)
        rep = O.replay_actions(z[b], y[b], res["actions"][b])
        assert abs(res["regret"][b] - rep["regret"]) <= 1e-10 * max(1.0, abs(rep["regret"]))
    last = engine.exact_ftl_general_batch(z, y, norm=norm, first_prefix=T)
    assert (last["actions"][:, T].view(np.uint64) == res["actions"][:, T].view(np.uint64)).all()
    starved = engine.exact_ftl_general_batch(z, y, norm=norm, max_iter=2)
    assert (starved["status"] > 0).all()
    if norm == "l2":     # sequences inside the linear regime keep the closed form; the others are solved
        zs, ys = O.gen_gpath(4, 40, 1, d)
        mixed = engine.exact_ftl_general_batch(np.stack([zs, zs * 3.0]), np.stack([ys, ys]))
        assert (mixed["status"] == 0).all()
        want = O.exact_ftl_l2_prefix(zs, ys)
        assert np.all(np.abs(mixed["actions"][0] - want) <= 1e-12)
        x3 = mixed["actions"][1, 40]
        assert _obj(x3, zs * 3.0, ys) <= _scipy_opt(zs * 3.0, ys, "l2") + 2e-6 * 20
        # the stateful twin (exact_ftl.py:130-193) on a sequence outside the regime
        sol = X.ExactFTLNoClip(d=d, T_max=40)
        xk = sol.solve_prefix_from_full(zs * 3.0, ys, 25)
        assert _obj(xk, zs[:25] * 3.0, ys[:25]) <= _scipy_opt(zs[:25] * 3.0, ys[:25], "l2") + 2e-6 * 12.5
    else:
        sol = X.ExactFTLNoClip(d=d, T_max=T, norm=norm)
        xk = sol.solve_prefix(z[0], y[0])
        assert _obj(xk, z[0], y[0]) <= _scipy_opt(z[0], y[0], norm) + 2e-8 * max(1.0, 0.5 * np.abs(y[0]).sum())
    if norm == "l1":     # split variables (x = u - w) cover the l1 ball up to d = 16; beyond: the reference's solver-failure exception
        with pytest.raises(RuntimeError):
            X.ExactFTLNoClip(d=17, T_max=4, norm="l1")


# ------------------------------------------------------------------ exact_ftl.py deterministic half, reference-pinned
def test_exact_ftl_twin_vs_reference_fixtures(golden_exact_det):
    """replay_exact_ftl / run_ftrl(comparator_action=...) / _simulate_ftrl of the twin against fixtures produced by
    RUNNING the reference's exact_ftl.py (exact_ftl.py:230-333, 399-420; tools/make_golden.py), d = 1 ... 33 including
    the reference's own d = 10 smoke configuration.  d <= 8: played losses and x_last bit-exact (sequential sums, same
    order); comparator loss 1e-12 (BLAS matvec + pairwise sum in the reference).  d > 8: warp reductions, 1e-12."""
    from conftest import exact_det_inputs
    from online_convex_optimization_b200 import exact_ftl as X
    n = 0
    for i, case in enumerate(golden_exact_det["cases"]):
        z, y, acts = exact_det_inputs(case, i)
        T, d = z.shape
        for k, a in acts.items():
            got = X.replay_exact_ftl(z, y, a)
            want = {f: hexf(v) for f, v in case["replay"][k].items() if f != "actions_seed"}
            if d <= 8:
                assert got.cum_loss == want["cum_loss"], (case["name"], k)
            assert abs(got.cum_loss - want["cum_loss"]) <= REL * max(1.0, want["cum_loss"])
            assert abs(got.comp_loss - want["comp_loss"]) <= REL * max(1.0, want["comp_loss"])
            assert abs(got.regret - want["regret"]) <= REL * max(1.0, want["cum_loss"])
            assert (got.x_last == a[T]).all()
            n += 1
        for key, rec in case["ftrl"].items():
            eta = float(key.split("_")[0][3:]); which = key.split("_")[1]
            got = X.run_ftrl(z, y, eta0=eta, comparator_action=acts[which][T])
            want_x = np.array([hexf(v) for v in rec["x_last"]])
            if d <= 8:
                assert got.cum_loss == hexf(rec["cum_loss"]), (case["name"], key)
                assert (got.x_last == want_x).all()
            assert abs(got.cum_loss - hexf(rec["cum_loss"])) <= REL * max(1.0, hexf(rec["cum_loss"]))
            assert abs(got.comp_loss - hexf(rec["comp_loss"])) <= REL * max(1.0, hexf(rec["comp_loss"]))
            assert abs(got.regret - hexf(rec["regret"])) <= REL * max(1.0, hexf(rec["cum_loss"]))
            assert np.all(np.abs(got.x_last - want_x) <= REL * np.maximum(1.0, np.abs(want_x)))
            n += 1
    assert n >= 80


def test_exact_ftl_reference_smoke_configuration_d10():
    """The reference module's __main__ (exact_ftl.py:460-475): T = 200, d = 10 -- exact FTL + FTRL through the twin,
    against the scipy oracle (the exact half has no reference fixture: CVXPY)."""
    from oracle import exact_ftl_scipy as XS
    from online_convex_optimization_b200 import exact_ftl as X
    rng = np.random.default_rng(0)
    T, d = 200, 10
    z = rng.normal(size=(T, d))
    z /= np.maximum(1.0, np.linalg.norm(z, axis=1, keepdims=True))
    w_true = rng.normal(size=d)
    w_true /= np.linalg.norm(w_true)
    y = np.sign(z @ w_true + 0.2 * rng.normal(size=T))
    res, actions = X.run_ftl_exact(z, y, norm="l2", solver="ECOS", return_actions=True)
    assert actions.shape == (T + 1, d) and np.all(np.linalg.norm(actions, axis=1) <= 1 + 1e-12)
    for k in (1, 7, 50, 200):                                    # prefix optima: never worse than the SOCP restatement
        val = 0.5 * np.abs(z[:k] @ actions[k] - y[:k]).sum()
        ref = XS.solve(z[:k], y[:k], "l2")[0]
        assert val <= ref + 1e-6 * max(1.0, ref), (k, val, ref)
    assert abs(res.regret - (res.cum_loss - res.comp_loss)) < 1e-9
    ftrl = X.run_ftrl(z, y, eta0=1.0)
    assert abs(ftrl.comp_loss - res.comp_loss) <= 1e-9 * max(1.0, res.comp_loss)   # same exact comparator
    X._smoke()
    solver = X.ExactFTLNoClip(d=d, T_max=T)                      # stateful API at d = 10
    a1 = solver.append_row(z[0], float(y[0]))
    assert np.allclose(a1, actions[1]) and solver._last_length == 1
    # outside the linear regime: the general solver takes over up to d = 32 ...
    res3, act3 = X.run_ftl_exact(z[:60] * 3.0, y[:60], return_actions=True)
    for k in (5, 60):
        val = 0.5 * np.abs(z[:k] * 3.0 @ act3[k] - y[:k]).sum()
        ref = XS.solve(z[:k] * 3.0, y[:k], "l2")[0]
        assert val <= ref + 2e-6 * max(1.0, 0.5 * k), (k, val, ref)
    assert abs(res3.regret - (res3.cum_loss - res3.comp_loss)) < 1e-9
    z40 = rng.normal(size=(30, 40))
    with pytest.raises(RuntimeError):                            # ... beyond it: the reference's solver-failure exception
        X.run_ftl_exact(z40, np.sign(rng.normal(size=30)))


def test_driver_twin_vs_reference_driver_fixtures(golden_driver32):
    """driver.py twin against the float32 statistics the REFERENCE's driver.py produced (driver.py:70-129 through
    algorithms.py, tools/make_golden.py): g(T) and the per-case means / CIs, 1e-5 (north_star's fp32 tolerance)."""
    from online_convex_optimization_b200 import algorithms as alg32
    from online_convex_optimization_b200 import driver as dr
    from online_convex_optimization_b200 import sequence_generation as sg
    grid = np.array([100, 200, 300])
    g_want = {int(k): hexf(v) for k, v in golden_driver32["g_emp_runs40"].items()}
    g_got = alg32.empirical_worst_case_thresholds(grid, runs=40, base_seed=0)
    for T in grid:
        assert abs(g_got[int(T)] - g_want[int(T)]) <= 1e-5 * g_want[int(T)]
    for title, rec in golden_driver32["stats_small"].items():
        got = dr.evaluate_stream_with_stats(sg.CASES[title], grid, g_want, runs=rec["runs"], replicates=rec["replicates"], base_seed=0)
        for k in ("FTRL", "FTL", "SMART", "EMP"):
            want_m = np.array([hexf(v) for v in rec["stats"][k][0]]); want_c = np.array([hexf(v) for v in rec["stats"][k][1]])
            assert np.all(np.abs(got[k][0] - want_m) <= 1e-5 * np.maximum(1.0, np.abs(want_m))), (title, k, got[k][0], want_m)
            assert np.all(np.abs(got[k][1] - want_c) <= 1e-5 * np.maximum(1.0, np.abs(want_c)) + 1e-5), (title, k)


@pytest.mark.parametrize("norm", ["linf", "l1"])
def test_exact_driver_twin_other_norms(norm):
    """exact_ftl_driver with norm != 'l2' solves on THAT ball (exact_ftl_driver.py:64-206 passes norm to every
    solver): g(T) and the per-case regrets against the scipy oracle (HiGHS), not the l2 closed form."""
    from oracle import exact_ftl_scipy as XS
    from online_convex_optimization_b200 import exact_ftl_driver as XD
    from online_convex_optimization_b200 import sequence_generation as sg
    grid = np.array([12, 30])
    cfg = XD.ExperimentConfig(T_grid=grid, g_runs=3, norm=norm, solver_opts={"tol": 1e-9})
    g = XD.empirical_worst_case_thresholds(grid, runs=3, base_seed=0, norm=norm, solver=None, solver_opts={"tol": 1e-9})
    for T in grid:
        best = 0.0
        for r in range(3):
            z, y = O.gen_gpath(0, int(T), r, 5)
            comp = XS.solve(z, y, norm)[0]
            cum = O.simulate_alg(z, y, 0, R2, full=True)["cum_loss"]
            best = max(best, cum - comp)
        assert abs(g[int(T)] - best) <= 1e-6 * max(1.0, best), (norm, T, g[int(T)], best)
    st = XD.evaluate_stream_with_stats(sg.CASES["Random i.i.d. (separable)"], grid, runs=2, replicates=1, cfg=cfg, stream_name="iid")
    # FTRL column = FTRL loss - OPTIMAL VALUE on the chosen ball (unique even where the minimiser is not)
    for ti, T in enumerate(grid):
        vals = []
        for run in range(2):
            z, y, _ = O.gen_iid(2025 * (run + 1), int(T), 0)
            z = np.asarray(z, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
            vals.append(O.simulate_alg(z, y, 0, R2, full=True)["cum_loss"] - XS.solve(z, y, norm)[0])
        assert abs(st["FTRL"][0][ti] - np.mean(vals)) <= 1e-6 * max(1.0, abs(np.mean(vals))), (norm, T)
    assert np.all(np.isfinite(st["FTL (exact)"][0]))
    l2 = XD.evaluate_stream_with_stats(sg.CASES["Random i.i.d. (separable)"], grid, runs=2, replicates=1,
                                       cfg=XD.ExperimentConfig(T_grid=grid, g_runs=3, norm="l2"), stream_name="iid")
    assert not np.allclose(l2["FTRL"][0], st["FTRL"][0])          # a different ball gives a different comparator
    empty = XD.evaluate_stream_with_stats(sg.CASES["Label flips"], grid, runs=0, replicates=1, cfg=cfg, stream_name="none")
    assert all((m == 0).all() and (c == 0).all() for m, c in empty.values())
