#!/usr/bin/env python
"""Generate tests/golden/* by RUNNING THE REFERENCE ITSELF in this container.

Usage (build container only; /root/reference does not exist on the GPU box):
    NUMBA_CACHE_DIR=/tmp/numba_cache python tools/make_golden.py

Imports /root/reference/{fast_algorithms,algorithms,sequence_generation}.py
read-only and loads fast_driver.py through the in-memory shim of SURVEY.md
Appendix D (matplotlib stubbed, dataclass default patched in memory).  Nothing
from the reference is copied: only its *outputs* on seeded inputs are stored.

Outputs
  tests/golden/rng_kat.npz        numpy SeedSequence/PCG64/ziggurat/choice/random known answers
  tests/golden/sim_cases.json     regrets of simulate_alg / simulate_SMART_like on seeded sequences
  tests/golden/iterates.npz       per-step iterates x_t built from the reference's own njit helpers
  tests/golden/driver_stats.json  empirical_worst_case_thresholds + evaluate_stream_with_stats outputs
  tests/golden/exact_det.json     exact_ftl.py's deterministic half (cvxpy stubbed, never called): replay_exact_ftl,
                                  _simulate_ftrl, run_ftrl(comparator_action=...) on seeded sequences incl. d = 10
  tests/golden/driver32_stats.json  driver.py (float32 algorithms.py path): g(T) and evaluate_stream_with_stats
Floats are stored with repr()/float.hex() so they round-trip bit-exactly.
"""
from __future__ import annotations

import json
import math
import os
import sys
import types
from unittest.mock import MagicMock

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
REF = "/root/reference"
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import fast_algorithms as fa  # noqa: E402
import algorithms as alg32  # noqa: E402
import sequence_generation as sg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
R2 = math.sqrt(2.0)


def load_fast_driver():
    sys.modules["matplotlib"] = MagicMock()
    sys.modules["matplotlib.pyplot"] = MagicMock()
    src = open(os.path.join(REF, "fast_driver.py")).read().replace(
        "T_grid: IntArray = np.arange(100, 1100, 100, dtype=int)",
        "T_grid: IntArray = __import__('dataclasses').field(default_factory=lambda: np.arange(100, 1100, 100, dtype=int))")
    m = types.ModuleType("fast_driver")
    sys.modules["fast_driver"] = m
    exec(compile(src, "fast_driver.py", "exec"), m.__dict__)
    return m


def gpath(base, T, r, d=5):
    """fast_algorithms.py:231-239 with d as a parameter (the reference hard-codes 5)."""
    gen = fa._rng(base, T, r)
    z = gen.standard_normal((T, d)).astype(np.float64, copy=False)
    norms = np.linalg.norm(z, axis=1, keepdims=True).astype(np.float64, copy=False)
    z *= (1.0 / np.maximum(norms, 1.0))
    y = gen.choice([-1.0, 1.0], size=T).astype(np.float64, copy=False)
    return z, y


def rng_kat():
    keys = [(0, 100, 0), (0, 1000, 999), (2025, 1000, 13), (4050, 0, 11), (7, 10000, 0), (2 ** 40 + 5, 7, 2 ** 33)]
    out = {"keys": np.array(keys, dtype=np.uint64)}
    for i, k in enumerate(keys):
        ss = np.random.SeedSequence([int(v) for v in k])
        out[f"state_{i}"] = ss.generate_state(4, np.uint64)
        out[f"raw_{i}"] = np.random.PCG64(ss).random_raw(64)
        out[f"normal_{i}"] = fa._rng(*k).standard_normal(4096)
        out[f"random_{i}"] = fa._rng(*k).random(256)
        g = fa._rng(*k)
        g.standard_normal(333)
        out[f"choice_after_333_{i}"] = g.choice([-1.0, 1.0], size=257)
    # tail draws: the first 64 |x| > R normals of a long stream and their positions
    x = fa._rng(11, 22, 33).standard_normal(2_000_000)
    pos = np.nonzero(np.abs(x) > 3.6541528853610088)[0][:64]
    out["tail_pos"] = pos
    out["tail_val"] = x[pos]
    out["long_sum_bits"] = np.array([np.bitwise_xor.reduce(x.view(np.uint64))], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "rng_kat.npz"), **out)


def sim_cases():
    cases = []

    def add(name, z, y, spec):
        rec = {"name": name, "spec": spec, "T": int(z.shape[0]), "d": int(z.shape[1]), "results": {}}
        for flag in (0, 1):
            for eta in (R2, 0.25, 1.0, 4.0):
                rec["results"][f"alg{flag}_eta{eta!r}"] = float(fa.simulate_alg(z, y, flag, eta)).hex()
        if z.shape[0] <= 1000:
            T = z.shape[0]
            for th in (math.sqrt(2 * T), 3.0, 0.5, 1e9):
                rec["results"][f"smart_th{th!r}"] = float(fa.simulate_SMART_like(z, y, th, R2)).hex()
        # float32 NumPy twin (algorithms.py) -- fp32-mode tolerance target, short T only
        if z.shape[0] <= 300 and z.shape[1] == 5:
            zf, yf = z.astype(np.float32), y.astype(np.float32)
            rec["results"]["f32_alg0"] = float(alg32.simulate_alg(zf, yf, 0, R2)).hex()
            rec["results"]["f32_alg1"] = float(alg32.simulate_alg(zf, yf, 1, R2)).hex()
        cases.append(rec)

    for (b, T, r, d) in [(0, 100, 0, 5), (0, 1000, 999, 5), (0, 1000, 448, 5), (7, 10000, 0, 1), (7, 2000, 3, 1),
                         (3, 500, 2, 64), (3, 300, 2, 8), (3, 300, 5, 13), (5, 257, 1, 2), (5, 64, 1, 3), (9, 1, 0, 5)]:
        z, y = gpath(b, T, r, d)
        add(f"gpath_{b}_{T}_{r}_d{d}", z, y, {"family": "gpath", "base_seed": b, "T": T, "run": r, "d": d})
    for run in (0, 1):
        run_seed = 2025 * (run + 1)
        for title, builder in sg.CASES.items():
            sampler = builder(run_seed=run_seed)
            for T in (100, 300):
                for rep in (0, 1):
                    z, y, _ = sampler(T, rep)
                    add(f"case_{title}_{run_seed}_{T}_{rep}", z, y,
                        {"family": title, "run_seed": run_seed, "T": T, "rep": rep, "d": 5})
    z, y = gpath(0, 100, 0, 5)
    extra = {"gpath_0_100_0_z0_hex": [float(v).hex() for v in z[0]], "gpath_0_100_0_y8": [float(v) for v in y[:8]]}
    json.dump({"cases": cases, "extra": extra}, open(os.path.join(OUT, "sim_cases.json"), "w"), indent=0)


def iterates():
    """x_t per step, produced by stepping the reference's own njit helpers
    (fast_algorithms._action_ftrl/_action_ftl/_dot/_compute_gradient) exactly as
    _simulate_alg_core does (fast_algorithms.py:99-111)."""
    out = {}
    for name, (b, T, r, d) in {"d1": (7, 400, 0, 1), "d5": (0, 200, 3, 5)}.items():
        z, y = gpath(b, T, r, d)
        for flag in (0, 1):
            theta = np.zeros(d)
            x = np.zeros(d)
            xs = np.zeros((T, d))
            for t in range(T):
                if flag == 0:
                    fa._action_ftrl(theta, t + 1, R2, x)
                else:
                    fa._action_ftl(theta, x)
                xs[t] = x
                q = fa._dot(z[t], x)
                g = fa._compute_gradient(q, y[t])
                for j in range(d):
                    theta[j] += g * z[t, j]
            out[f"{name}_alg{flag}_x"] = xs
            out[f"{name}_alg{flag}_theta"] = theta.copy()
        out[f"{name}_key"] = np.array([b, T, r, d], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "iterates.npz"), **out)


def driver_stats():
    fd = load_fast_driver()
    res = {}
    T_grid = np.arange(100, 1100, 100, dtype=int)
    g_emp = fa.empirical_worst_case_thresholds(T_grid, runs=1000, base_seed=0)
    res["g_emp_runs1000"] = {str(k): float(v).hex() for k, v in g_emp.items()}
    g5 = fa.empirical_worst_case_thresholds(np.array([100, 200, 300]), runs=5, base_seed=3)
    res["g_emp_runs5_seed3"] = {str(k): float(v).hex() for k, v in g5.items()}
    # reduced driver run (fast): T_grid 100..300, runs/replicates trimmed
    small_grid = np.array([100, 200, 300])
    res["stats_small"] = {}
    for title, builder in sg.CASES.items():
        runs = min(sg.RUNS_BY_TITLE[title], 3)
        reps = min(sg.REPLICATES_BY_TITLE[title], 2)
        st = fd.evaluate_stream_with_stats(builder, small_grid, g_emp, runs=runs, replicates=reps, base_seed=0,
                                           stream_name=title)
        res["stats_small"][title] = {"runs": runs, "replicates": reps,
                                     "stats": {k: [[float(v).hex() for v in m], [float(v).hex() for v in c]]
                                               for k, (m, c) in st.items()}}
    if os.environ.get("GOLDEN_FULL", "1") == "1":   # the reference's full default workload (~40 s)
        res["stats_full"] = {}
        for title, builder in sg.CASES.items():
            st = fd.evaluate_stream_with_stats(builder, T_grid, g_emp, runs=sg.RUNS_BY_TITLE[title],
                                               replicates=sg.REPLICATES_BY_TITLE[title], base_seed=0, stream_name=title)
            res["stats_full"][title] = {k: [[float(v).hex() for v in m], [float(v).hex() for v in c]]
                                        for k, (m, c) in st.items()}
    json.dump(res, open(os.path.join(OUT, "driver_stats.json"), "w"), indent=0)


def load_ref_module(name, patch_dataclass=True):
    """Load /root/reference/<name>.py in memory with matplotlib / cvxpy stubbed (never called by what we run)."""
    for mod in ("matplotlib", "matplotlib.pyplot", "cvxpy"):
        sys.modules.setdefault(mod, MagicMock())
    src = open(os.path.join(REF, name + ".py")).read()
    if patch_dataclass:
        src = src.replace(
            "T_grid: IntArray = np.arange(100, 1100, 100, dtype=int)",
            "T_grid: IntArray = __import__('dataclasses').field(default_factory=lambda: np.arange(100, 1100, 100, dtype=int))")
    m = types.ModuleType(name)
    sys.modules[name] = m
    exec(compile(src, name + ".py", "exec"), m.__dict__)
    return m


def closed_form_actions(z, y):
    """Prefix actions fed to the reference's replay: v_k/||v_k|| (0 on ties) -- any (T+1, d) array would do, the
    reference function only replays them (exact_ftl.py:306-333)."""
    T, d = z.shape
    a = np.zeros((T + 1, d))
    v = np.zeros(d)
    for k in range(T):
        v = v + y[k] * z[k]
        n = math.sqrt(float(v @ v))
        a[k + 1] = v / n if n > 0 else 0.0
    return a


def exact_det():
    sys.modules.setdefault("cvxpy", MagicMock())      # imported at exact_ftl.py:10, never called by what runs below
    import exact_ftl as ex                            # the reference module itself (njit cache -> NUMBA_CACHE_DIR)
    cases = []

    def add(name, z, y, spec):
        z = np.ascontiguousarray(z, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        T, d = z.shape
        rng = np.random.default_rng(len(cases) + 1)
        acts = {"closed": closed_form_actions(z, y)}
        r = rng.standard_normal((T + 1, d)); r /= np.maximum(1.0, np.linalg.norm(r, axis=1, keepdims=True)); r[0] = 0.0
        acts["random"] = r
        rec = {"name": name, "spec": spec, "T": T, "d": d, "replay": {}, "ftrl": {}}
        for k, a in acts.items():
            res = ex.replay_exact_ftl(z, y, a)
            rec["replay"][k] = {"actions_seed": len(cases) + 1, "cum_loss": float(res.cum_loss).hex(), "regret": float(res.regret).hex(),
                                "comp_loss": float(res.comp_loss).hex()}
        comp = acts["closed"][T]
        for eta in (1.0, R2, 0.3):
            for which, cvec in (("closed", comp), ("random", acts["random"][T])):
                res = ex.run_ftrl(z, y, eta0=eta, comparator_action=cvec)
                res2 = ex._simulate_ftrl(z, y, eta0=eta, comparator_action=cvec)
                assert res.regret == res2.regret and res.cum_loss == res2.cum_loss
                rec["ftrl"][f"eta{eta!r}_{which}"] = {"cum_loss": float(res.cum_loss).hex(), "regret": float(res.regret).hex(),
                                                     "comp_loss": float(res.comp_loss).hex(),
                                                     "x_last": [float(v).hex() for v in res.x_last]}
        cases.append(rec)

    for (b, T, r, d) in [(0, 200, 1, 5), (0, 150, 2, 10), (4, 120, 0, 1), (4, 90, 3, 33), (2, 64, 7, 8), (2, 1, 0, 10)]:
        z, y = gpath(b, T, r, d)
        add(f"gpath_{b}_{T}_{r}_d{d}", z, y, {"family": "gpath", "base_seed": b, "T": T, "run": r, "d": d})
    for title, builder in sg.CASES.items():
        sampler = builder(run_seed=2025)
        z, y, _ = sampler(100, 0)
        add(f"case_{title}", z, y, {"family": title, "run_seed": 2025, "T": 100, "rep": 0, "d": 5})
    # the reference module's own smoke configuration (exact_ftl.py:460-475), FTRL half (the exact half needs CVXPY)
    rng = np.random.default_rng(0)
    T, d = 200, 10
    z = rng.normal(size=(T, d))
    z /= np.maximum(1.0, np.linalg.norm(z, axis=1, keepdims=True))
    w_true = rng.normal(size=d)
    w_true /= np.linalg.norm(w_true)
    y = np.sign(z @ w_true + 0.2 * rng.normal(size=T))
    add("smoke_main_d10", z, y, {"family": "smoke_main"})
    json.dump({"cases": cases, "note": "actions['closed'] = v_k/||v_k||; actions['random'] = default_rng(actions_seed) rows clipped to the unit ball, row 0 zero"},
              open(os.path.join(OUT, "exact_det.json"), "w"), indent=0)


def driver32_stats():
    dr = load_ref_module("driver")
    res = {}
    # float32 twin (algorithms.py:28-54) on knife-edge shapes: d = 1 clipped rows (|z| = 1 on a third of the rounds, the
    # projected action is +-1 or +-(1 - 2^-24), q - y is exactly 0 or one ulp) and d = 8
    res["f32_small_d"] = []
    for (b, T, r, d) in [(7, 300, 0, 1), (7, 257, 3, 1), (3, 200, 2, 8), (5, 129, 1, 2)]:
        z, y = gpath(b, T, r, d)
        zf, yf = z.astype(np.float32), y.astype(np.float32)
        rec = {"spec": {"family": "gpath", "base_seed": b, "T": T, "run": r, "d": d}, "results": {}}
        for flag in (0, 1):
            for eta in (R2, 0.25, 4.0):
                rec["results"][f"alg{flag}_eta{eta!r}"] = float(alg32.simulate_alg(zf, yf, flag, eta)).hex()
        res["f32_small_d"].append(rec)
    small_grid = np.array([100, 200, 300])
    g32 = alg32.empirical_worst_case_thresholds(small_grid, runs=40, base_seed=0)
    res["g_emp_runs40"] = {str(k): float(v).hex() for k, v in g32.items()}
    res["stats_small"] = {}
    for title, builder in sg.CASES.items():
        runs = min(sg.RUNS_BY_TITLE[title], 3)
        reps = min(sg.REPLICATES_BY_TITLE[title], 2)
        st = dr.evaluate_stream_with_stats(builder, small_grid, g32, runs=runs, replicates=reps, base_seed=0, stream_name=title)
        res["stats_small"][title] = {"runs": runs, "replicates": reps,
                                     "stats": {k: [[float(v).hex() for v in m], [float(v).hex() for v in c]] for k, (m, c) in st.items()}}
    json.dump(res, open(os.path.join(OUT, "driver32_stats.json"), "w"), indent=0)


if __name__ == "__main__":
    if "--only-new" in sys.argv:
        exact_det(); driver32_stats()
        print("exact_det.json / driver32_stats.json written"); sys.exit(0)
    rng_kat()
    sim_cases()
    iterates()
    driver_stats()
    exact_det()
    driver32_stats()
    print("golden fixtures written to", os.path.normpath(OUT))
