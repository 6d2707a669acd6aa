"""CPU tests: the oracle (oracle/oco_oracle.c) against golden vectors produced by
running the reference itself (tools/make_golden.py) and against numpy's RNG."""
import math

import numpy as np
import pytest

from conftest import exact_det_inputs, hexf, make_sequence
from oracle import oracle as O

R2 = math.sqrt(2.0)


def test_rng_known_answers(golden_rng):
    g = golden_rng
    for i, key in enumerate(g["keys"]):
        key = tuple(int(v) for v in key)
        assert (O.seedseq3(*key) == g[f"state_{i}"]).all()
        assert (O.rng_raw(key, 64) == g[f"raw_{i}"]).all()
        assert (O.rng_normals(key, 4096).view(np.uint64) == g[f"normal_{i}"].view(np.uint64)).all()
        assert (O.rng_doubles(key, 256) == g[f"random_{i}"]).all()
        assert (O.rng_choice_pm1(key, 333, 257) == g[f"choice_after_333_{i}"]).all()


def test_rng_tail_draws_bit_exact(golden_rng):
    """Tail normals go through the log1p restatement (glibc FMA variant)."""
    x = O.rng_normals((11, 22, 33), 2_000_000)
    pos = golden_rng["tail_pos"]
    assert (x[pos].view(np.uint64) == golden_rng["tail_val"].view(np.uint64)).all()
    assert np.bitwise_xor.reduce(x.view(np.uint64)) == golden_rng["long_sum_bits"][0]


def test_rng_matches_live_numpy():
    for key in [(5, 6, 7), (2025 * 3, 400, 14)]:
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(list(key))))
        assert (gen.standard_normal(50_000).view(np.uint64) == O.rng_normals(key, 50_000).view(np.uint64)).all()


def test_sequences_first_row(golden_sim):
    z, y = O.gen_gpath(0, 100, 0, 5)
    assert [float(v).hex() for v in z[0]] == golden_sim["extra"]["gpath_0_100_0_z0_hex"]
    assert [float(v) for v in y[:8]] == golden_sim["extra"]["gpath_0_100_0_y8"]


def test_simulators_bit_exact_vs_reference(golden_sim):
    n = 0
    for case in golden_sim["cases"]:
        z, y = make_sequence(case["spec"])
        assert z.shape == (case["T"], case["d"])
        for k, v in case["results"].items():
            if k.startswith("alg"):
                flag = int(k[3])
                eta = float(k.split("eta")[1])
                assert O.simulate_alg(z, y, flag, eta) == hexf(v), (case["name"], k)
            elif k.startswith("smart_th"):
                th = float(k[len("smart_th"):])
                assert O.simulate_smart_like(z, y, th, R2) == hexf(v), (case["name"], k)
            n += 1
    assert n > 300


def test_iterates_bit_exact(golden_iter):
    for name in ("d1", "d5"):
        b, T, r, d = (int(v) for v in golden_iter[f"{name}_key"])
        z, y = O.gen_gpath(b, T, r, d)
        for flag in (0, 1):
            res = O.simulate_alg(z, y, flag, R2, trace=True)
            assert (res["x_trace"].view(np.uint64) == golden_iter[f"{name}_alg{flag}_x"].view(np.uint64)).all()
            assert (res["theta"].view(np.uint64) == golden_iter[f"{name}_alg{flag}_theta"].view(np.uint64)).all()


def test_empirical_thresholds(golden_stats):
    want = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs5_seed3"].items()}
    assert O.empirical_worst_case_thresholds(np.array([100, 200, 300]), runs=5, base_seed=3) == want
    full = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs1000"].items()}
    got = O.empirical_worst_case_thresholds(np.arange(100, 600, 100), runs=1000, base_seed=0)
    assert got == {k: full[k] for k in got}
    assert full[1000] == 19.088517830594924


def test_batch_and_sweep_consistency():
    zs, ys = zip(*[O.gen_gpath(0, 200, r, 5) for r in range(6)])
    z = np.stack(zs)
    y = np.stack(ys)
    flags, etas = [1, 0, 0], [R2, R2, 0.5]
    reg = O.simulate_batch(z, y, flags, etas)
    for b in range(6):
        for a in range(3):
            assert reg[b, a] == O.simulate_alg(z[b], y[b], flags[a], etas[a])
    assert (O.regret_sweep(0, 200, 0, 6, 5, flags, etas) == reg).all()
    reg32 = O.simulate_batch(z.astype(np.float32), y.astype(np.float32), flags, etas)
    assert reg32[2, 1] == O.simulate_alg(z[2].astype(np.float32), y[2].astype(np.float32), 0, R2)
    sm, sw = O.smart_batch(z, y, [3.0, 1e9], [R2, R2])
    assert sm[1, 0] == O.simulate_smart_like(z[1], y[1], 3.0, R2)
    assert (sw[:, 1] == -1).all()


def test_exact_ftl_closed_form_known_answers():
    """published exact-FTL curves -- label flips T/4, switching 1.5/6.5/12.5."""
    for T, want in ((100, 25.0), (500, 125.0), (1000, 250.0)):
        z, y, _ = O.gen_flip(T)
        act = O.exact_ftl_l2_prefix(z, y)
        assert O.replay_actions(z, y, act)["regret"] == want
    for T, want in ((100, 1.5), (500, 6.5), (1000, 12.5)):
        z, y, _ = O.gen_switching(T)
        act = O.exact_ftl_l2_prefix(z, y)
        assert O.replay_actions(z, y, act)["regret"] == want
    with pytest.raises(ValueError):
        O.exact_ftl_l2_prefix(np.full((3, 2), 2.0), np.ones(3))


def test_exact_ftl_deterministic_half_vs_reference(golden_exact_det):
    """Oracle replay / FTRL-with-fixed-comparator against fixtures produced by running the reference's exact_ftl.py
    (cvxpy stubbed, never called): replay_exact_ftl (:306-333) and _simulate_ftrl / run_ftrl(comparator_action=...)
    (:230-277, 399-420).  The played losses are sequential njit sums -> bit-exact; the comparator loss is a BLAS
    matvec + pairwise np.sum in the reference -> 1e-12."""
    n = 0
    for i, case in enumerate(golden_exact_det["cases"]):
        z, y, acts = exact_det_inputs(case, i)
        T, d = z.shape
        for k, a in acts.items():
            got = O.replay_actions(z, y, a)
            want = {f: hexf(v) for f, v in case["replay"][k].items() if f != "actions_seed"}
            assert got["cum_loss"] == want["cum_loss"], (case["name"], k)
            assert abs(got["comp_loss"] - want["comp_loss"]) <= 1e-12 * max(1.0, want["comp_loss"])
            assert abs(got["regret"] - want["regret"]) <= 1e-12 * max(1.0, want["cum_loss"])
            n += 1
        for key, rec in case["ftrl"].items():
            eta = float(key.split("_")[0][3:]); which = key.split("_")[1]
            full = O.simulate_alg(z, y, 0, eta, full=True)
            assert full["cum_loss"] == hexf(rec["cum_loss"]), (case["name"], key)      # same FTRL trajectory as fast_algorithms
            cvec = acts[which][T]
            comp = 0.5 * float(np.abs(z @ cvec - y).sum())
            assert abs(comp - hexf(rec["comp_loss"])) <= 1e-12 * max(1.0, comp)
            n += 1
    assert n >= 80
