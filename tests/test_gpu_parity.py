"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI / the drop-in Python
modules, against (a) golden vectors produced by running the reference (tests/golden) and
(b) the CPU oracle on the same seeded inputs.  Bit-exact wherever the summation order is the
reference's (every d <= 8 kernel and the whole device generator); 1e-12 relative for the
warp-parallel reductions (d > 8, SMART prefix loss) as north_star states."""
import math

import numpy as np
import pytest

from conftest import hexf, make_sequence
from oracle import oracle as O

pytestmark = pytest.mark.gpu

R2 = math.sqrt(2.0)
REL = 1e-12   # north_star tolerance where the reduction order changes


@pytest.fixture(scope="module")
def eng():
    from online_convex_optimization_b200 import engine
    return engine


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


# ------------------------------------------------------------------ device generator
def test_device_rng_known_answers(eng, golden_rng):
    g = golden_rng
    for i, key in enumerate(g["keys"]):
        key = tuple(int(v) for v in key)
        assert (eng.rng_probe(4, key, 4) == g[f"state_{i}"]).all()
        assert (eng.rng_probe(0, key, 64) == g[f"raw_{i}"]).all()
        assert (bits(eng.rng_probe(1, key, 4096)) == bits(g[f"normal_{i}"])).all()
        assert (eng.rng_probe(2, key, 256) == g[f"random_{i}"]).all()
        assert (eng.rng_probe(3, key, 257, n_skip=333) == g[f"choice_after_333_{i}"]).all()


def test_device_rng_tail_and_log1p(eng, golden_rng):
    x = eng.rng_probe(1, (11, 22, 33), 2_000_000)
    pos = golden_rng["tail_pos"]
    assert (bits(x[pos]) == bits(golden_rng["tail_val"])).all()
    assert np.bitwise_xor.reduce(bits(x)) == golden_rng["long_sum_bits"][0]
    u = O.rng_doubles((3, 4, 5), 20000)
    assert (bits(eng.rng_probe(5, (3, 4, 5), 20000)) == bits(O.log1p(-u))).all()


def test_rcp_inrange_matches_division(eng):
    """The branch-free reciprocal the d = 1 paths use (row clip, FTL action, FTRL projection) returns the bits of
    IEEE 1.0 / r: 2^29 pseudo-random operands, half over [2^-1000, 2^1000], half over [1, 2^12), both signs."""
    for seed in (1, 2):
        out = eng.rng_probe(6, (seed, 0, 0), 1 << 16, n_skip=1 << 12)
        assert int(out.view(np.uint64)[0]) == 0


@pytest.mark.parametrize("d", [1, 2, 5, 7, 8, 13, 64, 100, 128])
def test_gen_gpath_bit_exact(eng, d):
    """Warp-per-sequence generator (jump-ahead PCG64, role walk over the failing draws, numpy's pairwise row norm):
    rows and labels equal the oracle's bit for bit for any d <= 128; the round-1 thread-per-sequence kernel agrees."""
    T = 257 if d < 64 else 96
    z, y = eng.gen_sequences("gpath", 5, T, d, key_a=3, key_c0=7)
    for b in range(5):
        zo, yo = O.gen_gpath(3, T, 7 + b, d)
        assert (bits(z[b]) == bits(zo)).all() and (y[b] == yo).all()
    z32, y32 = eng.gen_sequences("gpath", 2, T, d, key_a=3, key_c0=7, dtype=np.float32)
    assert (z32[1] == O.gen_gpath(3, T, 8, d)[0].astype(np.float32)).all()
    if d in (1, 5, 64):                                           # legacy kernel (OCO_FILL_LEGACY=1) on the shapes it supports
        import os
        os.environ["OCO_FILL_LEGACY"] = "1"
        try:
            zl, yl = eng.gen_sequences("gpath", 5, T, d, key_a=3, key_c0=7)
        finally:
            del os.environ["OCO_FILL_LEGACY"]
        assert (bits(zl) == bits(z)).all() and (yl == y).all()


def test_gen_gpath_long_sequences_and_ragged_tails(eng):
    """Long streams (tail draws, carried wedge / tail candidates across 32-raw rounds) and the lane-0 finish of the
    last < 34 normals: T*d not a multiple of anything, T = 1, T*d < 34."""
    for d, T, B in ((64, 2000, 2), (5, 10_000, 2), (9, 3001, 2), (3, 1, 2), (2, 2, 1), (33, 1, 1), (1, 35, 3), (128, 517, 1)):
        z, y = eng.gen_sequences("gpath", B, T, d, key_a=11, key_c0=5)
        for b in range(B):
            zo, yo = O.gen_gpath(11, T, 5 + b, d)
            assert (bits(z[b]) == bits(zo)).all() and (y[b] == yo).all(), (d, T, b)


def test_gen_case_families_bit_exact(eng):
    T = 300
    keys = np.array([[2025, 13], [2025, 14], [4050, 13], [6075, 20]], dtype=np.uint64)
    z, y = eng.gen_sequences("iid", 4, T, 5, keys=keys)
    for b, (rs, st) in enumerate(keys):
        zo, yo, _ = O.gen_iid(int(rs), T, int(st) - 13)
        assert (z[b].view(np.uint32) == zo.view(np.uint32)).all() and (y[b] == yo).all()
    keys[:, 1] += 10
    z, y = eng.gen_sequences("noisy", 4, T, 5, keys=keys, p_flip=0.10)
    for b, (rs, st) in enumerate(keys):
        zo, yo, _ = O.gen_iid(int(rs), T, int(st) - 23, noisy_p=0.10)
        assert (z[b].view(np.uint32) == zo.view(np.uint32)).all() and (y[b] == yo).all()
    z, y = eng.gen_sequences("flip", 2, 101, 5)
    zo, yo, _ = O.gen_flip(101)
    assert (z[1] == zo).all() and (y[1] == yo).all()
    z, y = eng.gen_sequences("switch", 1, 333, 5, block_len=20)
    zo, yo, _ = O.gen_switching(333)
    assert (z[0] == zo).all() and (y[0] == yo).all()


# ------------------------------------------------------------------ FTL / FTRL
def test_simulate_alg_golden_bit_exact(eng, golden_sim):
    """Every golden case of d <= 8 through the C ABI: results equal the reference's bit for bit."""
    n = 0
    for case in golden_sim["cases"]:
        if case["d"] > 8:
            continue
        z, y = make_sequence(case["spec"])
        flags, etas, want = [], [], []
        for k, v in case["results"].items():
            if k.startswith("alg"):
                flags.append(int(k[3])); etas.append(float(k.split("eta")[1])); want.append(hexf(v))
        got = eng.simulate_batch(z, y, flags, etas)[0]
        assert (bits(got) == bits(np.array(want))).all(), case["name"]
        n += len(want)
    assert n > 200


def test_simulate_alg_wide_d_tolerance(eng, golden_sim):
    for case in golden_sim["cases"]:
        if case["d"] <= 8:
            continue
        z, y = make_sequence(case["spec"])
        flags, etas, want = [], [], []
        for k, v in case["results"].items():
            if k.startswith("alg"):
                flags.append(int(k[3])); etas.append(float(k.split("eta")[1])); want.append(hexf(v))
        want = np.array(want)
        for mode in ("two_pass", "linear"):
            got = eng.simulate_batch(z, y, flags, etas, comparator=mode)[0]
            assert np.all(np.abs(got - want) <= REL * np.abs(want)), (case["name"], mode, got, want)


def test_iterates_bit_exact(eng, golden_iter):
    """BASELINE config 2's 'bit-exact iterate check': played x_t and final theta, d = 1 and d = 5."""
    for name in ("d1", "d5"):
        b, T, r, d = (int(v) for v in golden_iter[f"{name}_key"])
        z, y = O.gen_gpath(b, T, r, d)
        res = eng.simulate_batch(z, y, [0, 1], [R2, R2], full=True, trace=True)
        for a, flag in enumerate((0, 1)):
            assert (bits(res["x_trace"][0, a]) == bits(golden_iter[f"{name}_alg{flag}_x"])).all()
            assert (bits(res["theta"][0, a]) == bits(golden_iter[f"{name}_alg{flag}_theta"])).all()


@pytest.mark.parametrize("d,T,B", [(1, 1000, 37), (3, 200, 130), (5, 1000, 300), (8, 64, 5), (5, 1, 3), (2, 7, 129)])
def test_simulate_batch_vs_oracle(eng, d, T, B):
    rng = np.random.default_rng(d * 1000 + T)
    z = rng.standard_normal((B, T, d)) * 0.7          # NOT clipped: exercises |q| > 1 and projection
    y = rng.choice([-1.0, 1.0, 0.0, 0.5], size=(B, T))
    flags = [0, 1, 0, 0, 1, 0, 0]
    etas = [R2, R2, 0.1, 10.0, 3.0, 1.0, 0.3]
    want = O.simulate_batch(z, y, flags, etas, full=True)
    got = eng.simulate_batch(z, y, flags, etas, full=True)
    for k in ("regret", "cum_loss", "comp_loss", "theta"):
        assert (bits(got[k]) == bits(want[k])).all(), k
    z32, y32 = z.astype(np.float32), y.astype(np.float32)
    assert (bits(eng.simulate_batch(z32, y32, flags[:2], etas[:2])) == bits(O.simulate_batch(z32, y32, flags[:2], etas[:2]))).all()


def test_simulate_wide_vs_oracle(eng):
    for d, T, B in ((64, 400, 9), (33, 100, 4), (100, 50, 3), (16, 300, 5)):
        zs, ys = zip(*[O.gen_gpath(5, T, r, d) for r in range(B)])
        z, y = np.stack(zs), np.stack(ys)
        flags, etas = [0, 1, 0], [R2, R2, 0.2]
        want = O.simulate_batch(z, y, flags, etas)
        for mode in ("two_pass", "linear"):
            got = eng.simulate_batch(z, y, flags, etas, comparator=mode)
            assert np.all(np.abs(got - want) <= REL * np.abs(want)), (d, mode)
        # rows outside the unit ball: linear mode must detect it and fall back
        z2 = z * 3.0
        want2 = O.simulate_batch(z2, y, flags, etas)
        got2 = eng.simulate_batch(z2, y, flags, etas, comparator="linear")
        assert np.all(np.abs(got2 - want2) <= REL * np.abs(want2))


# ------------------------------------------------------------------ SMART
def test_smart_golden(eng, golden_sim):
    n = 0
    for case in golden_sim["cases"]:
        ths = [k for k in case["results"] if k.startswith("smart_th")]
        if not ths or case["d"] > 8:
            continue
        z, y = make_sequence(case["spec"])
        th = [float(k[len("smart_th"):]) for k in ths]
        want = np.array([hexf(case["results"][k]) for k in ths])
        got, sw = eng.smart_batch(z, y, th, [R2] * len(th))
        assert np.all(np.abs(got[0] - want) <= REL * np.maximum(np.abs(want), 1.0)), (case["name"], got, want)
        for a, t in enumerate(th):
            assert sw[0, a] == O.simulate_smart_like(z, y, t, R2, full=True)["switch_t"]
        if case["spec"]["family"] in ("Label flips", "Switching leaders"):
            assert (bits(got[0]) == bits(want)).all()    # dyadic data: any summation order is exact
        n += len(th)
    assert n > 60


# ------------------------------------------------------------------ fused seeded sweep
@pytest.mark.parametrize("d,T,B,A", [(1, 500, 200, 9), (5, 300, 150, 1), (5, 200, 140, 2), (2, 100, 10, 4), (5, 100, 130, 9),
                                     (6, 90, 70, 2), (7, 65, 40, 5), (8, 120, 131, 9)])      # d = 6, 7, 8: two-pass fused sweep (round 2)
def test_regret_sweep_bit_exact(eng, d, T, B, A):
    flags = ([1] + [0] * 8)[:A]
    etas = [R2 * s for s in (1.0, 0.125, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0)][:A]
    want = O.regret_sweep(7, T, 11, B, d, flags, etas)
    got = eng.regret_sweep("gpath", B, T, d, flags, etas, key_a=7, key_c0=11)
    assert (bits(got["regret"]) == bits(want)).all()
    st = got["stats"]
    assert (st[:, 0] == want.max(axis=0)).all() and (st[:, 3] == B).all()
    assert np.allclose(st[:, 1], want.sum(axis=0), rtol=1e-13) and np.allclose(st[:, 2], (want ** 2).sum(axis=0), rtol=1e-13)


def test_regret_sweep_case_families(eng):
    T = 200
    keys = np.array([[2025 * (r + 1), 13 + rep] for r in range(3) for rep in range(4)], dtype=np.uint64)
    got = eng.regret_sweep("iid", len(keys), T, 5, [0, 1], [R2, R2], keys=keys)["regret"]
    for b, (rs, st) in enumerate(keys):
        z, y, _ = O.gen_iid(int(rs), T, int(st) - 13)
        assert got[b, 0] == O.simulate_alg(z, y, 0, R2) and got[b, 1] == O.simulate_alg(z, y, 1, R2)
    keys[:, 1] += 10
    got = eng.regret_sweep("noisy", len(keys), T, 5, [0, 1], [R2, R2], keys=keys, p_flip=0.10)["regret"]
    for b, (rs, st) in enumerate(keys):
        z, y, _ = O.gen_iid(int(rs), T, int(st) - 23, noisy_p=0.10)
        assert got[b, 0] == O.simulate_alg(z, y, 0, R2) and got[b, 1] == O.simulate_alg(z, y, 1, R2)


def test_empirical_thresholds_golden(eng, golden_stats):
    """The reference's own g(T) table (fast_driver default: runs=1000, T=100..1000), bit for bit."""
    from online_convex_optimization_b200 import fast_algorithms as fa
    full = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs1000"].items()}
    assert fa.empirical_worst_case_thresholds(np.arange(100, 1100, 100), runs=1000, base_seed=0) == full
    small = {int(k): hexf(v) for k, v in golden_stats["g_emp_runs5_seed3"].items()}
    assert fa.empirical_worst_case_thresholds(np.array([100, 200, 300]), runs=5, base_seed=3) == small


# ------------------------------------------------------------------ drop-in module + edges
def test_dropin_module_signatures(golden_sim):
    from online_convex_optimization_b200 import fast_algorithms as fa
    from online_convex_optimization_b200 import sequence_generation as sg
    case = next(c for c in golden_sim["cases"] if c["name"].startswith("case_Massart"))
    spec = case["spec"]
    sampler = sg.CASES[spec["family"]](run_seed=spec["run_seed"])
    z, y, u = sampler(spec["T"], spec["rep"])
    assert z.dtype == np.float32 and y.dtype == np.float32 and u.shape == (5,)
    assert fa.simulate_alg(z, y, alg_flag=0, eta0=R2) == hexf(case["results"][f"alg0_eta{R2!r}"])
    assert fa.simulate_alg(z, y, alg_flag=1, eta0=R2) == hexf(case["results"][f"alg1_eta{R2!r}"])
    want = hexf(case["results"][f"smart_th{math.sqrt(2 * spec['T'])!r}"])
    assert abs(fa.simulate_SMART(z, y) - want) <= REL * max(1.0, abs(want))
    want = hexf(case["results"]["smart_th3.0"])
    assert abs(fa.simulate_empirical_g_SMART(z, y, 3.0) - want) <= REL * max(1.0, abs(want))
    assert sg.RUNS_BY_TITLE["Massart noise 10%"] == 48 and sg.REPLICATES_BY_TITLE["Massart noise 10%"] == 20
    assert isinstance(fa._rng(0, 1, 2), np.random.Generator)


def test_edge_cases(eng):
    assert eng.simulate_batch(np.zeros((0, 5, 3)), np.zeros((0, 5)), [0], [1.0]).shape == (0, 1)
    r = eng.simulate_batch(np.zeros((2, 0, 3)), np.zeros((2, 0)), [0, 1], [1.0, 1.0])
    assert (r == 0.0).all()
    r = eng.simulate_batch(np.zeros((3, 4, 2)), np.zeros((3, 4)), [0, 1], [1.0, 1.0])   # all-zero rows: q == y
    assert (r == 0.0).all()
    got = eng.regret_sweep("gpath", 0, 10, 5, [0], [1.0])
    assert got["regret"].shape == (0, 1)
    g, _ = eng.g_sweep([10, 20], runs=0)
    assert (g == 0.0).all()
    assert (eng.regret_sweep("gpath", 3, 0, 5, [0, 1], [1.0, 1.0])["regret"] == 0.0).all()          # T = 0
    z0, y0 = eng.gen_sequences("gpath", 2, 0, 5)
    assert z0.shape == (2, 0, 5) and y0.shape == (2, 0)
    with pytest.raises(ValueError):
        eng.simulate_batch(np.zeros((2, 4, 3)), np.zeros((2, 5)), [0], [1.0])
    with pytest.raises(ValueError):
        eng.simulate_batch(np.zeros((2, 4, 300)), np.zeros((2, 4)), [0], [1.0])
    with pytest.raises(ValueError):
        eng.smart_batch(np.zeros((1, 4, 3)), np.zeros((1, 4)), [1.0] * 5, [1.0] * 5)
    with pytest.raises(ValueError):
        eng.regret_sweep("gpath", 4, 10, 9, [0], [1.0])          # fused sweeps: d <= 8 (g-path)


def test_full_size_properties(eng):
    """BASELINE-size (T = 1e4) checks that do not need the oracle: determinism, independence of
    batch position / grid shape, and the closed form FTL regret on label flips (T/4)."""
    T, B = 10_000, 2048
    a = eng.regret_sweep("gpath", B, T, 1, [1, 0, 0], [R2, R2, 0.5], key_a=0, key_c0=0)["regret"]
    b = eng.regret_sweep("gpath", 300, T, 1, [0], [R2], key_a=0, key_c0=1000)["regret"]
    assert (bits(a[1000:1300, 1]) == bits(b[:, 0])).all()
    assert (bits(a[:4]) == bits(O.regret_sweep(0, T, 0, 4, 1, [1, 0, 0], [R2, R2, 0.5]))).all()
    z, y = eng.gen_sequences("flip", 1, T, 5)
    assert eng.simulate_batch(z, y, [1], [R2])[0, 0] == T / 4


# ------------------------------------------------------------------ float32 compute mode (algorithms.py)
def test_float32_mode_matches_numpy_reference(eng, golden_sim):
    """The optional fp32 mode (algorithms.py semantics): 1e-5 relative to the reference's float32 results."""
    from online_convex_optimization_b200 import algorithms as alg32
    n = 0
    for case in golden_sim["cases"]:
        if "f32_alg0" not in case["results"]:
            continue
        z, y = make_sequence(case["spec"])
        z32, y32 = np.asarray(z, dtype=np.float32), np.asarray(y, dtype=np.float32)
        for flag in (0, 1):
            want = hexf(case["results"][f"f32_alg{flag}"])
            got = alg32.simulate_alg(z32, y32, flag, R2)
            assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (case["name"], flag, got, want)
            n += 1
    assert n >= 30
    with pytest.raises(ValueError):
        eng.simulate_batch(np.zeros((1, 4, 2)), np.zeros((1, 4)), [0], [1.0], compute="f32")   # needs float32 inputs
    g = alg32.empirical_worst_case_thresholds(np.array([100]), runs=50, base_seed=0)
    g64 = O.empirical_worst_case_thresholds(np.array([100]), runs=50, base_seed=0)
    assert abs(g[100] - g64[100]) <= 1e-5 * g64[100]


def test_small_d_linear_comparator(eng):
    """d <= 8 streamed kernel, single-pass comparator: 1e-12 on clipped rows, exact fallback otherwise."""
    for d, T, B in ((5, 400, 70), (1, 1000, 33), (8, 100, 40)):
        zs, ys = zip(*[O.gen_gpath(6, T, r, d) for r in range(B)])
        z, y = np.stack(zs), np.stack(ys)
        flags, etas = [0, 1, 0, 0], [R2, R2, 0.2, 5.0]
        want = O.simulate_batch(z, y, flags, etas)
        got = eng.simulate_batch(z, y, flags, etas, comparator="linear")
        assert np.all(np.abs(got - want) <= REL * np.maximum(1.0, np.abs(want))), d
        z2 = z * 2.5                                   # outside the unit ball: must fall back to the exact pass
        want2 = O.simulate_batch(z2, y, flags, etas)
        got2 = eng.simulate_batch(z2, y, flags, etas, comparator="linear")
        assert (bits(got2) == bits(want2)).all(), d


def test_smart_linear_mode(eng):
    """O(T) SMART (linear regime): same switch times and regrets (1e-12) as the oracle; exact fallback."""
    T, B = 600, 40
    zs, ys = zip(*[O.gen_iid(2025 * (r + 1), T, 0)[:2] for r in range(B // 2)] + [O.gen_gpath(2, T, r, 5) for r in range(B // 2)])
    z, y = np.stack([np.asarray(a, dtype=np.float64) for a in zs]), np.stack([np.asarray(a, dtype=np.float64) for a in ys])
    # (a threshold of exactly 0.5 would be a knife edge: with a unit-norm first row the switch test at t = 0 is
    #  0.5*||z_0|| >= 0.5, decided by the last bit of the row norm -- any reordering may flip it)
    th, et = [math.sqrt(2 * T), 3.0, 0.75], [R2, R2, R2]
    want, want_sw = O.smart_batch(z, y, th, et)
    got, sw = eng.smart_batch(z, y, th, et, mode="linear")
    assert (sw == want_sw).all()
    assert np.all(np.abs(got - want) <= REL * np.maximum(1.0, np.abs(want)))
    z2 = z.copy(); z2[::3] *= 1.7                        # some sequences leave the unit ball -> exact kernel
    want2, want_sw2 = O.smart_batch(z2, y, th, et)
    got2, sw2 = eng.smart_batch(z2, y, th, et, mode="linear")
    assert (sw2 == want_sw2).all()
    assert np.all(np.abs(got2 - want2) <= REL * np.maximum(1.0, np.abs(want2)))
    zf, yf, _ = O.gen_flip(1000)
    assert eng.smart_batch(zf, yf, [math.sqrt(2000)], [R2], mode="linear")[0][0, 0] == O.simulate_smart_like(zf, yf, math.sqrt(2000), R2)


def test_regret_sweep_many_variants_and_tiny_shapes(eng):
    """More FTRL variants than one thread holds (several groups), duplicate FTL columns, T = 1, B = 1."""
    flags = [0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0]          # 16 FTRL (2 groups of 8) + 3 FTL columns
    etas = [0.05 * (i + 1) * R2 for i in range(len(flags))]
    want = O.regret_sweep(3, 150, 5, 70, 2, flags, etas)
    got = eng.regret_sweep("gpath", 70, 150, 2, flags, etas, key_a=3, key_c0=5)
    assert (bits(got["regret"]) == bits(want)).all()
    assert (got["stats"][:, 0] == want.max(axis=0)).all()
    for T, B in ((1, 5), (2, 1), (33, 1)):
        want = O.regret_sweep(0, T, 0, B, 5, [0, 1], [R2, R2])
        got = eng.regret_sweep("gpath", B, T, 5, [0, 1], [R2, R2])["regret"]
        assert (bits(got) == bits(want)).all(), (T, B)
    keys = np.array([[2025, 13 + r] for r in range(9)], dtype=np.uint64)
    got = eng.regret_sweep("iid", 9, 120, 5, [1] + [0] * 8, [R2 * 2.0 ** (k - 4) for k in range(9)], keys=keys)["regret"]
    for b in range(9):
        z, y, _ = O.gen_iid(2025, 120, b)
        for a in range(9):
            assert got[b, a] == O.simulate_alg(z, y, 1 if a == 0 else 0, R2 * 2.0 ** (a - 4))


def test_sweep_d1_candidate_sums_equal_literal_second_pass(eng, monkeypatch):
    """d = 1: the comparator losses picked from the four candidate running sums are the bits of the literal second
    pass (which finite theta never needs; OCO_SWEEP_FORCE_SECOND_PASS=1 forces it, labels regenerated from the stream)."""
    alg, eta = [1] + [0] * 8, [R2] + [R2 * s for s in (0.125, 0.25, 0.5, 1, 2, 4, 8, 16)]
    for fam, T in (("gpath", 700), ("gpath", 33), ("iid", 257)):
        kw = dict(key_a=3, key_c0=5) if fam == "gpath" else dict(key_a=2025, key_c0=13)
        a = eng.regret_sweep(fam, 300, T, 1, alg, eta, full=True, **kw)
        monkeypatch.setenv("OCO_SWEEP_FORCE_SECOND_PASS", "1")
        b = eng.regret_sweep(fam, 300, T, 1, alg, eta, full=True, **kw)
        monkeypatch.delenv("OCO_SWEEP_FORCE_SECOND_PASS")
        for k in ("regret", "cum_loss", "comp_loss"):
            assert (a[k].view(np.uint64) == b[k].view(np.uint64)).all(), (fam, T, k)


def test_sweep_large_sample_bit_exact(eng):
    """10^8 device-generated normals (about 8.5e5 wedge tests, 2.8e4 tail draws through log1p) feeding
    2e5 fused FTRL trajectories: every regret equals the CPU oracle's bit for bit."""
    B, T, d = 200_000, 100, 5
    want = O.regret_sweep(12345, T, 777, B, d, [0], [R2])
    got = eng.regret_sweep("gpath", B, T, d, [0], [R2], key_a=12345, key_c0=777)["regret"]
    assert (bits(got) == bits(want)).all()


# ------------------------------------------------------------------ single-pass (linear) comparator in the fused sweep
def _sweep_linear_check(eng, d, T, B, flags, etas, key_a, key_c0):
    """cum_loss bit-exact; regret / comp_loss within 1e-12 * max(1, cum_loss) of the oracle (the reference's own
    sequential comparator sum carries that much rounding noise; north_star's tolerance for d > 1)."""
    want = O.regret_sweep(key_a, T, key_c0, B, d, flags, etas)
    got = eng.regret_sweep("gpath", B, T, d, flags, etas, key_a=key_a, key_c0=key_c0, full=True, comparator="linear")
    if d <= 5:
        ex = eng.regret_sweep("gpath", B, T, d, flags, etas, key_a=key_a, key_c0=key_c0, full=True)
        assert (bits(ex["regret"]) == bits(want)).all()
    else:   # d = 6, 7, 8: the fused sweep exists in the single-pass mode only; the exact reference values come from the oracle
        zs, ys = zip(*[O.gen_gpath(key_a, T, key_c0 + b, d) for b in range(B)])
        full = O.simulate_batch(np.stack(zs), np.stack(ys), flags, etas, full=True)
        ex = {"cum_loss": full["cum_loss"], "comp_loss": full["comp_loss"]}
        assert (bits(full["regret"]) == bits(want)).all()
    assert (bits(got["cum_loss"]) == bits(ex["cum_loss"])).all()
    scale = np.maximum(1.0, ex["cum_loss"])
    assert np.all(np.abs(got["regret"] - want) <= REL * scale), np.max(np.abs(got["regret"] - want) / scale)
    assert np.all(np.abs(got["comp_loss"] - ex["comp_loss"]) <= REL * scale)
    st = got["stats"]
    assert (st[:, 0] == got["regret"].max(axis=0)).all() and (st[:, 3] == B).all()
    assert np.allclose(st[:, 1], got["regret"].sum(axis=0), rtol=1e-13)
    return np.max(np.abs(got["regret"] - want) / np.maximum(np.abs(want), 1e-300))


@pytest.mark.parametrize("d,T,B,A", [(5, 300, 150, 1), (5, 201, 140, 2), (2, 100, 10, 4), (3, 257, 70, 9), (4, 64, 129, 5),
                                     (5, 1, 3, 1), (5, 2, 1, 2), (2, 150, 70, 19), (8, 120, 131, 2), (7, 90, 40, 1), (6, 64, 33, 5),
                                     (8, 70, 40, 12), (7, 65, 33, 9), (6, 97, 35, 10)])     # NF = 8: d = 7, 8 read the scale table directly
def test_regret_sweep_linear_comparator(eng, d, T, B, A):
    flags = ([0, 1] + [0] * 7 + [1] + [0] * 9)[:A]
    etas = [R2 * s for s in (1.0, 1.0, 0.125, 0.25, 0.5, 2.0, 4.0, 8.0, 16.0, 1.0, .1, .2, .3, .4, .6, .7, .8, .9, 1.1)][:A]
    _sweep_linear_check(eng, d, T, B, flags, etas, 7, 11)


def test_regret_sweep_linear_comparator_full_T(eng):
    """BASELINE config 5's shape at full length (T = 1e4, d = 5): both comparator modes against the oracle, several CTA
    rounds (the generate / play software pipeline: 1300 sequences = 11 chunks on few CTAs is exercised by B > 128)."""
    rel = _sweep_linear_check(eng, 5, 10_000, 12, [0], [R2], 0, 0)
    assert rel <= 5e-12          # observed ~6e-13 relative to the regret itself
    _sweep_linear_check(eng, 5, 10_000, 4, [0, 1], [R2, R2], 0, 999_996)
    # many rounds per CTA at a moderate length
    _sweep_linear_check(eng, 5, 64, 128 * 148 * 5 * 2 + 77, [0], [R2], 3, 0)


def test_regret_sweep_linear_forced_second_pass_is_bit_exact(eng, monkeypatch):
    """The fallback every unverified sequence takes (literal second pass over regenerated rows) gives the reference's bits."""
    monkeypatch.setenv("OCO_SWEEP_FORCE_SECOND_PASS", "1")
    for d, T, B, A in ((5, 120, 70, 2), (3, 90, 40, 4), (2, 33, 200, 1)):
        flags = ([0, 1] + [0] * 7)[:A]; etas = [R2 * s for s in (1.0, 1.0, 0.125, 0.25)][:A]
        want = O.regret_sweep(3, T, 5, B, d, flags, etas)
        got = eng.regret_sweep("gpath", B, T, d, flags, etas, key_a=3, key_c0=5, comparator="linear")
        assert (bits(got["regret"]) == bits(want)).all()
    monkeypatch.delenv("OCO_SWEEP_FORCE_SECOND_PASS")


def test_g_sweep_linear_matches_exact(eng):
    """empirical_worst_case_thresholds with comparator="linear": the g(T) maxima agree with the bit-exact table."""
    Tg = np.array([100, 300, 1000])
    g0, st0 = eng.g_sweep(Tg, runs=300, base_seed=0)
    g1, st1 = eng.g_sweep(Tg, runs=300, base_seed=0, comparator="linear")
    assert np.all(np.abs(g1 - g0) <= REL * Tg / 2) and np.all(g1 > 0)
    from online_convex_optimization_b200 import fast_algorithms as fa
    ga = fa.empirical_worst_case_thresholds(Tg, runs=300, base_seed=0)
    gb = fa.empirical_worst_case_thresholds(Tg, runs=300, base_seed=0, comparator="linear")
    assert [ga[int(t)] for t in Tg] == list(g0) and [gb[int(t)] for t in Tg] == list(g1)


# ------------------------------------------------------------------ linear modes: rows slightly outside the unit ball
@pytest.mark.parametrize("d", [2, 5, 64])
def test_linear_modes_reject_over_unit_rows(eng, d):
    """Rows with ||z||^2 = 1 + 9e-7 aligned with the comparator (the round-1 counter-example: the identity
    0.5|q - y| = 0.5(1 - y q) fails for |q| > 1 by ~4.5e-7 per row).  The precondition is an error BUDGET now
    (sum (||z||^2 - 1)+ <= 1e-13 T), so such sequences take the literal pass: every linear mode agrees with the
    oracle to 1e-12; float32-normalised rows (||z||^2 up to 1 + 1e-7) likewise."""
    T = 1001
    c = math.sqrt(1.0 + 9e-7)
    z = np.zeros((3, T, d)); z[:, :, 0] = c
    y = np.ones((3, T)); y[1, ::7] = -1.0; y[2, ::2] = -1.0
    flags, etas = [1, 0, 0], [R2, R2, 0.3]
    want = O.simulate_batch(z, y, flags, etas)
    for mode in ("two_pass", "linear"):
        got = eng.simulate_batch(z, y, flags, etas, comparator=mode)
        assert np.all(np.abs(got - want) <= REL * np.maximum(1.0, np.abs(want))), (d, mode, got, want)
    if d <= 8:
        th = [math.sqrt(2 * T), 3.0]
        wr, ws = O.smart_batch(z, y, th, [R2, R2])
        for mode in ("exact", "linear"):
            gr, gs = eng.smart_batch(z, y, th, [R2, R2], mode=mode)
            assert (gs == ws).all() and np.all(np.abs(gr - wr) <= REL * np.maximum(1.0, np.abs(wr))), (d, mode)
    # float32-normalised random rows widened to float64: norms up to 1 + 1e-7
    rng = np.random.default_rng(5)
    zf = rng.standard_normal((4, 300, d)).astype(np.float32)
    zf /= np.maximum(np.linalg.norm(zf, axis=2, keepdims=True), 1.0).astype(np.float32)
    yf = np.where(rng.random((4, 300)) < 0.5, -1.0, 1.0).astype(np.float32)
    want = O.simulate_batch(zf.astype(np.float64), yf.astype(np.float64), flags, etas)
    got = eng.simulate_batch(zf, yf, flags, etas, comparator="linear")
    assert np.all(np.abs(got - want) <= REL * np.maximum(1.0, np.abs(want)))


# ------------------------------------------------------------------ SMART at full length
def test_smart_full_length_T1e4(eng):
    """simulate_SMART_like at T = 1e4, d = 5 (fast_algorithms.py:118-164): regrets (1e-12) and SWITCH TIMES against the
    oracle's literal O(T^2) loop on 10 sequences -- random g-path rows with thresholds that switch early / mid /
    late / never, and the two adversarial families (label flips switch at ~4*thresh, one of them at t ~ 8000)."""
    T, d = 10_000, 5
    zs, ys = zip(*[O.gen_gpath(0, T, r, d) for r in range(8)])
    zf, yf, _ = O.gen_flip(T)
    zw, yw, _ = O.gen_switching(T)
    z = np.stack(list(zs) + [np.asarray(zf, dtype=np.float64), np.asarray(zw, dtype=np.float64)])
    y = np.stack(list(ys) + [np.asarray(yf, dtype=np.float64), np.asarray(yw, dtype=np.float64)])
    th = [3.0, 15.0, 2000.0, math.sqrt(2 * T)]
    et = [R2] * 4
    wr, ws = O.smart_batch(z, y, th, et)
    assert (ws[:8, 0] >= 0).all() and (ws[:8, 3] < 0).all()            # random rows: early switch / never
    assert 7000 < ws[8, 2] < 9000 and 0 <= ws[8, 3] < 1000                # label flips: late and early switches
    for mode in ("exact", "linear"):
        gr, gs = eng.smart_batch(z, y, th, et, mode=mode)
        assert (gs == ws).all(), (mode, gs, ws)
        assert np.all(np.abs(gr - wr) <= REL * np.maximum(1.0, np.abs(wr))), (mode, np.max(np.abs(gr - wr)))
    # the drop-in single-sequence entry points at full length
    from online_convex_optimization_b200 import fast_algorithms as fa
    assert abs(fa.simulate_SMART(z[8], y[8]) - wr[8, 3]) <= REL * max(1.0, abs(wr[8, 3]))
    assert abs(fa.simulate_empirical_g_SMART(z[0], y[0], 15.0) - wr[0, 1]) <= REL * max(1.0, abs(wr[0, 1]))


@pytest.mark.parametrize("d,T,B", [(10, 400, 9), (33, 300, 6), (64, 250, 5), (100, 120, 4), (256, 60, 3)])
def test_smart_wide_d(eng, d, T, B):
    """simulate_SMART_like for d > 8 (the reference has no limit on d; kernels_wide_smart.cu): regrets within 1e-12
    and SWITCH TIMES against the oracle's literal O(T^2) loop -- g-path rows (unit ball: the pruning bound is
    active) with thresholds that switch early / mid / never, the same rows scaled out of the ball with real-valued
    labels (no pruning: every round takes the literal prefix loss), float32 inputs, and both modes."""
    zs, ys = zip(*[O.gen_gpath(4, T, r, d) for r in range(B)])
    z, y = np.stack(zs), np.stack(ys)
    th, et = [1.5, 6.0, math.sqrt(2 * T), 1e9], [R2, 0.5 * R2, R2, R2]
    wr, ws = O.smart_batch(z, y, th, et)
    assert (ws[:, 0] >= 0).all() and (ws[:, 3] < 0).all()
    for mode in ("exact", "linear"):
        gr, gs = eng.smart_batch(z, y, th, et, mode=mode)
        assert (gs == ws).all(), (mode, gs, ws)
        assert np.all(np.abs(gr - wr) <= REL * np.maximum(1.0, np.abs(wr))), (mode, np.max(np.abs(gr - wr)))
    z2 = z * 1.9; y2 = y * np.linspace(0.5, 1.5, T)[None, :]
    wr2, ws2 = O.smart_batch(z2, y2, th, et)
    gr2, gs2 = eng.smart_batch(z2, y2, th, et)
    assert (gs2 == ws2).all()
    assert np.all(np.abs(gr2 - wr2) <= REL * np.maximum(1.0, np.abs(wr2)))
    z32, y32 = z.astype(np.float32), y.astype(np.float32)
    wr3, ws3 = O.smart_batch(z32, y32, th[:2], et[:2])
    gr3, gs3 = eng.smart_batch(z32, y32, th[:2], et[:2])
    assert (gs3 == ws3).all() and np.all(np.abs(gr3 - wr3) <= REL * np.maximum(1.0, np.abs(wr3)))
    # the drop-in single-sequence entry points
    from online_convex_optimization_b200 import fast_algorithms as fa
    assert abs(fa.simulate_SMART(z[0], y[0]) - wr[0, 2]) <= REL * max(1.0, abs(wr[0, 2]))
    assert abs(fa.simulate_SMART_like(z[1], y[1], 6.0, 0.5 * R2) - wr[1, 1]) <= REL * max(1.0, abs(wr[1, 1]))


def test_fused_sweep_random_shapes(eng):
    """Randomised shapes of the fused g-path sweeps (tools/fuzz_gsweep.py, 24 cases): T around the 32-step blocks of the
    packed labels / staged scale rows, B around the 128-sequence chunks, 1 ... 17 variants in 1 ... 3 groups with
    random FTL columns and eta0, random stream keys.  d = 1 and the two-pass mode (any d <= 8) bit-exact; the single-pass
    mode within 1e-12 * max(1, cum_loss) with cum_loss bit-exact."""
    rng = np.random.default_rng(2026)
    bits = lambda a: np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    for it in range(24):
        d = int(rng.choice([1, 1, 2, 3, 4, 5, 5, 6, 7, 8]))
        T = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 97, 128, 200, 513]))
        B = int(rng.choice([1, 31, 32, 33, 127, 128, 129, 257, 700]))
        A = int(rng.choice([1, 2, 3, 5, 8, 9, 12, 17]))
        flags = [int(rng.random() < 0.25) for _ in range(A)]
        etas = [float(R2 * 2.0 ** rng.integers(-4, 5) * (1.0 if rng.random() < 0.5 else rng.uniform(0.5, 1.5))) for _ in range(A)]
        ka, kc = int(rng.integers(0, 5000)), int(rng.integers(0, 10 ** 6))
        want = O.regret_sweep(ka, T, kc, B, d, flags, etas)
        tag = (it, d, T, B, A)
        got = eng.regret_sweep("gpath", B, T, d, flags, etas, key_a=ka, key_c0=kc, full=True)
        assert (bits(got["regret"]) == bits(want)).all(), tag
        assert got["stats"][:, 0].tolist() == want.max(axis=0).tolist(), tag
        if d >= 2:
            lin = eng.regret_sweep("gpath", B, T, d, flags, etas, key_a=ka, key_c0=kc, comparator="linear", full=True)
            assert (bits(lin["cum_loss"]) == bits(got["cum_loss"])).all(), tag
            assert np.all(np.abs(lin["regret"] - want) <= REL * np.maximum(1.0, lin["cum_loss"])), tag


def test_float32_mode_small_d_knife_edges(eng):
    """float32 mode against the reference's algorithms.py on d = 1 / 2 / 8 g-path rows cast to float32 (fixtures made by
    running the reference, tests/golden/driver32_stats.json): at d = 1 a third of the rows have |z| = 1 exactly and
    q - y is 0 or one float32 ulp, so only IEEE float32 sqrt / divide reproduce the reference's gradient steps."""
    import json, os
    from conftest import GOLDEN
    recs = json.load(open(os.path.join(GOLDEN, "driver32_stats.json")))["f32_small_d"]
    n = 0
    for rec in recs:
        z, y = make_sequence(rec["spec"])
        z32, y32 = np.asarray(z, dtype=np.float32)[None], np.asarray(y, dtype=np.float32)[None]
        for key, v in rec["results"].items():
            flag = int(key[3]); eta = float(key.split("eta")[1]); want = hexf(v)
            got = eng.simulate_batch(z32, y32, [flag], [eta], compute="f32")[0, 0]
            assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (rec["spec"], key, got, want)
            n += 1
    assert n == 24


def test_round2_paths_edge_cases(eng):
    """Empty / degenerate shapes through the round-2 kernels: T = 0, B = 0, B = 1, explicit keys, d at the limits."""
    for d in (1, 3, 5):
        for mode in ("two_pass", "linear"):
            r = eng.regret_sweep("gpath", 3, 0, d, [0, 1], [R2, R2], comparator=mode, full=True)
            assert (r["regret"] == 0).all() and (r["cum_loss"] == 0).all() and (r["comp_loss"] == 0).all()
            r = eng.regret_sweep("gpath", 0, 10, d, [0], [R2], comparator=mode)
            assert r["regret"].shape == (0, 1) and (r["stats"] == 0).all()
    keys = np.array([[7, 11 + b] for b in range(5)], dtype=np.uint64)      # explicit (a, c) keys == implicit key_a / key_c0 + b
    a = eng.regret_sweep("gpath", 5, 77, 4, [0, 1], [R2, 0.5], key_a=7, key_c0=11, comparator="linear")["regret"]
    b = eng.regret_sweep("gpath", 5, 77, 4, [0, 1], [R2, 0.5], keys=keys, comparator="linear")["regret"]
    assert (bits(a) == bits(b)).all()
    g, st = eng.g_sweep(np.array([], dtype=np.int64), runs=5, comparator="linear")
    assert g.shape == (0,)
    g, st = eng.g_sweep(np.array([50]), runs=0, comparator="linear")
    assert g[0] == 0.0
    z, y = eng.gen_sequences("gpath", 0, 10, 7)
    assert z.shape == (0, 10, 7)
    z, y = eng.gen_sequences("gpath", 2, 0, 7)
    assert z.shape == (2, 0, 7)
    with pytest.raises(ValueError):
        eng.gen_sequences("gpath", 1, 4, 129)
    with pytest.raises(ValueError):
        eng.regret_sweep("gpath", 1, 4, 2, [0], [R2], comparator="nonsense")
    # wide-d exact paths at the limits
    for d in (9, 256):
        zz, yy = O.gen_gpath(2, 30, 0, d) if d <= 128 else (None, None)
        if zz is None:
            rng = np.random.default_rng(1)
            zz = rng.standard_normal((30, d)); zz /= np.maximum(1.0, np.linalg.norm(zz, axis=1, keepdims=True)); yy = np.sign(rng.standard_normal(30))
        ex = eng.exact_ftl_prefix_batch(zz[None], yy[None])
        assert int(ex["status"][0]) == 0
        v = np.cumsum(yy[:, None] * zz, axis=0)
        want = v / np.linalg.norm(v, axis=1, keepdims=True)
        assert np.allclose(ex["actions"][0, 1:], want, rtol=1e-12, atol=1e-14)
        rp = eng.replay_actions(zz[None], yy[None], ex["actions"])
        assert abs(rp["regret"][0] - ex["regret"][0]) <= REL * max(1.0, abs(ex["cum_loss"][0]))
    ex = eng.exact_ftl_prefix_batch(np.zeros((2, 0, 12)), np.zeros((2, 0)))
    assert ex["actions"].shape == (2, 1, 12) and (ex["regret"] == 0).all()
    with pytest.raises(ValueError):
        eng.exact_ftl_prefix_batch(np.zeros((1, 3, 257)), np.ones((1, 3)))


def test_g_sweep_exact_max_is_bit_exact_at_single_pass_speed(eng):
    """comparator="exact_max": the single-pass regret is an upper bound of the literal one for every run, so only the
    candidates for the maximum take the literal second pass -- g(T) equals the two-pass (= reference) value bit for bit."""
    for d, Tg, runs, seed in ((5, [100, 257, 1000], 700, 0), (2, [64, 300], 300, 3), (8, [120], 200, 1), (3, [1, 2, 33], 40, 9)):
        Tg = np.array(Tg)
        if d <= 5:
            g0, st0 = eng.g_sweep(Tg, runs=runs, base_seed=seed, d=d)
        else:       # d = 6..8 has no fused two-pass sweep: the oracle gives the literal values
            g0 = np.array([max(0.0, O.regret_sweep(seed, int(T), 0, runs, d, [0], [R2]).max()) for T in Tg])
        g1, st1 = eng.g_sweep(Tg, runs=runs, base_seed=seed, d=d, comparator="exact_max")
        assert (bits(g1) == bits(g0)).all(), (d, g1, g0)
        assert (st1[:, 3] == runs).all()
    from online_convex_optimization_b200 import fast_algorithms as fa
    want = O.empirical_worst_case_thresholds(np.array([100, 400]), runs=300, base_seed=2)
    assert fa.empirical_worst_case_thresholds(np.array([100, 400]), runs=300, base_seed=2) == want           # default = exact_max
    assert fa.empirical_worst_case_thresholds(np.array([100, 400]), runs=300, base_seed=2, comparator="two_pass") == want
    # full length: 64 runs at T = 1e4 against the oracle
    g, _ = eng.g_sweep(np.array([10_000]), runs=64, base_seed=0, comparator="exact_max")
    assert g[0] == max(0.0, O.regret_sweep(0, 10_000, 0, 64, 5, [0], [R2]).max())
