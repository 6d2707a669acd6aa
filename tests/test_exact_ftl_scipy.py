"""CPU test: the closed form used for exact FTL in the linear regime (oracle_exact_ftl_l2_prefix, and
the CUDA kernel that the GPU tests compare with it) IS the optimum of the reference's SOCP
min_{||x||_2 <= 1} 0.5 * sum_i |z_i.x - y_i|   (exact_ftl.py:83-105),
checked against an independent general-purpose solver (scipy SLSQP on the epigraph form) -- the
reference's own solver stack (CVXPY + conic solver) is not installed here (parity unpinned)."""
import numpy as np
import pytest

from oracle import exact_ftl_scipy as XS
from oracle import oracle as O

_objective = XS.objective


def _scipy_solve(z, y):
    return XS.solve(z, y, "l2")


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_closed_form_is_the_socp_optimum(seed):
    T, d = 24, 3
    z, y = O.gen_gpath(seed, T, 0, d)
    actions = O.exact_ftl_l2_prefix(z, y)
    for k in (1, 2, 5, 11, 24):
        f_closed = _objective(actions[k], z[:k], y[:k])
        f_scipy, x_scipy = _scipy_solve(z[:k], y[:k])
        assert f_closed <= f_scipy + 1e-7, (k, f_closed, f_scipy)
        assert abs(f_closed - f_scipy) <= 1e-5 * max(1.0, f_scipy), (k, f_closed, f_scipy)
        assert np.linalg.norm(actions[k]) <= 1.0 + 1e-12


def test_tie_break_is_the_centre():
    """v = 0 (label flips at even prefixes): every feasible x is optimal; the closed form returns the centre."""
    z, y, _ = O.gen_flip(10)
    actions = O.exact_ftl_l2_prefix(z, y)
    for k in (2, 4, 10):
        assert (actions[k] == 0).all()
        for x0 in (-1.0, -0.3, 0.0, 0.7, 1.0):       # the objective is flat on the feasible segment
            x = np.zeros(5); x[0] = x0
            assert _objective(x, z[:k].astype(float), y[:k].astype(float)) == _objective(actions[k], z[:k].astype(float), y[:k].astype(float))


def test_scipy_oracle_norms_agree_where_they_must():
    """d = 1: the three unit balls coincide and the optimum is a clipped weighted median -- pins the LP / SLSQP
    restatements of exact_ftl.py:83-105 against a closed form (including rows outside the ball and real labels)."""
    rng = np.random.default_rng(3)
    for k in (1, 2, 7, 30):
        z = rng.standard_normal((k, 1)) * 2.0
        y = rng.standard_normal(k)
        want = XS.weighted_median_d1(z, y)
        for norm in XS.NORMS:
            got, x = XS.solve(z, y, norm)
            assert abs(got - want) <= 1e-7 * max(1.0, want), (k, norm, got, want)
            assert abs(x[0]) <= 1.0 + 1e-9
    # d > 1: l1 ball inside l2 ball inside linf ball -> optimal values are ordered the other way round
    z = rng.standard_normal((25, 3)); y = rng.standard_normal(25)
    f1, f2, fi = (XS.solve(z, y, n)[0] for n in ("l1", "l2", "linf"))
    assert fi <= f2 + 1e-6 and f2 <= f1 + 1e-6
