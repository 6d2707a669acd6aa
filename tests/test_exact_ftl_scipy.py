"""CPU test: the closed form used for exact FTL in the linear regime (oracle_exact_ftl_l2_prefix, and
the CUDA kernel that the GPU tests compare with it) IS the optimum of the reference's SOCP
min_{||x||_2 <= 1} 0.5 * sum_i |z_i.x - y_i|   (exact_ftl.py:83-105),
checked against an independent general-purpose solver (scipy SLSQP on the epigraph form) -- the
reference's own solver stack (CVXPY + conic solver) is not installed here (parity unpinned)."""
import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import oracle as O


def _objective(x, z, y):
    return 0.5 * np.abs(z @ x - y).sum()


def _scipy_solve(z, y):
    k, d = z.shape
    # variables (x, s): min 0.5*sum(s)  s.t.  s >= +-(z x - y),  ||x||^2 <= 1
    cons = [{"type": "ineq", "fun": lambda v: v[d:] - (z @ v[:d] - y)},
            {"type": "ineq", "fun": lambda v: v[d:] + (z @ v[:d] - y)},
            {"type": "ineq", "fun": lambda v: 1.0 - v[:d] @ v[:d]}]
    best = None
    for x0 in (np.zeros(d), np.full(d, 0.1), -np.full(d, 0.1)):
        v0 = np.concatenate([x0, np.abs(z @ x0 - y) + 0.1])
        r = minimize(lambda v: 0.5 * v[d:].sum(), v0, constraints=cons, method="SLSQP", options={"maxiter": 500, "ftol": 1e-12})
        x = r.x[:d]
        n = np.linalg.norm(x)
        if n > 1.0:
            x = x / n
        f = _objective(x, z, y)
        if best is None or f < best[0]:
            best = (f, x)
    return best


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
