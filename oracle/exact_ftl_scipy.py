"""TEST INFRASTRUCTURE -- scipy restatement of the reference's exact-FTL conic program (exact_ftl.py:83-105)

    minimise 0.5 * sum_i |z_i.x - y_i|   subject to  ||x||_p <= 1,   p in {2, inf, 1}

and of its prefix protocol (exact_ftl.py:280-303: actions[0] = 0, actions[k] solves the first k rows).
The reference solves this with CVXPY and whatever conic solver CVXPY picks; neither is installed in this
image and the reference pins no version, so this oracle is PARITY UNPINNED: it restates the *problem*, with
independent solvers (HiGHS LP for linf / l1, SLSQP on the epigraph form for l2), not the reference's
arithmetic.  Only tests/ may import it; the product path never does.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog, minimize

NORMS = ("l2", "linf", "l1")


def objective(x: np.ndarray, z: np.ndarray, y: np.ndarray) -> float:
    """exact_ftl.py:224-227 (`0.5*abs(z@x - y).sum()`)."""
    return float(0.5 * np.abs(z @ x - y).sum())


def solve(z: np.ndarray, y: np.ndarray, norm: str = "l2"):
    """(optimal value, a minimiser) of the program above for one prefix.  exact_ftl.py:83-105."""
    z = np.asarray(z, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    k, d = z.shape
    if norm == "linf":      # variables (x, s): s >= +-(z x - y), -1 <= x <= 1          (exact_ftl.py:91-92)
        c = np.concatenate([np.zeros(d), 0.5 * np.ones(k)])
        A = np.block([[z, -np.eye(k)], [-z, -np.eye(k)]])
        r = linprog(c, A_ub=A, b_ub=np.concatenate([y, -y]), bounds=[(-1, 1)] * d + [(0, None)] * k, method="highs")
        return float(r.fun), r.x[:d]
    if norm == "l1":        # x = p - q, p, q >= 0, sum(p + q) <= 1                      (exact_ftl.py:93-94)
        c = np.concatenate([np.zeros(2 * d), 0.5 * np.ones(k)])
        A = np.block([[z, -z, -np.eye(k)], [-z, z, -np.eye(k)], [np.ones((1, 2 * d)), np.zeros((1, k))]])
        r = linprog(c, A_ub=A, b_ub=np.concatenate([y, -y, [1.0]]), bounds=[(0, None)] * (2 * d + k), method="highs")
        return float(r.fun), r.x[:d] - r.x[d:2 * d]
    if norm != "l2":
        raise ValueError("norm must be one of {'l2','linf','l1'}")                       # exact_ftl.py:95-96
    cons = [{"type": "ineq", "fun": lambda v: v[d:] - (z @ v[:d] - y)},                  # s >= z x - y
            {"type": "ineq", "fun": lambda v: v[d:] + (z @ v[:d] - y)},                  # s >= -(z x - y)
            {"type": "ineq", "fun": lambda v: 1.0 - v[:d] @ v[:d]}]                      # ||x||_2 <= 1  (exact_ftl.py:89-90)
    best = (np.inf, np.zeros(d))
    for x0 in (np.zeros(d), np.full(d, 0.2), -np.full(d, 0.2)):                          # SLSQP is local: a few starts
        v0 = np.concatenate([x0, np.abs(z @ x0 - y) + 0.1])
        r = minimize(lambda v: 0.5 * v[d:].sum(), v0, constraints=cons, method="SLSQP", options={"maxiter": 1000, "ftol": 1e-13})
        x = r.x[:d] / max(1.0, float(np.linalg.norm(r.x[:d])))
        f = objective(x, z, y)
        if f < best[0]:
            best = (f, x)
    return best


def weighted_median_d1(z: np.ndarray, y: np.ndarray) -> float:
    """d = 1 optimum in closed form (all three balls are [-1, 1]): min sum |z_i| * |x - y_i/z_i| is a weighted
    median of the breakpoints y_i/z_i, clipped to [-1, 1].  Returns the optimal VALUE (minimisers may form a segment)."""
    z = np.asarray(z, dtype=np.float64).reshape(-1); y = np.asarray(y, dtype=np.float64).reshape(-1)
    cand = [-1.0, 1.0] + [float(np.clip(yi / zi, -1.0, 1.0)) for zi, yi in zip(z, y) if zi != 0.0]
    return min(float(0.5 * np.abs(z * c - y).sum()) for c in cand)


def compute_prefix_values(z: np.ndarray, y: np.ndarray, norm: str = "l2", prefixes=None):
    """Optimal value of every requested prefix length (exact_ftl.py:280-303 protocol; prefix 0 has value 0)."""
    T = z.shape[0]
    out = {}
    for k in (range(T + 1) if prefixes is None else prefixes):
        out[int(k)] = 0.0 if k == 0 else solve(z[:k], y[:k], norm)[0]
    return out
