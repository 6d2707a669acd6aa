"""Host prototype of the barrier / smoothing-Newton exact-FTL solver (development tool, not product path).
min_x 0.5*sum|z_i.x - y_i| s.t. ||x||_p <= 1.  Eliminating the epigraph variables t_i of the log barrier gives the
smooth convex F_mu(x) = sum_i [t_i - mu*log(2*mu*t_i)] + mu*B(x),  t_i = mu + sqrt(mu^2 + r_i^2)."""
import itertools
import sys
import numpy as np


def barrier(x, norm, SIG):
    d = x.size
    if norm == "l2":
        q = 1.0 - x @ x
        return 2 * x / q, 2 * np.eye(d) / q + 4 * np.outer(x, x) / q ** 2
    if norm == "linf":
        q = 1.0 - x * x
        return 2 * x / q, np.diag(2 * (1 + x * x) / q ** 2)
    s = 1.0 - SIG @ x
    return (SIG / s[:, None]).sum(0), (SIG / s[:, None]).T @ (SIG / s[:, None])


def feasible(x, norm):
    if norm == "l2":
        return x @ x < 1.0
    if norm == "linf":
        return np.abs(x).max() < 1.0
    return np.abs(x).sum() < 1.0


def grad(x, z, y, mu, norm, SIG, hess=False):
    r = z @ x - y
    s = np.sqrt(mu * mu + r * r)
    t = mu + s
    gB, HB = barrier(x, norm, SIG)
    g = z.T @ (r / t) + mu * gB
    if not hess:
        return g
    w = mu / (t * s)
    H = (z * w[:, None]).T @ z + mu * HB
    return g, H


def mod_chol_solve(H, b):
    d = b.size
    L = np.zeros((d, d))
    floor = 1e-13 * max(np.diag(H).max(), 1e-300)
    for j in range(d):
        piv = H[j, j] - L[j, :j] @ L[j, :j]
        if piv < floor:
            piv = floor
        L[j, j] = np.sqrt(piv)
        for i in range(j + 1, d):
            L[i, j] = (H[i, j] - L[i, :j] @ L[j, :j]) / L[j, j]
    yv = np.linalg.solve(L, b)
    return np.linalg.solve(L.T, yv)


def solve(z, y, norm="l2", tol=1e-8, max_newton=300, theta=0.15, center=0.05, verbose=False):
    k, d = z.shape
    SIG = np.array(list(itertools.product([-1.0, 1.0], repeat=d))) if norm == "l1" else None
    nu = {"l2": 1, "linf": 2 * d, "l1": 2 ** d}[norm]
    f0 = 0.5 * np.abs(y).sum()
    mu_f = tol * max(1.0, f0) / (k + nu)
    mu = max(np.abs(y).mean(), mu_f, 1e-3)
    x = np.zeros(d)
    nit = 0
    evals = 0
    while nit < max_newton:
        g, H = grad(x, z, y, mu, norm, SIG, hess=True)
        p = mod_chol_solve(H, -g)
        gp = g @ p
        lam2 = -gp / mu
        if lam2 <= center:
            if mu <= mu_f:
                return x, nit, evals, True
            mu = max(mu * theta, mu_f)
            continue
        nit += 1
        a = 1.0
        while not feasible(x + a * p, norm):
            a *= 0.5
        lo, hi, dlo, dhi = 0.0, None, gp, None
        for _ in range(12):
            evals += 1
            da = grad(x + a * p, z, y, mu, norm, SIG) @ p
            if abs(da) <= 0.2 * abs(gp) or (da <= 0 and hi is None):
                break
            if da > 0:
                hi, dhi = a, da
            else:
                lo, dlo = a, da
            # secant inside the bracket, safeguarded towards bisection
            a_new = lo - dlo * (hi - lo) / (dhi - dlo)
            a = min(max(a_new, lo + 0.1 * (hi - lo)), hi - 0.1 * (hi - lo))
        x = x + a * p
    return x, nit, evals, False


def obj(x, z, y):
    return 0.5 * np.abs(z @ x - y).sum()


if __name__ == "__main__":
    from scipy.optimize import linprog
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    worst = 0.0
    tot_n = tot_e = cnt = 0
    for trial in range(60):
        d = int(rng.integers(1, 9)); k = int(rng.choice([1, 2, 3, 5, 17, 40, 200, 1000]))
        scale = float(rng.choice([0.3, 0.9, 3.0]))
        z = rng.standard_normal((k, d)) * scale
        y = rng.standard_normal(k) if trial % 2 else np.sign(rng.standard_normal(k))
        for norm in ("linf", "l1", "l2"):
            x, nit, ev, ok = solve(z, y, norm)
            tot_n += nit; tot_e += ev; cnt += 1
            f = obj(x, z, y)
            if norm == "linf":
                c = np.concatenate([np.zeros(d), 0.5 * np.ones(k)])
                A = np.block([[z, -np.eye(k)], [-z, -np.eye(k)]])
                fr = linprog(c, A_ub=A, b_ub=np.concatenate([y, -y]), bounds=[(-1, 1)] * d + [(0, None)] * k, method="highs").fun
            elif norm == "l1":
                c = np.concatenate([np.zeros(2 * d), 0.5 * np.ones(k)])
                A = np.block([[z, -z, -np.eye(k)], [-z, z, -np.eye(k)], [np.ones((1, 2 * d)), np.zeros((1, k))]])
                fr = linprog(c, A_ub=A, b_ub=np.concatenate([y, -y, [1.0]]), bounds=[(0, None)] * (2 * d + k), method="highs").fun
            else:
                fr = None
            if fr is not None:
                gap = (f - fr) / max(1.0, 0.5 * np.abs(y).sum())
                worst = max(worst, gap)
                flag = "" if ok and gap < 2e-8 else "  <<<<"
                print(f"d={d} k={k} s={scale} {norm}: newton={nit} evals={ev} ok={ok} gap={gap:.2e}{flag}")
            else:
                print(f"d={d} k={k} s={scale} {norm}: newton={nit} evals={ev} ok={ok} f={f:.9f} |x|={np.linalg.norm(x):.6f}")
    print("worst gap", worst, "mean newton", tot_n / cnt, "mean evals", tot_e / cnt)
