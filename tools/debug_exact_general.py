"""GPU check of the general exact-FTL solver against HiGHS over random shapes (development tool)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from online_convex_optimization_b200 import engine
from scipy.optimize import linprog

rng = np.random.default_rng(0)
worst = 0.0
for trial in range(24):
    d = int(rng.integers(1, 9)); T = int(rng.choice([3, 17, 60, 300])); B = 3
    sc = float(rng.choice([0.3, 0.9, 3.0]))
    z = rng.standard_normal((B, T, d)) * sc
    y = rng.standard_normal((B, T)) if trial % 2 else np.sign(rng.standard_normal((B, T)))
    for norm in ("linf", "l1", "l2"):
        t0 = time.time(); res = engine.exact_ftl_general_batch(z, y, norm=norm); dt = time.time() - t0
        gaps = []
        for b in range(B):
            for k in sorted({1, min(2, T), T // 2 + 1, T}):
                zz, yy = z[b, :k], y[b, :k]
                f = 0.5 * np.abs(zz @ res["actions"][b, k] - yy).sum()
                if norm == "linf":
                    c = np.concatenate([np.zeros(d), 0.5 * np.ones(k)])
                    A = np.block([[zz, -np.eye(k)], [-zz, -np.eye(k)]])
                    fr = linprog(c, A_ub=A, b_ub=np.concatenate([yy, -yy]), bounds=[(-1, 1)] * d + [(0, None)] * k, method="highs").fun
                elif norm == "l1":
                    c = np.concatenate([np.zeros(2 * d), 0.5 * np.ones(k)])
                    A = np.block([[zz, -zz, -np.eye(k)], [-zz, zz, -np.eye(k)], [np.ones((1, 2 * d)), np.zeros((1, k))]])
                    fr = linprog(c, A_ub=A, b_ub=np.concatenate([yy, -yy, [1.0]]), bounds=[(0, None)] * (2 * d + k), method="highs").fun
                else:
                    continue
                gaps.append((f - fr) / max(1.0, 0.5 * np.abs(yy).sum()))
        g = max(gaps) if gaps else float("nan")
        worst = max(worst, g) if gaps else worst
        print(f"d={d} T={T} sc={sc} {norm}: status={res['status']} max gap={g:.2e} time={dt*1e3:.1f} ms", flush=True)
print("worst", worst)
# throughput: B sequences x T prefixes
for (B, T, d) in ((64, 200, 5), (8, 1000, 5), (1, 1000, 8)):
    z = rng.standard_normal((B, T, d)); y = np.sign(rng.standard_normal((B, T)))
    for norm in ("l2", "l1"):
        engine.exact_ftl_general_batch(z[:1, :10], y[:1, :10], norm=norm)
        t0 = time.time(); res = engine.exact_ftl_general_batch(z, y, norm=norm); dt = time.time() - t0
        print(f"B={B} T={T} d={d} {norm}: {dt:.3f} s, {B*T/dt:.0f} prefix solves/s, status max {res['status'].max()}", flush=True)
