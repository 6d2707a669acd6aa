"""Randomised parity check of the fused g-path sweeps against the CPU oracle (one-off sanity run on a GPU box):
random (d, T, B, A, keys) shapes; d = 1 bit-exact (both comparator modes), d >= 2 two-pass bit-exact and
single-pass within 1e-12 * max(1, cum_loss) with cum_loss bit-exact.  usage: python tools/fuzz_gsweep.py [n] [seed]"""
import math, sys
sys.path.insert(0, ".")
import numpy as np
from online_convex_optimization_b200 import engine
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bits = lambda a: np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
bad = 0
for it in range(n):
    d = int(rng.choice([1, 1, 2, 3, 4, 5, 5, 6, 7, 8]))
    T = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 97, 128, 200, 513, 1000, 4097, 10000]))
    B = int(rng.choice([1, 31, 32, 33, 127, 128, 129, 257, 700] if T < 4000 else [5, 130, 300]))
    A = int(rng.choice([1, 2, 3, 5, 8, 9, 12, 17]))
    flags = [int(rng.random() < 0.25) for _ in range(A)]
    etas = [float(math.sqrt(2.0) * 2.0 ** rng.integers(-4, 5) * (1.0 if rng.random() < 0.5 else rng.uniform(0.5, 1.5))) for _ in range(A)]
    ka, kc = int(rng.integers(0, 5000)), int(rng.integers(0, 10 ** 6))
    want = O.regret_sweep(ka, T, kc, B, d, flags, etas)
    ok = True
    got = engine.regret_sweep("gpath", B, T, d, flags, etas, key_a=ka, key_c0=kc, full=True)
    ok = ok and (bits(got["regret"]) == bits(want)).all() and got["stats"][:, 0].tolist() == want.max(axis=0).tolist()
    cum = got["cum_loss"]
    if d >= 2:
        lin = engine.regret_sweep("gpath", B, T, d, flags, etas, key_a=ka, key_c0=kc, comparator="linear", full=True)
        ok = ok and (bits(lin["cum_loss"]) == bits(cum)).all()
        ok = ok and bool(np.all(np.abs(lin["regret"] - want) <= 1e-12 * np.maximum(1.0, lin["cum_loss"])))
    bad += not ok
    print(f"{it:3d} d={d} T={T} B={B} A={A} ftl={sum(flags)} -> {'ok' if ok else 'MISMATCH'}", flush=True)
print("FUZZ", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
