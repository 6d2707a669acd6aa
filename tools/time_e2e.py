"""Where the host-API overhead of the C3 sweep goes: stats only vs per-trajectory regrets (pageable numpy output)."""
import sys, time, math
sys.path.insert(0, ".")
import numpy as np
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
alg, eta = [1] + [0] * 8, [R2] + [R2 * s for s in (0.125, 0.25, 0.5, 1, 2, 4, 8, 16)]
B, T = 1_000_000, 10_000
for want in (False, True, False, True):
    t0 = time.perf_counter()
    r = engine.regret_sweep("gpath", B, T, 1, alg, eta, want_regret=want)
    dt = time.perf_counter() - t0
    print(f"want_regret={want}: {dt*1e3:.1f} ms -> {B*T*9/dt:.3e}")
t0 = time.perf_counter(); a = np.empty((B, 9)); a[:] = 0; print(f"np.empty+touch 72 MB: {(time.perf_counter()-t0)*1e3:.1f} ms")
