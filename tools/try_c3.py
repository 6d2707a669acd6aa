"""C3 (d = 1, FTL + 8 FTRL) parity + timing: pipelined gsweep kernel vs the round-1 sweep kernel (OCO_SWEEP_LEGACY=1)."""
import math, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
alg, eta = [1] + [0] * 8, [R2] + [R2 * s for s in (0.125, 0.25, 0.5, 1, 2, 4, 8, 16)]
if "parity" in sys.argv:
    from oracle import oracle as O
    for T, B, A in ((500, 700, 9), (33, 130, 9), (1000, 300, 5), (257, 129, 1), (2, 3, 9), (700, 200, 17)):
        fl = (alg + [0] * 8)[:A]; et = (eta + [0.07 * k for k in range(1, 9)])[:A]
        want = O.regret_sweep(3, T, 5, B, 1, fl, et)
        got = engine.regret_sweep("gpath", B, T, 1, fl, et, key_a=3, key_c0=5, full=True)
        ok = (got["regret"].view(np.uint64) == want.view(np.uint64)).all()
        os.environ["OCO_SWEEP_FORCE_SECOND_PASS"] = "1"
        g2 = engine.regret_sweep("gpath", B, T, 1, fl, et, key_a=3, key_c0=5, full=True)
        del os.environ["OCO_SWEEP_FORCE_SECOND_PASS"]
        ok2 = all((got[k].view(np.uint64) == g2[k].view(np.uint64)).all() for k in ("regret", "cum_loss", "comp_loss"))
        print(f"T={T} B={B} A={A}: bit-exact vs oracle {ok}; candidate sums == literal pass {ok2}; stats max ok {(got['stats'][:,0]==want.max(axis=0)).all()}")
T = 10000
for B in [int(a) for a in sys.argv[1:] if a.isdigit()] or [1_000_000]:
    for legacy in ("1", "0"):
        os.environ["OCO_SWEEP_LEGACY"] = legacy
        for _ in range(2):
            reg, st = engine.regret_sweep_device("gpath", B, T, 1, alg, eta, want_regret=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            reg, st = engine.regret_sweep_device("gpath", B, T, 1, alg, eta, want_regret=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"legacy={legacy} B={B}: {ms:.2f} ms  {B*T*9/ms/1e-3:.4e} traj-steps/s  frac={B*T*9*15/ms/1e-3/1.857e13:.3f}  max0={float(st[0,0]):.6f}", flush=True)
