"""Timing experiments on gsweep_kernel: phases in isolation (OCO_GSWEEP_DEBUG) at slab_frac = 1."""
import math, os, sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
T, d = 10000, 5
os.environ["OCO_GSWEEP_SLAB_FRAC"] = "1"
for B in [int(a) for a in sys.argv[1:]] or [113664]:
    for dbg in ("0", "1", "2"):
        os.environ["OCO_GSWEEP_DEBUG"] = dbg
        for _ in range(2):
            reg, st = engine.regret_sweep_device("gpath", B, T, d, [0], [R2], comparator="linear", want_regret=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            reg, st = engine.regret_sweep_device("gpath", B, T, d, [0], [R2], comparator="linear", want_regret=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"B={B} debug_phase={dbg}: {ms:.2f} ms  {B*T/ms/1e-3:.3e} steps/s", flush=True)
