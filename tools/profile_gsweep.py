"""Small driver for ncu: one warm-up + one profiled launch of gsweep_kernel (comparator="linear", d = 5).
usage: python tools/profile_gsweep.py [B] [T] [slab_frac] [d]"""
import math, os, sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 113664
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
if len(sys.argv) > 3:
    os.environ["OCO_GSWEEP_SLAB_FRAC"] = sys.argv[3]
d = int(sys.argv[4]) if len(sys.argv) > 4 else 5
for _ in range(2):
    reg, st = engine.regret_sweep_device("gpath", B, T, d, [0], [math.sqrt(2.0)], comparator="linear", want_regret=False)
torch.cuda.synchronize()
print("ok", float(st[0, 0]))
