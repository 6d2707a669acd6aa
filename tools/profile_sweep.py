"""Small driver for ncu: one warm-up + one profiled launch of the fused sweep kernel."""
import math, sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 189440
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
if wl == "c3":
    d, alg, eta = 1, [1] + [0] * 8, [R2] + [R2 * s for s in (0.125, 0.25, 0.5, 1, 2, 4, 8, 16)]
else:
    d, alg, eta = 5, [0], [R2]
for _ in range(2):
    reg, st = engine.regret_sweep_device("gpath", B, T, d, alg, eta, key_a=0, key_c0=0)
torch.cuda.synchronize()
print("ok", float(st[0, 0]))
