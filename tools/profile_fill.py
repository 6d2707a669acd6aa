"""Small driver for ncu: the warp-cooperative g-path generator (d = 64)."""
import sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 14208
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 64
for _ in range(2):
    z, y = engine.gen_sequences_device("gpath", B, T, d, dtype=torch.float64)
torch.cuda.synchronize()
print("ok", float(z[0, 0, 0]))
