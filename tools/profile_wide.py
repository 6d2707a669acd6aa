"""Small driver for ncu: the wide-d TMA kernel on a resident d=64 chunk."""
import math, sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
mode = sys.argv[2] if len(sys.argv) > 2 else "linear"
T, d = 10000, 64
z, y = engine.gen_sequences_device("gpath", B, T, d, key_a=0, key_c0=0, dtype=torch.float64)
for _ in range(2):
    reg = engine.simulate_batch_device(z, y, [0], [math.sqrt(2.0)], comparator=mode)
torch.cuda.synchronize()
print("ok", float(reg.mean()))
