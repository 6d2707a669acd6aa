"""Small driver for ncu: warm-up + one profiled launch of the d <= 8 streamed kernel.
usage: profile_stream.py d T B f64|f32 A two_pass|linear"""
import math, sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
d, T, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dt = torch.float64 if sys.argv[4] == "f64" else torch.float32
A = int(sys.argv[5]); mode = sys.argv[6]
R2 = math.sqrt(2.0)
alg = [0] if A == 1 else ([0, 1] if A == 2 else [1] + [0] * (A - 1))
eta = [R2] * A
z, y = engine.gen_sequences_device("gpath", B, T, d, dtype=dt)
for _ in range(2):
    r = engine.simulate_batch_device(z, y, alg, eta, comparator=mode)
torch.cuda.synchronize()
print("ok", float(r[0, 0]))
