"""Driver for ncu / timing of the exact SMART kernel: profile_smart.py T B d"""
import sys, time, math
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
T, B, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
z, y = engine.gen_sequences_device("gpath", B, T, d, dtype=torch.float32)
th, et = [math.sqrt(2 * T), 0.75 * math.sqrt(T)], [math.sqrt(2)] * 2
for _ in range(2):
    r, sw = engine.smart_batch_device(z, y, th, et, mode="exact")
torch.cuda.synchronize()
t0 = time.perf_counter(); r, sw = engine.smart_batch_device(z, y, th, et, mode="exact"); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"SMART exact: {B} x T={T} d={d}: {dt*1e3:.2f} ms, {B*T*2/dt:.3e} traj-steps/s, row-ops/s {B*T*T/2/dt:.3e}, switched {(sw>=0).float().mean().item():.2f}")
