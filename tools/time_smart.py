import sys, time, math
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
T, B = 10000, 4096
z, y = engine.gen_sequences_device("gpath", B, T, 5, dtype=torch.float64)
th, et = [math.sqrt(2 * T), 60.0], [math.sqrt(2)] * 2
for mode in ("linear", "exact"):
    nb = B
    r, sw = engine.smart_batch_device(z[:nb], y[:nb], th, et, mode=mode); torch.cuda.synchronize()
    t0 = time.perf_counter(); r, sw = engine.smart_batch_device(z[:nb], y[:nb], th, et, mode=mode); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"SMART {mode}: {nb} sequences x T={T} x 2 thresholds in {dt*1e3:.1f} ms -> {nb*T*2/dt:.3e} trajectory-steps/s; switched: {(sw>=0).float().mean().item():.2f}")
