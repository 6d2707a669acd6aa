"""Throughput of the d <= 8 streamed kernel (sim_stream_kernel) on device-resident sequences."""
import math, sys, time
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
for (d, T, B, dt, alg, eta) in [(5, 10000, 131072, torch.float64, [0, 1], [R2, R2]), (5, 10000, 131072, torch.float32, [0, 1], [R2, R2]),
                                 (5, 1000, 1310720, torch.float32, [0, 1], [R2, R2]), (1, 10000, 262144, torch.float64, [1] + [0] * 8, [R2] * 9),
                                 (5, 10000, 131072, torch.float64, [0], [R2])]:
    z, y = engine.gen_sequences_device("gpath", B, T, d, dtype=dt)
    for mode in ("two_pass", "linear"):
        for _ in range(2):
            r = engine.simulate_batch_device(z, y, alg, eta, comparator=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = engine.simulate_batch_device(z, y, alg, eta, comparator=mode)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        es = 8 if dt == torch.float64 else 4
        seq_steps = B * T / (ms * 1e-3)
        print(f"d={d} T={T} B={B} {str(dt)[6:]} A={len(alg)} {mode}: {ms:.2f} ms  seq-steps/s={seq_steps:.3e} traj-steps/s={seq_steps*len(alg):.3e} "
              f"compulsory GB/s={(d+1)*es*seq_steps/1e9:.0f}")
    del z, y
