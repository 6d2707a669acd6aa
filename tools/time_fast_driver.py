import sys, time, math
sys.path.insert(0, ".")
import numpy as np, torch
from online_convex_optimization_b200 import fast_driver as fd, engine, sequence_generation as sg, fast_algorithms as fa
cfg = fd.ExperimentConfig()
fd.run_experiment()
def t(f, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
print("full experiment", t(fd.run_experiment))
print("g_emp", t(lambda: fa.empirical_worst_case_thresholds(cfg.T_grid, runs=1000)))
g = fa.empirical_worst_case_thresholds(cfg.T_grid, runs=1000)
for title, b in sg.CASES.items():
    print(title, t(lambda: fd.evaluate_stream_with_stats(b, cfg.T_grid, g, runs=sg.RUNS_BY_TITLE[title], replicates=sg.REPLICATES_BY_TITLE[title])))
# kernel-level: iid T=1000
keys = np.array([[2025 * (r + 1), 13 + rep] for r in range(48) for rep in range(16)], dtype=np.uint64)
z, y = engine.gen_sequences_device(1, 768, 1000, 5, keys=keys, dtype=torch.float32)
print("gen 768x1000", t(lambda: engine.gen_sequences_device(1, 768, 1000, 5, keys=keys, dtype=torch.float32), 10))
print("alg 768x1000", t(lambda: engine.simulate_batch_device(z, y, [0, 1], [math.sqrt(2)] * 2), 10))
print("smart 768x1000", t(lambda: engine.smart_batch_device(z, y, [math.sqrt(2000), 19.0], [math.sqrt(2)] * 2), 10))
