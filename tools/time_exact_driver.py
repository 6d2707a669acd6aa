"""Wall time of the exact_ftl_driver twin's default experiment (reference: exact_ftl_driver.py main, one CVXPY solve per prefix)."""
import sys, time
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import exact_ftl_driver as ed
ed.run_experiment()
torch.cuda.synchronize(); t0 = time.perf_counter()
res = ed.run_experiment()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
n_prefix = 0
cfg = ed.ExperimentConfig()
from online_convex_optimization_b200.sequence_generation import RUNS_BY_TITLE, REPLICATES_BY_TITLE, CASES
for T in cfg.T_grid:
    n_prefix += int(T) * cfg.g_runs
    for title in CASES:
        n_prefix += int(T) * RUNS_BY_TITLE.get(title, 1) * REPLICATES_BY_TITLE.get(title, 1)
print(f"exact_ftl_driver default experiment: {dt:.3f} s, {n_prefix} exact prefix solves ({n_prefix/dt:.3e}/s)")
print({k: float(v) for k, v in list(res['g_emp'].items())[:3]})
