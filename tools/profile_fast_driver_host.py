"""cProfile of the fast_driver twin's default experiment (host-side overhead of the C1 workload)."""
import sys, cProfile, pstats, io
sys.path.insert(0, ".")
from online_convex_optimization_b200 import fast_driver as fd
fd.run_experiment(); fd.run_experiment()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): fd.run_experiment()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
