"""Tiny pass through every kernel family for compute-sanitizer (memcheck)."""
import math, sys
sys.path.insert(0, ".")
import numpy as np
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
r = engine.regret_sweep("gpath", 130, 40, 1, [1] + [0] * 8, [R2] * 9)
r = engine.regret_sweep("gpath", 130, 40, 5, [0], [R2])
keys = np.array([[2025, 13 + i] for i in range(40)], dtype=np.uint64)
r = engine.regret_sweep("noisy", 40, 36, 5, [0, 1], [R2, R2], keys=keys, p_flip=0.1)
z, y = engine.gen_sequences("gpath", 37, 44, 5)
a = engine.simulate_batch(z, y, [0, 1, 0], [R2, R2, 0.3], full=True, trace=True)
a = engine.simulate_batch(z, y, [0, 1], [R2, R2], comparator="linear")
a = engine.simulate_batch(z[:, :43].copy(), y[:, :43].copy(), [0], [R2])             # odd T: direct-load kernel
a = engine.simulate_batch(z.astype(np.float32), y.astype(np.float32), [0, 1], [R2, R2], compute="f32")
s = engine.smart_batch(z, y, [3.0, 9.0], [R2, R2]); s = engine.smart_batch(z, y, [3.0], [R2], mode="linear")
e = engine.exact_ftl_prefix_batch(z, y); rp = engine.replay_actions(z, y, e["actions"])
zw, yw = engine.gen_sequences("gpath", 9, 42, 64)
w = engine.simulate_batch(zw, yw, [0, 1], [R2, R2], comparator="linear"); w = engine.simulate_batch(zw, yw, [0], [R2])
zw, yw = engine.gen_sequences("iid", 3, 20, 5, key_a=2025, key_c0=13); zf, yf = engine.gen_sequences("flip", 2, 9, 5)
g, st = engine.g_sweep([10, 20], runs=33)
print("sanitize smoke done", float(r["stats"][0, 0]), float(w[0, 0]))
