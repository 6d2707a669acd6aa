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
# round 2 kernels: pipelined sweep (d = 1 exact, d = 3 / 5 single-pass comparator, several rounds per CTA), warp-cooperative
# generator at an odd d, wide-d exact FTL / replay / fixed comparator, staged float32 kernel
r = engine.regret_sweep("gpath", 700, 33, 3, [0, 1, 0], [R2, R2, 0.2], comparator="linear", full=True)
r = engine.regret_sweep("gpath", 128 * 148 * 5 + 5, 6, 5, [0], [R2], comparator="linear")
g, st = engine.g_sweep([10, 21], runs=33, comparator="linear")
zo, yo = engine.gen_sequences("gpath", 5, 37, 13); zo, yo = engine.gen_sequences("gpath", 3, 9, 128, dtype=np.float32)
z10, y10 = engine.gen_sequences("gpath", 6, 40, 10)
e10 = engine.exact_ftl_prefix_batch(z10, y10); rp = engine.replay_actions(z10, y10, e10["actions"])
a10 = engine.simulate_batch(z10, y10, [0, 1], [R2, R2], full=True, comparator_action=e10["actions"][:, -1, :])
zf, yf = engine.gen_sequences("gpath", 70, 48, 5, dtype=np.float32)
a32 = engine.simulate_batch(zf, yf, [0, 1, 0, 0, 1], [R2, R2, 0.1, 3.0, 1.0], compute="f32", full=True)
print("sanitize smoke done", float(r["stats"][0, 0]), float(w[0, 0]))
# last session of round 2: packed labels / staged scale rows at ragged T (31, 33, 65), NF = 4 / 8 at d = 1 ... 8, the two-pass
# fused sweep at d = 6 ... 8, SMART for d > 8, the lane-per-variable interior-point kernel (l2 / linf / l1 split variables)
for d_, T_, A_ in ((1, 31, 9), (1, 65, 5), (2, 33, 9), (5, 97, 12), (6, 33, 5), (7, 40, 9), (8, 35, 2)):
    fl = ([0, 1] + [0] * 15)[:A_]; et = [R2 * (0.25 + 0.5 * i) for i in range(A_)]
    r = engine.regret_sweep("gpath", 257, T_, d_, fl, et, full=True)
    if d_ >= 2:
        r = engine.regret_sweep("gpath", 257, T_, d_, fl, et, comparator="linear", full=True)
z12, y12 = engine.gen_sequences("gpath", 5, 60, 12); z70, y70 = engine.gen_sequences("gpath", 3, 41, 70)
s = engine.smart_batch(z12, y12, [1.5, 6.0, 1e9], [R2, R2, R2]); s = engine.smart_batch(z70 * 1.7, y70, [2.0], [R2])
for nm, zz in (("l2", z12 * 3.0), ("linf", z12 * 2.0), ("l1", z12 * 2.0)):
    ge = engine.exact_ftl_general_batch(zz[:, :23].copy(), y12[:, :23].copy(), norm=nm)
print("sanitize smoke (round 2, last session) done", int(ge["status"].sum()))
