import sys, math
sys.path.insert(0, ".")
import numpy as np
from online_convex_optimization_b200 import engine as eng
from oracle import oracle as O
R2 = math.sqrt(2)
T, B = 600, 40
zs, ys = zip(*[O.gen_iid(2025 * (r + 1), T, 0)[:2] for r in range(B // 2)] + [O.gen_gpath(2, T, r, 5) for r in range(B // 2)])
z, y = np.stack([np.asarray(a, dtype=np.float64) for a in zs]), np.stack([np.asarray(a, dtype=np.float64) for a in ys])
th, et = [math.sqrt(2 * T), 3.0, 0.5], [R2, R2, R2]
want, want_sw = O.smart_batch(z, y, th, et)
got, sw = eng.smart_batch(z, y, th, et, mode="linear")
print("sw mismatch", np.argwhere(sw != want_sw)[:10], (sw != want_sw).sum())
err = np.abs(got - want) / np.maximum(1, np.abs(want)); print("max rel err", err.max(), np.argwhere(err > 1e-12)[:10])
for b, a in np.argwhere(sw != want_sw)[:5]:
    print(b, a, sw[b, a], want_sw[b, a], got[b, a], want[b, a])
z2 = z.copy(); z2[::3] *= 1.7
want2, want_sw2 = O.smart_batch(z2, y, th, et)
got2, sw2 = eng.smart_batch(z2, y, th, et, mode="linear")
print("fallback: sw mismatch", (sw2 != want_sw2).sum(), "max rel", (np.abs(got2 - want2) / np.maximum(1, np.abs(want2))).max())
