import sys, numpy as np
sys.path.insert(0, '.')
from online_convex_optimization_b200 import engine as eng
from oracle import oracle as O
for d in (1, 2, 5):
    T = 257
    z, y = eng.gen_sequences("gpath", 5, T, d, key_a=3, key_c0=7)
    for b in range(5):
        zo, yo = O.gen_gpath(3, T, 7 + b, d)
        bad = np.argwhere(z[b].view(np.uint64) != zo.view(np.uint64))
        print(d, b, "z mismatches", len(bad), "y mismatches", int((y[b] != yo).sum()))
        for (t, j) in bad[:5]:
            print("   t", t, "j", j, z[b][t, j].hex(), zo[t, j].hex(), "raw normal > 1?", abs(zo[t, j]))
