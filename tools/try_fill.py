"""Warp-cooperative g-path generator: bit-exact vs the CPU oracle for many (d, T), then timing vs the thread-per-sequence kernel."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from online_convex_optimization_b200 import engine
from oracle import oracle as O
ok = True
for d, T, B in ((5, 100, 9), (64, 50, 5), (1, 333, 4), (8, 77, 3), (16, 40, 3), (32, 33, 2), (7, 65, 3), (13, 120, 3), (100, 21, 2), (128, 9, 2),
                (3, 1, 2), (2, 2, 1), (64, 2000, 3), (5, 10000, 2), (9, 3000, 2)):
    z, y = engine.gen_sequences("gpath", B, T, d, key_a=11, key_c0=5)
    good = True
    for b in range(B):
        zo, yo = O.gen_gpath(11, T, 5 + b, d)
        good = good and (z[b].view(np.uint64) == zo.view(np.uint64)).all() and (y[b] == yo).all()
    print(f"d={d} T={T} B={B}: bit-exact {good}", flush=True)
    ok = ok and good
zf, yf = engine.gen_sequences("gpath", 3, 500, 64, key_a=1, key_c0=0, dtype=np.float32)
z6, y6 = engine.gen_sequences("gpath", 3, 500, 64, key_a=1, key_c0=0)
ok = ok and (zf == z6.astype(np.float32)).all() and (yf == y6).all()
print("PARITY", "OK" if ok else "FAILED", flush=True)
T, d, B = 10000, 64, 14208
for legacy in ("1", "0"):
    os.environ["OCO_FILL_LEGACY"] = legacy
    z, y = engine.gen_sequences_device("gpath", B, T, d, dtype=torch.float64); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z, y = engine.gen_sequences_device("gpath", B, T, d, dtype=torch.float64); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"legacy={legacy}: fill {B} x {T} x {d}: {ms:.1f} ms  {B*T*d/ms/1e-3:.3e} normals/s  {B*T*(d+1)*8/ms/1e6:.0f} GB/s written", flush=True)
    del z, y
