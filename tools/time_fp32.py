"""float32 compute mode (algorithms.py semantics) vs the float64 kernel on the same device-resident float32 sequences."""
import math, sys
sys.path.insert(0, ".")
import torch
from online_convex_optimization_b200 import engine
R2 = math.sqrt(2.0)
for d, A, B in ((5, 2, 65536), (5, 1, 131072), (1, 9, 65536), (8, 4, 65536)):
    T = 10000
    z, y = engine.gen_sequences_device("iid" if d == 5 else "gpath", B, T, d, key_a=2025, key_c0=13, dtype=torch.float32)
    alg = ([0, 1] + [0] * 7)[:A]; eta = [R2 * 2.0 ** (k - 2) for k in range(A)]
    res = {}
    for comp in ("f64", "f32"):
        for _ in range(2):
            r = engine.simulate_batch_device(z, y, alg, eta, compute=comp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = engine.simulate_batch_device(z, y, alg, eta, compute=comp)
        e1.record(); torch.cuda.synchronize()
        res[comp] = (e0.elapsed_time(e1) / 3, r)
    err = (res["f32"][1] - res["f64"][1]).abs() / res["f64"][1].abs().clamp(min=1.0)
    print(f"d={d} A={A} B={B}: f64 {res['f64'][0]:.2f} ms ({B*T*A/res['f64'][0]/1e-3:.3e}/s)  f32 {res['f32'][0]:.2f} ms ({B*T*A/res['f32'][0]/1e-3:.3e}/s) "
          f"speedup {res['f64'][0]/res['f32'][0]:.2f}x  max rel diff f32 vs f64 {float(err.max()):.2e}  input GB/s (f32) {B*T*(d+1)*4*2/res['f32'][0]/1e6:.0f}")
    del z, y
