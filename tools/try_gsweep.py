"""Parity + timing of the regenerating g-sweep kernel (comparator="linear") against the oracle and the slab kernel.
usage: python tools/try_gsweep.py [parity] [time] [B]"""
import math, os, sys, time
sys.path.insert(0, ".")
import numpy as np
R2 = math.sqrt(2.0)
from online_convex_optimization_b200 import engine
what = sys.argv[1:] or ["parity", "time"]

def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)

if "parity" in what:
    from oracle import oracle as O
    ok = True
    for frac in ("-",):
        for d, T, B, A in ((5, 300, 150, 1), (5, 201, 140, 2), (2, 100, 10, 4), (3, 257, 70, 9), (4, 64, 129, 5), (5, 1, 3, 1)):
            flags = ([0, 1] + [0] * 7)[:A]
            etas = [R2 * s for s in (1.0, 1.0, 0.125, 0.25, 0.5, 2.0, 4.0, 8.0, 16.0)][:A]
            want = O.regret_sweep(7, T, 11, B, d, flags, etas)
            got = engine.regret_sweep("gpath", B, T, d, flags, etas, key_a=7, key_c0=11, comparator="linear", full=True)
            ex = engine.regret_sweep("gpath", B, T, d, flags, etas, key_a=7, key_c0=11, full=True)
            assert (bits(ex["regret"]) == bits(want)).all()
            cum_ok = (bits(got["cum_loss"]) == bits(ex["cum_loss"])).all()
            err = np.abs(got["regret"] - want) / np.maximum(1.0, ex["cum_loss"])
            errc = np.abs(got["comp_loss"] - ex["comp_loss"]) / np.maximum(1.0, ex["cum_loss"])
            rel = np.abs(got["regret"] - want) / np.maximum(1e-300, np.abs(want))
            st_ok = got["stats"][:, 0].tolist() == got["regret"].max(axis=0).tolist()
            print(f"frac={frac} d={d} T={T} B={B} A={A}: cum bit-exact={cum_ok} max|dreg|/max(1,cum)={err.max():.2e} comp {errc.max():.2e} "
                  f"rel-to-regret={rel.max():.2e} stats_max_ok={st_ok}")
            ok = ok and cum_ok and err.max() <= 1e-12 and errc.max() <= 1e-12 and st_ok
        os.environ["OCO_SWEEP_FORCE_SECOND_PASS"] = "1"
        for d, T, B, A in ((5, 120, 70, 2), (3, 90, 40, 4)):
            flags = ([0, 1] + [0] * 7)[:A]; etas = [R2 * s for s in (1.0, 1.0, 0.125, 0.25)][:A]
            want = O.regret_sweep(3, T, 5, B, d, flags, etas)
            got = engine.regret_sweep("gpath", B, T, d, flags, etas, key_a=3, key_c0=5, comparator="linear")
            e = (bits(got["regret"]) == bits(want)).all()
            print(f"frac={frac} forced literal second pass d={d}: bit-exact={e}")
            ok = ok and e
        del os.environ["OCO_SWEEP_FORCE_SECOND_PASS"]
    # T = 1e4 spot check
    want = O.regret_sweep(0, 10000, 0, 8, 5, [0], [R2])
    for frac in ("-",):
        got = engine.regret_sweep("gpath", 8, 10000, 5, [0], [R2], comparator="linear", full=True)
        err = np.abs(got["regret"] - want)
        print(f"T=1e4 d=5 frac={frac}: max |dreg| = {err.max():.3e}  rel to regret {np.max(err/np.abs(want)):.2e}  rel to cum {np.max(err/got['cum_loss']):.2e}")
        ok = ok and np.max(err / got["cum_loss"]) <= 1e-12
    print("PARITY", "OK" if ok else "FAILED")

if "time" in what:
    import torch
    T, d = 10000, 5
    def run(mode, B):
        for _ in range(2):
            reg, st = engine.regret_sweep_device("gpath", B, T, d, [0], [R2], comparator=mode, want_regret=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            reg, st = engine.regret_sweep_device("gpath", B, T, d, [0], [R2], comparator=mode, want_regret=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(f"mode={mode} B={B}: {ms:.2f} ms/launch  {B*T/ms/1e-3:.3e} steps/s  max={float(st[0,0]):.6f}", flush=True)
    Bs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [94720, 378880, 1000000]
    for B in Bs:
        run("linear", B)
    run("two_pass", Bs[-1])
