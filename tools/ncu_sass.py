"""SASS-level view of an .ncu-rep: per instruction executed count per warp-step, samples and main stall reasons.
usage: python tools/ncu_sass.py rep warps steps [min_exec_per_step]"""
import csv, subprocess, sys
rep = sys.argv[1]; W = float(sys.argv[2]) * float(sys.argv[3]); thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[ix['# Samples']] or 0) for r in rows[2:] if len(r) > ix['# Samples'])
for r in rows[2:]:
    try: n = float(r[ix['Instructions Executed']])
    except Exception: continue
    if n / W < thr: continue
    s = float(r[ix['# Samples']] or 0)
    top = sorted(((float(r[ix[h]] or 0), h[6:]) for h in st), reverse=True)[:2]
    thr_ = float(r[ix['Avg. Threads Executed']] or 0)
    print(f"{r[ix['Address']][-5:]} {n / W:5.2f} thr={thr_:4.1f} {100 * s / tot:5.2f}%  {r[ix['Source']].strip()[:70]:70s} " +
          " ".join(f"{k}={100 * v / max(s, 1):.0f}%" for v, k in top if v > 0))
