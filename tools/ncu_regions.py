"""Aggregate an .ncu-rep source page by average active threads per instruction: how many issue slots and stall
samples go to low-occupancy (divergent slow-path) instructions.  usage: ncu_regions.py rep warps steps"""
import csv, subprocess, sys
rep = sys.argv[1]; W = float(sys.argv[2]) * float(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
bins = {}
for r in rows[2:]:
    try: n = float(r[ix['Instructions Executed']])
    except Exception: continue
    if n == 0: continue
    thr = float(r[ix['Avg. Threads Executed']] or 0)
    b = "<4" if thr < 4 else ("4-16" if thr < 16 else ("16-31" if thr < 31 else "31-32"))
    e = bins.setdefault(b, [0.0, 0.0, 0])
    e[0] += n; e[1] += float(r[ix['# Samples']] or 0); e[2] += 1
ti = sum(e[0] for e in bins.values()); ts = sum(e[1] for e in bins.values())
for b, e in sorted(bins.items()):
    print(f"active threads {b:6s}: {e[0] / W:7.1f} warp-instr/warp-step ({100 * e[0] / ti:4.1f} %), {100 * e[1] / ts:4.1f} % of samples, {e[2]} SASS instrs")
