"""Per-CUDA-source-line stall reasons from an .ncu-rep (--import-source on, -lineinfo).
Usage: python tools/ncu_stalls.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = ""; hdr = None; agg = []
keys = ["stall_wait", "stall_long_sb", "stall_math", "stall_short_sb", "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_selected", "stall_not_selected"]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] in ("Function Name", ""): continue
    try:
        vals = [float(r[hdr[k]] or 0) for k in keys]; n = float(r[hdr["Instructions Executed"]])
    except Exception:
        continue
    agg.append((sum(vals), vals, n, cur, r[0], r[1].strip()[:70]))
tot = [sum(a[1][i] for a in agg) for i in range(len(keys))]
T = sum(tot)
print("totals: " + "  ".join(f"{k[6:]}={100*v/T:.1f}%" for k, v in zip(keys, tot)))
print("share%  wait long math short disp  | line")
for s, vals, n, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{100*s/T:5.1f}  " + " ".join(f"{100*v/T:4.1f}" for v in vals[:5]) + f"  | {f}:{ln} {src}")
