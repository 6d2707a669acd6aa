"""Per-CUDA-source-line executed warp-instructions (and stall samples) from an .ncu-rep captured with
--import-source on and -lineinfo.  Usage: python tools/ncu_lines.py rep warps steps [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; W = float(sys.argv[2]) * float(sys.argv[3]); top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] in ("Function Name",) or r[0] == "": continue
    try:
        n = float(r[hdr["Instructions Executed"]]); s = float(r[hdr["# Samples"]] or 0)
    except Exception:
        continue
    agg.append((n, s, cur_file, r[0], r[1].strip()[:90]))
tot = sum(a[0] for a in agg); ts = sum(a[1] for a in agg)
print(f"total {tot / W:.1f} warp-instr per warp-step, {ts:.0f} samples")
for n, s, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{n / W:7.1f} {100 * s / ts:5.1f}%  {f}:{ln}  {src}")
