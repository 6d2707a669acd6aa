"""Regenerate the metric / opcode / region / stall tables of profiles/r02_ncu_summary.md from the two final ncu reports
(gpurun_out/r02_gsweep_c3.ncu-rep, gpurun_out/r02_gsweep_c5.ncu-rep); the prose of each section is kept."""
import subprocess
s = open('profiles/r02_ncu_summary.md').read()
def run(cmd): return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout.rstrip()
W3, W5 = 151552 // 32, 284160 // 32
sec = {}
for tag, rep, W in (("c3", "gpurun_out/r02_gsweep_c3.ncu-rep", W3), ("c5", "gpurun_out/r02_gsweep_c5.ncu-rep", W5)):
    sec[tag] = (run(f"python tools/ncu_summary.py {rep} {W} 10000"), run(f"python tools/ncu_regions.py {rep} {W} 10000"), run(f"python tools/ncu_stalls.py {rep} 18"))
h1 = s.index("## 1. gsweep_kernel<1,1,8>"); h2 = s.index("## 2. gsweep_kernel<5,0,1>")
head, s1, s2 = s[:h1], s[h1:h2], s[h2:]
def rebuild(sect, parts):
    lines = sect.split("\n")
    intro_end = next(i for i, l in enumerate(lines) if l.startswith("| metric | value |"))
    intro = "\n".join(lines[:intro_end]).rstrip("\n") + "\n\n"
    reg_line = next(l for l in lines if l.startswith("Share of issue slots"))
    return (intro + parts[0] + "\n\n" + reg_line + "\n```\n" + parts[1] + "\n```\n\nStall reasons per source line (tools/ncu_stalls.py):\n```\n" + parts[2] + "\n```\n\n")
out = head + rebuild(s1, sec["c3"]) + rebuild(s2, sec["c5"])
open('profiles/r02_ncu_summary.md', 'w').write(out)
