"""Summarise an .ncu-rep: key raw metrics + executed-instruction mix per warp-step (source page)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; warps = float(sys.argv[2]); steps = float(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = dict(zip(rows[h], rows[-1]))
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_mio_throttle']
print("| metric | value |\n|---|---|")
for k in keys:
    if k in d: print(f"| {k} | {d[k]} |")
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
W = warps * steps
byop = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    try: n = float(r[ix['Instructions Executed']])
    except Exception: continue
    src = r[ix['Source']].strip()
    toks = src.split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    byop[op] += n; tot += n; samp[op] += float(r[ix['# Samples']] or 0)
print(f"\nexecuted warp-instructions per warp-step: {tot / W:.1f}\n")
print("| opcode | per warp-step | stall samples |\n|---|---|---|")
for op, n in byop.most_common(24): print(f"| {op} | {n / W:.1f} | {samp[op]:.0f} |")
