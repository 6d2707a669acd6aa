"""Update profiles/r02_kernel_profiles.json from an ncu report (what bench.py's `roofline.traffic` / issue-slot bound scale).
usage: python tools/make_kernel_profiles.py REP KEY WARPS STEPS TRAJ_PER_THREAD
  REP    .ncu-rep of ONE launch (ncu --set full)         KEY   the entry name, e.g. "gsweep_kernel<1,1,8>"
  WARPS  warps of sequences in that launch (B / 32)      STEPS T      TRAJ_PER_THREAD trajectories per sequence"""
import csv, json, os, subprocess, sys
rep, key, warps, steps, tpt = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
units = dict(zip(rows[h], rows[h + 1])); d = dict(zip(rows[h], rows[-1]))
def val(k):
    v = float(d[k].replace(',', ''))
    u = units.get(k, '')
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}.get(u, 1.0)
traj_steps = warps * 32 * steps * tpt
rec = {
    "ncu_report": os.path.basename(rep),
    "dram_bytes_read": val('dram__bytes_read.sum'), "dram_bytes_write": val('dram__bytes_write.sum'),
    "trajectory_steps_per_launch": traj_steps,
    "dram_bytes_per_trajectory_step": (val('dram__bytes_read.sum') + val('dram__bytes_write.sum')) / traj_steps,
    "issue_active_pct": float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']),
    "fp64_pipe_active_pct": float(d['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']),
    "kernel_ms_under_ncu": float(d['gpu__time_duration.sum']) * {'ms': 1.0, 'us': 1e-3, 's': 1e3, 'ns': 1e-6}.get(units.get('gpu__time_duration.sum', 'ms'), 1.0),
    "warp_instr_per_warp_step": round(float(d['smsp__inst_executed.sum'].replace(',', '')) / (warps * steps), 1),
    "registers_per_thread": int(float(d['launch__registers_per_thread'])),
    "trajectories_per_thread": tpt,
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "r02_kernel_profiles.json")
js = json.load(open(path))
js[key] = rec
json.dump(js, open(path, "w"), indent=1)
print(key, json.dumps(rec))
