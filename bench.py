#!/usr/bin/env python
"""bench.py -- trajectory-steps/sec of the FTL/FTRL regret sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: `--trials` seeded sequences per GPU
(g-path family, keyed (base_seed, T, run) exactly like fast_algorithms.py:231-239) x T rounds
x A (algorithm, eta0) variants, generated + simulated + compared by the fused sweep kernel,
plus (N > 1) the NCCL all-reduce of the per-variant regret statistics.

  workload c3 (default, BASELINE config 3): d = 1, T = 1e4, FTL + FTRL at 8 eta0 scales (A = 9)
  workload c5 (BASELINE config 5 shape):     d = 5, T = 1e4, FTRL eta0 = sqrt 2 (A = 1), g(T) maximum

value  : device-timed (CUDA events on the launching stream), seeds already on the device.
e2e    : the same sweep through the public host API (engine.regret_sweep: host arguments in,
         per-trajectory regrets + statistics copied back to host memory inside the timed region).
         The reference API for this path (empirical_worst_case_thresholds(T_grid, runs, base_seed))
         takes seeds, not arrays, so the per-step host->device payload is the variant table only.
roofline: the sweep kernel generates its inputs on-chip, so its bound is the fp64 pipe, not HBM
         (SURVEY.md 8d): achieved = (9d+6) algorithmic non-fused fp64 ops per trajectory-step x
         trajectory-steps/s of the kernel; peak = non-FMA fp64 issue rate measured live by the
         library's microbenchmark (MEASURED_PEAKS.json has no fp64 entry).
cpu_baseline / --impl reference: the CPU oracle port of the same sweep (oracle/oco_oracle.c,
         generation + A simulations per sequence, OpenMP over all host threads) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R2 = math.sqrt(2.0)
WORKLOADS = {
    "c3": dict(d=1, T=10_000, alg=[1] + [0] * 8, eta=[R2] + [R2 * s for s in (0.125, 0.25, 0.5, 1, 2, 4, 8, 16)],
               trials=1_000_000, name="C3 batched d=1 regret sweep: T=1e4 x (FTL + FTRL x 8 eta0 scales)"),
    "c5": dict(d=5, T=10_000, alg=[0], eta=[R2], trials=1_000_000,
               name="C5 empirical g_T sweep: d=5, T=1e4, FTRL eta0=sqrt2"),
    # streamed: sequences resident in HBM in the reference layout z[B][T][d] (written once by the
    # generator kernel, untimed); one step = one pass of the TMA-fed FTRL kernel over the resident chunk
    "c4": dict(d=64, T=10_000, alg=[0], eta=[R2], trials=14208, streamed=True,   # 2 full waves of 3 CTAs x 148 SMs x 16 trajectories
               name="C4 d=64 constrained FTRL on streamed loss vectors (resident chunk of the 1e5-trial job), T=1e4"),
}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set(); self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_threads() -> int:
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: do not trust it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sweep_rate(wl, budget_s: float, nthreads: int = 0):
    """Oracle port timed on the host cores: returns (traj_steps_per_s, cores, sample description)."""
    from oracle import oracle as O
    A = len(wl["alg"])
    cores = nthreads or host_threads()
    n0 = max(cores, 8)
    t0 = time.perf_counter(); O.regret_sweep(0, wl["T"], 0, n0, wl["d"], wl["alg"], wl["eta"], nthreads=cores)
    dt0 = time.perf_counter() - t0
    n = int(max(n0, min(200_000, n0 * budget_s / max(dt0, 1e-6))))
    n -= n % cores or 0
    n = max(n, cores)
    t0 = time.perf_counter(); O.regret_sweep(0, wl["T"], n0, n, wl["d"], wl["alg"], wl["eta"], nthreads=cores)
    dt = time.perf_counter() - t0
    return n * A * wl["T"] / dt, cores, f"{n} of the workload's trials x T={wl['T']} x A={A} (generation + simulation), {dt:.1f} s"


def cpu_stream_rate(wl, budget_s: float):
    """Oracle port of the streamed workload (pre-generated sequences, simulation only)."""
    import numpy as np
    from oracle import oracle as O
    cores = host_threads()
    n = max(cores, 16)
    zs, ys = zip(*[O.gen_gpath(0, wl["T"], r, wl["d"]) for r in range(n)])
    z, y = np.stack(zs), np.stack(ys)
    O.simulate_batch(z[:cores], y[:cores], wl["alg"], wl["eta"], nthreads=cores)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        O.simulate_batch(z, y, wl["alg"], wl["eta"], nthreads=cores); reps += 1
    dt = time.perf_counter() - t0
    return reps * n * len(wl["alg"]) * wl["T"] / dt, cores, f"{n} resident sequences x T={wl['T']} x d={wl['d']}, {reps} passes, simulation only, {dt:.1f} s"


def run_streamed(args, wl, *, emit=True, init_dist=True):
    """BASELINE config 4: HBM-resident sequences streamed through the TMA-fed wide-d kernel."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from online_convex_optimization_b200 import engine
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and init_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, T, d, A = wl["trials"], wl["T"], wl["d"], len(wl["alg"])
    mode = args.comparator
    z, y = engine.gen_sequences_device("gpath", B, T, d, key_a=0, key_c0=rank * B, dtype=torch.float64)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        return engine.simulate_batch_device(z, y, wl["alg"], wl["eta"], comparator=mode)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        reg = step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record(); reg = step(); ev[k][1].record()
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = float(world) * B * A * T * args.steps / (ms * 1e-3)
    # e2e: host buffers (pinned) through the host API, PCIe copies inside the timed region, bounded sample
    nb = min(B, 512)
    zh = z[:nb].cpu().pin_memory().numpy(); yh = y[:nb].cpu().pin_memory().numpy()
    # warm-up at the timed size: the first full-size call grows the library's stream-ordered pool (device chunk buffers)
    engine.simulate_batch(zh, yh, wl["alg"], wl["eta"], comparator=mode, device=local)
    barrier()
    t0 = time.perf_counter()
    out = engine.simulate_batch(zh, yh, wl["alg"], wl["eta"], comparator=mode, device=local)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = float(world) * nb * A * T / float(t.item())
    del z, y, flush
    torch.cuda.empty_cache()
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        passes = 1 if mode == "linear" else 2
        bytes_per_step = (d + 1) * 8 / 1            # compulsory bytes, counted once
        kernel_rate = B * A * T * args.steps / (ms * 1e-3)
        achieved = kernel_rate * bytes_per_step / 1e9
        line = {
            "metric": "trajectory-steps/sec", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "trials_per_gpu": B, "T": T, "d": d, "variants": A, "comparator": mode,
                       "resident_bytes": int(B) * T * (d + 1) * 8,
                       "l2": "resident chunk (%.1f GB) >> 126 MB L2, plus a 256 MiB flush between timed steps" % (B * T * (d + 1) * 8 / 1e9),
                       "parallelism": f"dp{world} (independent sequences per rank, no collective)"},
            "e2e": {"value": e2e_val, "unit": "trajectory-steps/s", "h2d_bytes_per_step": int(nb) * T * (d + 1) * 8,
                    "d2h_bytes_per_step": int(nb) * A * 8, "sample": f"{nb} sequences from pinned host memory",
                    "api": "engine.simulate_batch (oco_simulate_batch_host)"},
            "gpu_launches": args.steps * 2, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (lambda v: None if v is None else v * passes)(ncu_traffic("sim_wide_tma_kernel<64,16,double>", float(B) * A * T)),
                         "traffic_note": "bytes per launch from the committed ncu capture (linear mode) x passes over the input",
                         "kernel": "sim_wide_tma_kernel<64,16,double>", "algorithmic_bytes_per_trajectory_step": bytes_per_step,
                         "passes_over_input": passes, "peak_nominal": 8000.0, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "result_check": {"mean_regret": float(reg.mean().item()), "max_regret": float(reg.max().item())},
        }
        if world == 1 and not args.no_cpu and emit:
            v, cores, sample = cpu_stream_rate(wl, min(args.cpu_budget, 10.0))
            line["cpu_baseline"] = {"value": v, "unit": "trajectory-steps/s", "cores": cores, "kind": "port", "sample": sample}
        if emit:
            emit_line(line)
    if world > 1 and init_dist:
        dist.destroy_process_group()
    return line


def run_reference(args, wl):
    """--impl reference: the CPU port of the same sweep, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    A = len(wl["alg"]); cores = host_threads()
    if wl.get("streamed"):      # pre-generated resident sequences, simulation only (what the streamed kernel does)
        v, cores, sample = cpu_stream_rate(wl, max(5.0, min(60.0, 10.0 * args.steps)))
        emit_line({"impl": "reference", "metric": "trajectory-steps/sec", "value": v, "unit": "trajectory-steps/s",
                   "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                   "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                   "config": {"workload": wl["name"], "T": wl["T"], "d": wl["d"], "variants": A},
                   "cpu_baseline": {"value": v, "unit": "trajectory-steps/s", "cores": cores, "kind": "port", "sample": sample},
                   "e2e": {"value": v, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    n0 = max(cores, 8)
    t0 = time.perf_counter(); O.regret_sweep(0, wl["T"], 0, n0, wl["d"], wl["alg"], wl["eta"], nthreads=cores)
    dt0 = time.perf_counter() - t0
    n = max(cores, int(n0 * per_step_budget / max(dt0, 1e-6)))
    for _ in range(args.warmup):
        O.regret_sweep(0, wl["T"], 0, n, wl["d"], wl["alg"], wl["eta"], nthreads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        O.regret_sweep(0, wl["T"], k * n, n, wl["d"], wl["alg"], wl["eta"], nthreads=cores)
    dt = time.perf_counter() - t0
    val = args.steps * n * A * wl["T"] / dt
    sample = f"{n} trials/step x T={wl['T']} x A={A}, oracle port (C, OpenMP), generation included"
    emit_line({
        "impl": "reference", "metric": "trajectory-steps/sec", "value": val, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "trials_per_step": n, "T": wl["T"], "d": wl["d"], "variants": A},
        "cpu_baseline": {"value": val, "unit": "trajectory-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def ncu_traffic(kernel: str, units_per_launch: float):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, scaled from the committed ncu capture
    (profiles/r01_dram_traffic.json: bytes per trajectory-step of that kernel), or None."""
    rec = profile_record(kernel)
    return rec["dram_bytes_per_trajectory_step"] * units_per_launch if rec else None


def ncu_utilisation(kernel: str):
    """Issue-slot and fp64-pipe utilisation of `kernel` from the committed ncu capture (profiles/r01_dram_traffic.json
    <- profiles/r01_ncu_summary.md): north_star's stated bound for sequences generated on the device."""
    rec = profile_record(kernel)
    if not rec:
        return None
    return {"issue_active_pct": rec["issue_active_pct"], "fp64_pipe_active_pct": rec["fp64_pipe_active_pct"],
            "warp_instr_per_warp_step": rec.get("warp_instr_per_warp_step"),
            "source": "committed ncu capture under profiles/ (ncu --set full, not measured in this run)"}


def hbm_view(traffic_bytes_per_step, ms_per_step):
    """The same kernel against the HBM roofline (why it is NOT the bound): DRAM bytes per step from the committed ncu
    capture divided by the measured step time, against MEASURED_PEAKS.json hbm_gbs."""
    if not traffic_bytes_per_step:
        return None
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    gbs = traffic_bytes_per_step / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
            "note": "slab traffic only; inputs are generated on chip, so HBM is not the limiter of this kernel"}


def _quiet_stdout():
    """Libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the duration of the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()



PROFILE_JSON = "r02_kernel_profiles.json"


def profile_record(kernel: str):
    """Committed ncu capture of `kernel` (profiles/r02_kernel_profiles.json, falling back to round 1's file)."""
    for name in (PROFILE_JSON, "r01_dram_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]
        except Exception:
            continue
    return None


def issue_bound(kernel: str, sm_count: int, sm_mhz: float):
    """Issue-slot bound of a generated-on-device kernel: SMs x 4 schedulers x clock / (warp-instructions per
    warp-step, from the committed ncu capture) x 32 trajectories per warp-step -- the ceiling the kernel would reach
    at 100 % issue utilisation WITH ITS CURRENT instruction count (north_star: 'issue-slot / FP64-pipe utilisation')."""
    rec = profile_record(kernel)
    if not rec or not rec.get("warp_instr_per_warp_step") or not sm_mhz:
        return None
    return sm_count * 4 * sm_mhz * 1e6 / rec["warp_instr_per_warp_step"] * 32 * rec.get("trajectories_per_thread", 1)


def timed(torch, fn, reps):
    """Mean device time (ms) of fn() over `reps` calls after one warm-up call, CUDA events on the current stream."""
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def run_c5(args, torch, dist, engine, odist, rank, world, local, dev, fp64_peak, sm_count, sm_mhz):
    """BASELINE config 5, the north-star target: empirical g(T) sweep, 1e7 runs x T = 1e4 x d = 5, FTRL eta0 = sqrt 2,
    STRONG-scaled: the 1e7 runs are sharded over the ranks (dist.shard_runs) and the per-T curve goes through ONE
    collective (dist.g_sweep_sharded == fast_algorithms.empirical_worst_case_thresholds, fast_algorithms.py:211-247).
    Timed on the device (events around sweep + collective), max over ranks.  e2e: the host API on host arguments
    (engine.g_sweep on this rank's shard, statistics back in host memory, then the collective)."""
    import numpy as np
    runs, T, d = args.c5_runs, 10_000, 5
    out = {}
    for mode in ("exact_max", "linear", "two_pass"):
        # exact_max: the reference's g(T) bit for bit at single-pass speed (single-pass upper bound + literal second pass on
        # the candidates for the maximum); linear: the single-pass value itself; two_pass: the literal pass for EVERY run
        n_timed = 1 if mode == "two_pass" else 2
        n_runs = runs
        odist.g_sweep_sharded([T], n_runs if mode != "two_pass" else min(n_runs, world * 128 * 148 * 6), comparator=mode)   # warm-up (slab allocation, pool growth)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_timed):
            g, st = odist.g_sweep_sharded([T], n_runs, comparator=mode)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n_timed], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rate = n_runs * T / (ms * 1e-3)
        kname = "gsweep_kernel<5,0,1>"       # every mode runs the pipelined kernel; two_pass adds the literal pass over the slab
        rec = profile_record(kname) or {}
        ib = issue_bound(kname, sm_count, sm_mhz)
        per_gpu = rate / world
        out[mode] = {
            "value": rate, "unit": "trajectory-steps/s", "runs": n_runs, "T": T, "d": d, "n_gpus": world, "scaling": "strong",
            "ms_per_sweep": ms, "g_T": float(g[0].item()), "count": float(st[0, 3].item()),
            "roofline": {"bound": "fp64-pipe", "achieved": per_gpu * 51 / 1e9, "peak": fp64_peak / 1e9,
                         "unit": "Gop/s (non-fused fp64)", "frac": per_gpu * 51 / fp64_peak, "kernel": kname,
                         "algorithmic_ops_per_trajectory_step": 51,
                         "traffic": (rec.get("dram_bytes_per_trajectory_step") or 0) * (n_runs / world) * T or None,
                         "traffic_from_profile": True,
                         "issue_slot_bound": None if ib is None else {
                             "steps_per_s_at_100pct_issue": ib, "frac": per_gpu / ib,
                             "warp_instr_per_warp_step": rec.get("warp_instr_per_warp_step"),
                             "note": "the kernel is integer-ALU / issue bound (numpy-exact PCG64 + ziggurat: ~400 of its "
                                     "~600 instructions per warp-step), not fp64-pipe bound; both views reported"}},
        }
    # e2e through the host API (host arguments in, per-T statistics in host memory), linear mode, full job
    lo, n = odist.shard_runs(runs, rank, world)
    engine.g_sweep(np.array([T]), runs=min(n, 1024), run_lo=lo, comparator="exact_max", device=local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    gh, sth = engine.g_sweep(np.array([T]), runs=n, run_lo=lo, comparator="exact_max", device=local)
    if n == 0:
        sth[:, 0] = -np.inf
    stt = torch.from_numpy(sth).to(dev)
    odist.allreduce_stats(stt)
    sth = stt.cpu().numpy()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["exact_max"]["e2e"] = {"value": runs * T / float(t.item()), "unit": "trajectory-steps/s", "h2d_bytes_per_step": 8 + 64,
                            "d2h_bytes_per_step": 32, "api": "engine.g_sweep(comparator='exact_max') (oco_g_sweep_mode_host) + one collective",
                            "g_T": float(max(sth[0, 0], 0.0))}
    # parity spot-check inside the measured job: this rank's first 4 runs at full length against the CPU oracle
    from oracle import oracle as O            # checker only (never on the measured path)
    want = O.regret_sweep(0, T, lo, 4, d, [0], [R2])
    err = {}
    for mode in ("two_pass", "linear"):
        got = engine.regret_sweep("gpath", 4, T, d, [0], [R2], key_a=0, key_c0=lo, comparator=mode, full=True, device=local)
        e = float(np.max(np.abs(got["regret"] - want) / np.maximum(1.0, got["cum_loss"])))
        err[mode] = e
    te = torch.tensor([err["two_pass"], err["linear"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    out["result_check"] = {"oracle_spot_check": f"4 runs per rank x T={T} x d={d} vs the CPU oracle, max over {world} rank(s)",
                           "max_abs_err_over_max_1_cum__two_pass": float(te[0].item()), "two_pass_bit_exact": float(te[0].item()) == 0.0,
                           "max_abs_err_over_max_1_cum__linear": float(te[1].item()), "tolerance": 1e-12,
                           "g_T_exact_max": out["exact_max"]["g_T"], "g_T_two_pass": out["two_pass"]["g_T"], "g_T_linear": out["linear"]["g_T"],
                           "g_T_exact_max_bit_identical_to_two_pass": out["exact_max"]["g_T"] == out["two_pass"]["g_T"]}
    return out


def run_other_single_gpu(args, torch, engine, local, fp64_peak):
    """Every other number the docs quote, measured in the driver-run line (N = 1 only): C1, C2, exact drivers,
    general exact solver, SMART at T = 1e4, the float32 mode, the full C4 job with generation."""
    import numpy as np
    ex = {}

    def wall(fn, reps=1):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, r
    try:    # C1: the reference's default fast_driver workload (fast_driver.py:201-220 minus plotting; 42.6 s on one CPU core)
        from online_convex_optimization_b200 import fast_driver as fd
        dt, res = wall(lambda: fd.run_experiment(), 2)
        ex["c1_fast_driver_default"] = {"wall_s": dt, "trajectory_steps": 4.36e7, "value": 4.36e7 / dt, "unit": "trajectory-steps/s (nominal)",
                                        "g_1000": float(res["g_emp"][1000]), "reference_cpu_1core_s": 42.6}
    except Exception as e:
        ex["c1_fast_driver_default"] = {"error": repr(e)}
    try:    # C2: exact FTL vs fast FTL on the same sequences, 1e4 trials (d = 5, T = 100), iterates compared on the device
        z, y = engine.gen_sequences("gpath", 10_000, 100, 5)
        dt, exr = wall(lambda: engine.exact_ftl_prefix_batch(z, y, want_actions=True))
        sim = engine.simulate_batch(z, y, [1], [R2], full=True, trace=True)
        same = bool((exr["actions"][:, :100].view(np.uint64) == sim["x_trace"][:, 0].view(np.uint64)).all())
        ex["c2_exact_vs_fast_ftl_1e4_trials"] = {"wall_s": dt, "prefix_solves": 1_000_000, "value": 1e6 / dt, "unit": "exact prefix solves/s (host API, copies included)",
                                                 "iterates_bit_identical_to_fast_ftl": same}
        from online_convex_optimization_b200 import exact_ftl_driver as xd
        dt, res = wall(lambda: xd.run_experiment())
        ex["c2_exact_ftl_driver_default"] = {"wall_s": dt, "prefix_solves": 1.06e7, "value": 1.06e7 / dt, "unit": "exact prefix solves/s", "g_100": float(res["g_emp"][100])}
        zz = z[:80] * 1.7          # outside the linear regime: the general interior-point solver
        dt, g = wall(lambda: engine.exact_ftl_general_batch(zz, y[:80], norm="l2"))
        ex["exact_general_solver"] = {"wall_s": dt, "prefix_solves": 8000, "value": 8000 / dt, "unit": "prefix solves/s (l2, d=5, rows scaled x1.7)",
                                      "unconverged": int(g["status"].sum())}
    except Exception as e:
        ex["c2_error"] = repr(e)
    try:    # SMART at T = 1e4 (fast_algorithms.py:118-164), d = 5, 2 thresholds
        T, B = 10_000, 4096
        zd, yd = engine.gen_sequences_device("gpath", B, T, 5, dtype=torch.float64)
        th, et = [math.sqrt(2 * T), 60.0], [R2, R2]
        ms, _ = timed(torch, lambda: engine.smart_batch_device(zd, yd, th, et, mode="linear"), 2)
        ms2, (_, sw2) = timed(torch, lambda: engine.smart_batch_device(zd, yd, th, et, mode="exact"), 2)
        ms3, _ = timed(torch, lambda: engine.smart_batch_device(zd[:592], yd[:592], th, et, mode="exact"), 2)
        ex["smart_t1e4"] = {"linear_mode": {"value": B * T * 2 / (ms * 1e-3), "sequences": B, "ms": ms},
                            "exact_mode": {"value": B * T * 2 / (ms2 * 1e-3), "sequences": B, "ms": ms2,
                                           "switched_fraction": float((sw2 >= 0).double().mean().item()),
                                           "note": "literal switch test; the O(t) prefix loss is evaluated only in rounds the exact bound "
                                                   "0.5((t+1) - s.vy) <= s_loss cannot rule out (kernels_warp.cu: smart_prune_margin)"},
                            "exact_mode_592_sequences_cta_kernel": {"value": 592 * T * 2 / (ms3 * 1e-3), "ms": ms3},
                            "unit": "trajectory-steps/s", "T": T, "d": 5, "thresholds": 2}
        del zd, yd
    except Exception as e:
        ex["smart_t1e4"] = {"error": repr(e)}
    try:    # float32 compute mode (algorithms.py:28-54) vs the float64 kernel on the same device-resident float32 sequences
        T, B = 10_000, 65_536
        zf, yf = engine.gen_sequences_device("iid", B, T, 5, key_a=2025, key_c0=13, dtype=torch.float32)
        ms64, _ = timed(torch, lambda: engine.simulate_batch_device(zf, yf, [0, 1], [R2, R2]), 2)
        ms32, _ = timed(torch, lambda: engine.simulate_batch_device(zf, yf, [0, 1], [R2, R2], compute="f32"), 2)
        ex["fp32_mode_d5"] = {"f32_state": {"value": B * T * 2 / (ms32 * 1e-3), "ms": ms32}, "f64_state": {"value": B * T * 2 / (ms64 * 1e-3), "ms": ms64},
                              "speedup_f32_over_f64": ms64 / ms32, "unit": "trajectory-steps/s", "sequences": B, "T": T, "variants": 2,
                              "input_bytes_per_trajectory_step": 12}
        del zf, yf
    except Exception as e:
        ex["fp32_mode_d5"] = {"error": repr(e)}
    torch.cuda.empty_cache(); engine.release_scratch(local)
    try:    # the FULL C4 job: 1e5 trials x T = 1e4 x d = 64 = 520 GB of sequences, generated AND simulated chunk by chunk:
        # two resident chunks of 7104 sequences (one wave of the TMA kernel each, 37 GB), the generator of chunk k + 1
        # (issue bound) on a second stream under the simulation of chunk k (HBM bound)
        T, d, chunk, total = 10_000, 64, 7_104, 100_000
        dev = torch.device("cuda", local)
        bufs = [(torch.empty((chunk, T, d), dtype=torch.float64, device=dev), torch.empty((chunk, T), dtype=torch.float64, device=dev)) for _ in range(2)]
        s_sim, s_gen = torch.cuda.current_stream(), torch.cuda.Stream()
        starts = list(range(0, total, chunk))

        def job(mode):
            best = torch.full((), -float("inf"), dtype=torch.float64, device=dev)
            ev_fill = [torch.cuda.Event() for _ in starts]; ev_sim = [torch.cuda.Event() for _ in starts]

            def fill(k):
                nb = min(chunk, total - starts[k]); z, y = bufs[k & 1]
                with torch.cuda.stream(s_gen):
                    if k >= 2:
                        s_gen.wait_event(ev_sim[k - 2])                       # the buffer's previous chunk has been simulated
                    engine.gen_sequences_device("gpath", nb, T, d, key_a=0, key_c0=starts[k], out=(z[:nb], y[:nb]))
                    ev_fill[k].record(s_gen)
            s_gen.wait_stream(s_sim)
            fill(0)
            for k in range(len(starts)):
                if k + 1 < len(starts):
                    fill(k + 1)
                nb = min(chunk, total - starts[k]); z, y = bufs[k & 1]
                s_sim.wait_event(ev_fill[k])
                reg = engine.simulate_batch_device(z[:nb], y[:nb], [0], [R2], comparator=mode)
                best = torch.maximum(best, reg.max())
                ev_sim[k].record(s_sim)
            s_sim.wait_stream(s_gen)
            return best
        job_ms = {}
        for mode in ("linear", "two_pass"):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m = job(mode); e1.record(); torch.cuda.synchronize()
            job_ms[mode] = (e0.elapsed_time(e1), float(m.item()))
        ex["c4_full_job_1e5_trials"] = {m: {"wall_ms": v[0], "value": total * T / (v[0] * 1e-3), "max_regret": v[1]} for m, v in job_ms.items()}
        ex["c4_full_job_1e5_trials"].update({"unit": "trajectory-steps/s", "trials": total, "T": T, "d": d,
                                             "includes": "on-device generation (warp-per-sequence generator) of all 15 chunks of 7104 sequences, "
                                                         "overlapped on a second stream with the simulation of the previous chunk"})
        del bufs
    except Exception as e:
        ex["c4_full_job_1e5_trials"] = {"error": repr(e)}
    torch.cuda.empty_cache(); engine.release_scratch(local)
    return ex


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--trials", type=int, default=0, help="sequences per GPU per step (default: the workload's)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU baseline work (rank 0, N=1 only)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the brief c5 / c4 measurements appended to the default c3 line")
    ap.add_argument("--comparator", default="linear", choices=["linear", "two_pass"], help="c4 only")
    ap.add_argument("--c5-runs", type=int, default=10_000_000, help="total runs of the strong-scaled C5 g(T) sweep (other_workloads)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.trials:
        wl["trials"] = args.trials
    if args.impl == "reference":
        return run_reference(args, wl)
    if wl.get("streamed"):
        return run_streamed(args, wl)

    import numpy as np
    import torch
    import torch.distributed as dist
    from online_convex_optimization_b200 import dist as odist
    from online_convex_optimization_b200 import engine

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, T, d, A = wl["trials"], wl["T"], wl["d"], len(wl["alg"])
    run_lo = rank * B                                   # weak scaling: every rank owns its own block of runs
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(k):
        reg, stats = engine.regret_sweep_device("gpath", B, T, d, wl["alg"], wl["eta"], key_a=0, key_c0=run_lo)
        odist.allreduce_stats(stats)
        return reg, stats

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xff)                            # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        reg, stats = step(k)
        ev[k][1].record()
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_steps = float(world) * B * A * T * args.steps
    value = total_steps / (ms * 1e-3)

    # ---- e2e through the public host API (host args in, regrets + stats back in host memory)
    e2e_steps = max(1, min(args.steps, 3))
    # warm-up with TWO live results: `out = f()` keeps the previous result alive while the next one is produced, so the
    # steady state of the loop below uses two output buffers (page-locked, from torch's caching host allocator)
    w1 = engine.regret_sweep("gpath", B, T, d, wl["alg"], wl["eta"], key_a=0, key_c0=run_lo, device=local)
    w2 = engine.regret_sweep("gpath", B, T, d, wl["alg"], wl["eta"], key_a=0, key_c0=run_lo, device=local)
    del w1, w2
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        out = engine.regret_sweep("gpath", B, T, d, wl["alg"], wl["eta"], key_a=0, key_c0=run_lo, device=local)
        if world > 1:
            st = torch.from_numpy(out["stats"]).to(dev); odist.allreduce_stats(st); st.cpu()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = float(world) * B * A * T * e2e_steps / (float(t.item()) * 1e-3)

    if rank == 0:
        st = stats.cpu().numpy()
        fp64_peak = engine.fp64_peak_ops(local)
        kernel_steps_per_s = B * A * T * args.steps / (ms * 1e-3)           # this GPU's kernel rate
        alg_ops = (9 * d + 6)
        # the kernel this workload launches (d = 1 g-path: the pipelined gsweep_kernel; profiles/r02_kernel_profiles.json)
        kname = "gsweep_kernel<%d,%d,%d>" % (
            d, int(any(f != 0 for f in wl["alg"])),
            (lambda n: 1 if n <= 1 else (4 if n <= 4 else 8))(sum(1 for f in wl["alg"] if f == 0)))
        sm_count0 = torch.cuda.get_device_properties(local).multi_processor_count
        ib0 = issue_bound(kname, sm_count0, (clocks or {}).get("sm_max_mhz"))
        line = {
            "metric": "trajectory-steps/sec", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "trials_per_gpu": B, "T": T, "d": d, "variants": A,
                       "sequence_source": "generated on device from numpy stream keys (base_seed=0, T, run)",
                       "l2": "256 MiB flush between timed steps; inputs are generated on-chip (no HBM input traffic)",
                       "parallelism": f"dp{world} (runs sharded, one all-reduce of [A,4] statistics per step)"},
            "e2e": {"value": e2e_val, "unit": "trajectory-steps/s", "h2d_bytes_per_step": A * 12 + 64,
                    "d2h_bytes_per_step": B * A * 8 + A * 32, "steps": e2e_steps,
                    "api": "engine.regret_sweep (oco_regret_sweep_host): per-trajectory regrets + stats to host"},
            "gpu_launches": args.steps * 3 + (args.steps if world > 1 else 0),
            "clocks": clocks,
            "roofline": {"bound": "fp64-pipe", "achieved": kernel_steps_per_s * alg_ops / 1e9, "peak": fp64_peak / 1e9,
                         "unit": "Gop/s (non-fused fp64)", "frac": kernel_steps_per_s * alg_ops / fp64_peak,
                         "traffic": ncu_traffic(kname, float(B) * A * T), "traffic_from_profile": True, "kernel": kname,
                         "algorithmic_ops_per_trajectory_step": alg_ops,
                         "peak_source": "oco_fp64_peak_probe_host measured live (alternating DMUL/DADD, no FMA)",
                         "ncu": ncu_utilisation(kname),
                         "issue_slot_bound": None if ib0 is None else {
                             "steps_per_s_at_100pct_issue": ib0, "frac": kernel_steps_per_s / ib0,
                             "note": "SMs x 4 schedulers x clock / warp-instructions per warp-step (committed ncu capture) x "
                                     "32 lanes x trajectories per thread: the ceiling at the kernel's current instruction count"},
                         "hbm_view": hbm_view(ncu_traffic(kname, float(B) * A * T), ms / args.steps)},
            "result_check": {"max_regret_per_variant": [float(v) for v in st[:, 0]],
                             "mean_regret_per_variant": [float(v) for v in st[:, 1] / np.maximum(st[:, 3], 1)]},
        }
        if world == 1 and not args.no_cpu:
            v, cores, sample = cpu_sweep_rate(wl, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": "trajectory-steps/s", "cores": cores, "kind": "port", "sample": sample}
    # ---- the north-star configuration (C5: 1e7 runs x T = 1e4 x d = 5 g(T) sweep, strong-scaled over the N ranks, one
    # collective of the per-T curve) is measured at EVERY N; the remaining quoted workloads at N = 1
    if args.workload == "c3" and not args.no_extras:
        del flush
        torch.cuda.empty_cache()
        engine.release_scratch(local)
        extras = {}
        fp64_peak_all = engine.fp64_peak_ops(local)
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz = None
        try:
            sm_mhz = float(subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"],
                                          capture_output=True, text=True, timeout=5).stdout.strip())
        except Exception:
            pass
        try:
            c5 = run_c5(args, torch, dist, engine, odist, rank, world, local, dev, fp64_peak_all, sm_count, sm_mhz)
            extras["c5_g_sweep_1e7_runs_strong_scaled"] = c5
        except Exception as ex:      # extras must never cost the headline number
            extras["c5_error"] = repr(ex)
        engine.release_scratch(local)
        if world == 1:
            try:
                extras.update(run_other_single_gpu(args, torch, engine, local, fp64_peak_all))
                sargs = argparse.Namespace(**vars(args)); sargs.steps, sargs.warmup = 3, 2
                for mode in ("linear", "two_pass"):
                    sargs.comparator = mode
                    l4 = run_streamed(sargs, dict(WORKLOADS["c4"]), emit=False, init_dist=False)
                    extras["c4_d64_streamed_" + mode] = {k: l4[k] for k in ("value", "unit", "ms_per_step", "roofline", "e2e", "config")}
            except Exception as ex:
                extras["error"] = repr(ex)
        if rank == 0:
            line["other_workloads"] = extras
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
