"""ctypes front-end of the CPU oracle (oracle/oco_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(online_convex_optimization_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboco_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)
c_i64p = C.POINTER(C.c_int64)
c_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile liboco_oracle.so with the committed Makefile (gcc, seconds)."""
    src = os.path.join(_HERE, "oco_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- RNG probes
def seedseq3(a, b, c):
    out = np.zeros(4, np.uint64)
    lib().oracle_seedseq3(C.c_uint64(a), C.c_uint64(b), C.c_uint64(c), _p(out, c_u64p))
    return out


def rng_raw(key, n):
    out = np.zeros(n, np.uint64)
    lib().oracle_rng_raw(C.c_uint64(key[0]), C.c_uint64(key[1]), C.c_uint64(key[2]), C.c_int64(n), _p(out, c_u64p))
    return out


def rng_normals(key, n):
    out = np.zeros(n, np.float64)
    lib().oracle_rng_normals(C.c_uint64(key[0]), C.c_uint64(key[1]), C.c_uint64(key[2]), C.c_int64(n), _p(out, c_dp))
    return out


def rng_doubles(key, n):
    out = np.zeros(n, np.float64)
    lib().oracle_rng_doubles(C.c_uint64(key[0]), C.c_uint64(key[1]), C.c_uint64(key[2]), C.c_int64(n), _p(out, c_dp))
    return out


def rng_choice_pm1(key, n_normals, n):
    out = np.zeros(n, np.float64)
    lib().oracle_rng_choice_pm1(C.c_uint64(key[0]), C.c_uint64(key[1]), C.c_uint64(key[2]),
                                C.c_int64(n_normals), C.c_int64(n), _p(out, c_dp))
    return out


def log1p(x):
    x = _f64(x)
    out = np.empty_like(x)
    lib().oracle_log1p(_p(x, c_dp), C.c_int64(x.size), _p(out, c_dp))
    return out


# ------------------------------------------------------------ sequence families
def gen_gpath(base_seed, T, run, d=5):
    """fast_algorithms.py:231-239 -> (z (T,d) f64, y (T,) f64)."""
    z = np.empty((T, d), np.float64)
    y = np.empty(T, np.float64)
    lib().oracle_gen_gpath(C.c_uint64(base_seed), C.c_int64(T), C.c_uint64(run), C.c_int(d), _p(z, c_dp), _p(y, c_dp))
    return z, y


def gen_task_u(run_seed, stream_id, d=5):
    u = np.empty(d, np.float32)
    lib().oracle_gen_task_u(C.c_uint64(run_seed), C.c_uint64(stream_id), C.c_int(d), _p(u, c_fp))
    return u


def gen_iid(run_seed, T, rep, d=5, noisy_p=None):
    """sequence_generation.py:54-90 -> (z f32, y f32, u f32)."""
    u = gen_task_u(run_seed, 21 if noisy_p is not None else 11, d)
    z = np.empty((T, d), np.float32)
    y = np.empty(T, np.float32)
    stream = (23 if noisy_p is not None else 13) + rep
    lib().oracle_gen_iid_f32(C.c_uint64(run_seed), C.c_int64(T), C.c_uint64(stream), C.c_int(d), _p(u, c_fp),
                             C.c_double(-1.0 if noisy_p is None else noisy_p), _p(z, c_fp), _p(y, c_fp))
    return z, y, u


def gen_flip(T, d=5):
    z = np.empty((T, d), np.float32)
    y = np.empty(T, np.float32)
    lib().oracle_gen_flip_f32(C.c_int64(T), C.c_int(d), _p(z, c_fp), _p(y, c_fp))
    return z, y, np.zeros(d, np.float32)


def gen_switching(T, block_len=20, d=5):
    z = np.empty((T, d), np.float32)
    y = np.empty(T, np.float32)
    lib().oracle_gen_switching_f32(C.c_int64(T), C.c_int64(block_len), C.c_int(d), _p(z, c_fp), _p(y, c_fp))
    return z, y, np.zeros(d, np.float32)


# ------------------------------------------------------------------ simulators
def simulate_alg(z, y, alg_flag, eta0, *, full=False, trace=False):
    """fast_algorithms.py:171-177.  full=True -> dict with cum/comp/theta (and x trace)."""
    z = _f64(z)
    y = _f64(y)
    T, d = z.shape
    out = np.zeros(3)
    th = np.zeros(d)
    xt = np.zeros((T, d)) if trace else None
    lib().oracle_simulate_alg(_p(z, c_dp), _p(y, c_dp), C.c_int64(T), C.c_int(d), C.c_int(int(alg_flag)),
                              C.c_double(float(eta0)), _p(out, c_dp), _p(th, c_dp),
                              _p(xt, c_dp) if trace else None)
    if not full and not trace:
        return float(out[0])
    return {"regret": float(out[0]), "cum_loss": float(out[1]), "comp_loss": float(out[2]), "theta": th, "x_trace": xt}


def simulate_smart_like(z, y, theta_thresh, eta0, *, full=False):
    """fast_algorithms.py:184-195."""
    z = _f64(z)
    y = _f64(y)
    T, d = z.shape
    out = np.zeros(3)
    sw = C.c_int64(-1)
    lib().oracle_simulate_smart(_p(z, c_dp), _p(y, c_dp), C.c_int64(T), C.c_int(d), C.c_double(float(theta_thresh)),
                                C.c_double(float(eta0)), _p(out, c_dp), C.byref(sw))
    if not full:
        return float(out[0])
    return {"regret": float(out[0]), "total_loss": float(out[1]), "comp_loss": float(out[2]), "switch_t": int(sw.value)}


def simulate_batch(z, y, alg_flags, eta0s, *, nthreads=0, full=False):
    """z (B,T,d), y (B,T) f32 or f64 -> regret (B,A) [and cum, comp, theta]."""
    z = np.ascontiguousarray(z)
    y = np.ascontiguousarray(y)
    f32 = z.dtype == np.float32
    if not f32:
        z = _f64(z)
    y = np.ascontiguousarray(y, dtype=z.dtype)
    B, T, d = z.shape
    af = np.ascontiguousarray(alg_flags, dtype=np.int32)
    et = _f64(eta0s)
    A = af.size
    reg = np.zeros((B, A))
    cum = np.zeros((B, A))
    comp = np.zeros((B, A))
    th = np.zeros((B, A, d))
    lib().oracle_simulate_batch(z.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int(int(f32)),
                                C.c_int64(B), C.c_int64(T), C.c_int(d), _p(af, c_ip), _p(et, c_dp), C.c_int(A),
                                _p(reg, c_dp), _p(cum, c_dp), _p(comp, c_dp), _p(th, c_dp), C.c_int(nthreads))
    if full:
        return {"regret": reg, "cum_loss": cum, "comp_loss": comp, "theta": th}
    return reg


def smart_batch(z, y, threshs, eta0s, *, nthreads=0):
    z = np.ascontiguousarray(z)
    f32 = z.dtype == np.float32
    if not f32:
        z = _f64(z)
    y = np.ascontiguousarray(y, dtype=z.dtype)
    B, T, d = z.shape
    th = _f64(threshs)
    et = _f64(eta0s)
    A = th.size
    reg = np.zeros((B, A))
    sw = np.zeros((B, A), np.int64)
    lib().oracle_smart_batch(z.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int(int(f32)),
                             C.c_int64(B), C.c_int64(T), C.c_int(d), _p(th, c_dp), _p(et, c_dp), C.c_int(A),
                             _p(reg, c_dp), _p(sw, c_i64p), C.c_int(nthreads))
    return reg, sw


def regret_sweep(base_seed, T, run_lo, n_runs, d, alg_flags, eta0s, *, nthreads=0):
    af = np.ascontiguousarray(alg_flags, dtype=np.int32)
    et = _f64(eta0s)
    A = af.size
    reg = np.zeros((n_runs, A))
    lib().oracle_regret_sweep(C.c_uint64(base_seed), C.c_int64(T), C.c_uint64(run_lo), C.c_int64(n_runs), C.c_int(d),
                              _p(af, c_ip), _p(et, c_dp), C.c_int(A), _p(reg, c_dp), C.c_int(nthreads))
    return reg


def empirical_worst_case_thresholds(T_grid, *, runs=5, base_seed=0, nthreads=0):
    """fast_algorithms.py:211-247."""
    tg = np.ascontiguousarray(T_grid, dtype=np.int64)
    g = np.zeros(tg.size)
    lib().oracle_g_thresholds(_p(tg, c_i64p), C.c_int(tg.size), C.c_int64(runs), C.c_uint64(base_seed),
                              _p(g, c_dp), C.c_int(nthreads))
    return {int(t): float(v) for t, v in zip(tg, g)}


def exact_ftl_l2_prefix(z, y):
    """Closed-form exact-FTL prefix actions (linear regime); PARITY UNPINNED."""
    z = _f64(z)
    y = _f64(y)
    T, d = z.shape
    act = np.zeros((T + 1, d))
    rc = lib().oracle_exact_ftl_l2_prefix(_p(z, c_dp), _p(y, c_dp), C.c_int64(T), C.c_int(d), _p(act, c_dp))
    if rc != 0:
        raise ValueError("rows outside the linear regime (||z||<=1, y=+-1)")
    return act


def replay_actions(z, y, actions):
    z = _f64(z)
    y = _f64(y)
    a = _f64(actions)
    T, d = z.shape
    out = np.zeros(3)
    lib().oracle_replay_actions(_p(z, c_dp), _p(y, c_dp), C.c_int64(T), C.c_int(d), _p(a, c_dp), _p(out, c_dp))
    return {"regret": float(out[0]), "cum_loss": float(out[1]), "comp_loss": float(out[2])}


def num_threads() -> int:
    return int(lib().oracle_num_threads())
