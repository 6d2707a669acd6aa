"""Batched entry points over the C ABI (include/oco_b200.h).

Host API (numpy in / numpy out, copies done by the library's own double-buffered pipeline):
    simulate_batch, smart_batch, gen_sequences, regret_sweep, g_sweep
Device API (torch CUDA tensors in / out, enqueued on torch's current stream, no sync):
    simulate_batch_device, smart_batch_device, gen_sequences_device, regret_sweep_device

Everything here is argument marshalling; the arithmetic is in csrc/*.cu.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (COMPARATOR_EXACT_MAX, COMPARATOR_LINEAR, COMPARATOR_TWO_PASS, FAM_FLIP, FAM_GPATH, FAM_IID, FAM_NOISY, FAM_SWITCH,
                   OCO_F32, OCO_F64)

SQRT2 = math.sqrt(2.0)
FAMILIES = {"gpath": FAM_GPATH, "iid": FAM_IID, "noisy": FAM_NOISY, "flip": FAM_FLIP, "switch": FAM_SWITCH}
class _Comparators(dict):
    def __missing__(self, key):
        raise ValueError(f"comparator must be 'two_pass', 'linear' or (g sweeps only) 'exact_max' (got {key!r})")


COMPARATORS = _Comparators({"two_pass": COMPARATOR_TWO_PASS, "linear": COMPARATOR_LINEAR, "exact_max": COMPARATOR_EXACT_MAX})


def _vp(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(None)
    return C.c_void_p(a.ctypes.data)


def _variants(alg_flags, eta0s):
    af = np.ascontiguousarray(np.atleast_1d(alg_flags), dtype=np.int32)
    et = np.ascontiguousarray(np.atleast_1d(eta0s), dtype=np.float64)
    if af.shape != et.shape or af.ndim != 1:
        raise ValueError("alg_flags and eta0s must be 1-D and of equal length")
    return af, et


def _prep_zy(z, y):
    z = np.asarray(z)
    y = np.asarray(y)
    if z.dtype == np.float32 and y.dtype == np.float32:
        dt = np.float32
    else:
        dt = np.float64
    z = np.ascontiguousarray(z, dtype=dt)
    y = np.ascontiguousarray(y, dtype=dt)
    if z.ndim == 2:
        z = z[None]
        y = y[None]
    if z.ndim != 3 or y.ndim != 2 or y.shape != z.shape[:2]:
        raise ValueError(f"expected z (B,T,d) and y (B,T); got {z.shape} and {y.shape}")
    return z, y, (OCO_F32 if dt == np.float32 else OCO_F64)


def _keys(keys, B):
    if keys is None:
        return None
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    if k.shape != (B, 2):
        raise ValueError(f"keys must have shape ({B}, 2)")
    return k


# ----------------------------------------------------------------------------- host API
def simulate_batch(z, y, alg_flags, eta0s, *, full: bool = False, trace: bool = False,
                   comparator: str = "two_pass", comparator_action=None, compute: str = "f64", device: int = 0):
    """Batched fast_algorithms.simulate_alg: z (B,T,d), y (B,T) -> regret (B,A).
    full=True returns a dict with cum_loss, comp_loss, theta (B,A,d), x_last (B,A,d; d <= 8, or any d <= 256 with comparator_action)
    [and x_trace (B,A,T,d)].  comparator_action (B,d): fixed comparator (exact_ftl.run_ftrl).
    compute="f32": the float32-state semantics of the reference's algorithms.py (float32 inputs)."""
    z, y, dt = _prep_zy(z, y)
    B, T, d = z.shape
    af, et = _variants(alg_flags, eta0s)
    A = af.size
    reg = np.empty((B, A))
    cum = np.empty((B, A)) if full else None
    comp = np.empty((B, A)) if full else None
    th = np.empty((B, A, d)) if full else None
    xl = np.zeros((B, A, d)) if (full and (d <= 8 or comparator_action is not None)) else None
    xt = np.empty((B, A, T, d)) if trace else None
    ca = None
    if comparator_action is not None:
        ca = np.ascontiguousarray(comparator_action, dtype=np.float64).reshape(-1, d)
        if ca.shape != (B, d):
            raise ValueError(f"comparator_action must have shape ({B}, {d})")
    mode = COMPARATORS[comparator]
    cdt = {"f64": OCO_F64, "f32": OCO_F32}[compute]
    _lib.check(_lib.load().oco_simulate_batch_host(_vp(z), _vp(y), dt, cdt, B, T, d, _vp(af), _vp(et), A, _vp(ca),
                                                   _vp(reg), _vp(cum), _vp(comp), _vp(th), _vp(xl), _vp(xt), mode, device))
    if not full and not trace:
        return reg
    return {"regret": reg, "cum_loss": cum, "comp_loss": comp, "theta": th, "x_last": xl, "x_trace": xt}


def exact_ftl_prefix_batch(z, y, *, want_actions: bool = True, device: int = 0):
    """Batched exact-FTL prefix actions + replay (unit l2 ball, linear regime; see oco_b200.h).
    Returns dict(actions (B,T+1,d) | None, cum_loss, comp_loss, regret (B,), status (B,) int32)."""
    z, y, dt = _prep_zy(z, y)
    B, T, d = z.shape
    act = np.empty((B, T + 1, d)) if want_actions else None
    cum, comp, reg = np.empty(B), np.empty(B), np.empty(B)
    status = np.zeros(B, dtype=np.int32)
    _lib.check(_lib.load().oco_exact_ftl_prefix_batch_host(_vp(z), _vp(y), dt, B, T, d, _vp(act), _vp(cum), _vp(comp),
                                                           _vp(reg), _vp(status), device))
    return {"actions": act, "cum_loss": cum, "comp_loss": comp, "regret": reg, "status": status}


NORMS = {"l2": 0, "linf": 1, "l1": 2}


def exact_ftl_general_batch(z, y, *, norm: str = "l2", max_iter: int = 300, tol: float = 1e-8, first_prefix: int = 0,
                            device: int = 0):
    """General exact-FTL prefix actions + replay (any rows / labels, unit l2 / linf / l1 ball; see oco_b200.h).
    first_prefix = T solves only the full sequence (shorter prefixes' actions are zeros unless the closed form gave them).
    Returns dict(actions (B,T+1,d), cum_loss, comp_loss, regret (B,), status (B,) = prefixes that used up max_iter Newton steps)."""
    z, y, dt = _prep_zy(z, y)
    B, T, d = z.shape
    act = np.empty((B, T + 1, d))
    cum, comp, reg = np.empty(B), np.empty(B), np.empty(B)
    status = np.zeros(B, dtype=np.int32)
    _lib.check(_lib.load().oco_exact_ftl_general_batch_host(_vp(z), _vp(y), dt, B, T, d, NORMS[norm], int(max_iter), float(tol), int(first_prefix),
                                                            _vp(act), _vp(cum), _vp(comp), _vp(reg), _vp(status), device))
    return {"actions": act, "cum_loss": cum, "comp_loss": comp, "regret": reg, "status": status}


def replay_actions(z, y, actions, *, device: int = 0):
    """Batched exact_ftl.replay_exact_ftl: z (B,T,d), y (B,T), actions (B,T+1,d) -> dict of (B,) arrays."""
    z, y, dt = _prep_zy(z, y)
    B, T, d = z.shape
    act = np.ascontiguousarray(actions, dtype=np.float64)
    if act.shape != (B, T + 1, d):
        raise ValueError("actions must have shape (B, T+1, d)")
    cum, comp, reg = np.empty(B), np.empty(B), np.empty(B)
    _lib.check(_lib.load().oco_replay_actions_batch_host(_vp(z), _vp(y), dt, B, T, d, _vp(act), _vp(cum), _vp(comp), _vp(reg), device))
    return {"cum_loss": cum, "comp_loss": comp, "regret": reg}


SMART_MODES = {"exact": 0, "linear": 1}


def smart_batch(z, y, theta_threshs, eta0s, *, mode: str = "exact", device: int = 0):
    """Batched fast_algorithms.simulate_SMART_like -> (regret (B,A), switch_t (B,A)).
    mode="linear": O(T) prefix loss in the linear regime (see oco_b200.h), exact fallback per sequence."""
    z, y, dt = _prep_zy(z, y)
    B, T, d = z.shape
    th = np.ascontiguousarray(np.atleast_1d(theta_threshs), dtype=np.float64)
    et = np.ascontiguousarray(np.atleast_1d(eta0s), dtype=np.float64)
    if th.shape != et.shape or th.ndim != 1:
        raise ValueError("theta_threshs and eta0s must be 1-D and of equal length")
    A = th.size
    reg = np.empty((B, A))
    sw = np.empty((B, A), dtype=np.int64)
    _lib.check(_lib.load().oco_smart_batch_host(_vp(z), _vp(y), dt, B, T, d, _vp(th), _vp(et), A, _vp(reg), _vp(sw),
                                                SMART_MODES[mode], device))
    return reg, sw


def gen_sequences(family, B: int, T: int, d: int = 5, *, key_a: int = 0, key_c0: int = 0, keys=None,
                  p_flip: float = -1.0, block_len: int = 20, dtype=None, device: int = 0):
    """Sequences of one family as numpy arrays z (B,T,d), y (B,T) (see oco_gen_sequences)."""
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    if dtype is None:
        dtype = np.float64 if fam == FAM_GPATH else np.float32
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("dtype must be float32 or float64")
    z = np.empty((B, T, d), dtype=dtype)
    y = np.empty((B, T), dtype=dtype)
    k = _keys(keys, B)
    _lib.check(_lib.load().oco_gen_sequences_host(fam, key_a, key_c0, _vp(k), B, T, d, float(p_flip), block_len,
                                                  OCO_F32 if dtype == np.float32 else OCO_F64, _vp(z), _vp(y), device))
    return z, y


_PINNED_MIN_BYTES = 8 << 20


def _host_out(shape):
    """Float64 output array for a host-API call.  Large outputs (>= 8 MB: the per-trajectory regrets of a big sweep) are
    backed by page-locked memory from torch's caching host allocator when CUDA is available, so the library's
    device-to-host copy is one DMA at PCIe speed; the driver's staged pageable copy depends on the host (measured 4 to
    45 ms for 72 MB from box to box).  OCO_PINNED_OUTPUTS=0 keeps plain numpy arrays."""
    n = int(np.prod(shape))
    if n * 8 >= _PINNED_MIN_BYTES and os.environ.get("OCO_PINNED_OUTPUTS", "1") != "0":
        try:
            torch = _torch()
            if torch.cuda.is_available():
                return torch.empty(tuple(int(v) for v in shape), dtype=torch.float64, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape)


def regret_sweep(family, B: int, T: int, d: int, alg_flags, eta0s, *, key_a: int = 0, key_c0: int = 0, keys=None,
                 p_flip: float = -1.0, want_regret: bool = True, full: bool = False, comparator: str = "two_pass",
                 device: int = 0):
    """Fused generate+simulate+compare.  Returns dict(regret (B,A) | None, stats (A,4) = max,sum,sumsq,count
    [, cum_loss, comp_loss]).  comparator="linear" (g-path, 2 <= d <= 8): no second pass, no scratch traffic,
    regrets within 1e-12 * max(1, cum_loss) of the reference (see include/oco_b200.h)."""
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    af, et = _variants(alg_flags, eta0s)
    A = af.size
    reg = _host_out((B, A)) if want_regret else None
    cum = _host_out((B, A)) if full else None
    comp = _host_out((B, A)) if full else None
    stats = np.zeros((A, 4))
    k = _keys(keys, B)
    _lib.check(_lib.load().oco_regret_sweep_mode_host(fam, key_a, key_c0, _vp(k), B, T, d, float(p_flip), _vp(af), _vp(et), A,
                                                      _vp(reg), _vp(cum), _vp(comp), _vp(stats), COMPARATORS[comparator], device))
    return {"regret": reg, "stats": stats, "cum_loss": cum, "comp_loss": comp}


def g_sweep(T_grid, *, runs: int, base_seed: int = 0, run_lo: int = 0, d: int = 5, eta0: float = SQRT2,
            comparator: str = "two_pass", device: int = 0):
    """empirical_worst_case_thresholds for runs [run_lo, run_lo+runs) -> (g_max (nT,), stats (nT,4)).
    comparator: "two_pass" (literal second pass for every run), "linear" (single pass, 1e-12 class) or "exact_max":
    the single-pass regret bounds the literal one from above for every run (|q - y| >= 1 - y q), so only the runs whose
    bound reaches the largest bound take the literal second pass -- g_max equals "two_pass" bit for bit at "linear" speed
    (the sum / sum of squares in stats are the single-pass ones)."""
    tg = np.ascontiguousarray(T_grid, dtype=np.int64)
    g = np.zeros(tg.size)
    st = np.zeros((tg.size, 4))
    _lib.check(_lib.load().oco_g_sweep_mode_host(_vp(tg), tg.size, run_lo, runs, base_seed, d, float(eta0), _vp(g), _vp(st),
                                                 COMPARATORS[comparator], device))
    return g, st


def rng_probe(what: int, key, n: int, n_skip: int = 0, device: int = 0):
    out = np.zeros(n, dtype=np.uint64 if what in (0, 4) else np.float64)
    _lib.check(_lib.load().oco_rng_probe_host(what, int(key[0]), int(key[1]), int(key[2]), n_skip, n, _vp(out), device))
    return out


def release_scratch(device: int = 0) -> None:
    """Return the library's cached scratch memory (sweep slab, tables) to the driver."""
    _lib.check(_lib.load().oco_release_scratch(device))


def fp64_peak_ops(device: int = 0) -> float:
    """Measured issue rate (ops/s) of independent, non-fused fp64 multiplies and adds."""
    v = C.c_double(0.0)
    _lib.check(_lib.load().oco_fp64_peak_probe_host(device, C.byref(v)))
    return v.value


# ----------------------------------------------------------------------------- device API (torch)
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: online_convex_optimization_b200 has no CPU fallback")
    return torch


def _stream_ptr(torch) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dptr(t) -> C.c_void_p:
    return C.c_void_p(None if t is None else t.data_ptr())


def _check_device_inputs(torch, z, y):
    """The kernels read raw pointers: z (B,T,d) and y (B,T) must be contiguous CUDA tensors of ONE float dtype on one device."""
    if z.dtype != y.dtype or z.dtype not in (torch.float32, torch.float64):
        raise ValueError("z and y must both be float32 or float64")
    if not (z.is_cuda and y.is_cuda and z.is_contiguous() and y.is_contiguous()):
        raise ValueError("z and y must be contiguous CUDA tensors")
    if z.device != y.device or z.dim() != 3 or y.dim() != 2 or tuple(y.shape) != tuple(z.shape[:2]):
        raise ValueError("need z (B,T,d) and y (B,T) on the same device")
    return z.shape


def simulate_batch_device(z, y, alg_flags, eta0s, *, full: bool = False, comparator: str = "two_pass", compute: str = "f64"):
    """z (B,T,d), y (B,T): contiguous CUDA tensors (float32 or float64).  Returns CUDA tensors."""
    torch = _torch()
    B, T, d = _check_device_inputs(torch, z, y)
    af, et = _variants(alg_flags, eta0s)
    A = af.size
    reg = torch.empty((B, A), dtype=torch.float64, device=z.device)
    cum = torch.empty_like(reg) if full else None
    comp = torch.empty_like(reg) if full else None
    th = torch.empty((B, A, d), dtype=torch.float64, device=z.device) if full else None
    mode = COMPARATORS[comparator]
    with torch.cuda.device(z.device):
        _lib.check(_lib.load().oco_simulate_batch(_dptr(z), _dptr(y), OCO_F32 if z.dtype == torch.float32 else OCO_F64,
                                                  {"f64": OCO_F64, "f32": OCO_F32}[compute], B, T, d, _vp(af), _vp(et), A, None, _dptr(reg), _dptr(cum), _dptr(comp),
                                                  _dptr(th), None, None, mode, _stream_ptr(torch)))
    if not full:
        return reg
    return {"regret": reg, "cum_loss": cum, "comp_loss": comp, "theta": th}


def smart_batch_device(z, y, theta_threshs, eta0s, *, mode: str = "exact"):
    torch = _torch()
    B, T, d = _check_device_inputs(torch, z, y)
    th = np.ascontiguousarray(np.atleast_1d(theta_threshs), dtype=np.float64)
    et = np.ascontiguousarray(np.atleast_1d(eta0s), dtype=np.float64)
    A = th.size
    reg = torch.empty((B, A), dtype=torch.float64, device=z.device)
    sw = torch.empty((B, A), dtype=torch.int64, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(_lib.load().oco_smart_batch(_dptr(z), _dptr(y), OCO_F32 if z.dtype == torch.float32 else OCO_F64,
                                               B, T, d, _vp(th), _vp(et), A, _dptr(reg), _dptr(sw), SMART_MODES[mode],
                                               _stream_ptr(torch)))
    return reg, sw


def gen_sequences_device(family, B: int, T: int, d: int = 5, *, key_a: int = 0, key_c0: int = 0, keys=None,
                         p_flip: float = -1.0, block_len: int = 20, dtype=None, device=None, out=None):
    """Sequences of one family written to device memory (enqueued on the current stream, not synchronised).
    out=(z, y): existing contiguous CUDA tensors of shape (B,T,d) / (B,T) to fill (double-buffered pipelines)."""
    torch = _torch()
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    if out is not None:
        z, y = out
        if tuple(z.shape) != (B, T, d):
            raise ValueError(f"out[0] must have shape ({B}, {T}, {d})")
        _check_device_inputs(torch, z, y)
        dtype, dev = z.dtype, z.device
    else:
        if dtype is None:
            dtype = torch.float64 if fam == FAM_GPATH else torch.float32
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        z = torch.empty((B, T, d), dtype=dtype, device=dev)
        y = torch.empty((B, T), dtype=dtype, device=dev)
    k = _keys(keys, B)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().oco_gen_sequences(fam, key_a, key_c0, _vp(k), B, T, d, float(p_flip), block_len,
                                                 OCO_F32 if dtype == torch.float32 else OCO_F64, _dptr(z), _dptr(y),
                                                 _stream_ptr(torch)))
    return z, y


def regret_sweep_device(family, B: int, T: int, d: int, alg_flags, eta0s, *, key_a: int = 0, key_c0: int = 0, keys=None,
                        p_flip: float = -1.0, want_regret: bool = True, comparator: str = "two_pass", device=None):
    """Returns (regret (B,A) CUDA tensor or None, stats (A,4) CUDA tensor); enqueued, not synchronised."""
    torch = _torch()
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    af, et = _variants(alg_flags, eta0s)
    A = af.size
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    reg = torch.empty((B, A), dtype=torch.float64, device=dev) if want_regret else None
    stats = torch.empty((A, 4), dtype=torch.float64, device=dev)
    k = _keys(keys, B)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().oco_regret_sweep_mode(fam, key_a, key_c0, _vp(k), B, T, d, float(p_flip), _vp(af), _vp(et), A,
                                                     _dptr(reg), None, None, _dptr(stats), COMPARATORS[comparator],
                                                     _stream_ptr(torch)))
    return reg, stats
