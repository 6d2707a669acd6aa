"""ctypes binding of liboco_b200.so (C ABI: include/oco_b200.h).  Fails loudly when the
library is missing and cannot be built; never falls back to a CPU implementation."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liboco_b200.so")

OCO_OK, OCO_ERR_VALUE, OCO_ERR_RUNTIME = 0, 1, 2
OCO_F32, OCO_F64 = 0, 1
FAM_GPATH, FAM_IID, FAM_NOISY, FAM_FLIP, FAM_SWITCH = 0, 1, 2, 3, 4
COMPARATOR_TWO_PASS, COMPARATOR_LINEAR, COMPARATOR_EXACT_MAX = 0, 1, 2
MAX_VARIANTS, MAX_SMART_THRESHOLDS = 64, 4

_vp, _i, _i64, _u64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double

# name -> argtypes, exactly the prototypes of include/oco_b200.h
SIGNATURES = {
    "oco_last_error": ([], C.c_char_p),
    "oco_abi_version": ([], _i),
    "oco_device_count": ([C.POINTER(_i)], _i),
    "oco_simulate_batch": ([_vp, _vp, _i, _i, _i64, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "oco_simulate_batch_host": ([_vp, _vp, _i, _i, _i64, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i], _i),
    "oco_exact_ftl_prefix_batch": ([_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "oco_exact_ftl_general_batch": ([_vp, _vp, _i, _i64, _i64, _i, _i, _i, _dbl, _i64, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "oco_exact_ftl_general_batch_host": ([_vp, _vp, _i, _i64, _i64, _i, _i, _i, _dbl, _i64, _vp, _vp, _vp, _vp, _vp, _i], _i),
    "oco_replay_actions_batch_host": ([_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _vp, _i], _i),
    "oco_exact_ftl_prefix_batch_host": ([_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i], _i),
    "oco_smart_batch": ([_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _i, _vp, _vp, _i, _vp], _i),
    "oco_smart_batch_host": ([_vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _i, _vp, _vp, _i, _i], _i),
    "oco_gen_sequences": ([_i, _u64, _u64, _vp, _i64, _i64, _i, _dbl, _i64, _i, _vp, _vp, _vp], _i),
    "oco_gen_sequences_host": ([_i, _u64, _u64, _vp, _i64, _i64, _i, _dbl, _i64, _i, _vp, _vp, _i], _i),
    "oco_regret_sweep": ([_i, _u64, _u64, _vp, _i64, _i64, _i, _dbl, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "oco_regret_sweep_host": ([_i, _u64, _u64, _vp, _i64, _i64, _i, _dbl, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i], _i),
    "oco_g_sweep_host": ([_vp, _i, _u64, _i64, _u64, _i, _dbl, _vp, _vp, _i], _i),
    "oco_regret_sweep_mode": ([_i, _u64, _u64, _vp, _i64, _i64, _i, _dbl, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "oco_regret_sweep_mode_host": ([_i, _u64, _u64, _vp, _i64, _i64, _i, _dbl, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i], _i),
    "oco_g_sweep_mode_host": ([_vp, _i, _u64, _i64, _u64, _i, _dbl, _vp, _vp, _i, _i], _i),
    "oco_rng_probe_host": ([_i, _u64, _u64, _u64, _i64, _i64, _vp, _i], _i),
    "oco_release_scratch": ([_i], _i),
    "oco_fp64_peak_probe_host": ([_i, C.POINTER(_dbl)], _i),
}

_lib = None


def load() -> C.CDLL:
    """Load (building first if the .so is absent and nvcc exists).  Raises RuntimeError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # no nvcc / compile error: fail loudly, no fallback
            raise RuntimeError(
                f"liboco_b200.so is missing at {LIB_PATH} and could not be built ({e}); "
                "run `python -m online_convex_optimization_b200.build`. There is no CPU fallback.") from e
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map ABI status codes to the reference's exception types (exact_ftl.py:125-126, 142-143)."""
    if rc == OCO_OK:
        return
    msg = (load().oco_last_error() or b"").decode()
    if rc == OCO_ERR_VALUE:
        raise ValueError(msg)
    raise RuntimeError(msg)


def device_count() -> int:
    n = C.c_int(0)
    check(load().oco_device_count(C.byref(n)))
    return n.value
