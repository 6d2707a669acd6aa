"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/oco_b200.h
declares with the declared arity, validates arguments, and fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "oco_b200.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"OCO_API\s+[\w\s\*]+?\b(oco_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_header_symbols_exported_and_bound():
    from online_convex_optimization_b200 import _lib
    decl = declared()
    assert len(decl) >= 14
    lib = _lib.load()
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
        assert len(_lib.SIGNATURES[name][0]) == nargs, f"{name}: header has {nargs} args"
    assert set(_lib.SIGNATURES) == set(decl)
    assert lib.oco_abi_version() == 1


def test_no_cpu_fallback_and_error_mapping():
    import torch
    from online_convex_optimization_b200 import _lib, engine, fast_algorithms as fa
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path is not reachable")
    assert _lib.device_count() == 0
    z, y = np.zeros((4, 2)), np.zeros(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fa.simulate_alg(z, y, 0, 1.0)
    with pytest.raises(RuntimeError):
        fa.empirical_worst_case_thresholds(np.array([10]), runs=2)
    with pytest.raises(RuntimeError):
        engine.regret_sweep("gpath", 4, 10, 1, [0], [1.0])


def test_argument_validation_is_value_error():
    from online_convex_optimization_b200 import _lib, engine
    with pytest.raises(ValueError):
        engine.simulate_batch(np.zeros((2, 4, 3)), np.zeros((2, 5)), [0], [1.0])
    with pytest.raises(ValueError):
        engine.simulate_batch(np.zeros((2, 4, 3)), np.zeros((2, 4)), [0, 1], [1.0])
    with pytest.raises(ValueError):
        engine.gen_sequences("gpath", 2, 4, 5, keys=np.zeros((3, 2)))
    lib = _lib.load()
    # straight through the C ABI: bad dtype / negative sizes / NULL pointers are OCO_ERR_VALUE
    rc = lib.oco_simulate_batch(None, None, 7, 1, 1, 1, 1, None, None, 1, None, None, None, None, None, None, None, 0, None)
    assert rc == _lib.OCO_ERR_VALUE and b"dtype" in lib.oco_last_error()
    rc = lib.oco_simulate_batch(None, None, 1, 1, -1, 1, 1, None, None, 1, None, None, None, None, None, None, None, 0, None)
    assert rc == _lib.OCO_ERR_VALUE
    af = (ctypes.c_int * 1)(0)
    et = (ctypes.c_double * 1)(1.0)
    rc = lib.oco_regret_sweep(0, 0, 0, None, 4, 10, 9, -1.0, af, et, 1, None, None, None, None, None)      # g-path: d <= 8
    assert rc == _lib.OCO_ERR_VALUE and b"d <= 8" in lib.oco_last_error()
    rc = lib.oco_regret_sweep(1, 0, 0, None, 4, 10, 6, -1.0, af, et, 1, None, None, None, None, None)      # iid family: d <= 5
    assert rc == _lib.OCO_ERR_VALUE and b"d <= 5" in lib.oco_last_error()
    rc = lib.oco_smart_batch(None, None, 1, 0, 4, 3, et, et, 9, None, None, 0, None)
    assert rc == _lib.OCO_ERR_VALUE
    # shape limits of the round-2 kernels (argument validation comes before any device work)
    buf = np.zeros(8)
    dummy = buf.ctypes.data_as(ctypes.c_void_p)
    rc = lib.oco_smart_batch(dummy, dummy, 1, 1, 4, 257, et, et, 1, dummy, None, 0, None)                    # SMART: d <= 256
    assert rc == _lib.OCO_ERR_VALUE and b"d <= 256" in lib.oco_last_error()
    st = (ctypes.c_int * 1)(0)
    for norm, d_bad, msg in ((2, 17, b"l1 ball: d <= 16"), (1, 33, b"linf ball: d <= 32")):                    # general exact FTL
        rc = lib.oco_exact_ftl_general_batch(dummy, dummy, 1, 1, 4, d_bad, norm, 50, 1e-8, 0, dummy, None, None, None, st, None)
        assert rc == _lib.OCO_ERR_VALUE and msg in lib.oco_last_error(), lib.oco_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "online_convex_optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oco_oracle" not in txt, f


def test_header_is_plain_c_and_links(tmp_path):
    """include/oco_b200.h compiles as C99 and a C program links against the library."""
    import shutil
    import subprocess
    from online_convex_optimization_b200 import _lib
    _lib.load()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi_check.c"
    src.write_text('#include <stdio.h>\n#include "oco_b200.h"\n'
                   'int main(void) { int n = -1; int rc = oco_device_count(&n);\n'
                   '  int bad = oco_simulate_batch_host(0, 0, 9, OCO_F64, 1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0);\n'
                   '  printf("%d %d %d %d %s\\n", oco_abi_version(), rc, n >= 0, bad, oco_last_error()); return 0; }\n')
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-loco_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(None, 4)
    assert out[0] == "1" and out[1] == "0" and out[2] == "1" and out[3] == "1" and "dtype" in out[4]
