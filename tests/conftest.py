import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def hexf(s):
    return float.fromhex(s)


@pytest.fixture(scope="session")
def golden_sim():
    return json.load(open(os.path.join(GOLDEN, "sim_cases.json")))


@pytest.fixture(scope="session")
def golden_rng():
    return np.load(os.path.join(GOLDEN, "rng_kat.npz"))


@pytest.fixture(scope="session")
def golden_iter():
    return np.load(os.path.join(GOLDEN, "iterates.npz"))


@pytest.fixture(scope="session")
def golden_stats():
    return json.load(open(os.path.join(GOLDEN, "driver_stats.json")))


@pytest.fixture(scope="session")
def golden_exact_det():
    return json.load(open(os.path.join(GOLDEN, "exact_det.json")))


@pytest.fixture(scope="session")
def golden_driver32():
    return json.load(open(os.path.join(GOLDEN, "driver32_stats.json")))


def exact_det_inputs(case, index):
    """Inputs of one tests/golden/exact_det.json case: (z, y, {"closed": actions, "random": actions}); the recipe is
    the one tools/make_golden.py used (actions are INPUTS of the reference's replay, not its outputs)."""
    import math
    spec = case["spec"]
    if spec["family"] == "smoke_main":                  # exact_ftl.py:460-468
        rng = np.random.default_rng(0)
        T, d = 200, 10
        z = rng.normal(size=(T, d))
        z /= np.maximum(1.0, np.linalg.norm(z, axis=1, keepdims=True))
        w_true = rng.normal(size=d)
        w_true /= np.linalg.norm(w_true)
        y = np.sign(z @ w_true + 0.2 * rng.normal(size=T))
    else:
        z, y = make_sequence(spec)
    z = np.ascontiguousarray(z, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    T, d = z.shape
    closed = np.zeros((T + 1, d)); v = np.zeros(d)
    for k in range(T):
        v = v + y[k] * z[k]
        n = math.sqrt(float(v @ v))
        closed[k + 1] = v / n if n > 0 else 0.0
    r = np.random.default_rng(index + 1).standard_normal((T + 1, d))
    r /= np.maximum(1.0, np.linalg.norm(r, axis=1, keepdims=True)); r[0] = 0.0
    return z, y, {"closed": closed, "random": r}


def make_sequence(spec):
    """Rebuild a golden case's input sequence with the ORACLE generators (pinned
    against the reference's numpy generators by the tests that compare results)."""
    from oracle import oracle as O
    fam = spec["family"]
    if fam == "gpath":
        return O.gen_gpath(spec["base_seed"], spec["T"], spec["run"], spec["d"])
    if fam == "Random i.i.d. (separable)":
        return O.gen_iid(spec["run_seed"], spec["T"], spec["rep"])[:2]
    if fam == "Massart noise 10%":
        return O.gen_iid(spec["run_seed"], spec["T"], spec["rep"], noisy_p=0.10)[:2]
    if fam == "Label flips":
        return O.gen_flip(spec["T"])[:2]
    if fam == "Switching leaders":
        return O.gen_switching(spec["T"])[:2]
    raise KeyError(fam)
