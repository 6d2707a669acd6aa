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
