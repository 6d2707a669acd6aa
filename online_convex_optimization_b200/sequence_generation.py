"""Drop-in twin of the reference's ``sequence_generation`` module: the same builders, case
registry and per-case run / replicate counts (sequence_generation.py:24-123), with the
sequences produced by the device generators (csrc/streams.cuh) that restate numpy's
SeedSequence -> PCG64 -> ziggurat draws for the same stream keys.

Each sampler also exposes ``.spec`` (family id, run_seed, p) so the batched driver
(``fast_driver.evaluate_stream_with_stats``) can fuse generation into the simulation kernels
instead of materialising arrays one (T, rep) at a time.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np

from . import engine
from ._lib import FAM_FLIP, FAM_IID, FAM_NOISY, FAM_SWITCH


def flip_sequence(T: int, d: int = 5):
    """sequence_generation.py:24-28."""
    z, y = engine.gen_sequences(FAM_FLIP, 1, T, d, dtype=np.float32)
    return z[0], y[0], np.zeros(d, dtype=np.float32)


def switching_two_leaders_sequence(T: int, *, block_len: int = 20, d: int = 5):
    """sequence_generation.py:36-47."""
    z, y = engine.gen_sequences(FAM_SWITCH, 1, T, d, block_len=block_len, dtype=np.float32)
    return z[0], y[0], np.zeros(d, dtype=np.float32)


def _task_vector(run_seed: int, stream_id: int, d: int) -> np.ndarray:
    """u of sequence_generation.py:55-59 / :73-77 (host numpy; it is returned for API parity only --
    the drivers ignore it and the device generator derives its own copy for the labels)."""
    gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence([run_seed, 0, stream_id])))
    u = gen.standard_normal(d).astype(np.float32, copy=False)
    n = float(np.linalg.norm(u))
    if n > 0:
        u /= n
    return u


def make_random_iid_stream(*, d: int = 5, run_seed: int = 0):
    """sequence_generation.py:54-70."""
    u_cache = []                                  # u is formed on first use: the batched drivers never ask for it

    def sample(T: int, rep: int = 0):
        if not u_cache:
            u_cache.append(_task_vector(run_seed, 11, d))
        z, y = engine.gen_sequences(FAM_IID, 1, T, d, key_a=run_seed, key_c0=13 + rep, dtype=np.float32)
        return z[0], y[0], u_cache[0]
    sample.spec = {"family": FAM_IID, "run_seed": run_seed, "d": d, "p": -1.0, "stream0": 13}
    return sample


def make_noisy_iid_stream(*, p: float, d: int = 5, run_seed: int = 0):
    """sequence_generation.py:72-90."""
    u_cache = []

    def sample(T: int, rep: int = 0):
        if not u_cache:
            u_cache.append(_task_vector(run_seed, 21, d))
        z, y = engine.gen_sequences(FAM_NOISY, 1, T, d, key_a=run_seed, key_c0=23 + rep, p_flip=p, dtype=np.float32)
        return z[0], y[0], u_cache[0]
    sample.spec = {"family": FAM_NOISY, "run_seed": run_seed, "d": d, "p": float(p), "stream0": 23}
    return sample


def make_flip_stream(*, d: int = 5, run_seed: int = 0):
    """sequence_generation.py:92-95."""
    def sample(T: int, rep: int = 0):
        return flip_sequence(T, d=d)
    sample.spec = {"family": FAM_FLIP, "run_seed": run_seed, "d": d, "p": -1.0, "stream0": 0}
    return sample


def make_switching_two_leaders_stream(*, block_len: int = 20, d: int = 5, run_seed: int = 0):
    """sequence_generation.py:97-100."""
    def sample(T: int, rep: int = 0):
        return switching_two_leaders_sequence(T, block_len=block_len, d=d)
    sample.spec = {"family": FAM_SWITCH, "run_seed": run_seed, "d": d, "p": -1.0, "stream0": 0, "block_len": block_len}
    return sample


# sequence_generation.py:104-123 -- same titles, same counts
CASES: Dict[str, Callable[..., Callable[[int, int], Tuple[np.ndarray, np.ndarray, np.ndarray]]]] = {
    "Random i.i.d. (separable)": lambda *, run_seed: make_random_iid_stream(d=5, run_seed=run_seed),
    "Massart noise 10%": lambda *, run_seed: make_noisy_iid_stream(p=0.10, d=5, run_seed=run_seed),
    "Label flips": lambda *, run_seed: make_flip_stream(d=5, run_seed=run_seed),
    "Switching leaders": lambda *, run_seed: make_switching_two_leaders_stream(block_len=20, d=5, run_seed=run_seed),
}
RUNS_BY_TITLE = {"Random i.i.d. (separable)": 48, "Massart noise 10%": 48, "Label flips": 1, "Switching leaders": 1}
REPLICATES_BY_TITLE = {"Random i.i.d. (separable)": 16, "Massart noise 10%": 20, "Label flips": 1, "Switching leaders": 1}
