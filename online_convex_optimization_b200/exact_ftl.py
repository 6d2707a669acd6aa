"""Twin of the reference's ``exact_ftl`` module (exact_ftl.py:62-453): same class / function
names, arguments, return types and exceptions, with the solves done by the batched CUDA path.

Two device solvers sit behind ``ExactFTLNoClip`` (DESIGN.md "Exact FTL"):

* unit **l2** ball in the *linear regime* (every row ||z_t||_2 <= 1 and y_t in {-1,+1} -- what
  every generator in sequence_generation.py produces, hence everything exact_ftl_driver.py
  runs): the reference's SOCP (exact_ftl.py:83-105) has the closed-form optimum v/||v||,
  v = sum y_i z_i (0 on ties), computed for all prefixes by ``oco_exact_ftl_prefix_batch`` for any
  d <= 256 (thread per sequence up to d = 8, warp per sequence beyond; the reference's own smoke run,
  exact_ftl.py:460-475, is d = 10 -- ``python -m online_convex_optimization_b200.exact_ftl`` runs it);
* everything else (rows outside the ball, real-valued labels, the linf / l1 balls of
  exact_ftl.py:91-96) with d <= 32 (l1: d <= 16, split variables u - w beyond d = 8): the interior-point solver of
  ``oco_exact_ftl_general_batch`` (per-lane registers up to d = 8, lane-per-variable with the Hessian in shared
  memory beyond),
  duality gap <= ``solver_opts["tol"]`` * max(1, f(0)) (default 1e-8) within
  ``solver_opts["max_iter"]`` (default 300) Newton steps per prefix; a prefix that does not converge raises the
  ``RuntimeError`` exact_ftl.py:125-126 raises when its conic solver fails.

The deterministic half of the module -- ``replay_exact_ftl``, ``_simulate_ftrl`` / ``run_ftrl(comparator_action=...)``
(exact_ftl.py:230-333, 399-420: njit helpers + NumPy, no CVXPY) -- is pinned to fixtures produced by running the
reference itself (tools/make_golden.py -> tests/golden/exact_det.json).  The conic solve is not:
PARITY UNPINNED for ``ExactFTLNoClip`` solves: the reference's arithmetic there is CVXPY's (not installed,
unpinned); where the optimum is not unique (k < d rows, ties) the two solvers may return different minimisers of
the same objective.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional, Tuple

import numpy as np

from . import engine


GENERAL_MAX_D = 32      # oco_exact_ftl_general_batch: interior-point solver, l2 / linf balls (l1: d <= 16)


def _ensure_float64_contiguous(arr: np.ndarray) -> np.ndarray:
    """exact_ftl.py:52-55."""
    if arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]:
        return arr
    return np.ascontiguousarray(arr, dtype=np.float64)


def _solve_prefixes(z: np.ndarray, y: np.ndarray, *, norm: str = "l2", solver_opts: Optional[dict] = None,
                    only_last: bool = False) -> dict:
    """All prefix optima of one sequence (only the full-length one when ``only_last``)."""
    opts = solver_opts or {}
    if norm == "l2" or z.shape[1] == 1:            # for d = 1 the three unit balls coincide
        res = engine.exact_ftl_prefix_batch(z[None], y[None], want_actions=True)
        if int(res["status"][0]) == 0:
            return res
    if z.shape[1] > GENERAL_MAX_D:
        raise RuntimeError("Exact FTL (no clip) failed: status=outside_linear_regime "
                           f"(a row has ||z||_2 > 1 or a label is not +-1, and the general solver covers d <= {GENERAL_MAX_D})")
    res = engine.exact_ftl_general_batch(z[None], y[None], norm=norm, max_iter=int(opts.get("max_iter", 300)),
                                         tol=float(opts.get("tol", 1e-8)), first_prefix=z.shape[0] if only_last else 0)
    if int(res["status"][0]) != 0:
        raise RuntimeError(f"Exact FTL (no clip) failed: status=max_iter_reached ({int(res['status'][0])} prefixes)")
    return res


class ExactFTLNoClip:
    """exact_ftl.py:62-193 -- build once, solve many; stateful row buffer with a prefix mask."""

    def __init__(self, d: int, T_max: int, *, norm: Literal["l2", "linf", "l1"] = "l2",
                 solver: Optional[str] = None, solver_opts: Optional[dict] = None) -> None:
        self.d = int(d)
        self.T_max = int(T_max)
        self.norm = norm
        self.solver = solver                    # accepted for signature parity; there is one solver here
        self.solver_opts = {} if solver_opts is None else dict(solver_opts)
        if self.norm not in ("l2", "linf", "l1"):
            raise ValueError("norm must be one of {'l2','linf','l1'}")
        if (self.norm == "l1" and self.d > GENERAL_MAX_D // 2) or (self.norm == "linf" and self.d > GENERAL_MAX_D):
            raise RuntimeError(f"Exact FTL (no clip) failed: status=unsupported_norm_{self.norm} "
                               f"(the linf solver covers d <= {GENERAL_MAX_D}, the l1 solver d <= {GENERAL_MAX_D // 2})")
        self._Z_buf = np.zeros((self.T_max, self.d), dtype=np.float64)
        self._y_buf = np.zeros(self.T_max, dtype=np.float64)
        self._w_buf = np.zeros(self.T_max, dtype=np.float64)
        self._last_length = 0

    def _solve_current(self) -> np.ndarray:
        t = self._last_length
        if t == 0:
            return np.zeros(self.d, dtype=np.float64)
        res = _solve_prefixes(self._Z_buf[:t], self._y_buf[:t], norm=self.norm, solver_opts=self.solver_opts, only_last=True)
        return np.asarray(res["actions"][0, t], dtype=np.float64)

    def reset_buffers(self) -> None:
        """exact_ftl.py:130-138."""
        self._Z_buf.fill(0.0)
        self._y_buf.fill(0.0)
        self._w_buf.fill(0.0)
        self._last_length = 0

    def append_row(self, z_row: np.ndarray, y_val: float) -> np.ndarray:
        """exact_ftl.py:140-149."""
        if self._last_length >= self.T_max:
            raise ValueError("sequence longer than T_max")
        idx = self._last_length
        self._Z_buf[idx] = z_row
        self._y_buf[idx] = y_val
        self._w_buf[idx] = 1.0
        self._last_length += 1
        return self._solve_current()

    def _set_prefix(self, z_source: np.ndarray, y_source: np.ndarray, length: int) -> None:
        """exact_ftl.py:151-169."""
        t = int(length)
        if t < 0 or t > self.T_max:
            raise ValueError("length must be between 0 and T_max inclusive")
        if t > 0:
            np.copyto(self._Z_buf[:t], z_source[:t])
            np.copyto(self._y_buf[:t], y_source[:t])
            self._w_buf[:t] = 1.0
        if t < self._last_length:
            tail = slice(t, self._last_length)
            self._Z_buf[tail] = 0.0
            self._y_buf[tail] = 0.0
            self._w_buf[tail] = 0.0
        self._last_length = t

    def solve_prefix_from_full(self, z_full: np.ndarray, y_full: np.ndarray, length: int) -> np.ndarray:
        """exact_ftl.py:171-181."""
        z_src = _ensure_float64_contiguous(z_full)
        y_src = _ensure_float64_contiguous(y_full)
        self._set_prefix(z_src, y_src, length)
        return self._solve_current()

    def solve_prefix(self, z_prefix: np.ndarray, y_prefix: np.ndarray) -> np.ndarray:
        """exact_ftl.py:183-193."""
        z_src = _ensure_float64_contiguous(z_prefix)
        y_src = _ensure_float64_contiguous(y_prefix)
        t, d = z_src.shape
        if d != self.d:
            raise ValueError(f"Expected {self.d}-dimensional data, got {d}")
        if t > self.T_max:
            raise ValueError("prefix longer than T_max")
        self._set_prefix(z_src, y_src, t)
        return self._solve_current()


@dataclass
class RunResult:
    """exact_ftl.py:217-222."""
    cum_loss: float
    regret: float
    comp_loss: float
    x_last: np.ndarray


def compute_prefix_actions(solver: ExactFTLNoClip, z: np.ndarray, y: np.ndarray) -> np.ndarray:
    """exact_ftl.py:280-303 -- exact FTL solutions for every prefix length 0..T, one launch
    (the reference performs T conic solves).  Leaves the solver's buffers as T append_row calls would."""
    z_arr = _ensure_float64_contiguous(z)
    y_arr = _ensure_float64_contiguous(y)
    T, d = z_arr.shape
    if solver.d != d:
        raise ValueError(f"Solver dimension {solver.d} incompatible with data dimension {d}")
    if solver.T_max < T:
        raise ValueError("Solver T_max is smaller than sequence length")
    solver.reset_buffers()
    solver._set_prefix(z_arr, y_arr, T)
    if T == 0:
        return np.zeros((1, d), dtype=np.float64)
    res = _solve_prefixes(z_arr, y_arr, norm=solver.norm, solver_opts=solver.solver_opts)
    return np.asarray(res["actions"][0], dtype=np.float64)


def replay_exact_ftl(z: np.ndarray, y: np.ndarray, actions: np.ndarray) -> RunResult:
    """exact_ftl.py:306-333 -- cumulative loss of playing actions[t] at round t, comparator actions[T]."""
    z_arr = _ensure_float64_contiguous(z)
    y_arr = _ensure_float64_contiguous(y)
    T, d = z_arr.shape
    if actions.shape != (T + 1, d):
        raise ValueError("actions must have shape (T+1, d)")
    res = engine.replay_actions(z_arr[None], y_arr[None], np.asarray(actions, dtype=np.float64)[None])
    return RunResult(cum_loss=float(res["cum_loss"][0]), regret=float(res["regret"][0]),
                     comp_loss=float(res["comp_loss"][0]), x_last=np.array(actions[T], dtype=np.float64, copy=True))


def simulate(z: np.ndarray, y: np.ndarray, *, algo: Literal["ftrl", "ftl_exact"] = "ftl_exact", eta0: float = 1.0,
             norm: Literal["l2", "linf", "l1"] = "l2", solver: Optional[str] = None, solver_opts: Optional[dict] = None,
             ftl_solver: Optional[ExactFTLNoClip] = None, comparator_solver: Optional[ExactFTLNoClip] = None,
             prefix_actions: Optional[np.ndarray] = None, comparator_action: Optional[np.ndarray] = None) -> RunResult:
    """exact_ftl.py:336-392."""
    z_arr = _ensure_float64_contiguous(z)
    y_arr = _ensure_float64_contiguous(y)
    T, d = z_arr.shape
    if algo == "ftl_exact":
        solver_obj = ftl_solver or ExactFTLNoClip(d=d, T_max=T, norm=norm, solver=solver, solver_opts=solver_opts)
        actions = prefix_actions if prefix_actions is not None else compute_prefix_actions(solver_obj, z_arr, y_arr)
        return replay_exact_ftl(z_arr, y_arr, actions)
    if algo == "ftrl":
        if comparator_action is None:
            solver_for_comp = comparator_solver or ExactFTLNoClip(d=d, T_max=T, norm=norm, solver=solver, solver_opts=solver_opts)
            comp_vec = solver_for_comp.solve_prefix_from_full(z_arr, y_arr, T)
        else:
            comp_vec = _ensure_float64_contiguous(np.asarray(comparator_action))
        res = engine.simulate_batch(z_arr, y_arr, [0], [float(eta0)], full=True, comparator_action=comp_vec[None])
        return RunResult(cum_loss=float(res["cum_loss"][0, 0]), regret=float(res["regret"][0, 0]),
                         comp_loss=float(res["comp_loss"][0, 0]), x_last=res["x_last"][0, 0].copy())
    raise ValueError("algo must be either 'ftrl' or 'ftl_exact'")


def run_ftrl(z: np.ndarray, y: np.ndarray, *, eta0: float = 1.0, norm: Literal["l2", "linf", "l1"] = "l2",
             solver: Optional[str] = None, solver_opts: Optional[dict] = None,
             comparator_solver: Optional[ExactFTLNoClip] = None,
             comparator_action: Optional[np.ndarray] = None) -> RunResult:
    """exact_ftl.py:399-420."""
    return simulate(z, y, algo="ftrl", eta0=eta0, norm=norm, solver=solver, solver_opts=solver_opts,
                    comparator_solver=comparator_solver, comparator_action=comparator_action)


def run_ftl_exact(z: np.ndarray, y: np.ndarray, *, norm: Literal["l2", "linf", "l1"] = "l2", solver: Optional[str] = None,
                  solver_opts: Optional[dict] = None, ftl_solver: Optional[ExactFTLNoClip] = None,
                  prefix_actions: Optional[np.ndarray] = None, return_actions: bool = False):
    """exact_ftl.py:423-453."""
    z_arr = _ensure_float64_contiguous(z)
    y_arr = _ensure_float64_contiguous(y)
    T, d = z_arr.shape
    solver_obj = ftl_solver or ExactFTLNoClip(d=d, T_max=T, norm=norm, solver=solver, solver_opts=solver_opts)
    actions = prefix_actions if prefix_actions is not None else compute_prefix_actions(solver_obj, z_arr, y_arr)
    result = replay_exact_ftl(z_arr, y_arr, actions)
    if return_actions:
        return result, actions
    return result


def _smoke() -> None:
    """The reference module's own smoke run (exact_ftl.py:460-475): T = 200, d = 10, rows clipped to the unit ball,
    labels sign(z.w + noise) -- exact FTL (l2) and the FTRL baseline, printed in the reference's format."""
    rng = np.random.default_rng(0)
    T, d = 200, 10
    z = rng.normal(size=(T, d))
    z /= np.maximum(1.0, np.linalg.norm(z, axis=1, keepdims=True))
    w_true = rng.normal(size=d)
    w_true /= np.linalg.norm(w_true)
    y = np.sign(z @ w_true + 0.2 * rng.normal(size=T))
    print("FTL exact (l2)...")
    res_ftl = run_ftl_exact(z, y, norm="l2", solver="ECOS")
    print(f"cum_loss={res_ftl.cum_loss:.4f}  regret={res_ftl.regret:.4f}  comp={res_ftl.comp_loss:.4f}")
    print("FTRL baseline...")
    res_ftrl = run_ftrl(z, y, eta0=1.0)
    print(f"cum_loss={res_ftrl.cum_loss:.4f}  regret={res_ftrl.regret:.4f}  comp={res_ftrl.comp_loss:.4f}")


if __name__ == "__main__":
    _smoke()
