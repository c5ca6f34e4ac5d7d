"""Host-side mirror of the generated solver module's interface.

The reference loads ``mpc_solver/<optimizer_name>/<optimizer_name>.so`` and calls
``built_solver.solver().run(p)`` once per timestep
(``trajectory_tracker.py:56-62,362``), reading ``.solution / .cost / .exit_status /
.solve_time_ms`` (``:364-367``).  ``Solver`` keeps that call; ``BatchedSolver`` is
the batched form on torch CUDA tensors.  Both go through the C-ABI in
``csrc/libmpcb.so`` — there is no CPU path.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List

import numpy as np

from . import _lib
from .problem import Dims, RobotSpec, SolverSettings, EXIT_STATUS_NAMES


@dataclass
class SolverStatus:
    """What ``Solver.run`` returns — the attribute set of OpEn's ``OptimizerSolution``."""
    exit_status: str
    num_outer_iterations: int
    num_inner_iterations: int
    last_problem_norm_fpr: float
    f1_infeasibility: float
    f2_norm: float
    solve_time_ms: float
    penalty: float
    solution: List[float]
    lagrange_multipliers: List[float]
    cost: float


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class BatchedSolver:
    """Solve ``n_p * starts`` independent instances per call on one GPU.

    ``P`` is ``[n_p, np]`` float64 CUDA (reference layout, one row per scenario);
    instance ``b`` uses row ``b // starts`` and the initial guess ``U0[b]``.
    """

    def __init__(self, dims: Dims = Dims(), robot: RobotSpec = RobotSpec(),
                 settings: SolverSettings = SolverSettings(), device=None):
        import torch
        self._torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedSolver needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.dims, self.robot, self.settings = dims, robot, settings
        self._cd, self._cr, self._cc = dims.to_c(), robot.to_c(), settings.to_c()
        self._ws = None

    # -- helpers
    def _workspace(self, n_p: int, starts: int):
        torch = self._torch
        need = ctypes.c_size_t()
        _lib.check(self.lib.mpcb_workspace_bytes(ctypes.byref(self._cd), n_p, starts,
                                                 ctypes.byref(need)), "mpcb_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != self.device:
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def _chk(self, t, shape, name, dtype=None):
        torch = self._torch
        dtype = dtype or torch.float64
        if t is None:
            return None
        if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous {dtype} CUDA tensor")
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t

    def evaluate(self, P, U, Y=None, C=None, starts: int = 1):
        """psi, grad psi, f, F1, F2 at ``U`` (parity-test entry: ``mpcb_eval_f64``)."""
        torch = self._torch
        d = self.dims
        n_p = P.shape[0]
        B = n_p * starts
        self._chk(P, (n_p, d.np), "P")
        self._chk(U, (B, d.nu_total), "U")
        self._chk(Y, (B, d.n1), "Y")
        self._chk(C, (B,), "C")
        with torch.cuda.device(self.device):
            ws = self._workspace(n_p, starts)
            kw = dict(dtype=torch.float64, device=self.device)
            out = dict(f=torch.empty(B, **kw), psi=torch.empty(B, **kw),
                       grad=torch.empty(B, d.nu_total, **kw), F1=torch.empty(B, d.n1, **kw),
                       F2=torch.empty(B, d.n2, **kw))
            st = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.mpcb_eval_f64(ctypes.byref(self._cd), ctypes.byref(self._cr),
                                        ctypes.byref(self._cc), n_p, starts, _ptr(P), _ptr(U),
                                        _ptr(Y), _ptr(C), _ptr(out["f"]), _ptr(out["psi"]),
                                        _ptr(out["grad"]), _ptr(out["F1"]), _ptr(out["F2"]),
                                        _ptr(ws), ws.numel(), ctypes.c_void_p(st))
        _lib.check(rc, "mpcb_eval_f64")
        return out

    def alloc_outputs(self, B: int):
        torch = self._torch
        d = self.dims
        f64 = dict(device=self.device, dtype=torch.float64)
        i32 = dict(device=self.device, dtype=torch.int32)
        return dict(
            u=torch.empty(B, d.nu_total, **f64), cost=torch.empty(B, **f64),
            exit_status=torch.empty(B, **i32), n_outer=torch.empty(B, **i32),
            n_inner=torch.empty(B, **i32), fpr=torch.empty(B, **f64),
            f1_infeas=torch.empty(B, **f64), f2_norm=torch.empty(B, **f64),
            penalty=torch.empty(B, **f64), y=torch.empty(B, d.n1, **f64),
            evals=torch.empty(B, 4, **i32))

    def run_batch(self, P, U0=None, Y0=None, C0=None, starts: int = 1, out=None):
        """Solve; returns a dict of CUDA tensors (asynchronous on the current stream)."""
        torch = self._torch
        d = self.dims
        n_p = P.shape[0]
        B = n_p * starts
        self._chk(P, (n_p, d.np), "P")
        self._chk(U0, (B, d.nu_total), "U0")
        self._chk(Y0, (B, d.n1), "Y0")
        self._chk(C0, (B,), "C0")
        with torch.cuda.device(self.device):
            ws = self._workspace(n_p, starts)
            o = out if out is not None else self.alloc_outputs(B)
            st = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.mpcb_solve_f64(
                ctypes.byref(self._cd), ctypes.byref(self._cr), ctypes.byref(self._cc), n_p, starts,
                _ptr(P), _ptr(U0), _ptr(Y0), _ptr(C0), _ptr(o["u"]), _ptr(o["cost"]),
                _ptr(o["exit_status"]), _ptr(o["n_outer"]), _ptr(o["n_inner"]), _ptr(o["fpr"]),
                _ptr(o["f1_infeas"]), _ptr(o["f2_norm"]), _ptr(o["penalty"]), _ptr(o["y"]),
                _ptr(o["evals"]), _ptr(ws), ws.numel(), ctypes.c_void_p(st))
        _lib.check(rc, "mpcb_solve_f64")
        return o


    # -- f32 boundary mode (include/mpcb.h "f32 twins"): float32 tensors in and out, f64 arithmetic
    def _workspace_f32(self, n_p: int, starts: int, extra: int = 0):
        torch = self._torch
        need = ctypes.c_size_t()
        _lib.check(self.lib.mpcb_workspace_bytes_f32(ctypes.byref(self._cd), n_p, starts, ctypes.byref(need)),
                   "mpcb_workspace_bytes_f32")
        total = need.value + extra + 256
        if getattr(self, "_ws32", None) is None or self._ws32.numel() < total or self._ws32.device != self.device:
            self._ws32 = torch.empty(total, dtype=torch.uint8, device=self.device)
        off = (-self._ws32.data_ptr()) % 256              # the entry points want 256-byte alignment
        return self._ws32[off:off + need.value + extra]

    def run_batch_f32(self, P, U0=None, Y0=None, C0=None, starts: int = 1):
        """``run_batch`` with float32 CUDA tensors at the boundary (``mpcb_solve_f32``): the results
        equal the f64 solve of the widened inputs, rounded to float32."""
        torch = self._torch
        d = self.dims
        n_p = P.shape[0]
        B = n_p * starts
        f32 = torch.float32
        self._chk(P, (n_p, d.np), "P", f32)
        self._chk(U0, (B, d.nu_total), "U0", f32)
        self._chk(Y0, (B, d.n1), "Y0", f32)
        self._chk(C0, (B,), "C0", f32)
        with torch.cuda.device(self.device):
            ws = self._workspace_f32(n_p, starts)
            kf = dict(device=self.device, dtype=f32)
            ki = dict(device=self.device, dtype=torch.int32)
            o = dict(u=torch.empty(B, d.nu_total, **kf), cost=torch.empty(B, **kf),
                     exit_status=torch.empty(B, **ki), n_outer=torch.empty(B, **ki), n_inner=torch.empty(B, **ki),
                     fpr=torch.empty(B, **kf), f1_infeas=torch.empty(B, **kf), f2_norm=torch.empty(B, **kf),
                     penalty=torch.empty(B, **kf), y=torch.empty(B, d.n1, **kf), evals=torch.empty(B, 4, **ki))
            st = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.mpcb_solve_f32(
                ctypes.byref(self._cd), ctypes.byref(self._cr), ctypes.byref(self._cc), n_p, starts,
                _ptr(P), _ptr(U0), _ptr(Y0), _ptr(C0), _ptr(o["u"]), _ptr(o["cost"]), _ptr(o["exit_status"]),
                _ptr(o["n_outer"]), _ptr(o["n_inner"]), _ptr(o["fpr"]), _ptr(o["f1_infeas"]), _ptr(o["f2_norm"]),
                _ptr(o["penalty"]), _ptr(o["y"]), _ptr(o["evals"]), _ptr(ws), ws.numel(), ctypes.c_void_p(st))
        _lib.check(rc, "mpcb_solve_f32")
        return o

    def evaluate_f32(self, P, U, Y=None, C=None, starts: int = 1):
        """``evaluate`` with float32 CUDA tensors at the boundary (``mpcb_eval_f32``)."""
        torch = self._torch
        d = self.dims
        n_p = P.shape[0]
        B = n_p * starts
        f32 = torch.float32
        self._chk(P, (n_p, d.np), "P", f32)
        self._chk(U, (B, d.nu_total), "U", f32)
        self._chk(Y, (B, d.n1), "Y", f32)
        self._chk(C, (B,), "C", f32)
        with torch.cuda.device(self.device):
            ws = self._workspace_f32(n_p, starts, extra=((B * d.n2 * 8 + 255) // 256) * 256)
            kf = dict(device=self.device, dtype=f32)
            out = dict(f=torch.empty(B, **kf), psi=torch.empty(B, **kf), grad=torch.empty(B, d.nu_total, **kf),
                       F1=torch.empty(B, d.n1, **kf), F2=torch.empty(B, d.n2, **kf))
            st = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.mpcb_eval_f32(ctypes.byref(self._cd), ctypes.byref(self._cr), ctypes.byref(self._cc),
                                        n_p, starts, _ptr(P), _ptr(U), _ptr(Y), _ptr(C), _ptr(out["f"]),
                                        _ptr(out["psi"]), _ptr(out["grad"]), _ptr(out["F1"]), _ptr(out["F2"]),
                                        _ptr(ws), ws.numel(), ctypes.c_void_p(st))
        _lib.check(rc, "mpcb_eval_f32")
        return out


class Solver:
    """Drop-in for the object ``<optimizer_name>.solver()`` returns (one solve per call).

    ``run(p, initial_guess=None, initial_lagrange_multipliers=None, initial_penalty=None)``
    takes Python lists exactly like the PyO3 binding; bad dimensions raise
    ``RuntimeError`` (the caller's ``except RuntimeError`` at trajectory_tracker.py:321).
    """

    def __init__(self, dims: Dims = Dims(), robot: RobotSpec = RobotSpec(),
                 settings: SolverSettings = SolverSettings()):
        self.lib = _lib.load()
        self.dims, self.robot, self.settings = dims, robot, settings
        self._cd, self._cr, self._cc = dims.to_c(), robot.to_c(), settings.to_c()

    def run(self, p, initial_guess=None, initial_lagrange_multipliers=None, initial_penalty=None):
        d = self.dims
        pa = np.ascontiguousarray(p, dtype=np.float64)
        if pa.shape != (d.np,):
            raise RuntimeError(f"3003 -> wrong number of parameters: got {pa.size}, expected {d.np}")
        u0 = y0 = c0 = None
        if initial_guess is not None:
            u0 = np.ascontiguousarray(initial_guess, dtype=np.float64)
            if u0.shape != (d.nu_total,):
                raise RuntimeError("1600 -> Initial guess has incompatible dimensions")
        if initial_lagrange_multipliers is not None:
            y0 = np.ascontiguousarray(initial_lagrange_multipliers, dtype=np.float64)
            if y0.shape != (d.n1,):
                raise RuntimeError("1700 -> wrong dimension of Langrange multipliers")
        if initial_penalty is not None:
            c0 = np.array([float(initial_penalty)], dtype=np.float64)
        u = np.empty(d.nu_total)
        y = np.empty(d.n1)
        status = np.full(1, -1, dtype=np.int32)
        sc = np.zeros(8)

        def vp(a):
            return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

        rc = self.lib.mpcb_solve_one_host(ctypes.byref(self._cd), ctypes.byref(self._cr),
                                          ctypes.byref(self._cc), vp(pa), vp(u0), vp(y0), vp(c0),
                                          vp(u), vp(y), vp(status), vp(sc))
        _lib.check(rc, "mpcb_solve_one_host")
        return SolverStatus(
            exit_status=EXIT_STATUS_NAMES[int(status[0])], num_outer_iterations=int(sc[5]),
            num_inner_iterations=int(sc[6]), last_problem_norm_fpr=float(sc[1]),
            f1_infeasibility=float(sc[2]), f2_norm=float(sc[3]), solve_time_ms=float(sc[7]),
            penalty=float(sc[4]), solution=u.tolist(), lagrange_multipliers=y.tolist(),
            cost=float(sc[0]))


def solver(dims: Dims = Dims(), robot: RobotSpec = RobotSpec(),
           settings: SolverSettings = SolverSettings()) -> Solver:
    """Same spelling as the generated module's factory (``built_solver.solver()``)."""
    return Solver(dims, robot, settings)
