"""ctypes wrapper of the CPU oracle (oracle/mpc_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs; never by the product.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmpc_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mpc_oracle.c", "mpc_oracle_laned.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.mpco_param_len.restype = ctypes.c_int32
        L.mpco_eval.restype = ctypes.c_int32
        L.mpco_solve.restype = ctypes.c_int32
        L.mpco_solve_batch.restype = ctypes.c_int32
        L.mpcl_eval.restype = ctypes.c_int32
        L.mpcl_solve.restype = ctypes.c_int32
        L.mpcl_solve_batch.restype = ctypes.c_int32
        L.mpcl_team_groups.restype = ctypes.c_int32
        L.mpcl_team_groups_cfg.restype = ctypes.c_int32
        for f in (L.mpco_dist_to_lineseg, L.mpco_inside_ellipse, L.mpco_inside_cvx_polygon):
            f.restype = ctypes.c_double
        L.mpco_dist_to_lineseg.argtypes = [ctypes.c_double] * 6
        L.mpco_inside_ellipse.argtypes = [ctypes.c_double] * 7
        L.mpco_inside_cvx_polygon.argtypes = [ctypes.c_double, ctypes.c_double, _dp, _dp, _dp,
                                              ctypes.c_int]
        L.mpco_unicycle_rk4.argtypes = [_dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp]
        L.mpco_unicycle_rk4.restype = None
        _LIB = L
    return _LIB


def set_variant(akkt: int = 0, ls: int = 0, last: int = 0):
    """Sensitivity switches of the reference-order oracle (mpc_oracle.c "sensitivity switches");
    (0, 0, 0) is the restatement everything else is checked against."""
    lib().mpco_set_variant(ctypes.c_int(akkt), ctypes.c_int(ls), ctypes.c_int(last))


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def evaluate(dims, robot, p, u, y=None, c=10.0, want_grad=True, laned=False, team=False):
    """f, psi, grad psi, F1, F2 for one instance (laned=True: the kernel's operation order;
    team=True: the team contract forced, as cfg.team_mode = 1 does for the kernels)."""
    L = lib()
    L.mpcl_set_team_mode(ctypes.c_int(1 if team else 0))
    fn = L.mpcl_eval if laned else L.mpco_eval
    cd, cr = dims.to_c(), robot.to_c()
    p, u, y = _c(p), _c(u), _c(y)
    assert p.shape == (dims.np,) and u.shape == (dims.nu_total,)
    f = ctypes.c_double()
    psi = ctypes.c_double()
    g = np.zeros(dims.nu_total) if want_grad else None
    F1 = np.zeros(dims.n1)
    F2 = np.zeros(dims.n2)
    rc = fn(ctypes.byref(cd), ctypes.byref(cr), _p(p), _p(u), _p(y),
                     ctypes.c_double(c), ctypes.byref(f), ctypes.byref(psi), _p(g), _p(F1), _p(F2))
    L.mpcl_set_team_mode(ctypes.c_int(0))
    if rc:
        raise RuntimeError(f"mpco_eval failed: {rc}")
    return dict(f=f.value, psi=psi.value, grad=g, F1=F1, F2=F2)


SCALARS = ("cost", "fpr", "f1_infeas", "f2_norm", "penalty", "n_outer", "n_inner",
           "n_cost", "n_grad", "exit_status", "n_small")


def solve(dims, robot, cfg, p, u0=None, y0=None, c0=None, laned=False):
    L = lib()
    fn = L.mpcl_solve if laned else L.mpco_solve
    cd, cr, cc = dims.to_c(), robot.to_c(), cfg.to_c()
    p, u0, y0 = _c(p), _c(u0), _c(y0)
    u = np.zeros(dims.nu_total)
    y = np.zeros(dims.n1)
    sc = np.zeros(11)
    c0p = None if c0 is None else ctypes.byref(ctypes.c_double(c0))
    rc = fn(ctypes.byref(cd), ctypes.byref(cr), ctypes.byref(cc), _p(p), _p(u0),
                      _p(y0), c0p, _p(u), _p(y), _p(sc))
    if rc:
        raise RuntimeError(f"mpco_solve failed: {rc}")
    out = dict(zip(SCALARS, sc.tolist()))
    for k in ("n_outer", "n_inner", "n_cost", "n_grad", "exit_status", "n_small"):
        out[k] = int(out[k])
    out["u"] = u
    out["y"] = y
    return out


def solve_batch(dims, robot, cfg, P, U0=None, starts=1, threads=1, laned=False):
    """OpenMP batch driver; P [n_p, np], U0 [n_p*starts, 2N] or None."""
    L = lib()
    cd, cr, cc = dims.to_c(), robot.to_c(), cfg.to_c()
    P, U0 = _c(P), _c(U0)
    n_p = P.shape[0]
    B = n_p * starts
    U = np.zeros((B, dims.nu_total))
    SC = np.zeros((B, 11))
    if laned:
        rc = L.mpcl_solve_batch(ctypes.byref(cd), ctypes.byref(cr), ctypes.byref(cc),
                                ctypes.c_int32(n_p), ctypes.c_int32(starts), _p(P), _p(U0),
                                _p(U), _p(SC), ctypes.c_int32(dims.np), ctypes.c_int32(threads))
    else:
        rc = L.mpco_solve_batch(ctypes.byref(cd), ctypes.byref(cr), ctypes.byref(cc),
                                ctypes.c_int32(n_p), ctypes.c_int32(starts), _p(P), _p(U0),
                                _p(U), _p(SC), ctypes.c_int32(threads))
    if rc:
        raise RuntimeError(f"mpco_solve_batch failed: {rc}")
    return U, SC
