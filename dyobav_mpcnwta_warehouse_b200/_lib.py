"""ctypes binding of libmpcb.so (include/mpcb.h).  No CPU fallback: a missing
library or a missing GPU raises."""
from __future__ import annotations

import ctypes
import os

from .problem import CDims, CRobot, CSolverCfg

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPCB_LIB_PATH") or os.path.join(_HERE, "csrc", "libmpcb.so")
_lib = None

ERRORS = {-1: "unsupported or inconsistent dimensions", -2: "required pointer is NULL",
          -3: "workspace too small", -4: "CUDA error", -5: "misaligned pointer",
          -6: "no CUDA device"}

# every symbol include/mpcb.h declares
EXPORTS = ("mpcb_abi_version", "mpcb_last_error", "mpcb_param_len", "mpcb_num_decision", "mpcb_n1",
           "mpcb_n2", "mpcb_team_groups", "mpcb_team_groups_cfg", "mpcb_default_robot", "mpcb_default_solver_cfg", "mpcb_workspace_bytes",
           "mpcb_eval_f64", "mpcb_solve_f64", "mpcb_workspace_bytes_f32", "mpcb_eval_f32", "mpcb_solve_f32",
           "mpcb_solve_one_host", "mpcb_pack_f64",
           "mpcb_plant_step_f64", "mpcb_sincos_host", "mpcb_cluster_f64", "mpcb_fp64_peak_tflops")


class CSim(ctypes.Structure):
    """mpcb_sim (include/mpcb.h): device pointers as integers."""
    _fields_ = ([(n, ctypes.c_int32) for n in ("n", "T", "Kp", "Pd", "M")]
                + [(n, ctypes.c_double) for n in ("base_speed", "lin_vel_max", "ped_size", "stc_w", "dyn_w", "ts")]
                + [("tuning", ctypes.c_double * 10)]
                + [(n, ctypes.c_void_p) for n in ("state", "last_u", "ref_traj", "ref_len", "idx_ref", "goal",
                                                   "polys", "n_poly", "ped_pos", "ped_vel", "done", "od_in")])


def load():
    """Load the CUDA library; raise RuntimeError (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m dyobav_mpcnwta_warehouse_b200.csrc.build`"
            " (or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, dp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p
    pd, pr, pc = ctypes.POINTER(CDims), ctypes.POINTER(CRobot), ctypes.POINTER(CSolverCfg)
    L.mpcb_abi_version.restype = i32
    L.mpcb_last_error.restype = ctypes.c_char_p
    for name in ("mpcb_param_len", "mpcb_num_decision", "mpcb_n1", "mpcb_n2", "mpcb_team_groups"):
        getattr(L, name).restype = i32
        getattr(L, name).argtypes = [pd]
    L.mpcb_team_groups_cfg.restype = i32
    L.mpcb_team_groups_cfg.argtypes = [pd, pc]
    L.mpcb_default_robot.argtypes = [pr]
    L.mpcb_default_solver_cfg.argtypes = [pc]
    L.mpcb_workspace_bytes.restype = i32
    L.mpcb_workspace_bytes.argtypes = [pd, i32, i32, ctypes.POINTER(ctypes.c_size_t)]
    L.mpcb_eval_f64.restype = i32
    L.mpcb_eval_f64.argtypes = [pd, pr, pc, i32, i32] + [dp] * 9 + [vp, ctypes.c_size_t, vp]
    L.mpcb_solve_f64.restype = i32
    L.mpcb_solve_f64.argtypes = [pd, pr, pc, i32, i32] + [dp] * 15 + [vp, ctypes.c_size_t, vp]
    L.mpcb_workspace_bytes_f32.restype = i32
    L.mpcb_workspace_bytes_f32.argtypes = [pd, i32, i32, ctypes.POINTER(ctypes.c_size_t)]
    L.mpcb_eval_f32.restype = i32
    L.mpcb_eval_f32.argtypes = [pd, pr, pc, i32, i32] + [dp] * 9 + [vp, ctypes.c_size_t, vp]
    L.mpcb_solve_f32.restype = i32
    L.mpcb_solve_f32.argtypes = [pd, pr, pc, i32, i32] + [dp] * 15 + [vp, ctypes.c_size_t, vp]
    L.mpcb_solve_one_host.restype = i32
    L.mpcb_solve_one_host.argtypes = [pd, pr, pc] + [dp] * 8
    L.mpcb_pack_f64.restype = i32
    L.mpcb_pack_f64.argtypes = [pd, ctypes.POINTER(CSim), dp, vp]
    L.mpcb_plant_step_f64.restype = i32
    L.mpcb_plant_step_f64.argtypes = [pd, ctypes.POINTER(CSim), dp, vp]
    L.mpcb_cluster_f64.restype = i32
    L.mpcb_cluster_f64.argtypes = [pd, i32, i32, i32, ctypes.c_double, i32, ctypes.c_double, ctypes.c_double,
                                   dp, dp, dp, dp, dp, vp]
    L.mpcb_fp64_peak_tflops.restype = i32
    L.mpcb_fp64_peak_tflops.argtypes = [ctypes.POINTER(ctypes.c_double)]
    L.mpcb_sincos_host.restype = None
    L.mpcb_sincos_host.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        detail = ""
        if rc == -4 and _lib is not None:
            detail = ": " + _lib.mpcb_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({rc}: {ERRORS.get(rc, 'unknown error')}{detail})")


def sincos_host(x: float):
    """The kernels' portable sin/cos evaluated on the host (bit-identical to the device)."""
    sn, cs = ctypes.c_double(), ctypes.c_double()
    load().mpcb_sincos_host(float(x), ctypes.byref(sn), ctypes.byref(cs))
    return sn.value, cs.value


def cluster_hypotheses(dims, hyp, cur_pos, n_hyp=None, eps=1.0, min_samples=2, enlarge=2.0, human_size=0.2):
    """Device-side SWTA hypotheses -> ``o_d`` [n, Ndyn, N+1, 6] (torch CUDA tensors in and out)."""
    import torch
    L = load()
    n, N, K, _ = hyp.shape
    H = cur_pos.shape[1]
    if N != dims.N or not hyp.is_cuda or hyp.dtype != torch.float64 or not hyp.is_contiguous():
        raise RuntimeError("hyp must be a contiguous float64 CUDA tensor [n, N, K, 2]")
    od = torch.empty(n, dims.Ndyn, N + 1, 6, dtype=torch.float64, device=hyp.device)
    scratch = torch.empty(n, N + 1, dtype=torch.int32, device=hyp.device)
    cd = dims.to_c()
    ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream(hyp.device).cuda_stream)
    check(L.mpcb_cluster_f64(ctypes.byref(cd), n, K, H, eps, min_samples, enlarge, human_size, ptr(hyp),
                             ptr(n_hyp), ptr(cur_pos.contiguous()), ptr(od), ptr(scratch), st), "mpcb_cluster_f64")
    return od


def fp64_peak_tflops() -> float:
    """Measured FP64 FMA throughput of the current CUDA device in TFLOP/s."""
    v = ctypes.c_double()
    check(load().mpcb_fp64_peak_tflops(ctypes.byref(v)), "mpcb_fp64_peak_tflops")
    return v.value
