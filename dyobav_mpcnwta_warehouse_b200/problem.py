"""Problem dimensions, the parameter-vector layout and the solver settings.

Mirrors, for the NMPC hot path only, what the reference fixes at build time:

* the ``z`` / ``p`` layout of ``MpcModule.build``
  (``src/pkg_mpc_tracker/solver_build/mpc_builder.py:47-60``), assembled at
  run time by ``TrajectoryTracker.run_step`` (``trajectory_tracker.py:315-317``);
* the keys ``MpcConfiguration`` / ``CircularRobotSpecification`` read from
  ``config/mpc_default.yaml`` / ``config/mpc_fast.yaml``
  (``src/configs.py:93-103,145-176``);
* the OpEn ``SolverConfiguration`` the builder leaves at opengen defaults except
  ``initial_penalty=10`` (``mpc_builder.py:187-195``).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field, fields
from typing import Dict, Tuple

NU = 2        # (v, w)
NS = 3        # (x, y, theta)
NQ = 10       # penalty weights block
NDYNPAR = 6   # (x, y, rx, ry, angle, alpha)
MAX_N = 64
MAX_EDGE = 8
MAX_LBFGS = 10

EXIT_STATUS_NAMES = (
    "Converged",
    "NotConvergedIterations",
    "NotConvergedOutOfTime",
    "NotFiniteComputation",
)


class CDims(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("Nother", ctypes.c_int32), ("Nstc", ctypes.c_int32),
                ("nedge", ctypes.c_int32), ("Ndyn", ctypes.c_int32)]


class CRobot(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "ts", "vehicle_width", "vehicle_margin", "social_margin", "lin_vel_min",
        "lin_vel_max", "ang_vel_max", "lin_acc_min", "lin_acc_max", "ang_acc_max")]


class CSolverCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "tolerance", "initial_tolerance", "delta_tolerance", "inner_tol_update",
        "penalty_update", "sufficient_decrease", "initial_penalty", "sy_epsilon",
        "cbfgs_epsilon", "cbfgs_alpha")] + [
        ("max_inner", ctypes.c_int32), ("max_outer", ctypes.c_int32),
        ("lbfgs_mem", ctypes.c_int32), ("max_inner_total", ctypes.c_int32),
        ("team_mode", ctypes.c_int32), ("max_time_us", ctypes.c_int32)]


@dataclass(frozen=True)
class Dims:
    """Run-time problem dimensions (yaml keys ``N_hor, Nother, Nstcobs, nstcobs, Ndynobs``)."""
    N: int = 20
    Nother: int = 10
    Nstc: int = 10
    nedge: int = 4
    Ndyn: int = 15

    def __post_init__(self):
        if not (1 <= self.N <= MAX_N):
            raise ValueError(f"N_hor must be in 1..{MAX_N}, got {self.N}")
        if not (1 <= self.nedge <= MAX_EDGE):
            raise ValueError(f"nstcobs/3 must be in 1..{MAX_EDGE}, got {self.nedge}")
        if min(self.Nother, self.Nstc, self.Ndyn) < 0:
            raise ValueError("negative obstacle/robot count")

    @property
    def nu_total(self) -> int:
        return NU * self.N

    @property
    def n1(self) -> int:       # ALM constraints F1 (accelerations), mpc_builder.py:160
        return 2 * self.N

    @property
    def n2(self) -> int:       # penalty constraints F2 (broadcast to Ndyn), :72,119
        return self.Ndyn if self.Ndyn > 0 else 1

    def layout(self) -> Dict[str, Tuple[int, int]]:
        """``name -> (offset, length)`` of every block of ``p`` (mpc_builder.py:47-60)."""
        N = self.N
        blocks = [("u_m1", NU), ("s_0", NS), ("s_N", NS), ("q", NQ), ("r_s", NS * N),
                  ("r_v", N), ("c_0", NS * self.Nother), ("c", NS * N * self.Nother),
                  ("o_s", 3 * self.nedge * self.Nstc),
                  ("o_d", NDYNPAR * (N + 1) * self.Ndyn), ("q_stc", N), ("q_dyn", N)]
        out, off = {}, 0
        for name, ln in blocks:
            out[name] = (off, ln)
            off += ln
        return out

    @property
    def np(self) -> int:
        off, ln = self.layout()["q_dyn"]
        return off + ln

    def to_c(self) -> CDims:
        return CDims(self.N, self.Nother, self.Nstc, self.nedge, self.Ndyn)


@dataclass(frozen=True)
class RobotSpec:
    """``CircularRobotSpecification`` (configs.py:93-103); defaults = mpc_fast.yaml:6-18."""
    ts: float = 0.2
    vehicle_width: float = 0.5
    vehicle_margin: float = 0.2
    social_margin: float = 0.2
    lin_vel_min: float = -0.5
    lin_vel_max: float = 1.5
    ang_vel_max: float = 0.5
    lin_acc_min: float = -1.0
    lin_acc_max: float = 1.0
    ang_acc_max: float = 3.0

    def to_c(self) -> CRobot:
        return CRobot(*[float(getattr(self, f.name)) for f in fields(self)])


@dataclass(frozen=True)
class SolverSettings:
    """OpEn solver settings as the reference builds them (mpc_builder.py:187-195)."""
    tolerance: float = 1e-4
    initial_tolerance: float = 1e-4
    delta_tolerance: float = 1e-4
    inner_tol_update: float = 0.1
    penalty_update: float = 5.0
    sufficient_decrease: float = 0.1
    initial_penalty: float = 10.0
    sy_epsilon: float = 1e-10
    cbfgs_epsilon: float = 1e-8
    cbfgs_alpha: float = 1.0
    max_inner: int = 500
    max_outer: int = 10
    lbfgs_mem: int = 10
    # 0 = off.  Budget on the inner iterations of one solve, the batch analogue of the reference's
    # wall-clock cap ``max_solver_time`` (mpc_fast.yaml:45): exhausted -> "NotConvergedOutOfTime"
    max_inner_total: int = 0
    # 0 = kernels chosen by the dimensions (team kernels from 64 ellipses on); 1 = team kernels for any
    # dimensions (one instance gets a solver warp plus the worker pool of a CTA instead of one warp).
    # Selects the team arithmetic contract, so results differ from mode 0 by round-off; both are mirrored
    # by the laned oracle.  Meant as a latency mode for single solves; measured on B200 it only pays when
    # the per-step part of an evaluation is large (for the reference's dims it is 1.4x SLOWER).
    team_mode: int = 0
    # 0 = off.  Wall-clock cap on one solve in microseconds: the reference's ``max_solver_time``
    # (mpc_builder.py:189) itself.  Honoured for small batches (the latency kernel: the single solve per
    # timestep); exhausted -> "NotConvergedOutOfTime".  Results then depend on timing, as the reference's do.
    max_time_us: int = 0

    def __post_init__(self):
        if not (1 <= self.lbfgs_mem <= MAX_LBFGS):
            raise ValueError(f"lbfgs_mem must be in 1..{MAX_LBFGS}")

    def to_c(self) -> CSolverCfg:
        return CSolverCfg(self.tolerance, self.initial_tolerance, self.delta_tolerance,
                          self.inner_tol_update, self.penalty_update,
                          self.sufficient_decrease, self.initial_penalty, self.sy_epsilon,
                          self.cbfgs_epsilon, self.cbfgs_alpha, self.max_inner,
                          self.max_outer, self.lbfgs_mem, int(self.max_inner_total), int(self.team_mode),
                          int(self.max_time_us))


# The yaml keys of config/mpc_fast.yaml / mpc_default.yaml with their shipped
# values (the two files differ only in max_solver_time and optimizer_name).
MPC_YAML_DEFAULTS = {
    "ts": 0.2,
    "vehicle_width": 0.5, "vehicle_margin": 0.2, "social_margin": 0.2,
    "lin_vel_min": -0.5, "lin_vel_max": 1.5, "lin_acc_min": -1, "lin_acc_max": 1,
    "ang_vel_max": 0.5, "ang_acc_max": 3,
    "N_hor": 20, "action_steps": 1,
    "nu": 2, "ns": 3, "nq": 10, "Nother": 10, "Nstcobs": 10, "nstcobs": 12,
    "Ndynobs": 15, "ndynobs": 6,
    "lin_vel_penalty": 0, "lin_acc_penalty": 10.0, "ang_vel_penalty": 0,
    "ang_acc_penalty": 20.0, "qrpd": 100.0, "qpos": 0.0, "qvel": 10.0, "qtheta": 0.0,
    "qpN": 0.0, "qthetaN": 0.0,
    "max_solver_time": 100_000, "build_type": "release", "build_directory": "mpc_solver",
    "bad_exit_codes": ["NotConvergedIterations", "NotConvergedOutOfTime"],
    "optimizer_name": "navi_fast",
}


@dataclass
class MpcConfig:
    """The attribute bag ``MpcConfiguration`` exposes (configs.py:145-176), plus views."""
    values: dict = field(default_factory=lambda: dict(MPC_YAML_DEFAULTS))

    @classmethod
    def from_yaml(cls, path: str) -> "MpcConfig":
        import yaml
        with open(path, "r") as fh:
            loaded = yaml.safe_load(fh)
        missing = [k for k in MPC_YAML_DEFAULTS if k not in loaded]
        if missing:
            raise KeyError(f"{path}: missing MPC config keys {missing}")
        return cls(dict(loaded))

    def __getattr__(self, name):
        try:
            return self.__dict__["values"][name]
        except KeyError as e:
            raise AttributeError(name) from e

    def dims(self) -> Dims:
        v = self.values
        if v["nu"] != NU or v["ns"] != NS or v["nq"] != NQ or v["ndynobs"] != NDYNPAR:
            raise ValueError("only the unicycle problem (nu=2, ns=3, nq=10, ndynobs=6) is supported")
        if v["nstcobs"] % 3:
            raise ValueError("nstcobs must be 3 * (number of edges)")
        return Dims(N=v["N_hor"], Nother=v["Nother"], Nstc=v["Nstcobs"],
                    nedge=v["nstcobs"] // 3, Ndyn=v["Ndynobs"])

    def robot(self) -> RobotSpec:
        v = self.values
        return RobotSpec(**{f.name: float(v[f.name]) for f in fields(RobotSpec)})

    def tuning_params(self):
        """``TrajectoryTracker.set_work_mode`` weights for any non-'aligning' mode (:138-139)."""
        v = self.values
        return [v["qpos"], v["qvel"], v["qtheta"], v["lin_vel_penalty"], v["ang_vel_penalty"],
                v["qpN"], v["qthetaN"], v["qrpd"], v["lin_acc_penalty"], v["ang_acc_penalty"]]
