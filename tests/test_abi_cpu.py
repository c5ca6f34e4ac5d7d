"""CPU: the C-ABI library builds (cross-compiled), loads, exports every symbol include/mpcb.h
declares, and its host-side argument checks work.  No kernel is launched here."""
import ctypes
import os
import re

import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, _lib
from dyobav_mpcnwta_warehouse_b200.problem import CRobot, CSolverCfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(cuda_solver_lib):
    hdr = open(os.path.join(ROOT, "include", "mpcb.h")).read()
    declared = set(re.findall(r"\b(mpcb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(cuda_solver_lib, name), name
    assert cuda_solver_lib.mpcb_abi_version() == 4


def test_param_len_matches_reference_layout(cuda_solver_lib):
    # 2778 at the yaml defaults (mpc_builder.py:47-60 with mpc_fast.yaml:21-31)
    assert cuda_solver_lib.mpcb_param_len(ctypes.byref(Dims().to_c())) == 2778 == Dims().np
    for d in (Dims(Ndyn=40), Dims(N=40, Ndyn=160), Dims(N=5, Nother=2, Nstc=2, Ndyn=3)):
        cd = d.to_c()
        assert cuda_solver_lib.mpcb_param_len(ctypes.byref(cd)) == d.np
        assert cuda_solver_lib.mpcb_num_decision(ctypes.byref(cd)) == d.nu_total
        assert cuda_solver_lib.mpcb_n1(ctypes.byref(cd)) == d.n1
        assert cuda_solver_lib.mpcb_n2(ctypes.byref(cd)) == d.n2
    assert Dims(Ndyn=40).np == 5928 and Dims(N=40, Ndyn=160).np == 40968


def test_defaults_match_python_side(cuda_solver_lib):
    r, c = CRobot(), CSolverCfg()
    cuda_solver_lib.mpcb_default_robot(ctypes.byref(r))
    cuda_solver_lib.mpcb_default_solver_cfg(ctypes.byref(c))
    pr, pc = RobotSpec().to_c(), SolverSettings().to_c()
    for name, _ in CRobot._fields_:
        assert getattr(r, name) == getattr(pr, name), name
    for name, _ in CSolverCfg._fields_:
        assert getattr(c, name) == getattr(pc, name), name


def test_argument_errors_are_codes_not_crashes(cuda_solver_lib):
    L = cuda_solver_lib
    need = ctypes.c_size_t()
    assert L.mpcb_workspace_bytes(ctypes.byref(Dims().to_c()), 4, 2, ctypes.byref(need)) == 0
    assert need.value > 4 * 2778 * 8
    bad = Dims().to_c()
    bad.N = 0
    assert L.mpcb_workspace_bytes(ctypes.byref(bad), 4, 2, ctypes.byref(need)) == -1
    assert L.mpcb_param_len(ctypes.byref(bad)) == -1
    cd, cr, cc = Dims().to_c(), RobotSpec().to_c(), SolverSettings().to_c()
    # NULL p / u_out
    args = [ctypes.byref(cd), ctypes.byref(cr), ctypes.byref(cc), 1, 1] + [None] * 15 + [None, 0, None]
    assert L.mpcb_solve_f64(*args) == -2


def test_team_rule_is_the_same_in_library_and_oracle(cuda_solver_lib):
    """Team mode changes the summation order of the ellipse cost terms: the laned oracle must use
    the same number of worker groups as the kernels for every dimension set."""
    from oracle import oracle
    OL = oracle.lib()
    for N in (1, 7, 11, 12, 20, 32, 33, 40, 64):
        for Ndyn in (0, 15, 40, 63, 64, 70, 160, 256):
            cd = Dims(N=N, Ndyn=Ndyn).to_c()
            g = cuda_solver_lib.mpcb_team_groups(ctypes.byref(cd))
            assert g == OL.mpcl_team_groups(ctypes.byref(cd)), (N, Ndyn)
            assert g == 0 or (g & (g - 1) == 0 and g <= 32 and g * N <= 320), (N, Ndyn, g)
    assert cuda_solver_lib.mpcb_team_groups(ctypes.byref(Dims().to_c())) == 0
    assert cuda_solver_lib.mpcb_team_groups(ctypes.byref(Dims(Ndyn=40).to_c())) == 0
    assert cuda_solver_lib.mpcb_team_groups(ctypes.byref(Dims(N=40, Ndyn=160).to_c())) == 8
    # latency mode (cfg.team_mode = 1): team kernels for any dimensions, same rule in both
    from dyobav_mpcnwta_warehouse_b200 import SolverSettings as SS
    cfg = SS(team_mode=1).to_c()
    cuda_solver_lib.mpcb_team_groups_cfg.restype = ctypes.c_int32
    for N in (1, 7, 20, 33, 64):
        for Ndyn in (0, 15, 40, 160):
            cd = Dims(N=N, Ndyn=Ndyn).to_c()
            g = cuda_solver_lib.mpcb_team_groups_cfg(ctypes.byref(cd), ctypes.byref(cfg))
            assert g > 0 and g == OL.mpcl_team_groups_cfg(ctypes.byref(cd), ctypes.byref(cfg)), (N, Ndyn)
    assert cuda_solver_lib.mpcb_team_groups_cfg(ctypes.byref(Dims().to_c()), ctypes.byref(cfg)) == 16
    assert cuda_solver_lib.mpcb_team_groups_cfg(ctypes.byref(Dims().to_c()), ctypes.byref(SS().to_c())) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_batched_solver_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    with pytest.raises(RuntimeError, match="CUDA"):
        BatchedSolver()
