"""Whole-solve regression vectors (tests/golden/solve_regression.npz, written by
tests/golden/gen_solve_regression.py from the laned oracle): the CPU twin and the CUDA kernel
must both reproduce the committed bits — solution, cost, status, every counter."""
import os

import numpy as np
import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
import sys  # noqa: E402
sys.path.insert(0, os.path.join(HERE, "golden"))
from gen_solve_regression import CASES  # noqa: E402


@pytest.fixture(scope="module")
def fixture():
    return np.load(os.path.join(HERE, "golden", "solve_regression.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_laned_oracle_reproduces_committed_solves(fixture, name):
    dk, n_p, starts, modes, sk, seed = CASES[name]
    dims = Dims(**dk)
    U, SC = oracle.solve_batch(dims, RobotSpec(), SolverSettings(**sk), fixture[name + "/P"],
                               fixture[name + "/U0"], starts=starts, threads=os.cpu_count() or 1, laned=True)
    np.testing.assert_array_equal(U, fixture[name + "/U"])
    np.testing.assert_array_equal(SC[:, :10], fixture[name + "/SC"])   # (column 10, n_small, was added later)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_committed_solves(fixture, name):
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    assert torch.cuda.is_available()
    dk, n_p, starts, modes, sk, seed = CASES[name]
    dims = Dims(**dk)
    s = BatchedSolver(dims, RobotSpec(), SolverSettings(**sk))
    dev = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda").contiguous()  # noqa: E731
    o = {k: v.cpu().numpy() for k, v in
         s.run_batch(dev(fixture[name + "/P"]), dev(fixture[name + "/U0"]), starts=starts).items()}
    SC = fixture[name + "/SC"]
    np.testing.assert_array_equal(o["u"], fixture[name + "/U"])
    np.testing.assert_array_equal(o["cost"], SC[:, 0])
    np.testing.assert_array_equal(o["fpr"], SC[:, 1])
    np.testing.assert_array_equal(o["f2_norm"], SC[:, 3])
    np.testing.assert_array_equal(o["penalty"], SC[:, 4])
    np.testing.assert_array_equal(o["n_outer"], SC[:, 5].astype(np.int32))
    np.testing.assert_array_equal(o["n_inner"], SC[:, 6].astype(np.int32))
    np.testing.assert_array_equal(o["evals"][:, 0], SC[:, 7].astype(np.int32))
    np.testing.assert_array_equal(o["evals"][:, 1], SC[:, 8].astype(np.int32))
    np.testing.assert_array_equal(o["exit_status"], SC[:, 9].astype(np.int32))
