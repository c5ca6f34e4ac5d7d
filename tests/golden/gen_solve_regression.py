#!/usr/bin/env python
"""Regression vectors for WHOLE SOLVES, produced by the laned oracle (oracle/mpc_oracle_laned.c:
the kernel's documented operation order) — NOT by the reference: the reference's generated
solver cannot be built here, so these pin the arithmetic contract between
the CUDA kernel and its CPU twin across rounds, nothing more.  Anything that changes a single
bit of a solve (a reordered sum, a different fma placement) shows up as a diff against this file
and has to be a decision, not an accident.

Output: tests/golden/solve_regression.npz.  Usage: python tests/golden/gen_solve_regression.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {   # name: (dims kwargs, n_p, starts, modes, solver settings, seed)
    "default_full": (dict(), 6, 2, 3, dict(), 11),
    "ndyn40_short": (dict(Ndyn=40), 4, 1, 20, dict(max_inner=100, max_outer=3), 12),
    "n40_two_rows": (dict(N=40, Nother=4, Nstc=6, Ndyn=24), 3, 2, 3, dict(max_inner=60, max_outer=3), 13),
}


def main():
    out = {}
    for name, (dk, n_p, starts, modes, sk, seed) in CASES.items():
        dims = Dims(**dk)
        P = instances.generate(dims, n_p, seed=seed, modes=modes)
        U0 = instances.multistart_guesses(dims, P, starts, seed)
        U, SC = oracle.solve_batch(dims, RobotSpec(), SolverSettings(**sk), P, U0, starts=starts,
                                   threads=os.cpu_count() or 1, laned=True)
        out[name + "/P"] = P
        out[name + "/U0"] = U0
        out[name + "/U"] = U
        out[name + "/SC"] = SC
        print(name, "solves", n_p * starts, "mean inner", SC[:, 6].mean(), "status", SC[:, 9].astype(int).tolist())
    np.savez_compressed(os.path.join(HERE, "solve_regression.npz"), **out)


if __name__ == "__main__":
    main()
