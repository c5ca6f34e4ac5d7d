"""How often the latency kernel's deferred Lipschitz test fails after all (the discarded-batch path), in a
-DMPCB_SPEC_PROF build: the parity tests must exercise it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
dims, starts = Dims(), 8
for n_p, seed, kw in ((8, 3, {}), (6, 5, {}), (24, 23, {})):
    P = instances.generate(dims, n_p, seed=seed, modes=3)
    U0 = instances.multistart_guesses(dims, P, starts, seed)
    s = BatchedSolver(dims, RobotSpec(), SolverSettings(**kw))
    o = s.run_batch(torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda"), starts=starts)
    torch.cuda.synchronize()
    cnt = int(s._ws[4096 + (1024 + 16008) * 8:4096 + (1024 + 16009) * 8].view(torch.int64).item())
    print(f"n_p {n_p} seed {seed}: {n_p * starts} solves, {int(o['n_inner'].sum())} inner iterations, deferred Lipschitz failures: {cnt}")
