"""One solve step for ncu: python scripts/profile_step.py <scenarios> <starts> [max_inner max_outer]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
n_p, starts = int(sys.argv[1]), int(sys.argv[2])
kw = {}
if len(sys.argv) > 4:
    kw = dict(max_inner=int(sys.argv[3]), max_outer=int(sys.argv[4]))
dims = Dims()
P = instances.generate(dims, n_p, seed=instances.BASE_SEED + 4)
U0 = instances.multistart_guesses(dims, P, starts, instances.BASE_SEED + 4)
s = BatchedSolver(dims, RobotSpec(), SolverSettings(**kw))
dev = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda").contiguous()
Pd, Ud = dev(P), dev(U0)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=starts); e1.record(); torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1), "solves/s", n_p * starts / e0.elapsed_time(e1) * 1e3,
          "mean evals", o["evals"].double().mean(0).tolist(), "mean inner", o["n_inner"].double().mean().item())
