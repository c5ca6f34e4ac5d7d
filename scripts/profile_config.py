"""One solve step of a named workload (for ncu / timing):
python scripts/profile_config.py <workload> <scenarios> [max_inner max_outer [passes]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1])
n = int(sys.argv[2])
kw = {}
if len(sys.argv) > 4:
    kw = dict(max_inner=int(sys.argv[3]), max_outer=int(sys.argv[4]))
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 2
P = instances.generate(wl.dims, n, seed=wl.seed, pedestrians=wl.pedestrians, modes=wl.modes)
U0 = instances.multistart_guesses(wl.dims, P, wl.starts, wl.seed)
s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings(**kw))
Pd = torch.as_tensor(P, device="cuda"); Ud = torch.as_tensor(U0, device="cuda")
for it in range(passes):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ev = o["evals"][:, :2].double().sum(0)
    print(f"{wl.name} n={n}x{wl.starts} {kw}: {ms:.1f} ms  {n*wl.starts/ms*1e3:.1f} solves/s  "
          f"{float(ev.sum())/ms*1e3/1e6:.3f} Mevals/s  conv {float((o['exit_status']==0).double().mean()):.3f} "
          f"inner {o['n_inner'].double().mean().item():.0f} evals {o['evals'][:, :2].double().mean(0).tolist()}")
