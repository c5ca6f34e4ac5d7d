"""Headline workload, one GPU: solves/s and launch tail with and without hardest-first ordering."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1] if len(sys.argv) > 1 else "warehouse_b65536x8")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
P, U0 = bench.make_inputs(wl, n)
Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")
for order in ("0", "1", "0", "1"):
    os.environ["MPCB_ORDER"] = "2" if order == "1" else "0"
    s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
    s.run_batch(Pd, Ud, starts=wl.starts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("hardest-first order", order, f"{ms:.1f} ms {n*wl.starts/ms*1e3:.0f} solves/s", bench.launch_tail(s._ws))
