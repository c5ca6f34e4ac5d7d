import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload("warehouse_b65536x8")
for n in [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else '1,18,37,74'.split(','))]:
    P, U0 = bench.make_inputs(wl, n)
    Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")
    s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
    s.run_batch(Pd, Ud, starts=wl.starts)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("B", n * wl.starts, "ms", round(min(ts), 1))
