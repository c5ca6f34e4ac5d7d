"""Where one horizon evaluation spends its time (a -DMPCB_EVAL_PROF build loaded through MPCB_LIB_PATH):
clock64 segments summed over every evaluation of a launch.
python scripts/eval_prof.py <workload> <scenarios>   (small batch: latency kernel; large: one-warp kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1] if len(sys.argv) > 1 else "warehouse_b65536x8")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
P, U0 = bench.make_inputs(wl, n)
Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")
s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
o = s.run_batch(Pd, Ud, starts=wl.starts); torch.cuda.synchronize()
CB = 4096
t = s._ws[CB + (1024 + 16000) * 8:CB + (1024 + 16008) * 8].view(torch.int64).cpu().numpy().astype(float)
names = ["rollout", "reference path", "speed, control, fleet", "polygons", "ellipses + F2", "terminal, F2 grad, accel, totals", "adjoint"]
ne = t[7]
print(f"{wl.name} B = {n * wl.starts}: {ne:.0f} evaluations, {t[:7].sum() / ne:.0f} cycles each")
for i, nm in enumerate(names):
    print(f"  {nm:34s} {t[i] / ne:7.0f} cycles  {100 * t[i] / t[:7].sum():5.1f} %")
