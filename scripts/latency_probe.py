"""Single-solve latency of the drop-in Solver.run, one-warp kernel vs latency mode (team kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import solver
wl = instances.workload("warehouse_b65536x8")
P, _ = bench.make_inputs(wl, 8192)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rows = [P[i].tolist() for i in range(n)]
for mode in (0, 1, 0, 1):
    s = solver(wl.dims, RobotSpec(), SolverSettings(team_mode=mode))
    for r in rows[:20]:
        s.run(r)
    wall, dev, its = [], [], []
    for r in rows:
        t0 = time.perf_counter(); o = s.run(r); wall.append(1e3 * (time.perf_counter() - t0)); dev.append(o.solve_time_ms); its.append(o.num_inner_iterations)
    print(f"team_mode {mode}: p50 {np.percentile(wall,50):.2f} ms p95 {np.percentile(wall,95):.2f} ms  device p50 {np.percentile(dev,50):.2f}  "
          f"us per inner iteration {1e3*np.sum(dev)/np.sum(its):.2f}  mean iters {np.mean(its):.0f}")
