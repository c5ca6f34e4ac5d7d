import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import solver
d = Dims(); P = instances.generate(d, 40, seed=instances.BASE_SEED + 2)
s = solver(d, RobotSpec(), SolverSettings())
rows = [p.tolist() for p in P]
s.run(rows[0])
w, dv, it = [], [], []
for r in rows:
    t0 = time.perf_counter(); o = s.run(r); w.append(1e3 * (time.perf_counter() - t0)); dv.append(o.solve_time_ms); it.append(o.num_inner_iterations)
w, dv, it = np.array(w), np.array(dv), np.array(it)
print("wall p50 %.2f p95 %.2f | device p50 %.2f | overhead p50 %.3f max %.3f | us per inner iteration p50 %.1f" % (np.percentile(w, 50), np.percentile(w, 95), np.percentile(dv, 50), np.percentile(w - dv, 50), (w - dv).max(), np.percentile(1e3 * dv / np.maximum(it, 1), 50)))
