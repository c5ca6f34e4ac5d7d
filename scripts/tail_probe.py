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
    os.environ["MPCB_ORDER"] = order
    s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
    s.run_batch(Pd, Ud, starts=wl.starts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("MPCB_ORDER", order, f"{ms:.1f} ms {n*wl.starts/ms*1e3:.0f} solves/s", bench.launch_tail(s._ws))
# iterations by position in the solve order (last run: MPCB_ORDER=1)
import numpy as np
hdr = 4096 + 1024 * 17 * 8
order = s._ws[hdr:hdr + 4 * n].view(torch.int32).cpu().numpy()
keys = s._ws[hdr + 4 * n:hdr + 8 * n].view(torch.float32).cpu().numpy()
it = o["n_inner"].cpu().numpy().reshape(n, wl.starts).mean(1)
print("order is a permutation:", np.array_equal(np.sort(order), np.arange(n)), "keys finite:", np.isfinite(keys).mean(),
      "key quantiles", np.quantile(keys[np.isfinite(keys)], [0, .1, .25, .5, .75, .9, 1]).round(2))
print("mean inner iterations by decile of the solve order:", [int(it[order[i * n // 10:(i + 1) * n // 10]].mean()) for i in range(10)])
