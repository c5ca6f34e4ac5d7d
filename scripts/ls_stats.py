import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
for name, n in (("warehouse_b4096_ndyn40", 148), ("warehouse_b65536x8", 18)):
    wl = instances.workload(name)
    P, U0 = bench.make_inputs(wl, n)
    Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")
    for spec in ("0", "1", "0", "1"):
        os.environ["MPCB_SPEC"] = spec
        s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
        s.run_batch(Pd, Ud, starts=wl.starts)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
        ev = o["evals"].cpu().numpy(); it = o["n_inner"].cpu().numpy()
        print(name, "B", n * wl.starts, "spec", spec, f"{e0.elapsed_time(e1):.1f} ms", "cost evals/iter", ev[:, 0].sum() / it.sum(),
              "grad evals/iter", ev[:, 1].sum() / it.sum(), "max iters", it.max())
