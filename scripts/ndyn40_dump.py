"""Per-instance statistics of a configs[2] launch (iterations, evaluations, difficulty key, queue order)
and what a 50-iteration budgeted pre-solve sees, for scheduling studies.  Writes gpurun_out/ndyn40_stats.npz."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1] if len(sys.argv) > 1 else "warehouse_b4096_ndyn40")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
P, U0 = bench.make_inputs(wl, n)
Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")
out = {}
for tag, cfg in (("full", SolverSettings()), ("b50", SolverSettings(max_inner_total=50)), ("b150", SolverSettings(max_inner_total=150))):
    s = BatchedSolver(wl.dims, RobotSpec(), cfg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
    print(tag, f"{e0.elapsed_time(e1):.1f} ms")
    for k in ("n_inner", "n_outer", "evals", "exit_status", "fpr", "f2_norm", "f1_infeas", "penalty", "cost"):
        out[f"{tag}_{k}"] = o[k].cpu().numpy()
    if tag == "full":
        CB = 4096 + 1024 * 17 * 8
        ws = s._ws
        out["order"] = ws[CB:CB + 4 * n].view(torch.int32).cpu().numpy()
        out["keys"] = ws[CB + 4 * n:CB + 8 * n].view(torch.float32).cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/ndyn40_stats.npz", **out)
