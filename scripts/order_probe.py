"""configs[2]: how much of the launch tail a better queue order removes.  Orders tried (as a row
permutation of P with the kernel's own ordering switched off): natural, the static difficulty key,
what a 150-iteration budgeted pre-solve predicts (penalty constraints active or not), and - as the
bound - longest-first by the true amount of work."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1] if len(sys.argv) > 1 else "warehouse_b4096_ndyn40")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
P, U0 = bench.make_inputs(wl, n)
assert wl.starts == 1
Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")

def timed(cfg, Pd, Ud):
    s = BatchedSolver(wl.dims, RobotSpec(), cfg)
    s.run_batch(Pd[:296], Ud[:296], starts=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=1); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), o, s

ms, full, s = timed(SolverSettings(), Pd, Ud)
print(f"kernel's own order (static key): {ms:.1f} ms")
CB = 4096 + 1024 * 17 * 8
keys = s._ws[CB + 4 * n:CB + 8 * n].view(torch.float32).cpu().numpy().copy()
ev = full["evals"].cpu().numpy(); work = ev[:, 0] + 3.0 * ev[:, 1]
os.environ["MPCB_ORDER"] = "0"
ms, _, _ = timed(SolverSettings(), Pd, Ud)
print(f"natural order: {ms:.1f} ms")
for budget in (50, 150):
    msb, pre, _ = timed(SolverSettings(max_inner_total=budget), Pd, Ud)
    f2 = pre["f2_norm"].cpu().numpy(); st = pre["exit_status"].cpu().numpy()
    k = np.nan_to_num(keys, posinf=99.0)
    for name, pred in (("active then static key", (f2 > 0) * 100.0 - k), ("active then violation", (f2 > 0) * 100.0 + np.minimum(f2, 50.0))):
        pr = np.where(st != 2, -1e30, pred)
        perm = torch.as_tensor(np.argsort(-pr, kind="stable"), device="cuda")
        ms, o, _ = timed(SolverSettings(), Pd[perm].contiguous(), Ud[perm].contiguous())
        same = torch.equal(o["u"], full["u"][perm]) and torch.equal(o["n_inner"], full["n_inner"][perm])
        print(f"pre-solve {budget} iterations ({msb:.1f} ms), {name}: {ms:.1f} ms  same results {same}")
perm = torch.as_tensor(np.argsort(-work, kind="stable"), device="cuda")
ms, _, _ = timed(SolverSettings(), Pd[perm].contiguous(), Ud[perm].contiguous())
print(f"longest first by true work: {ms:.1f} ms")
