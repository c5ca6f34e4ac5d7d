"""Where an inner iteration of the latency kernel spends its time (debug build with -DMPCB_SPEC_PROF,
loaded through MPCB_LIB_PATH): clock64 segments of the solving warp, summed over a solve."""
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
s.run_batch(Pd, Ud, starts=wl.starts)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
B = n * wl.starts
CB = 4096 + 1024 * 8
t = s._ws[CB:CB + 8 * 8 * min(B, 128)].view(torch.int64).cpu().numpy().reshape(-1, 8).astype(float)
it = o["n_inner"].cpu().numpy()[:t.shape[0]].astype(float)
names = ["lip check, envelope, publish", "trial wait (bar1..bar2)", "accept, fpr, AKKT", "two-loop recursion", "wait look-ahead (bar3)", "-", "L-BFGS update", "-"]
tot = t.sum()
print(f"{B} solves, {e0.elapsed_time(e1):.1f} ms, mean iterations {it.mean():.0f}; cycles per inner iteration {tot / it.sum():.0f}")
for i in (1, 2, 6, 3, 4, 0):
    print(f"  {names[i]:32s} {t[:, i].sum() / it.sum():8.0f} cycles/iteration  {100 * t[:, i].sum() / tot:5.1f} %")
