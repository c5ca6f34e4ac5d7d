"""Where the worker pool of the team kernel spends its time (a -DMPCB_TEAM_PROF build loaded through
MPCB_LIB_PATH): idle (polling for a request), pass 1 (until the slowest worker arrives), pass 2, and how
busy each worker warp is inside pass 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload("dense_crowd_n40")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = instances.generate(wl.dims, n, seed=wl.seed, pedestrians=wl.pedestrians, modes=wl.modes)
U0 = instances.multistart_guesses(wl.dims, P, wl.starts, wl.seed)
s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
Pd = torch.as_tensor(P, device="cuda"); Ud = torch.as_tensor(U0, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
CT, WP, CB = 1024, 16, 4096
hdr = s._ws[CB:CB + CT * (1 + WP) * 8].view(torch.int64).cpu().numpy()
w = hdr[CT:].reshape(CT, WP)[:148].astype(float)
idle, p1, p2, nreq = w[:, 12], w[:, 13], w[:, 14], w[:, 15]
tot = idle + p1 + p2
own = w[:, ns:12 if ns == 2 else 16]
own = own[:, :10] if ns == 2 else own[:, :12]
print(f"{n} solves {ms:.0f} ms {n / ms * 1e3:.1f} solves/s; requests per CTA {nreq.mean():.0f}; cycles per request {tot.sum() / nreq.sum():.0f}")
print(f"  pool time: idle {100 * idle.sum() / tot.sum():.1f} %  pass 1 {100 * p1.sum() / tot.sum():.1f} %  pass 2 {100 * p2.sum() / tot.sum():.1f} %")
print(f"  per request: idle {idle.sum() / nreq.sum():.0f}  pass 1 {p1.sum() / nreq.sum():.0f}  pass 2 {p2.sum() / nreq.sum():.0f} cycles")
print("  busy share of pass 1 per worker warp (own work / wall):", np.round(own.sum(0) / p1.sum(), 2).tolist())
