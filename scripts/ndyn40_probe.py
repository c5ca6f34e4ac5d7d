"""configs[2] (4096 scenarios x 40 ellipses), one GPU: solves/s, launch tail and idle warp-time,
one-warp kernel vs team kernels (cfg.team_mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1] if len(sys.argv) > 1 else "warehouse_b4096_ndyn40")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
P, U0 = bench.make_inputs(wl, n)
Pd, Ud = torch.as_tensor(P, device="cuda"), torch.as_tensor(U0, device="cuda")
ref = None
for mode in modes:
    # modes >= 10: one-warp kernel, environment knob MPCB_PROBE_ENV=NAME set to mode - 10 (A/B runs
    # of an experimental switch; outputs compared with the first run)
    if mode >= 10 and os.environ.get("MPCB_PROBE_ENV"):
        os.environ[os.environ["MPCB_PROBE_ENV"]] = str(mode - 10)
    s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings(team_mode=mode if mode < 10 else 0))
    s.run_batch(Pd[:296], Ud[:296 * wl.starts], starts=wl.starts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    CT, WP, CB = 1024, 16, 4096
    hdr = s._ws[CB:CB + CT * (1 + WP) * 8].view(torch.int64).cpu().numpy()
    start, fin = hdr[:CT], hdr[CT:].reshape(CT, WP)
    fin = fin[fin > 0]; start = start[start > 0]
    t0, t1 = float(start.min()), float(fin.max())
    idle = float(np.mean((t1 - fin) / (t1 - t0)))
    q = np.percentile((fin - t0) / (t1 - t0), [10, 25, 50, 75, 90])
    it = o["n_inner"].cpu().numpy()
    if mode >= 10:
        if ref is None:
            ref = {k: v.clone() for k, v in o.items()}
        else:
            print("  same bits as the first run:", all(torch.equal(torch.nan_to_num(ref[k].double(), nan=-7.0), torch.nan_to_num(o[k].double(), nan=-7.0))
                                                    for k in ("u", "cost", "exit_status", "n_inner", "evals", "y")))
    print(f"team_mode {mode}: {ms:.1f} ms {n*wl.starts/ms*1e3:.0f} solves/s  idle warp-time {idle:.3f}  "
          f"warp finish quantiles {np.round(q,3).tolist()}  warps {fin.size}  mean iters {it.mean():.0f} max {it.max()}")
