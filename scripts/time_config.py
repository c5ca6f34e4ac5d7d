"""Time one BASELINE config on one GPU: python scripts/time_config.py <workload> [n_scenarios] [check]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
wl = instances.workload(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl.n_scenarios
check = len(sys.argv) > 3
t = time.time()
P = instances.generate(wl.dims, n, seed=wl.seed, pedestrians=wl.pedestrians, modes=wl.modes)
U0 = instances.multistart_guesses(wl.dims, P, wl.starts, wl.seed)
print(f"{wl.name}: dims {wl.dims}, np={wl.dims.np}, {n} scenarios x {wl.starts} starts, generated in {time.time()-t:.1f}s")
s = BatchedSolver(wl.dims, RobotSpec(), SolverSettings())
Pd = torch.as_tensor(P, device="cuda"); Ud = torch.as_tensor(U0, device="cuda")
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.run_batch(Pd, Ud, starts=wl.starts); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"  pass {it}: {ms:.1f} ms  {n*wl.starts/ms*1e3:.1f} solves/s  converged {float((o['exit_status']==0).double().mean()):.3f} "
          f"mean inner {o['n_inner'].double().mean().item():.0f} evals {o['evals'].double().mean(0).tolist()}")
if check:
    from oracle import oracle
    m = min(8, n)
    Uo, SC = oracle.solve_batch(wl.dims, RobotSpec(), SolverSettings(), P[:m], U0[:m*wl.starts], starts=wl.starts, threads=os.cpu_count(), laned=True)
    print("  bit-exact vs laned oracle on first", m*wl.starts, "solves:", bool(np.array_equal(o["u"][:m*wl.starts].cpu().numpy(), Uo)))
