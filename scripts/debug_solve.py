import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
from oracle import oracle
dims = Dims()
n_p, starts = int(sys.argv[1]), int(sys.argv[2])
mi, mo = int(sys.argv[3]), int(sys.argv[4])
cfg = SolverSettings(max_inner=mi, max_outer=mo)
P = instances.generate(dims, n_p, seed=3)
U0 = instances.multistart_guesses(dims, P, starts, 3)
s = BatchedSolver(dims, RobotSpec(), cfg)
dev = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda").contiguous()
o = {k: v.cpu().numpy() for k, v in s.run_batch(dev(P), dev(U0), starts=starts).items()}
Uo, SC = oracle.solve_batch(dims, RobotSpec(), cfg, P, U0, starts=starts, threads=8)
du = np.max(np.abs(o["u"] - Uo), axis=1)
for b in range(n_p * starts):
    print(b, "du %.2e" % du[b], "gpu st/out/in/ev", o["exit_status"][b], o["n_outer"][b], o["n_inner"][b], o["evals"][b],
          "cpu", int(SC[b, 9]), int(SC[b, 5]), int(SC[b, 6]), int(SC[b, 7]), int(SC[b, 8]),
          "cost %.6f %.6f" % (o["cost"][b], SC[b, 0]), "c %.0f %.0f" % (o["penalty"][b], SC[b, 4]), "fpr %.2e %.2e" % (o["fpr"][b], SC[b,1]))
