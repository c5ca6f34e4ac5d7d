"""How much do the three unverifiable details of the PANOC/ALM restatement matter?  

Solves the SAME bench instances (headline workload, first scenarios x 8 starts) with the
reference-order CPU oracle under each sensitivity switch (oracle/mpc_oracle.c) and reports, against
the restatement as shipped: the fraction of solves whose exit status changes and whose solution
moves by more than the north-star tolerance (1e-4, inf-norm), overall and among the solves that
report Converged in both.  Writes profiles/r2_solver_sensitivity.json.
Usage: python scripts/solver_sensitivity.py [n_solves=2048] [threads]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dyobav_mpcnwta_warehouse_b200 import RobotSpec, SolverSettings, instances
from oracle import oracle

n_solves = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
wl = instances.workload("warehouse_b65536x8")
P, U0 = bench.make_inputs(wl, 8192)
n_p = n_solves // wl.starts
P, U0 = P[:n_p], U0[:n_p * wl.starts]
VARIANTS = {
    "akkt_test_off": dict(akkt=1), "akkt_uses_gradient_of_previous_iterate": dict(akkt=2),
    "linesearch_fallback_to_forward_backward_step": dict(ls=1),
    "converged_allowed_at_last_outer_iteration": dict(last=1),
}
def run(**kw):
    oracle.set_variant(**kw)
    t = time.time()
    U, SC = oracle.solve_batch(wl.dims, RobotSpec(), SolverSettings(), P, U0, starts=wl.starts, threads=threads)
    oracle.set_variant()
    return U, SC, time.time() - t
U0_, S0, t0 = run()
res = {"workload": wl.name, "solves": int(n_p * wl.starts), "baseline": {
    "exit_status_histogram": bench.histogram(S0[:, 9]), "mean_inner_iters": float(S0[:, 6].mean()), "seconds": t0}}
print("baseline", res["baseline"])
for name, kw in VARIANTS.items():
    U, S, t = run(**kw)
    du = np.abs(U - U0_).max(1)
    both = (S[:, 9] == 0) & (S0[:, 9] == 0)
    r = {"switch": kw, "status_changed_frac": float((S[:, 9] != S0[:, 9]).mean()),
         "u_moved_gt_1e-4_frac": float((du > 1e-4).mean()),
         "both_converged_frac": float(both.mean()),
         "u_moved_gt_1e-4_frac_among_both_converged": float((du[both] > 1e-4).mean()) if both.any() else None,
         "max_du_among_both_converged": float(du[both].max()) if both.any() else None,
         "cost_rel_change_median": float(np.median(np.abs(S[:, 0] - S0[:, 0]) / np.maximum(1e-12, np.abs(S0[:, 0])))),
         "exit_status_histogram": bench.histogram(S[:, 9]), "mean_inner_iters": float(S[:, 6].mean()), "seconds": t}
    res[name] = r
    print(name, r)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_solver_sensitivity.json")
json.dump(res, open(out, "w"), indent=1)
print("wrote", out)
