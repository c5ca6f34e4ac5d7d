#!/usr/bin/env python
"""bench.py — MPC solves/s of the batched PANOC/ALM solver (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: every rank solves
`--scenarios` (default 8192) scenarios x 8 multi-start guesses = 65 536 solves at the
reference's default dims (N=20, np=2778) — BASELINE.json configs[3], sharded so that
8 GPUs together cover 65 536 scenarios x 8 starts (weak scaling: per-GPU work fixed).

  value        solves/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through BatchedSolver.run_batch with HOST (pinned) buffers:
               H2D of P and U0 and D2H of (u, cost, exit_status) inside the timed region
  roofline     the solve kernel against the FP64 pipe (this path is sequential small-state
               recursion + reductions: neither HBM- nor tensor-bound)
  cpu_baseline the CPU oracle port on this box's host cores, bounded sample
  --impl reference   the same oracle port timed alone (the reference's own solver is
               generated Rust/C that cannot be built without cargo/casadi)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic work of one horizon evaluation at the default dims
W_PSI_FLOP = {"default": 36.0e3}
W_LBFGS_FLOP = 1.84e3


def w_psi(dims):
    """SURVEY §8(d) per-unit figure, generalised over dims (same counting convention)."""
    N = dims.N
    per_step = 70 + 9 + 8 * (2 * dims.Nother - 1) + 2 + 29 * dims.Nstc + 66 * dims.Ndyn
    return N * per_step + 25 * (N * (N + 1) // 2) + 10 * 2 * N + 2 * dims.n2


def hbm_roofline(algorithmic_bytes, seconds):
    """The HBM side of the roofline, for the record: this path moves ~3 KB per solve."""
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    ach = algorithmic_bytes / seconds / 1e9
    return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
            "note": "algorithmic bytes 8*(np/starts + 2*2N + 10) per solve; far from the bound by design"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_baseline(dims, robot, cfg, P, U0, starts, n_solves, threads):
    """Time the oracle port on `n_solves` instances of the same workload with `threads` cores."""
    from oracle import oracle
    n_p = max(1, n_solves // starts)
    t0 = time.perf_counter()
    U, SC = oracle.solve_batch(dims, robot, cfg, P[:n_p], U0[:n_p * starts], starts=starts, threads=threads)
    dt = time.perf_counter() - t0
    return n_p * starts / dt, n_p * starts, dt, U, SC


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=8192, help="scenarios per GPU (x8 starts)")
    ap.add_argument("--starts", type=int, default=8)
    ap.add_argument("--cpu-solves", type=int, default=0, help="CPU sample size (0: 24 per core)")
    ap.add_argument("--latency-solves", type=int, default=24, help="single-solve latency sample (0: skip)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
    dims, robot, cfg = Dims(), RobotSpec(), SolverSettings()
    starts = args.starts
    cores = os.cpu_count() or 1
    config = {"workload": "warehouse_b65536x8 (BASELINE configs[3]): N=20, np=2778, "
                          f"{args.scenarios} scenarios x {starts} multi-start guesses per GPU",
              "solves_per_gpu": args.scenarios * starts, "settings": "tol 1e-4, <=10 outer x <=500 inner, c0=10",
              "l2": "inputs larger than L2 (P shard %.0f MB + staged copy)" % (args.scenarios * dims.np * 8 / 1e6)}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        n_solves = args.cpu_solves or max(16, 8 * cores)
        n_p = max(1, n_solves // starts)
        P = instances.generate(dims, n_p, seed=instances.BASE_SEED + 4)
        U0 = instances.multistart_guesses(dims, P, starts, instances.BASE_SEED + 4)
        for _ in range(args.warmup):
            cpu_baseline(dims, robot, cfg, P, U0, starts, min(cores, n_p * starts), cores)
        t_tot, n_tot = 0.0, 0
        for _ in range(args.steps):
            v, n, dt, _, _ = cpu_baseline(dims, robot, cfg, P, U0, starts, n_p * starts, cores)
            t_tot += dt
            n_tot += n
        val = n_tot / t_tot
        line = {"impl": "reference", "metric": "MPC solves/sec", "value": val, "unit": "solves/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_p * starts} solves per step of the same workload "
                                           "(oracle C port of OpEn PANOC/ALM, OpenMP over instances; the reference's "
                                           "generated Rust solver cannot be built here)"},
                "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    from dyobav_mpcnwta_warehouse_b200.csrc import build as cbuild
    cbuild.build()

    from dyobav_mpcnwta_warehouse_b200 import sharding
    n_global = args.scenarios * world                     # weak scaling: 8192 scenarios per GPU
    lo, hi = sharding.shard_range(n_global, rank, world)  # contiguous slice, all starts together
    seed = instances.BASE_SEED + 4 + 1000 * rank          # each rank generates its own slice
    P = instances.generate(dims, hi - lo, seed=seed)
    U0 = instances.multistart_guesses(dims, P, starts, seed)
    B = (hi - lo) * starts
    solver = BatchedSolver(dims, robot, cfg, device=dev)
    P_h = torch.from_numpy(P).pin_memory()
    U0_h = torch.from_numpy(U0).pin_memory()
    P_d, U0_d = P_h.to(dev), U0_h.to(dev)
    out = solver.alloc_outputs(B)
    u_h = torch.empty((B, dims.nu_total), dtype=torch.float64).pin_memory()
    cost_h = torch.empty(B, dtype=torch.float64).pin_memory()
    st_h = torch.empty(B, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        solver.run_batch(P_d, U0_d, starts=starts, out=out)

    def step_e2e():
        # host buffers in, host buffers out; the only inter-GPU traffic of the whole job is
        # the final gather of the best-of-starts solutions
        Pd = P_h.to(dev, non_blocking=True)
        Ud = U0_h.to(dev, non_blocking=True)
        o = solver.run_batch(Pd, Ud, starts=starts, out=out)
        bc, bu, bs, _ = sharding.best_of_starts(o["cost"], o["u"], o["exit_status"], starts)
        sharding.gather_results({"u": bu, "cost": bc, "exit_status": bs}, n_global)
        u_h.copy_(o["u"], non_blocking=True)
        cost_h.copy_(o["cost"], non_blocking=True)
        st_h.copy_(o["exit_status"], non_blocking=True)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_resident()
    with ClockSampler(local_rank) as clk:
        ms = timed(step_resident, args.steps)
    # the dominant kernel alone, on the launching stream (roofline): K3 staging is
    # ~0.1 % of a step, so the step time is the solve kernel's time to that precision
    ev = out["evals"].to(torch.float64).sum(0)
    n_inner = out["n_inner"].to(torch.float64).sum()
    wps = w_psi(dims)
    flop_per_launch = float(ev[0]) * wps + float(ev[1]) * 3 * wps + float(n_inner) * W_LBFGS_FLOP
    ms_e2e = timed(step_e2e, args.steps)
    total_solves = B * world
    value = total_solves * args.steps / (ms / 1e3)
    e2e = total_solves * args.steps / (ms_e2e / 1e3)

    status = out["exit_status"]
    stats = {"converged_frac": float((status == 0).double().mean()),
             "mean_inner_iters": float(out["n_inner"].double().mean()),
             "mean_cost_evals": float(out["evals"][:, 0].double().mean()),
             "mean_costgrad_evals": float(out["evals"][:, 1].double().mean())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp64_nominal = 148 * 64 * 2 * 1.965e9 / 1e12  # 64 FP64 FMA/clk/SM at max clock
    from dyobav_mpcnwta_warehouse_b200 import _lib
    try:
        fp64_peak, peak_src = _lib.fp64_peak_tflops(), "measured live: 8 independent DFMA chains/thread (mpcb_fp64_peak_tflops)"
    except Exception as exc:                      # never fall back silently
        raise RuntimeError(f"FP64 peak probe failed: {exc}")
    kernel_s = ms / 1e3 / args.steps
    achieved = flop_per_launch / kernel_s / 1e12
    line = {
        "metric": "MPC solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "e2e": {"value": e2e, "unit": "solves/s",
                "h2d_bytes_per_step": int(P_h.numel() * 8 + U0_h.numel() * 8),
                "d2h_bytes_per_step": int(u_h.numel() * 8 + cost_h.numel() * 8 + st_h.numel() * 4)},
        "gpu_launches": 2 * args.steps,
        "clocks": clk.summary(),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak,
                     # ncu dram__bytes_read+write of this kernel: 43 384 064 B for 16 384 solves
                     # (profiles/r1_traffic_metrics.csv) = 2 648 B per solve,
                     # scaled to this launch
                     "traffic": 2648.0 * B,
                     "peak_source": peak_src, "peak_nominal": fp64_nominal,
                     "note": "MEASURED_PEAKS.json holds no FP64 figure, so the FMA pipe is probed in this run "
                             "(nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2). achieved = SURVEY 8(d) work "
                             f"(W_psi={wps} flop, grad=3x, L-BFGS {W_LBFGS_FLOP:.0f}/iter) x evaluations "
                             "counted by the kernel / CUDA-event time",
                     "hbm": hbm_roofline((B * 8 * (dims.np / starts + 2 * dims.nu_total + 10)), kernel_s)},
        "solve_stats": stats,
    }
    if args.latency_solves > 0:
        # BASELINE configs[1]: one solve per call through the drop-in `solver().run(p)`, host lists
        from dyobav_mpcnwta_warehouse_b200.solver import solver as make_solver
        s1 = make_solver(dims, robot, cfg)
        wall, dev_ms = [], []
        for i in range(min(args.latency_solves, P.shape[0])):
            row = P[i].tolist()
            t0 = time.perf_counter()
            r = s1.run(row)
            wall.append(1e3 * (time.perf_counter() - t0))
            dev_ms.append(r.solve_time_ms)
        line["latency"] = {"p50_ms": float(np.percentile(wall, 50)), "p95_ms": float(np.percentile(wall, 95)),
                           "p50_device_ms": float(np.percentile(dev_ms, 50)), "solves": len(wall),
                           "what": "wall clock of solver().run(p) incl. H2D/D2H, one instance per call, "
                                   "reference settings (no wall-clock cap)"}
    if not args.no_cpu and world >= 1:
        n_cpu = args.cpu_solves or max(16, 24 * cores)
        v, n, dt, _, _ = cpu_baseline(dims, robot, cfg, P, U0, starts, n_cpu, cores)
        line["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                                "sample": f"first {n} solves of rank 0's batch, {dt:.1f} s wall, "
                                          "oracle C port (OpenMP over instances)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
