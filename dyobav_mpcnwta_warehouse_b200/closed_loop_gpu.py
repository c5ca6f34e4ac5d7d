"""Closed loop entirely on the device: per timestep one packer
launch (K4), one batched solve (K3 + K1) and one plant launch (K5) — no host work, no
host<->device traffic between steps.  Built from the same ``Episode`` objects as the host
harness (``closed_loop.ClosedLoopBatch``), which it reproduces bit for bit when the host
plant uses ``_lib.sincos_host``."""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from . import _lib, packing
from .closed_loop import Episode
from .problem import Dims, MpcConfig, SolverSettings
from .solver import BatchedSolver


class ClosedLoopGPU:
    def __init__(self, dims: Dims, cfg: MpcConfig, episodes: Sequence[Episode],
                 settings: SolverSettings = SolverSettings(), device=None):
        import torch
        self.torch = torch
        self.dims, self.cfg = dims, cfg
        rb = cfg.robot()
        self.solver = BatchedSolver(dims, rb, settings, device=device)
        self.device = self.solver.device
        self.lib = self.solver.lib
        eps = list(episodes)
        n = len(eps)
        base_speed = rb.lin_vel_max * 0.8
        trajs = [packing.ref_traj_from_path(rb.ts, e.ref_path, e.state, base_speed) for e in eps]
        T = max(len(t) for t in trajs)
        Kp = max(1, max(len(e.polygons) for e in eps))
        Pd = max(1, max(len(e.pedestrians) for e in eps))
        M = max(1, max((len(p.mode_velocities) for e in eps for p in e.pedestrians), default=1))
        if any(len(e.pedestrians) not in (0, Pd) for e in eps):
            raise ValueError("all episodes must have the same number of pedestrians")
        ref = np.zeros((n, T, 3)); ref_len = np.zeros(n, np.int32)
        polys = np.zeros((n, Kp, 4, 2)); n_poly = np.zeros(n, np.int32)
        ped_pos = np.zeros((n, Pd, 2)); ped_vel = np.zeros((n, Pd, M, 2))
        for i, (e, t) in enumerate(zip(eps, trajs)):
            ref[i, :len(t)] = np.asarray(t); ref_len[i] = len(t)
            for k, poly in enumerate(e.polygons):
                polys[i, k] = np.asarray(poly, dtype=np.float64)
            n_poly[i] = len(e.polygons)
            for k, p in enumerate(e.pedestrians):
                ped_pos[i, k] = p.position
                ped_vel[i, k, :len(p.mode_velocities)] = p.mode_velocities
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=self.device)  # noqa: E731
        self.state = dev(np.array([e.state for e in eps], dtype=np.float64))
        self.last_u = torch.zeros(n, 2, dtype=torch.float64, device=self.device)
        self.ref_traj, self.ref_len = dev(ref), dev(ref_len)
        self.idx_ref = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.goal = dev(np.array([e.ref_path[-1][:2] for e in eps], dtype=np.float64))
        self.polys, self.n_poly = dev(polys), dev(n_poly)
        self.ped_pos, self.ped_vel = dev(ped_pos), dev(ped_vel)
        self.done = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.P = torch.empty(n, dims.np, dtype=torch.float64, device=self.device)
        self.out = self.solver.alloc_outputs(n)
        sim = _lib.CSim()
        sim.n, sim.T, sim.Kp, sim.Pd, sim.M = n, T, Kp, (Pd if any(e.pedestrians for e in eps) else 0), M
        sim.base_speed, sim.lin_vel_max, sim.ped_size = base_speed, rb.lin_vel_max, 0.2
        sim.stc_w, sim.dyn_w, sim.ts = 10.0, 10.0, rb.ts
        for i, v in enumerate(cfg.tuning_params()):
            sim.tuning[i] = float(v)
        for name in ("state", "last_u", "ref_traj", "ref_len", "idx_ref", "goal", "polys", "n_poly",
                     "ped_pos", "ped_vel", "done"):
            setattr(sim, name, getattr(self, name).data_ptr())
        self.sim = sim
        self._cd = dims.to_c()
        self.n = n

    def step(self):
        """One control period for every episode: pack -> solve -> plant, all on the device."""
        torch = self.torch
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.mpcb_pack_f64(ctypes.byref(self._cd), ctypes.byref(self.sim),
                                              ctypes.c_void_p(self.P.data_ptr()), st), "mpcb_pack_f64")
            self.solver.run_batch(self.P, out=self.out)
            _lib.check(self.lib.mpcb_plant_step_f64(ctypes.byref(self._cd), ctypes.byref(self.sim),
                                                    ctypes.c_void_p(self.out["u"].data_ptr()), st),
                       "mpcb_plant_step_f64")

    def run(self, steps: int, record: bool = False):
        traj = [self.state.clone()] if record else None
        for _ in range(steps):
            self.step()
            if record:
                traj.append(self.state.clone())
        return traj
