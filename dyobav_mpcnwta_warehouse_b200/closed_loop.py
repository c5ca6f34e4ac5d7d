"""Batched closed-loop (receding-horizon) harness — SURVEY §8(f-2).

Mirrors, for E independent episodes at once, the per-timestep sequence of
``MainBase.run_one_step`` (main_base.py:267-346) around the solver call:

  reference window   ``TrajectoryTracker.get_ref_states``      (trajectory_tracker.py:243-270)
  speed reference    ``run_step`` incl. its max() quirk        (:304-310, SURVEY C-8)
  dynamic obstacles  per-offset (mu, std) -> ``[mx,my,sx,sy,0,1]`` (main_base.py:293-302), here
                     from constant-velocity pedestrian modes (the CNN predictor is out of scope)
  static obstacles   the Nstcobs closest polygons as half-spaces (mpc_interface.py:73-100)
  parameter row      ``params = last_u + state + finish_state + …`` (:315-317)
  solve              ONE batched call for all episodes  <- the hot path
  plant              RK4 unicycle step with the first action (basic_agent.py:106,
                     motion_model.py:141-163), pedestrians advance, termination test (:191-199)

The solver is a callable ``P[E, np] -> (u[E, 2N], cost[E], exit_status[E])`` so that the GPU
solver and the CPU oracle can drive the identical loop (tests compare the trajectories).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Sequence

import numpy as np

from . import packing
from .problem import Dims, MpcConfig


def _libm_sincos(x: float):
    return math.sin(x), math.cos(x)


def unicycle_rk4(state: np.ndarray, action: np.ndarray, ts: float, sincos=_libm_sincos) -> np.ndarray:
    """The reference's numpy branch of ``unicycle_model`` (motion_model.py:141-163), RK4.
    ``sincos`` defaults to libm; the device plant kernel uses the library's portable routine
    (``_lib.sincos_host``), which can be passed here to reproduce it bit for bit."""
    def d(s):
        sn, cs = sincos(float(s[2]))
        return ts * np.array([action[0] * cs, action[0] * sn, action[1]])
    k1 = d(state)
    k2 = d(state + 0.5 * k1)
    k3 = d(state + 0.5 * k2)
    k4 = d(state + k3)
    return state + (1 / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


@dataclass
class Pedestrian:
    position: np.ndarray                 # (2,)
    mode_velocities: np.ndarray          # (M, 2): predicted modes; mode 0 is what it really does
    size: float = 0.2                    # HUMAN_SIZE (main_base.py:187-188)


@dataclass
class Episode:
    state: np.ndarray                    # (3,) robot x, y, theta
    ref_path: List[tuple]                # waypoints; the last one is the goal
    polygons: List[np.ndarray] = field(default_factory=list)   # inflated static obstacles, 4 vertices each
    pedestrians: List[Pedestrian] = field(default_factory=list)
    # filled by ClosedLoopBatch
    ref_traj: list = None
    idx_ref: int = 0
    last_u: np.ndarray = None
    done: bool = False
    states: list = None
    actions: list = None
    costs: list = None
    statuses: list = None


class ClosedLoopBatch:
    def __init__(self, dims: Dims, cfg: MpcConfig, episodes: Sequence[Episode],
                 solve: Callable[[np.ndarray], tuple], sincos=_libm_sincos):
        self.dims, self.cfg, self.eps, self.solve = dims, cfg, list(episodes), solve
        self.sincos = sincos
        rb = cfg.robot()
        self.ts = rb.ts
        self.base_speed = rb.lin_vel_max * 0.8            # 'work' mode (trajectory_tracker.py:142-143)
        self.lin_vel_max = rb.lin_vel_max
        for e in self.eps:
            e.state = np.asarray(e.state, dtype=np.float64)
            e.ref_traj = packing.ref_traj_from_path(self.ts, e.ref_path, e.state, self.base_speed)
            e.idx_ref, e.last_u, e.done = 0, np.zeros(2), False
            e.states, e.actions, e.costs, e.statuses = [e.state.copy()], [], [], []
            e._halfspaces = [packing.polygon_halfspaces(np.asarray(p)) for p in e.polygons]

    # -- one parameter row, the way the reference's host code builds it
    def _row(self, e: Episode) -> List[float]:
        d, N = self.dims, self.dims.N
        # the reference passes N_hor into `action_steps` (trajectory_tracker.py:187): window [idx-N, idx+5N)
        ref_states, e.idx_ref = packing.ref_states_window(e.idx_ref, e.ref_traj, e.state, N, N)
        goal = e.ref_path[-1]
        gx, gy = e.state[0] - goal[0], e.state[1] - goal[1]
        dist_to_goal = math.sqrt(gx * gx + gy * gy)
        if dist_to_goal >= self.base_speed * N * self.ts:
            speed_ref = self.base_speed
        else:                                              # the reference's max() (SURVEY C-8)
            speed_ref = max(dist_to_goal / N / self.ts, self.lin_vel_max)
        # dynamic obstacles: every (pedestrian, mode) is one obstacle; offset t = position + t*ts*v
        obs = []
        for p in e.pedestrians:
            for v in p.mode_velocities:
                slots = [[p.position[0], p.position[1], p.size, p.size, 0, 1]]
                for t in range(1, N + 1):
                    c = p.position + t * self.ts * v
                    r = p.size + 0.03 * t                  # growing std of the prediction
                    slots.append([c[0], c[1], r, r, 0, 1])
                obs.append(slots)
        obs = obs[: d.Ndyn]
        # static obstacles: the Nstcobs polygons closest by EDGE distance (mpc_interface.py:90-100,
        # utils_geo.py:6-33), nearest first
        stc = [0.0] * (d.Nstc * 3 * d.nedge)
        if e.polygons:
            for slot, i in enumerate(packing.closest_polygons(e.state, e.polygons, d.Nstc)):
                b, a0, a1 = e._halfspaces[i]
                stc[slot * 3 * d.nedge:(slot + 1) * 3 * d.nedge] = b + a0 + a1
        return packing.assemble_params(d, self.cfg, e.state, ref_states, speed_ref, last_u=e.last_u,
                                       stc_constraints=stc,
                                       dyn_constraints=packing.pack_dyn_constraints(d, obs))

    def step(self) -> int:
        """Advance every unfinished episode by one timestep; returns how many are still running."""
        live = [e for e in self.eps if not e.done]
        if not live:
            return 0
        P = np.array([self._row(e) for e in live], dtype=np.float64)
        u, cost, status = self.solve(P)
        for e, ue, ce, se in zip(live, np.asarray(u), np.asarray(cost), np.asarray(status)):
            a = np.array(ue[:2], dtype=np.float64)         # action_steps = 1
            e.last_u = a                                    # past_actions keeps the solver's action (:329)
            if a[0] < 0:                                    # no-backward rule of the sim loop (main_base.py:320-321)
                a = np.zeros(2)
            e.state = unicycle_rk4(e.state, a, self.ts, self.sincos)
            e.states.append(e.state.copy()); e.actions.append(a); e.costs.append(float(ce)); e.statuses.append(int(se))
            for p in e.pedestrians:
                p.position = p.position + self.ts * p.mode_velocities[0]
            goal = e.ref_path[-1]
            # check_termination_condition (trajectory_tracker.py:191-199)
            if abs(e.state[0] - goal[0]) <= 0.5 and abs(e.state[1] - goal[1]) <= 0.5 and abs(a[0]) < 0.4:
                e.done = True
        return sum(not e.done for e in self.eps)

    def run(self, max_steps: int) -> None:
        for _ in range(max_steps):
            if self.step() == 0:
                break


# ---------------------------------------------------------------------------------------------
# Episode metrics of the evaluation loop (main_eva.py: main_base.py:327-336,426-434; main_pre.py:20-53)
def point_polygon_distance(pt, poly) -> float:
    """Distance from a point to a convex polygon, 0 inside (shapely's ``Polygon.distance(Point)``)."""
    V = np.asarray(poly, dtype=np.float64)
    m = V.shape[0]
    sg, inside, best = 0, True, math.inf
    for i in range(m):
        x0, y0 = V[i]
        x1, y1 = V[(i + 1) % m]
        dx, dy = x1 - x0, y1 - y0
        cr = dx * (pt[1] - y0) - dy * (pt[0] - x0)
        if cr == 0:
            inside = False
        elif sg == 0:
            sg = 1 if cr > 0 else -1
        elif (cr > 0) != (sg > 0):
            inside = False
        t = min(1.0, max(0.0, ((pt[0] - x0) * dx + (pt[1] - y0) * dy) / (dx * dx + dy * dy)))
        best = min(best, math.hypot(x0 + t * dx - pt[0], y0 + t * dy - pt[1]))
    return 0.0 if inside else best


def check_collision(state, polygons, pedestrians, human_size: float = 0.2) -> bool:
    """``check_collision`` (main_pre.py:20-31): strictly inside an (inflated) polygon, or within a
    pedestrian's radius."""
    if any(point_polygon_distance(state, p) == 0.0 and _strictly_inside(state, p) for p in polygons):
        return True
    return any(math.hypot(state[0] - q[0], state[1] - q[1]) <= human_size for q in pedestrians)


def _strictly_inside(pt, poly) -> bool:
    V = np.asarray(poly, dtype=np.float64)
    m = V.shape[0]
    sg = 0
    for i in range(m):
        x0, y0 = V[i]
        x1, y1 = V[(i + 1) % m]
        cr = (x1 - x0) * (pt[1] - y0) - (y1 - y0) * (pt[0] - x0)
        if cr == 0:
            return False
        if sg == 0:
            sg = 1 if cr > 0 else -1
        elif (cr > 0) != (sg > 0):
            return False
    return True


def episode_metrics(actions, trajectory, ref_traj, polygons, dyn_clearances):
    """The four evaluation metrics of a finished episode (main_base.py:426-434):
    smoothness = mean |second difference| of the speeds / angular speeds (main_pre.py:33-36),
    clearance = smallest distance of the driven positions to the inflated polygons (:38-42),
    deviation = [mean, max] distance of the driven positions to the reference trajectory (:48-52),
    dynamic clearance = smallest per-step distance to a pedestrian (:44-46, main_base.py:329)."""
    import statistics
    a = np.asarray(actions, dtype=np.float64)
    smooth = [statistics.mean(np.abs(np.diff(a[:, 0], n=2))), statistics.mean(np.abs(np.diff(a[:, 1], n=2)))]
    clearance = min(min(point_polygon_distance(pos, ob) for ob in polygons) for pos in trajectory)
    dev = [min(math.hypot(r[0] - pos[0], r[1] - pos[1]) for r in ref_traj) for pos in trajectory]
    return dict(smoothness=smooth, clearance=clearance, deviation=[statistics.mean(dev), max(dev)],
                clearance_dyn=min(dyn_clearances))


def make_episodes(n: int, seed: int, n_ped: int = 2, n_modes: int = 3) -> List[Episode]:
    """Small synthetic warehouse episodes: an L-shaped route, two rectangles beside it,
    pedestrians crossing ahead."""
    rng = np.random.Generator(np.random.PCG64(seed))
    eps = []
    for _ in range(n):
        x0, y0 = rng.uniform(3, 8), rng.uniform(3, 8)
        l1, l2 = rng.uniform(4, 7), rng.uniform(3, 6)
        path = [(x0 + l1, y0), (x0 + l1, y0 + l2)]
        state = np.array([x0, y0 + rng.normal(0, 0.1), rng.normal(0, 0.15)])
        polys = []
        for cx, cy in ((x0 + 0.5 * l1, y0 - 2.2), (x0 + l1 + 2.4, y0 + 0.5 * l2)):
            w, h = rng.uniform(1.5, 3), rng.uniform(1.5, 3)
            polys.append(np.array([[cx - w / 2, cy - h / 2], [cx + w / 2, cy - h / 2],
                                   [cx + w / 2, cy + h / 2], [cx - w / 2, cy + h / 2]]))
        peds = []
        for _ in range(n_ped):
            pos = np.array([x0 + rng.uniform(2.5, l1), y0 + rng.uniform(-2.5, 2.5)])
            hd = rng.uniform(-np.pi, np.pi)
            sp = rng.uniform(0.4, 1.0)
            modes = np.array([[sp * math.cos(hd + dh), sp * math.sin(hd + dh)]
                              for dh in np.linspace(0, 0.5, n_modes)])
            peds.append(Pedestrian(pos, modes))
        eps.append(Episode(state, path, polys, peds))
    return eps
