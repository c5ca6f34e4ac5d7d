"""Host-side packers that produce the parameter vector ``p`` in the reference layout.

These mirror, for the solve path only, what the reference does on the host
before every ``solver.run(p)``:

* ``assemble_params``        <- ``TrajectoryTracker.run_step`` (trajectory_tracker.py:291-317)
* ``pack_dyn_constraints``   <- ``MpcInterface.get_dyn_constraints`` (mpc_interface.py:82-88)
  fed by ``MainBase.run_one_step`` (main_base.py:293-302)
* ``pack_stc_constraints``   <- ``MpcInterface.get_stc_constraints`` (mpc_interface.py:73-80)
* ``polygon_halfspaces``     <- ``utils_geo.polygon_halfspace_representation`` (utils_geo.py:35-62)
* ``ref_traj_from_path`` / ``ref_states_window`` <- ``TrajectoryTracker.get_ref_traj`` /
  ``get_ref_states`` (trajectory_tracker.py:202-270)
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .problem import Dims, MpcConfig, NDYNPAR, NS


def polygon_halfspaces(vertices: np.ndarray) -> Tuple[List[float], List[float], List[float]]:
    """H-representation ``b - a0*x - a1*y > 0`` of a convex polygon.

    Each edge is scaled so that ``b - a.x`` is 1 at the vertex centroid and 0 on
    the edge (the normalisation of utils_geo.py:44-62).  Vertices may be in
    either orientation; they must be in boundary order.
    """
    V = np.asarray(vertices, dtype=np.float64)
    if V.ndim != 2 or V.shape[1] != 2 or V.shape[0] < 3:
        raise ValueError("polygon must be [m>=3, 2]")
    c = V.mean(axis=0)
    b, a0, a1 = [], [], []
    m = V.shape[0]
    for i in range(m):
        p0, p1 = V[i] - c, V[(i + 1) % m] - c
        # solve [p0; p1] a = [1; 1]
        det = p0[0] * p1[1] - p0[1] * p1[0]
        if abs(det) < 1e-14:
            continue  # degenerate edge through the centroid (rank deficient: skipped upstream too)
        ax = (p1[1] - p0[1]) / det
        ay = (p0[0] - p1[0]) / det
        a0.append(float(ax))
        a1.append(float(ay))
        b.append(float(ax * c[0] + ay * c[1] + 1.0))
    return b, a0, a1


def pack_stc_constraints(dims: Dims, polygons: Sequence[np.ndarray]) -> List[float]:
    """``o_s`` block: per polygon ``[b(nedge), a0(nedge), a1(nedge)]``, zero rows when unused."""
    ne = dims.nedge
    out = [0.0] * (dims.Nstc * 3 * ne)
    if len(polygons) > dims.Nstc:
        raise ValueError(f"at most Nstcobs={dims.Nstc} polygons")
    for i, poly in enumerate(polygons):
        b, a0, a1 = polygon_halfspaces(np.asarray(poly))
        if len(b) != ne:
            # the reference silently corrupts o_s here (SURVEY C-12); refuse instead
            raise ValueError(f"polygon {i} has {len(b)} edges, the solver was built for {ne}")
        out[i * 3 * ne:(i + 1) * 3 * ne] = b + a0 + a1
    return out


def pack_dyn_constraints(dims: Dims, full_dyn_obstacle_list=None) -> List[float]:
    """``o_d`` block: obstacle-major, then time 0..N, then (x, y, rx, ry, angle, alpha)."""
    per = (dims.N + 1) * NDYNPAR
    out = [0.0] * (dims.Ndyn * per)
    if full_dyn_obstacle_list is not None:
        if len(full_dyn_obstacle_list) > dims.Ndyn:
            raise ValueError(f"at most Ndynobs={dims.Ndyn} dynamic obstacles")
        for i, obs in enumerate(full_dyn_obstacle_list):
            flat = [float(v) for slot in obs for v in slot]
            if len(flat) != per:
                raise ValueError(f"obstacle {i}: expected {(dims.N + 1)} slots of {NDYNPAR}")
            out[i * per:(i + 1) * per] = flat
    return out


def dyn_obstacles_from_predictions(dims: Dims, mu_list_list, std_list_list):
    """(mu, std) per time offset -> per-obstacle slot lists, as main_base.py:293-302.

    Unfilled slots of a used obstacle are ``[0,0,0,0,0,1]`` (the reference's quirk,
    SURVEY C-2).
    """
    n_obs = max((len(m) for m in mu_list_list), default=0)
    out = [[[0, 0, 0, 0, 0, 1] for _ in range(dims.N + 1)] for _ in range(n_obs)]
    for t, (mus, stds) in enumerate(zip(mu_list_list, std_list_list)):
        for n, (mu, std) in enumerate(zip(mus, stds)):
            out[n][t] = [mu[0], mu[1], std[0], std[1], 0, 1]
    return out


def ref_traj_from_path(ts: float, ref_path: Sequence[Tuple[float, float]], state, speed: float):
    """Sample the reference path at ``speed*ts`` spacing starting from ``state``."""
    x, y = float(state[0]), float(state[1])
    idx = 0
    nx, ny = ref_path[0][0], ref_path[0][1]
    traj = []
    travelling = True
    while travelling:
        appended_dir = None
        while True:
            d = math.hypot(nx - x, ny - y)
            if d < 1e-9:
                idx += 1
                if idx > len(ref_path) - 1:
                    travelling = False
                    break
                nx, ny = ref_path[idx][0], ref_path[idx][1]
                break
            dx, dy = (nx - x) / d, (ny - y) / d
            appended_dir = (dx, dy)
            eta = d / speed
            if eta > ts:
                x, y = x + dx * speed * ts, y + dy * speed * ts
                break
            x, y = x + dx * speed * eta, y + dy * speed * eta
            idx += 1
            if idx > len(ref_path) - 1:
                travelling = False
                break
            nx, ny = ref_path[idx][0], ref_path[idx][1]
        if appended_dir is not None and not d < 1e-9:
            traj.append((x, y, math.atan2(appended_dir[1], appended_dir[0])))
    return traj


def ref_states_window(idx_ref_traj: int, ref_traj, state, action_steps: int = 1,
                      horizon: int = 20):
    """Next ``horizon`` reference states from the closest trajectory sample, padded with the last
    (``TrajectoryTracker.get_ref_states``, trajectory_tracker.py:243-270).

    The closest sample is searched in ``[idx - action_steps, idx + 5*action_steps)``.  NOTE the
    reference's own call (``set_ref_states``, :187) passes ``N_hor`` POSITIONALLY into
    ``action_steps`` and leaves ``horizon`` at its default 20: the effective window is
    ``[idx - N_hor, idx + 5*N_hor)``.  Callers that mirror the reference (closed_loop.py, the device
    packer K4) therefore pass ``action_steps=N_hor``."""
    lb = max(0, idx_ref_traj - action_steps)
    ub = min(len(ref_traj), idx_ref_traj + 5 * action_steps)
    # sqrt(dx*dx + dy*dy) instead of the reference's math.hypot (trajectory_tracker.py:257): the
    # same selection up to last-bit ties, and reproducible bit for bit by the device packer
    dists = [math.sqrt((state[0] - r[0]) * (state[0] - r[0]) + (state[1] - r[1]) * (state[1] - r[1]))
             for r in ref_traj[lb:ub]]
    idx = dists.index(min(dists)) + lb
    win = list(ref_traj[idx:idx + horizon])
    while len(win) < horizon:
        win.append(ref_traj[-1])
    return np.asarray(win, dtype=np.float64), idx


def point_to_polygon_distance(px: float, py: float, poly) -> float:
    """Smallest distance from a point to the EDGES of a polygon, as ``utils_geo.lineseg_dists``
    computes it for ``MpcInterface.get_closest_n_stc_obstacles`` (utils_geo.py:6-33,
    mpc_interface.py:90-100): along-edge overshoot ``h = max(s, t, 0)`` and perpendicular offset
    ``c`` of the unit tangent, ``d = |(h, c)|``.  (sqrt of the sum of squares where the reference
    uses np.hypot: the device packer reproduces it bit for bit.)"""
    V = np.asarray(poly, dtype=np.float64)
    m = V.shape[0]
    best = math.inf
    for i in range(m):
        ax, ay = float(V[i, 0]), float(V[i, 1])
        bx, by = float(V[(i + 1) % m, 0]), float(V[(i + 1) % m, 1])
        ex, ey = bx - ax, by - ay
        ln = math.sqrt(ex * ex + ey * ey)
        dx, dy = ex / ln, ey / ln
        s_ = (ax - px) * dx + (ay - py) * dy
        t_ = (px - bx) * dx + (py - by) * dy
        h = max(s_, t_, 0.0)
        c = (px - ax) * dy - (py - ay) * dx
        d = math.sqrt(h * h + c * c)
        best = d if d < best else best
    return best


def closest_polygons(state, polygons, n_keep: int) -> List[int]:
    """Indices of the ``n_keep`` polygons closest (by edge distance) to the robot, nearest first.
    The reference keeps the same SET (np.argpartition, mpc_interface.py:97) in numpy's unspecified
    partition order; the order only permutes the polygon slots of ``o_s`` (a sum over polygons)."""
    d = [point_to_polygon_distance(float(state[0]), float(state[1]), p) for p in polygons]
    order = sorted(range(len(polygons)), key=lambda i: (d[i], i))
    return order[:n_keep]


def assemble_params(dims: Dims, cfg: MpcConfig, state, ref_states: np.ndarray,
                    speed_ref: float, last_u=None, other_robot_states=None,
                    stc_constraints=None, dyn_constraints=None,
                    stc_weights=10.0, dyn_weights=10.0, tuning_params=None) -> List[float]:
    """Concatenate the blocks of ``p`` in the order of trajectory_tracker.py:315-317."""
    N = dims.N
    ref_states = np.asarray(ref_states, dtype=np.float64)
    if ref_states.shape != (N, NS):
        raise ValueError(f"ref_states must be [{N},{NS}]")
    if stc_constraints is None:
        stc_constraints = [0.0] * (dims.Nstc * 3 * dims.nedge)
    if dyn_constraints is None:
        dyn_constraints = [0.0] * (dims.Ndyn * NDYNPAR * (N + 1))
    if other_robot_states is None:
        other_robot_states = [0.0] * (NS * (N + 1) * dims.Nother)
    if last_u is None:
        last_u = [0.0, 0.0]
    tp = list(cfg.tuning_params()) if tuning_params is None else list(tuning_params)
    sw = [float(stc_weights)] * N if np.isscalar(stc_weights) else list(stc_weights)
    dw = [float(dyn_weights)] * N if np.isscalar(dyn_weights) else list(dyn_weights)
    params = (list(map(float, last_u)) + list(map(float, state)) + ref_states[-1, :].tolist()
              + [float(t) for t in tp] + ref_states.reshape(-1).tolist()
              + [float(speed_ref)] * N + list(map(float, other_robot_states))
              + list(map(float, stc_constraints)) + list(map(float, dyn_constraints)) + sw + dw)
    if len(params) != dims.np:
        raise ValueError(f"assembled {len(params)} parameters, layout needs {dims.np}")
    return params


# ---------------------------------------------------------------------------------------------
# SWTA hypotheses -> ellipses: MainBase.run_wta_prediction (main_base.py:175-208)
# clusters the predictor's position hypotheses of every time offset with
# DBSCAN(eps=1, min_samples=2) (utils_test.py:133-143) and fits (mean, 2*std) per cluster
# (utils_test.py:145-151); run_one_step turns them into obstacle slots (main_base.py:293-302).
def dbscan_labels(points: np.ndarray, eps: float = 1.0, min_samples: int = 2) -> np.ndarray:
    """Labels as sklearn.cluster.DBSCAN assigns them (clusters numbered by their first core
    point in index order, noise = -1), restated without sklearn."""
    X = np.asarray(points, dtype=np.float64)
    n = X.shape[0]
    d2 = (X[:, None, 0] - X[None, :, 0]) * (X[:, None, 0] - X[None, :, 0]) + \
         (X[:, None, 1] - X[None, :, 1]) * (X[:, None, 1] - X[None, :, 1])
    nb = d2 <= eps * eps
    core = nb.sum(1) >= min_samples
    labels = -np.ones(n, dtype=np.int64)
    c = 0
    for i in range(n):
        if labels[i] != -1 or not core[i]:
            continue
        labels[i] = c
        stack = [i]
        while stack:
            j = stack.pop()
            if core[j]:
                for v in np.nonzero(nb[j])[0]:
                    if labels[v] == -1:
                        labels[v] = c
                        stack.append(v)
        c += 1
    return labels


def hypotheses_to_obstacles(dims: Dims, current_positions, hypotheses, human_size: float = 0.2,
                            eps: float = 1.0, min_samples: int = 2, enlarge: float = 2.0):
    """``o_d`` block [Ndyn, N+1, 6] from the current pedestrian positions and per-offset hypotheses
    ``hypotheses[t]`` = array [K_t, 2] (t = 0..N-1 for offsets 1..N)."""
    N = dims.N
    mu_ll = [[tuple(p) for p in current_positions]]
    sd_ll = [[(human_size, human_size) for _ in current_positions]]
    for t in range(N):
        H = np.asarray(hypotheses[t], dtype=np.float64)
        lab = dbscan_labels(H, eps, min_samples)
        mus, sds = [], []
        for c in range(int(lab.max()) + 1 if lab.size else 0):
            pts = H[lab == c]
            n = pts.shape[0]
            m0 = m1 = 0.0
            for q in pts:                                   # np.mean(axis=0): rows added in order
                m0 += q[0]; m1 += q[1]
            m0 /= n; m1 /= n
            v0 = v1 = 0.0
            for q in pts:                                   # np.std: sqrt(sum((x-mean)^2)/n)
                v0 += (q[0] - m0) * (q[0] - m0); v1 += (q[1] - m1) * (q[1] - m1)
            mus.append((m0, m1))
            sds.append((math.sqrt(v0 / n) * enlarge + 0.0, math.sqrt(v1 / n) * enlarge + 0.0))
        mu_ll.append(mus)
        sd_ll.append(sds)
    obs = dyn_obstacles_from_predictions(dims, mu_ll, sd_ll)[: dims.Ndyn]
    out = np.zeros((dims.Ndyn, N + 1, NDYNPAR))
    for i, o in enumerate(obs):
        out[i] = np.asarray(o, dtype=np.float64)
    return out
