"""Seeded synthetic warehouse-shaped MPC instances.

Every instance is one ``p`` row in the reference layout (mpc_builder.py:47-60)
built the way the reference's host code builds it: corridor reference path
sampled at ``base_speed*ts`` ('work' mode, trajectory_tracker.py:142-143,202-240),
static rectangles inflated and encoded as half-spaces (main_base.py:125,
utils_geo.py:35-62), pedestrians' multimodal predictions as ellipses
``[x, y, rx, ry, 0, 1]`` per obstacle per time offset (main_base.py:293-302),
other robots all zero (trajectory_tracker.py:295-296).

Vectorised over the batch with numpy so the 65 536-scenario workload is
generated in seconds; generation is outside every timed region.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .problem import Dims, MpcConfig, NDYNPAR

WORLD = (-15.0, 18.0, -15.0, 14.3)   # main_base.py:105-108 / global_setting_warehouse.yaml
BASE_SEED = 20231017


@dataclass(frozen=True)
class Workload:
    """A named benchmark configuration (BASELINE.json ``configs``)."""
    name: str
    dims: Dims
    n_scenarios: int
    starts: int
    pedestrians: int
    modes: int
    seed: int


def workload(name: str) -> Workload:
    if name == "warehouse_b4096_ndyn40":          # configs[2]
        return Workload(name, Dims(Ndyn=40), 4096, 1, 2, 20, BASE_SEED + 3)
    if name == "warehouse_b65536x8":              # configs[3] — the headline metric
        return Workload(name, Dims(), 65536, 8, 2, 3, BASE_SEED + 4)
    if name == "dense_crowd_n40":                 # configs[4]
        return Workload(name, Dims(N=40, Ndyn=160), 16384, 1, 32, 5, BASE_SEED + 5)
    if name == "single_latency":                  # configs[1]
        return Workload(name, Dims(), 1, 1, 2, 3, BASE_SEED + 2)
    raise KeyError(name)


def _polyline_points(start, headings, lengths, s):
    """Points at arc length ``s`` [B,K] along per-instance polylines of L legs (clamped at the end)."""
    B, Lg = lengths.shape
    cum = np.concatenate([np.zeros((B, 1)), np.cumsum(lengths, axis=1)], axis=1)      # [B,L+1]
    dirs = np.stack([np.cos(headings), np.sin(headings)], axis=-1)                     # [B,L,2]
    verts = np.concatenate([start[:, None, :],
                            start[:, None, :] + np.cumsum(dirs * lengths[..., None], axis=1)], axis=1)
    s = np.minimum(s, cum[:, -1:])
    leg = np.clip((s[:, :, None] >= cum[:, None, 1:]).sum(-1), 0, Lg - 1)              # [B,K]
    bi = np.arange(B)[:, None]
    base = verts[bi, leg]
    off = s - cum[bi, leg]
    pts = base + dirs[bi, leg] * off[..., None]
    return pts, headings[bi, leg], verts


def generate(dims: Dims, n: int, seed: int, cfg: MpcConfig | None = None,
             pedestrians: int = 2, modes: int = 3, origin_fraction: float = 0.01):
    """Return ``P`` [n, np] float64 — one reference-layout parameter row per scenario."""
    cfg = cfg or MpcConfig()
    rb = cfg.robot()
    rng = np.random.Generator(np.random.PCG64(seed))
    N, ts = dims.N, rb.ts
    lay = dims.layout()
    P = np.zeros((n, dims.np))
    base_speed = rb.lin_vel_max * 0.8                      # 'work' mode
    step = base_speed * ts

    # ---- corridor path: 2-4 axis-aligned legs --------------------------------
    Lg = 4
    n_legs = rng.integers(2, 5, size=n)
    lengths = rng.uniform(3.0, 12.0, size=(n, Lg))
    lengths[np.arange(Lg)[None, :] >= n_legs[:, None]] = 0.0
    h0 = rng.integers(0, 4, size=n) * (np.pi / 2)
    turns = rng.choice([-1.0, 1.0], size=(n, Lg - 1)) * (np.pi / 2)
    headings = np.concatenate([h0[:, None], h0[:, None] + np.cumsum(turns, axis=1)], axis=1)
    start = np.stack([rng.uniform(WORLD[0] + 3, WORLD[1] - 3, size=n),
                      rng.uniform(WORLD[2] + 3, WORLD[3] - 3, size=n)], axis=-1)
    # keep geometry >= 1.5 m from the world origin (phantom obstacles there, SURVEY C-2)
    # except for a deliberate sub-population that starts right next to it
    near = np.hypot(start[:, 0], start[:, 1]) < 4.0
    start[near] += 6.0 * np.sign(start[near] + 1e-9)
    crossing = rng.uniform(size=n) < origin_fraction
    start[crossing] = rng.uniform(-0.6, 0.6, size=(int(crossing.sum()), 2))

    s0 = rng.uniform(0.0, 1.0, size=n)
    lat = rng.normal(0.0, 0.15, size=n)
    dpsi = rng.normal(0.0, 0.2, size=n)
    p0, hd0, _ = _polyline_points(start, headings, lengths, s0[:, None])
    nrm0 = np.stack([-np.sin(hd0[:, 0]), np.cos(hd0[:, 0])], axis=-1)
    state = np.concatenate([p0[:, 0] + nrm0 * lat[:, None], (hd0[:, 0] + dpsi)[:, None]], axis=1)

    s_ref = s0[:, None] + step * (np.arange(N)[None, :] + 1)
    ref_pts, ref_hd, _ = _polyline_points(start, headings, lengths, s_ref)
    r_s = np.concatenate([ref_pts, ref_hd[..., None]], axis=-1)                 # [n,N,3]

    o, ln = lay["u_m1"]
    P[:, o] = rng.uniform(0.0, 1.2, size=n)
    P[:, o + 1] = rng.uniform(-0.3, 0.3, size=n)
    o, ln = lay["s_0"]
    P[:, o:o + ln] = state
    o, ln = lay["s_N"]
    P[:, o:o + ln] = r_s[:, -1, :]
    o, ln = lay["q"]
    P[:, o:o + ln] = np.asarray(cfg.tuning_params(), dtype=np.float64)[None, :]
    o, ln = lay["r_s"]
    P[:, o:o + ln] = r_s.reshape(n, -1)
    o, ln = lay["r_v"]
    P[:, o:o + ln] = base_speed
    o, ln = lay["q_stc"]
    P[:, o:o + ln] = 10.0
    o, ln = lay["q_dyn"]
    P[:, o:o + ln] = 10.0
    # c_0 / c stay zero: the reference's default other_robot_states

    # ---- static rectangles, inflated 0.7, >= 1.0 m lateral clearance ----------
    if dims.Nstc > 0:
        if dims.nedge != 4:
            raise ValueError("the synthetic generator emits rectangles (nedge=4)")
        K = dims.Nstc
        s_c = rng.uniform(0.0, 1.0, size=(n, K)) * (lengths.sum(1, keepdims=True) + 4.0) - 2.0
        s_c = np.maximum(s_c, 0.0)
        w = rng.uniform(1.0, 4.0, size=(n, K)) + 2 * 0.7
        h = rng.uniform(1.0, 8.0, size=(n, K)) + 2 * 0.7
        side = rng.choice([-1.0, 1.0], size=(n, K))
        gap = rng.uniform(1.0, 3.0, size=(n, K))
        cpt, chd, _ = _polyline_points(start, headings, lengths, s_c)
        nrm = np.stack([-np.sin(chd), np.cos(chd)], axis=-1)
        # rectangle is axis-aligned; put its centre so the nearest face is `gap` from the path line
        half_across = np.where(np.abs(np.cos(chd)) > 0.5, h / 2, w / 2)
        ctr = cpt + nrm * (side * (gap + half_across))[..., None]
        # half-space rows, each = 1 at the centre, 0 on the edge: a = n_e / half_extent
        b = np.empty((n, K, 4)); a0 = np.zeros((n, K, 4)); a1 = np.zeros((n, K, 4))
        a0[..., 0] = 2.0 / w;  a0[..., 1] = -2.0 / w
        a1[..., 2] = 2.0 / h;  a1[..., 3] = -2.0 / h
        b[...] = a0 * ctr[..., 0:1] + a1 * ctr[..., 1:2] + 1.0
        # a rectangle placed next to one leg may cover another leg after a turn: drop every
        # rectangle whose (0.3 m grown) interior contains the start or a reference sample —
        # the map's obstacles never cover the planned path (unused polygon slots are zero rows)
        chk = np.concatenate([state[:, None, :2], ref_pts], axis=1)                   # [n,N+1,2]
        inx = np.abs(chk[:, None, :, 0] - ctr[:, :, None, 0]) < (w / 2 + 0.3)[:, :, None]
        iny = np.abs(chk[:, None, :, 1] - ctr[:, :, None, 1]) < (h / 2 + 0.3)[:, :, None]
        covered = (inx & iny).any(-1)                                                # [n,K]
        b[covered] = 0.0; a0[covered] = 0.0; a1[covered] = 0.0
        o, ln = lay["o_s"]
        P[:, o:o + ln] = np.concatenate([b, a0, a1], axis=-1).reshape(n, -1)

    # ---- pedestrians x modes -> ellipses --------------------------------------
    n_used = min(dims.Ndyn, pedestrians * modes)
    if n_used > 0:
        ahead = rng.uniform(3.0, 8.0, size=(n, pedestrians))
        plat = rng.uniform(-2.0, 2.0, size=(n, pedestrians))
        ppt, phd, _ = _polyline_points(start, headings, lengths, s0[:, None] + ahead)
        pn = np.stack([-np.sin(phd), np.cos(phd)], axis=-1)
        ppos = ppt + pn * plat[..., None]                                        # [n,Pd,2]
        phead = rng.uniform(-np.pi, np.pi, size=(n, pedestrians))
        pspeed = rng.uniform(0.8, 1.5, size=(n, pedestrians))
        mh = phead[..., None] + rng.normal(0.0, 0.3, size=(n, pedestrians, modes))
        ms = pspeed[..., None] * rng.uniform(0.7, 1.3, size=(n, pedestrians, modes))
        vel = np.stack([np.cos(mh), np.sin(mh)], axis=-1) * ms[..., None]        # [n,Pd,M,2]
        t = np.arange(N + 1, dtype=np.float64)
        ctr = ppos[:, :, None, None, :] + vel[:, :, :, None, :] * (t * ts)[None, None, None, :, None]
        rad = np.where(t == 0, 0.2, 0.2 + 0.03 * t)
        slots = np.zeros((n, pedestrians, modes, N + 1, NDYNPAR))
        slots[..., 0:2] = ctr
        slots[..., 2] = rad
        slots[..., 3] = rad
        slots[..., 5] = 1.0
        slots = slots.reshape(n, pedestrians * modes, N + 1, NDYNPAR)[:, :n_used]
        o, ln = lay["o_d"]
        od = np.zeros((n, dims.Ndyn, N + 1, NDYNPAR))
        od[:, :n_used] = slots
        P[:, o:o + ln] = od.reshape(n, -1)
    return P


def multistart_guesses(dims: Dims, P: np.ndarray, starts: int, seed: int) -> np.ndarray:
    """``U0`` [n*starts, 2N]: #0 zeros (the reference's behaviour), #1 (r_v, 0) x N,
    #2.. constant (v, w) with v~U[0,1.5], w in +-{0.15, 0.3, 0.5}."""
    n = P.shape[0]
    N = dims.N
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    U0 = np.zeros((n, starts, N, 2))
    if starts > 1:
        o, _ = dims.layout()["r_v"]
        U0[:, 1, :, 0] = P[:, o:o + N]
    wset = np.array([0.15, -0.15, 0.3, -0.3, 0.5, -0.5])
    for s in range(2, starts):
        U0[:, s, :, 0] = rng.uniform(0.0, 1.5, size=(n, 1))
        U0[:, s, :, 1] = wset[(s - 2) % len(wset)]
    return U0.reshape(n * starts, 2 * N)
