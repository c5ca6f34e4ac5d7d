"""The host side of the solve path against fixtures produced by the REFERENCE'S OWN per-timestep code
(tests/golden/gen_host_golden.py: unmodified MpcInterface.run_step + TrajectoryTracker.run_step on the
reference's warehouse map, scenario 0, constant-velocity predictor — BASELINE configs[0]; the
generated mpc_solver/navi_fast module loaded by the reference's own __import_solver).

CPU tests pin the host mirrors (packing.py, closed_loop.py): SURVEY 8 f-1 (static/dynamic obstacle
packing), f-2 (plant step, no-backward rule), f-4 (reference trajectory and window).  GPU tests pin the
device packer K4, the plant K5 and the CUDA solver on the same recorded steps: every recorded parameter
vector must give the recorded solution back bit for bit, so the recorded closed loop IS the loop the
reference drives with the CUDA module."""
import math
import os

import numpy as np
import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, packing
from dyobav_mpcnwta_warehouse_b200.closed_loop import unicycle_rk4
from dyobav_mpcnwta_warehouse_b200.problem import MpcConfig

HERE = os.path.dirname(os.path.abspath(__file__))


class _Episode(dict):
    """One recorded episode: its own arrays plus the map shared by all of them."""


@pytest.fixture(scope="module")
def all_episodes():
    z = np.load(os.path.join(HERE, "golden", "host_cases.npz"))
    eps = []
    for name, scn, mode in zip(z["episodes"], z["episode_scenarios"], z["episode_modes"]):
        e = _Episode({k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(str(name) + "/")})
        for k in ("map_obstacles", "map_halfspaces", "map_obstacles_px"):
            e[k] = z[k]
        e["scenario"], e["mode"], e["name"] = int(scn), str(mode), str(name)
        eps.append(e)
    return eps


@pytest.fixture(scope="module", params=[0, 1, 2, 3], ids=["main_s0", "eva_s0", "eva_s1", "eva_s2"])
def host(request, all_episodes):
    return all_episodes[request.param]


def _blocks(dims, p):
    return {k: p[o:o + n] for k, (o, n) in dims.layout().items()}


def _rows12(os_block):
    """o_s as a set of per-polygon records, each a sorted tuple of its (b, a0, a1) edge rows."""
    out = []
    for q in np.asarray(os_block).reshape(-1, 3, 4):
        rows = sorted((round(float(q[0, e]), 9), round(float(q[1, e]), 9), round(float(q[2, e]), 9)) for e in range(4))
        out.append(tuple(rows))
    return sorted(out)


def test_fixture_is_the_reference_scenario(all_episodes):
    e0 = all_episodes[0]
    assert e0["scenario"] == 0 and e0["mode"] == "main"                   # BASELINE configs[0]
    assert e0["p"].shape[1] == Dims().np == 2778 and e0["u"].shape[1] == 40
    assert e0["p"].shape[0] >= 40 and bool(e0["terminated"])              # the robot reaches its goal
    assert e0["map_obstacles"].shape[0] > 10                              # more polygons than Nstcobs
    assert set(np.unique(e0["status"])) == {0, 1}                         # blocked stretch in the middle
    assert [e["scenario"] for e in all_episodes] == [0, 0, 1, 2]          # main.py + main_eva.py episodes
    assert sum(e["p"].shape[0] for e in all_episodes) >= 180


def test_reference_trajectory_generator(host):
    """packing.ref_traj_from_path == TrajectoryTracker.get_ref_traj (trajectory_tracker.py:202-240)."""
    traj = packing.ref_traj_from_path(0.2, [tuple(x) for x in host["ref_path"]], host["start_state"],
                                      float(host["base_speed"]))
    np.testing.assert_array_equal(np.asarray(traj), host["ref_traj"])


def test_reference_window_speed_and_header_blocks(host):
    """r_s / s_N / r_v / u_m1 / s_0 / q / weights of every recorded step (trajectory_tracker.py:285-317),
    incl. the [idx - N_hor, idx + 5 N_hor) search window the reference's call produces (:187)."""
    d, cfg = Dims(), MpcConfig()
    ref_traj = [tuple(x) for x in host["ref_traj"]]
    goal = host["ref_path"][-1]
    idx = 0
    last_u = np.zeros(2)
    for t in range(host["p"].shape[0]):
        B = _blocks(d, host["p"][t])
        state = host["state"][t]
        win, idx = packing.ref_states_window(idx, ref_traj, state, d.N, d.N)
        assert idx == int(host["idx_ref"][t]), t
        np.testing.assert_array_equal(win.reshape(-1), B["r_s"])
        np.testing.assert_array_equal(win[-1], B["s_N"])
        dist = math.hypot(state[0] - goal[0], state[1] - goal[1])
        base = float(host["base_speed"])
        sr = base if dist >= base * d.N * 0.2 else max(dist / d.N / 0.2, 1.5)
        np.testing.assert_array_equal(B["r_v"], np.full(d.N, sr))
        np.testing.assert_array_equal(B["u_m1"], last_u)
        np.testing.assert_array_equal(B["s_0"], state)
        np.testing.assert_array_equal(B["q"], np.asarray(cfg.tuning_params(), dtype=np.float64))
        np.testing.assert_array_equal(B["q_stc"], np.full(d.N, 10.0))
        np.testing.assert_array_equal(B["q_dyn"], np.full(d.N, 10.0))
        assert not B["c_0"].any() and not B["c"].any()
        last_u = host["u"][t][:2]                      # past_actions keeps the solver's own action


def test_dynamic_obstacle_block(host):
    """(mu, std) lists -> slots (main_base.py:293-302) -> o_d (mpc_interface.py:82-88)."""
    d = Dims()
    for t in range(host["p"].shape[0]):
        mu_ll = [[tuple(m)] for m in host["mu"][t]]
        sd_ll = [[tuple(s)] for s in host["std"][t]]
        obs = packing.dyn_obstacles_from_predictions(d, mu_ll, sd_ll)
        np.testing.assert_array_equal(np.asarray(obs, dtype=np.float64), host["dyn_obs"][t])
        od = packing.pack_dyn_constraints(d, obs)
        np.testing.assert_array_equal(np.asarray(od), _blocks(d, host["p"][t])["o_d"])


def test_static_obstacle_block(host):
    """The Nstcobs polygons closest by EDGE distance (mpc_interface.py:90-100, utils_geo.py:6-33) as
    half-spaces (utils_geo.py:35-62).  The reference keeps the set in numpy's unspecified argpartition
    order and qhull's edge order, so polygons and their edge rows are compared as sets."""
    d = Dims()
    polys = [np.asarray(p) for p in host["map_obstacles"]]
    changed = 0
    prev = None
    for t in range(host["p"].shape[0]):
        sel = packing.closest_polygons(host["state"][t], polys, d.Nstc)
        # the same polygons the reference's interface returned as `closest_obstacle_list`
        want = sorted(tuple(np.round(p.reshape(-1), 9)) for p in host["closest"][t])
        got = sorted(tuple(np.round(polys[i].reshape(-1), 9)) for i in sel)
        assert got == want, t
        stc = packing.pack_stc_constraints(d, [polys[i] for i in sel])
        assert _rows12(stc) == _rows12(_blocks(d, host["p"][t])["o_s"]), t
        changed += prev is not None and set(sel) != prev
        prev = set(sel)
    assert changed >= 1            # the selection really changes along the route


def test_halfspaces_of_every_map_polygon(host):
    """packing.polygon_halfspaces == utils_geo.polygon_halfspace_representation on all map polygons."""
    for poly, hs in zip(host["map_obstacles"], host["map_halfspaces"]):
        b, a0, a1 = packing.polygon_halfspaces(poly)
        got = sorted(zip(np.round(b, 9), np.round(a0, 9), np.round(a1, 9)))
        want = sorted(zip(np.round(hs[0:4], 9), np.round(hs[4:8], 9), np.round(hs[8:12], 9)))
        assert got == want


def test_edge_distance_differs_from_vertex_distance_on_this_map(all_episodes):
    """Regression for the round-1 deviation (closest VERTEX): beside a long shelf the two rankings pick
    different polygons, and only the edge ranking is the reference's."""
    differs = 0
    for host in all_episodes:
        polys = [np.asarray(p) for p in host["map_obstacles"]]
        for t in range(host["p"].shape[0]):
            s = host["state"][t]
            by_vertex = sorted(range(len(polys)),
                               key=lambda i: (min(math.hypot(v[0] - s[0], v[1] - s[1]) for v in polys[i]), i))[:10]
            differs += set(by_vertex) != set(packing.closest_polygons(s, polys, 10))
    assert differs > 0


def test_plant_step_and_no_backward_rule(host):
    """Robot.one_step = RK4 unicycle (basic_agent.py:106, motion_model.py:141-163) with the sim loop's
    no-backward rule (main_base.py:320-321)."""
    for t in range(host["p"].shape[0]):
        a = host["u"][t][:2].copy()
        if a[0] < 0:
            a[:] = 0.0
        np.testing.assert_array_equal(a, host["action_applied"][t])
        nxt = unicycle_rk4(host["state"][t], a, 0.2)
        np.testing.assert_allclose(nxt, host["next_state"][t], rtol=0, atol=1e-14)   # libm vs numpy sin/cos
        if t + 1 < host["p"].shape[0]:
            np.testing.assert_array_equal(host["next_state"][t], host["state"][t + 1])


def test_episode_metrics_and_collision_flags(host):
    """The evaluation metrics main_eva.py prints (main_base.py:426-434, main_pre.py:20-53), computed by
    the reference's own functions on the recorded episode, against closed_loop.episode_metrics."""
    from dyobav_mpcnwta_warehouse_b200.closed_loop import episode_metrics, check_collision
    polys = [np.asarray(p) for p in host["map_obstacles"]]
    m = episode_metrics(host["u"][:, :2], host["robot_past_traj"], host["ref_traj"], polys, host["dyn_clearance"])
    np.testing.assert_allclose(m["smoothness"], host["metric_smoothness"], rtol=1e-13)
    np.testing.assert_allclose(m["clearance"], float(host["metric_clearance"]), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(m["deviation"], host["metric_deviation"], rtol=1e-13)
    assert m["clearance_dyn"] == float(host["metric_clearance_dyn"])
    # the per-step pedestrian clearance and the collision flag of the last step
    T = host["p"].shape[0]
    for t in range(T):
        # humans[t] was recorded BEFORE their step; the clearance of step t uses their next position
        if t + 1 < T:
            d = min(math.hypot(host["next_state"][t][0] - h[0], host["next_state"][t][1] - h[1]) for h in host["human"][t + 1])
            assert d == pytest.approx(float(host["dyn_clearance"][t]), rel=1e-13)
    if host["mode"] == "eva" and T > 1:
        assert bool(host["collision"]) != bool(host["terminated"])
        for t in range(T - 2):             # no collision before the last step of an evaluation episode
            assert not check_collision(host["next_state"][t], polys, host["human"][t + 1])


def test_recorded_solutions_are_the_laned_oracle(host):
    """The generator's solver was the laned oracle: re-solving a few recorded parameter vectors here
    reproduces the recorded solutions bit for bit (the GPU test below does all of them on the CUDA path)."""
    from oracle import oracle
    for t in (0, host["p"].shape[0] - 1):
        r = oracle.solve(Dims(), RobotSpec(), SolverSettings(), host["p"][t], laned=True)
        np.testing.assert_array_equal(r["u"], host["u"][t])
        assert r["exit_status"] == int(host["status"][t]) and r["n_inner"] == int(host["n_inner"][t])


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_solver_reproduces_the_recorded_closed_loop(host):
    """Every parameter vector the reference's TrajectoryTracker assembled, through the CUDA drop-in
    (`solver().run(p)`, the call at trajectory_tracker.py:362) and through the batched entry point:
    recorded solution, cost, status, iteration count — bit for bit."""
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import solver, BatchedSolver
    assert torch.cuda.is_available()
    T = host["p"].shape[0]
    bs = BatchedSolver(Dims(), RobotSpec(), SolverSettings())
    o = bs.run_batch(torch.as_tensor(host["p"], device="cuda").contiguous())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o["u"].cpu().numpy(), host["u"])
    np.testing.assert_array_equal(o["cost"].cpu().numpy(), host["cost"])
    np.testing.assert_array_equal(o["exit_status"].cpu().numpy(), host["status"].astype(np.int32))
    np.testing.assert_array_equal(o["n_inner"].cpu().numpy(), host["n_inner"].astype(np.int32))
    s = solver()
    for t in range(0, T, 6):
        sol = s.run(host["p"][t].tolist())
        assert sol.solution == host["u"][t].tolist()
        assert sol.exit_status == ("Converged", "NotConvergedIterations")[int(host["status"][t])]


@pytest.mark.gpu
def test_device_packer_and_plant_follow_the_reference_loop(host):
    """K4 (mpcb_pack_f64) and K5 (mpcb_plant_step_f64) driven along the recorded loop: the packer's
    parameter vector equals the reference's (o_s as a set of polygons, see test_static_obstacle_block),
    the plant's next state equals the reference robot's."""
    import ctypes
    import torch
    from dyobav_mpcnwta_warehouse_b200 import _lib
    assert torch.cuda.is_available()
    L = _lib.load()
    d, cfg = Dims(), MpcConfig()
    T = host["p"].shape[0]
    dev = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")  # noqa: E731
    ref = host["ref_traj"]
    polys = host["map_obstacles"]
    state = dev(host["state"][0][None])
    last_u = torch.zeros(1, 2, dtype=torch.float64, device="cuda")
    idx_ref = torch.zeros(1, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    ref_traj, ref_len = dev(ref[None]), dev([ref.shape[0]], torch.int32)
    goal = dev(host["ref_path"][-1][None])
    polys_d, n_poly = dev(polys[None]), dev([polys.shape[0]], torch.int32)
    od_in = torch.zeros(1, d.Ndyn, d.N + 1, 6, dtype=torch.float64, device="cuda")
    P = torch.empty(1, d.np, dtype=torch.float64, device="cuda")
    sim = _lib.CSim()
    sim.n, sim.T, sim.Kp, sim.Pd, sim.M = 1, ref.shape[0], polys.shape[0], 0, 1
    sim.base_speed, sim.lin_vel_max, sim.ped_size = float(host["base_speed"]), 1.5, 0.2
    sim.stc_w, sim.dyn_w, sim.ts = 10.0, 10.0, 0.2
    for i, v in enumerate(cfg.tuning_params()):
        sim.tuning[i] = float(v)
    for name, t in (("state", state), ("last_u", last_u), ("ref_traj", ref_traj), ("ref_len", ref_len),
                    ("idx_ref", idx_ref), ("goal", goal), ("polys", polys_d), ("n_poly", n_poly), ("done", done),
                    ("od_in", od_in)):
        setattr(sim, name, t.data_ptr())
    cd = d.to_c()
    lay = d.layout()
    o_s = slice(lay["o_s"][0], lay["o_s"][0] + lay["o_s"][1])
    for t in range(T):
        od_in.zero_()
        od_in[0, :host["dyn_obs"][t].shape[0]] = dev(host["dyn_obs"][t])
        _lib.check(L.mpcb_pack_f64(ctypes.byref(cd), ctypes.byref(sim), ctypes.c_void_p(P.data_ptr()), None), "pack")
        p = P[0].cpu().numpy()
        want = host["p"][t]
        mask = np.ones(d.np, bool)
        mask[o_s] = False
        np.testing.assert_array_equal(p[mask], want[mask], err_msg=f"step {t}")
        assert _rows12(p[o_s]) == _rows12(want[o_s]), t
        assert int(idx_ref.item()) == int(host["idx_ref"][t])
        # plant with the recorded solution (the solver itself is pinned by the test above)
        u = dev(host["u"][t][None])
        _lib.check(L.mpcb_plant_step_f64(ctypes.byref(cd), ctypes.byref(sim), ctypes.c_void_p(u.data_ptr()), None), "plant")
        torch.cuda.synchronize()
        np.testing.assert_allclose(state[0].cpu().numpy(), host["next_state"][t], rtol=0, atol=1e-14)
        np.testing.assert_array_equal(last_u[0].cpu().numpy(), host["u"][t][:2])
        state.copy_(dev(host["next_state"][t][None]))        # stay on the recorded states (1-ulp sin/cos)
    assert int(done.item()) == int(bool(host["terminated"]))  # at the goal iff the recorded run ended there
