"""Closed loop (north_star "closed-loop trajectories"): the receding-horizon
harness driven by the CPU oracles (CPU test) and by the GPU solver vs the laned oracle (GPU)."""
import copy
import os

import numpy as np
import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, MpcConfig, RobotSpec, SolverSettings
from dyobav_mpcnwta_warehouse_b200.closed_loop import ClosedLoopBatch, Episode, make_episodes, unicycle_rk4
from oracle import oracle

CFG = SolverSettings(max_inner=150, max_outer=4)     # iteration budget stands in for the 0.1 s cap


def _oracle_solve(laned):
    def solve(P):
        U, SC = oracle.solve_batch(Dims(), RobotSpec(), CFG, P, None, threads=os.cpu_count(), laned=laned)
        return U, SC[:, 0], SC[:, 9].astype(np.int32)
    return solve


def test_unicycle_rk4_matches_oracle_rollout():
    import ctypes
    s = np.array([0.3, -0.2, 0.4])
    out = (ctypes.c_double * 3)()
    oracle.lib().mpco_unicycle_rk4((ctypes.c_double * 3)(*s), 1.1, 0.35, 0.2, out)
    np.testing.assert_allclose(unicycle_rk4(s, np.array([1.1, 0.35]), 0.2), list(out), rtol=0, atol=1e-15)


def test_robot_reaches_goal_on_free_straight_path():
    ep = Episode(np.array([0.0, 5.0, 0.0]) + np.array([3.0, 0.0, 0.0]), [(9.0, 5.0)])
    sim = ClosedLoopBatch(Dims(), MpcConfig(), [ep], _oracle_solve(False))
    sim.run(60)
    assert ep.done, (ep.states[-1], len(ep.states))
    xs = np.array(ep.states)
    assert np.abs(xs[:, 1] - 5.0).max() < 0.05          # stays on the line
    v = np.array(ep.actions)[:, 0]
    assert (np.diff(np.concatenate([[0.0], v])) <= 1.0 * 0.2 + 0.1).all()    # acceleration bound: soft (ALM, iteration budget)
    assert v.max() <= 1.5 + 1e-9 and v.max() > 1.0      # speeds up to about the reference speed


def test_closed_loop_avoids_pedestrians_and_rectangles():
    eps = make_episodes(4, seed=11)
    sim = ClosedLoopBatch(Dims(), MpcConfig(), eps, _oracle_solve(True))
    sim.run(25)
    for e in eps:
        xs = np.array(e.states)[:, :2]
        assert np.isfinite(xs).all()
        for poly in e.polygons:                          # never inside an (inflated) rectangle
            lo, hi = poly.min(0), poly.max(0)
            inside = ((xs > lo + 0.05) & (xs < hi - 0.05)).all(1)
            assert not inside.any()


@pytest.mark.gpu
def test_gpu_closed_loop_equals_laned_oracle_bitwise():
    """Same closed loop, GPU solver vs laned oracle: every state of every episode identical."""
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    bs = BatchedSolver(Dims(), RobotSpec(), CFG)

    def gpu_solve(P):
        o = bs.run_batch(torch.as_tensor(P, device="cuda"))
        torch.cuda.synchronize()
        return o["u"].cpu().numpy(), o["cost"].cpu().numpy(), o["exit_status"].cpu().numpy()

    eps_a = make_episodes(12, seed=5)
    eps_b = copy.deepcopy(eps_a)
    a = ClosedLoopBatch(Dims(), MpcConfig(), eps_a, gpu_solve)
    b = ClosedLoopBatch(Dims(), MpcConfig(), eps_b, _oracle_solve(True))
    a.run(20)
    b.run(20)
    for ea, eb in zip(eps_a, eps_b):
        np.testing.assert_array_equal(np.array(ea.states), np.array(eb.states))
        np.testing.assert_array_equal(np.array(ea.actions), np.array(eb.actions))
        assert ea.statuses == eb.statuses and ea.costs == eb.costs


@pytest.mark.gpu
def test_gpu_closed_loop_close_to_reference_order_oracle():
    """Against the reference-order oracle the closed-loop trajectories stay within a stated
    tolerance (0.05 m over 15 steps): feedback re-anchors the solve every step, so round-off
    level differences in the applied input do not accumulate."""
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    bs = BatchedSolver(Dims(), RobotSpec(), SolverSettings())

    def gpu_solve(P):
        o = bs.run_batch(torch.as_tensor(P, device="cuda"))
        torch.cuda.synchronize()
        return o["u"].cpu().numpy(), o["cost"].cpu().numpy(), o["exit_status"].cpu().numpy()

    def ref_solve(P):
        U, SC = oracle.solve_batch(Dims(), RobotSpec(), SolverSettings(), P, None, threads=os.cpu_count())
        return U, SC[:, 0], SC[:, 9].astype(np.int32)

    eps_a = [Episode(np.array([3.0, 5.0, 0.1]), [(9.0, 5.0), (9.0, 9.0)]) for _ in range(2)]
    eps_b = copy.deepcopy(eps_a)
    a = ClosedLoopBatch(Dims(), MpcConfig(), eps_a, gpu_solve)
    b = ClosedLoopBatch(Dims(), MpcConfig(), eps_b, ref_solve)
    a.run(15)
    b.run(15)
    for ea, eb in zip(eps_a, eps_b):
        n = min(len(ea.states), len(eb.states))
        assert np.abs(np.array(ea.states[:n]) - np.array(eb.states[:n]))[:, :2].max() < 0.05


@pytest.mark.gpu
def test_device_packer_row_equals_host_packer_bitwise():
    """K4 (device-side packer) writes the same parameter rows as the host
    mirror of the reference's packing code, bit for bit."""
    import torch
    from dyobav_mpcnwta_warehouse_b200.closed_loop_gpu import ClosedLoopGPU
    eps_a = make_episodes(16, seed=2)
    eps_b = copy.deepcopy(eps_a)
    g = ClosedLoopGPU(Dims(), MpcConfig(), eps_a, CFG)
    import ctypes
    from dyobav_mpcnwta_warehouse_b200 import _lib
    _lib.check(g.lib.mpcb_pack_f64(ctypes.byref(g._cd), ctypes.byref(g.sim), ctypes.c_void_p(g.P.data_ptr()), None),
               "pack")
    torch.cuda.synchronize()
    h = ClosedLoopBatch(Dims(), MpcConfig(), eps_b, _oracle_solve(True))
    rows = np.array([h._row(e) for e in eps_b])
    np.testing.assert_array_equal(g.P.cpu().numpy(), rows)


@pytest.mark.gpu
def test_device_closed_loop_equals_host_closed_loop_bitwise():
    """pack -> solve -> plant entirely on the GPU for 25 control periods vs the host harness
    driven by the laned oracle (and the library's portable sin/cos in the plant): identical
    robot states, reference indices and termination flags."""
    import torch
    from dyobav_mpcnwta_warehouse_b200 import _lib
    from dyobav_mpcnwta_warehouse_b200.closed_loop_gpu import ClosedLoopGPU
    eps_a = make_episodes(10, seed=4)
    eps_b = copy.deepcopy(eps_a)
    g = ClosedLoopGPU(Dims(), MpcConfig(), eps_a, CFG)
    h = ClosedLoopBatch(Dims(), MpcConfig(), eps_b, _oracle_solve(True), sincos=_lib.sincos_host)
    steps = 25
    traj = g.run(steps, record=True)
    torch.cuda.synchronize()
    h.run(steps)
    traj = torch.stack(traj).cpu().numpy()                 # [steps+1, n, 3]
    for i, e in enumerate(eps_b):
        hs = np.array(e.states)
        np.testing.assert_array_equal(traj[:len(hs), i], hs)
        assert bool(g.done[i].item()) == e.done
        if not e.done:
            assert int(g.idx_ref[i].item()) == e.idx_ref


@pytest.mark.gpu
def test_device_clustering_equals_host_restatement_bitwise():
    """K6: SWTA hypotheses -> o_d on the device equals the host restatement of the
    reference's DBSCAN + Gaussian fit (itself pinned to sklearn/numpy on the CPU) bit for bit."""
    import torch
    from dyobav_mpcnwta_warehouse_b200 import _lib, packing
    d = Dims()
    rng = np.random.default_rng(3)
    n, K, H = 12, 40, 2
    cur = rng.uniform(-5, 5, (n, H, 2))
    hyp = np.zeros((n, d.N, K, 2))
    for e in range(n):
        vel = rng.uniform(-1, 1, (H, 3, 2))                      # 3 modes per pedestrian
        for t in range(d.N):
            pts = []
            for h in range(H):
                modes = rng.integers(0, 3, K // H)
                pts.append(cur[e, h] + (t + 1) * 0.2 * vel[h, modes] + rng.normal(0, 0.25, (K // H, 2)))
            hyp[e, t] = np.concatenate(pts)
    od = _lib.cluster_hypotheses(d, torch.as_tensor(hyp, device="cuda"), torch.as_tensor(cur, device="cuda"))
    torch.cuda.synchronize()
    od = od.cpu().numpy()
    for e in range(n):
        ref = packing.hypotheses_to_obstacles(d, [tuple(c) for c in cur[e]], [hyp[e, t] for t in range(d.N)])
        np.testing.assert_array_equal(od[e], ref)
