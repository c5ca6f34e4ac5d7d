"""CPU: host-side mirror of the reference interface — config keys, p packing, the generated
drop-in module, error behaviour."""
import importlib
import sys

import numpy as np
import pytest
import yaml

from dyobav_mpcnwta_warehouse_b200 import Dims, MpcConfig, RobotSpec, SolverSettings, instances, packing
from dyobav_mpcnwta_warehouse_b200.problem import MPC_YAML_DEFAULTS
from dyobav_mpcnwta_warehouse_b200.solver_build import MpcModule


def test_yaml_roundtrip_and_views(tmp_path):
    path = tmp_path / "mpc_fast.yaml"
    path.write_text(yaml.safe_dump(MPC_YAML_DEFAULTS))
    cfg = MpcConfig.from_yaml(str(path))
    assert cfg.dims() == Dims() and cfg.robot() == RobotSpec()
    # set_work_mode weights (trajectory_tracker.py:138-139)
    assert cfg.tuning_params() == [0.0, 10.0, 0.0, 0, 0, 0.0, 0.0, 100.0, 10.0, 20.0]
    assert cfg.optimizer_name == "navi_fast" and cfg.N_hor == 20
    bad = dict(MPC_YAML_DEFAULTS)
    del bad["Ndynobs"]
    path.write_text(yaml.safe_dump(bad))
    with pytest.raises(KeyError):
        MpcConfig.from_yaml(str(path))


def test_layout_offsets_match_survey_table():
    lay = Dims().layout()
    assert lay["o_s"] == (728, 120) and lay["o_d"] == (848, 1890) and lay["q_dyn"] == (2758, 20)
    assert Dims().np == 2778


def test_assemble_params_order_and_lengths():
    d, cfg = Dims(), MpcConfig()
    ref = np.arange(60, dtype=float).reshape(20, 3)
    p = packing.assemble_params(d, cfg, state=[1.0, 2.0, 0.5], ref_states=ref, speed_ref=1.2,
                                last_u=[0.3, -0.1])
    assert len(p) == 2778
    assert p[:2] == [0.3, -0.1] and p[2:5] == [1.0, 2.0, 0.5] and p[5:8] == ref[-1].tolist()
    assert p[8:18] == [float(x) for x in cfg.tuning_params()]
    assert p[18:78] == ref.reshape(-1).tolist() and p[78:98] == [1.2] * 20
    assert p[2738:2758] == [10.0] * 20 and p[2758:] == [10.0] * 20
    with pytest.raises(ValueError):
        packing.assemble_params(d, cfg, [0, 0, 0], ref[:-1], 1.2)


def test_polygon_halfspaces_normalisation():
    sq = np.array([[0, 0], [2, 0], [2, 1], [0, 1]], float)
    b, a0, a1 = packing.polygon_halfspaces(sq)
    c = sq.mean(0)
    for bi, x, y in zip(b, a0, a1):
        assert bi - x * c[0] - y * c[1] == pytest.approx(1.0)          # 1 at the centroid
    edge_mid = [(1, 0), (2, .5), (1, 1), (0, .5)]
    vals = [[bi - x * px - y * py for bi, x, y in zip(b, a0, a1)] for px, py in edge_mid]
    assert all(min(abs(v) for v in row) < 1e-12 for row in vals)      # 0 on its edge
    with pytest.raises(ValueError):
        packing.pack_stc_constraints(Dims(), [np.array([[0, 0], [1, 0], [0, 1]], float)])  # 3 edges


def test_dyn_constraints_packing_and_padding_quirk():
    d = Dims()
    mu = [[(1.0, 2.0)]] + [[(1.0 + t, 2.0), (5.0, 5.0)] for t in range(1, 21)]
    sd = [[(0.2, 0.2)]] + [[(0.3, 0.3), (0.4, 0.4)] for _ in range(1, 21)]
    obs = packing.dyn_obstacles_from_predictions(d, mu, sd)
    assert len(obs) == 2 and obs[1][0] == [0, 0, 0, 0, 0, 1]           # SURVEY C-2 padding
    flat = packing.pack_dyn_constraints(d, obs)
    assert len(flat) == 1890 and flat[:6] == [1.0, 2.0, 0.2, 0.2, 0.0, 1.0]
    assert flat[126:132] == [0, 0, 0, 0, 0, 1] and flat[2 * 126:] == [0.0] * (13 * 126)


def test_ref_traj_sampling_spacing():
    traj = packing.ref_traj_from_path(0.2, [(0.0, 0.0), (3.0, 0.0), (3.0, 2.0)], (0.0, 0.0, 0.0), 1.2)
    xs = np.array(traj)
    assert np.allclose(np.hypot(*np.diff(xs[:5, :2], axis=0).T), 0.24)
    win, idx = packing.ref_states_window(0, traj, (0.0, 0.0, 0.0), horizon=20)
    assert win.shape == (20, 3) and idx == 0


def test_generated_dropin_module(tmp_path, monkeypatch):
    cfg = MpcConfig()
    path = MpcModule(cfg).build(root_dir=str(tmp_path))
    assert path.endswith("mpc_solver/navi_fast/navi_fast.py")
    assert MpcModule(cfg).build(test=True) == 1
    # load it the way trajectory_tracker.py:56-62 does
    monkeypatch.syspath_prepend(str(tmp_path / "mpc_solver" / "navi_fast"))
    sys.modules.pop("navi_fast", None)
    built = importlib.import_module("navi_fast")
    s = built.solver()
    assert hasattr(s, "run") and s.dims == Dims()
    # the yaml's wall-clock cap (max_solver_time, microseconds: mpc_builder.py:189) travels into the module
    assert s.settings.max_time_us == 100_000 and s.settings.max_inner_total == 0
    with pytest.raises(RuntimeError, match="3003"):
        s.run([0.0] * 5)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
            s.run([0.0] * 2778)


def test_generator_is_deterministic_and_sane():
    d = Dims()
    a = instances.generate(d, 64, seed=5)
    b = instances.generate(d, 64, seed=5)
    assert np.array_equal(a, b) and np.isfinite(a).all()
    lay = d.layout()
    o, ln = lay["o_d"]
    od = a[:, o:o + ln].reshape(64, d.Ndyn, d.N + 1, 6)
    assert (od[:, 6:] == 0).all() and (od[:, :6, :, 5] == 1).all() and (od[:, :6, :, 4] == 0).all()
    u0 = instances.multistart_guesses(d, a, 8, 5)
    assert u0.shape == (512, 40) and (u0[0::8] == 0).all()
    assert np.allclose(u0[1::8, 0::2], 1.2) and (u0[1::8, 1::2] == 0).all()


def test_dbscan_restatement_matches_sklearn():
    """The reference clusters hypotheses with sklearn.cluster.DBSCAN(eps=1, min_samples=2)
    (utils_test.py:133-143); the restated labelling must agree with it."""
    from sklearn.cluster import DBSCAN
    rng = np.random.default_rng(0)
    for trial in range(40):
        k = int(rng.integers(2, 41))
        centres = rng.uniform(-4, 4, (int(rng.integers(1, 4)), 2))
        X = centres[rng.integers(0, len(centres), k)] + rng.normal(0, rng.uniform(0.1, 0.8), (k, 2))
        for ms in (2, 3):
            ref = DBSCAN(eps=1.0, min_samples=ms).fit(X).labels_
            got = packing.dbscan_labels(X, 1.0, ms)
            assert np.array_equal(ref, got), (trial, ms)


def test_hypotheses_to_obstacles_matches_numpy_fit():
    rng = np.random.default_rng(1)
    d = Dims()
    cur = [(1.0, 2.0), (4.0, -1.0)]
    hyp = [np.concatenate([np.array(cur[0]) + 0.2 * t + rng.normal(0, 0.2, (20, 2)),
                           np.array(cur[1]) - 0.1 * t + rng.normal(0, 0.3, (20, 2))]) for t in range(d.N)]
    od = packing.hypotheses_to_obstacles(d, cur, hyp)
    assert od.shape == (15, 21, 6)
    assert od[0, 0].tolist() == [1.0, 2.0, 0.2, 0.2, 0.0, 1.0] and od[1, 0].tolist() == [4.0, -1.0, 0.2, 0.2, 0.0, 1.0]
    lab = packing.dbscan_labels(hyp[3])
    for c in range(lab.max() + 1):
        pts = hyp[3][lab == c]
        np.testing.assert_array_equal(od[c, 4, :2], np.mean(pts, axis=0))          # bit-exact vs numpy
        np.testing.assert_array_equal(od[c, 4, 2:4], np.std(pts, axis=0) * 2 + 0)
    assert (od[:, :, 5][(od[:, :, :4] != 0).any(-1)] == 1).all()


@pytest.mark.parametrize("n_p,starts,nq", [(1, 1, 148), (301, 3, 148), (8192, 8, 148), (7, 5, 4), (148, 8, 148),
                                           (149, 1, 148), (5, 8, 1)])
def test_per_cta_queue_index_map_is_a_bijection(n_p, starts, nq):
    """Host mirror of the index arithmetic of the solve kernel's work queues
    (csrc/mpcb_solver.cuh, L_fetch): queue q owns scenarios q, q+nq, ...; ticket t of queue q is
    start t % starts of that queue's (t // starts)-th scenario.  Every instance must be handed
    out exactly once whatever the batch, the number of starts and the grid."""
    seen = np.zeros(n_p * starts, dtype=np.int64)
    for q in range(nq):
        nsc_q = (n_p - q + nq - 1) // nq if q < n_p else 0
        for t in range(nsc_q * starts):
            sc = q + (t // starts) * nq
            assert sc < n_p
            seen[sc * starts + t % starts] += 1
    assert (seen == 1).all()
