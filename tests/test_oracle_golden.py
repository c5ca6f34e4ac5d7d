"""CPU: the oracle's problem-definition half against (a) goldens produced by executing the
reference's own mpc_builder/mpc_cost/mpc_helper/motion_model code (tests/golden/gen_golden.py)
and (b) the reference's own known-answer vectors (src/tests/test_mpc_builder.py:16-253)."""
import ctypes
import math

import numpy as np
import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec
from oracle import oracle


def _cases(golden):
    return [str(k) for k in golden["cases"]]


def test_golden_file_has_cases(golden):
    assert len(_cases(golden)) >= 10


def test_oracle_matches_reference_goldens(golden):
    for key in _cases(golden):
        d = Dims(*[int(v) for v in golden[f"{key}/dims"]])
        r = oracle.evaluate(d, RobotSpec(), golden[f"{key}/p"], golden[f"{key}/u"],
                            golden[f"{key}/y"], float(golden[f"{key}/c"]))
        # tolerance: f64 round-off of a few thousand term sums (relative)
        assert r["f"] == pytest.approx(float(golden[f"{key}/f"]), rel=1e-12), key
        assert r["psi"] == pytest.approx(float(golden[f"{key}/psi"]), rel=1e-12), key
        g = golden[f"{key}/grad"]
        assert np.max(np.abs(r["grad"] - g)) <= 1e-11 * max(1.0, np.max(np.abs(g))), key
        np.testing.assert_allclose(r["F1"], golden[f"{key}/F1"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(r["F2"], golden[f"{key}/F2"], rtol=1e-13, atol=1e-13)


def test_constraint_sets_match_reference(golden):
    """Input box U and ALM set C the builder hands to OpEn (mpc_builder.py:151-166)."""
    rb = RobotSpec()
    key = _cases(golden)[0]
    N = int(golden[f"{key}/dims"][0])
    np.testing.assert_array_equal(golden[f"{key}/umin"], [rb.lin_vel_min, -rb.ang_vel_max] * N)
    np.testing.assert_array_equal(golden[f"{key}/umax"], [rb.lin_vel_max, rb.ang_vel_max] * N)
    np.testing.assert_array_equal(golden[f"{key}/cmin"], [rb.lin_acc_min] * N + [-rb.ang_acc_max] * N)
    np.testing.assert_array_equal(golden[f"{key}/cmax"], [rb.lin_acc_max] * N + [rb.ang_acc_max] * N)


def test_gradient_against_finite_differences(golden):
    key = "small_1"
    d = Dims(*[int(v) for v in golden[f"{key}/dims"]])
    p, u, y, c = golden[f"{key}/p"], golden[f"{key}/u"].copy(), golden[f"{key}/y"], float(golden[f"{key}/c"])
    g = oracle.evaluate(d, RobotSpec(), p, u, y, c)["grad"]
    h = 1e-6
    for i in range(d.nu_total):
        up, um = u.copy(), u.copy()
        up[i] += h
        um[i] -= h
        fd = (oracle.evaluate(d, RobotSpec(), p, up, y, c, False)["psi"]
              - oracle.evaluate(d, RobotSpec(), p, um, y, c, False)["psi"]) / (2 * h)
        assert g[i] == pytest.approx(fd, rel=2e-5, abs=2e-4)


# ---- the reference's known-answer vectors ----------------
def test_known_answer_dist_to_lineseg():
    L = oracle.lib()
    assert L.mpco_dist_to_lineseg(1, 2, 3, 2, 3, 0) == pytest.approx(2.0, abs=1e-12)
    assert L.mpco_dist_to_lineseg(1, 2, 3, 1, 3, 0) == pytest.approx(math.sqrt(5), abs=1e-12)


def test_known_answer_inside_ellipses():
    L = oracle.lib()
    assert L.mpco_inside_ellipse(1, 2, 1, 2, 1, 1, 0) == 1.0
    assert L.mpco_inside_ellipse(1, 2, 1, 4, 1, 1, 0) == pytest.approx(-3.0, abs=1e-3)


def _poly(x, y, b, a0, a1):
    L = oracle.lib()
    arr = lambda v: (ctypes.c_double * len(v))(*v)  # noqa: E731
    return L.mpco_inside_cvx_polygon(x, y, arr(b), arr(a0), arr(a1), len(b))


def test_known_answer_inside_cvx_polygon():
    assert _poly(1, 2, [0, 2, 1, 3], [-1, 1, 0, 0], [0, 0, -1, 1]) == 3.0
    assert _poly(1, 2, [0, 1, 0, 1], [-1, 1, 0, 0], [0, 0, -1, 1]) == 0.0
    # cost_inside_cvx_polygon, weight 2 -> 2 * 3^2 = 18 (test_mpc_builder.py:123-137)
    assert 2 * _poly(1, 2, [0, 2, 1, 3], [-1, 1, 0, 0], [0, 0, -1, 1]) ** 2 == 18.0


def test_known_answer_refpath_deviation():
    """cost_refpath_deviation: p=(1,2), polyline (0,0),(1,0),(3,2), weight 0.5 -> 1.0"""
    L = oracle.lib()
    d = min(L.mpco_dist_to_lineseg(1, 2, 0, 0, 1, 0) ** 2, L.mpco_dist_to_lineseg(1, 2, 1, 0, 3, 2) ** 2)
    assert 0.5 * d == pytest.approx(1.0, abs=1e-3)


def test_known_answer_fleet_and_points():
    """cost_fleet_collision p=(1,2), safe distance 2 vs (0,1),(2,0) -> sum of hinges 2 (the reference's
    test weighs it by 2 -> 4), vs (0,0),(2,0) -> 0 (test_mpc_builder.py:181-200), evaluated by BOTH
    oracles through the whole problem: a one-step horizon with every other weight zero, the robot at
    rest at (1,2), the predicted robots `c` at the test's points (weight 10, mpc_builder.py:93-97) and
    the t=0 robots `c_0` far away."""
    from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec
    d = Dims(N=1, Nother=3, Nstc=0, nedge=4, Ndyn=0)
    rb = RobotSpec(vehicle_width=2.0)
    lay = d.layout()
    for pts, hinge_sum in (([(0, 1), (2, 0)], 2.0), ([(0, 0), (2, 0)], 0.0)):
        p = np.zeros(d.np)
        p[lay["s_0"][0]:lay["s_0"][0] + 3] = [1.0, 2.0, 0.3]
        p[lay["r_s"][0]:lay["r_s"][0] + 3] = [1.0, 2.0, 0.3]
        c0 = np.full((d.Nother, 3), 50.0)
        p[lay["c_0"][0]:lay["c_0"][0] + 9] = c0.reshape(-1)
        c = np.full((d.Nother, d.N, 3), 50.0)
        for r, q in enumerate(pts):
            c[r, 0, :2] = q
        p[lay["c"][0]:lay["c"][0] + 9] = c.reshape(-1)
        for laned in (False, True):
            out = oracle.evaluate(d, rb, p, np.zeros(2), None, 10.0, laned=laned)
            assert out["f"] == 10.0 * hinge_sum, (pts, laned)
            # d cost / d v: the hinge is linear in the squared distance (mpc_cost.py:75)
            assert np.isfinite(out["grad"]).all()
    # dist_to_points_square p=(0,0) vs (1,0),(2,0) -> [1, 4] (test_mpc_builder.py:16-26): the same
    # squared distances through the hinge with safe distance^2 = 5: [5-1, 5-4]
    rb5 = RobotSpec(vehicle_width=math.sqrt(5.0))
    p = np.zeros(d.np)
    c = np.full((d.Nother, d.N, 3), 50.0)
    c[0, 0, :2] = (1, 0); c[1, 0, :2] = (2, 0)
    p[lay["c"][0]:lay["c"][0] + 9] = c.reshape(-1)
    p[lay["c_0"][0]:lay["c_0"][0] + 9] = 50.0
    out = oracle.evaluate(d, rb5, p, np.zeros(2), None, 10.0)
    assert out["f"] == pytest.approx(10.0 * ((5 - 1) + (5 - 4)), rel=1e-12)


def test_unicycle_rk4_matches_closed_form():
    L = oracle.lib()
    s = (ctypes.c_double * 3)(0.3, -0.2, 0.4)
    out = (ctypes.c_double * 3)()
    v, w, ts = 1.1, 0.35, 0.2
    L.mpco_unicycle_rk4(s, v, w, ts, out)
    th = 0.4
    cx = (ts * v / 6) * (math.cos(th) + 4 * math.cos(th + ts * w / 2) + math.cos(th + ts * w))
    sy = (ts * v / 6) * (math.sin(th) + 4 * math.sin(th + ts * w / 2) + math.sin(th + ts * w))
    assert out[0] == pytest.approx(0.3 + cx, abs=1e-15)
    assert out[1] == pytest.approx(-0.2 + sy, abs=1e-15)
    assert out[2] == pytest.approx(th + ts * w, abs=1e-15)


def test_committed_goldens_regenerate_from_their_own_inputs():
    """tests/golden/gen_golden.py re-runs the REFERENCE'S code on the committed inputs: the committed
    outputs must come back exactly (authoring container only: needs /root/reference)."""
    import os
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("the reference tree is only present in the authoring container")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "golden", "gen_golden.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("MAXDIFF")][-1]
    assert float(line.split()[1]) == 0.0, line
