"""CPU: the laned oracle (kernel operation order) against the reference-order oracle and the
goldens produced by the reference's own code; and the round-off sensitivity of PANOC that makes
a bit-exact mirror necessary."""
import numpy as np
import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
from oracle import oracle


def test_laned_eval_matches_reference_goldens(golden):
    for key in [str(k) for k in golden["cases"]]:
        d = Dims(*[int(v) for v in golden[f"{key}/dims"]])
        r = oracle.evaluate(d, RobotSpec(), golden[f"{key}/p"], golden[f"{key}/u"], golden[f"{key}/y"],
                            float(golden[f"{key}/c"]), laned=True)
        assert r["f"] == pytest.approx(float(golden[f"{key}/f"]), rel=1e-12), key
        assert r["psi"] == pytest.approx(float(golden[f"{key}/psi"]), rel=1e-12), key
        g = golden[f"{key}/grad"]
        assert np.max(np.abs(r["grad"] - g)) <= 1e-11 * max(1.0, np.max(np.abs(g))), key
        np.testing.assert_allclose(r["F1"], golden[f"{key}/F1"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["F2"], golden[f"{key}/F2"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("dims", [Dims(), Dims(Ndyn=40), Dims(N=40, Nother=3, Nstc=4, Ndyn=12),
                                  Dims(N=32, Nother=2, Nstc=2, Ndyn=3), Dims(N=1, Nother=1, Nstc=1, Ndyn=1)])
def test_laned_eval_matches_reference_order_eval(dims):
    P = instances.generate(dims, 4, seed=13, modes=3)
    rng = np.random.default_rng(1)
    for i in range(4):
        u = np.stack([rng.uniform(-0.5, 1.5, dims.N), rng.uniform(-0.5, 0.5, dims.N)], 1).reshape(-1)
        y = rng.normal(0, 2.0, dims.n1)
        a = oracle.evaluate(dims, RobotSpec(), P[i], u, y, 50.0)
        b = oracle.evaluate(dims, RobotSpec(), P[i], u, y, 50.0, laned=True)
        assert b["psi"] == pytest.approx(a["psi"], rel=1e-12)
        assert np.max(np.abs(a["grad"] - b["grad"])) <= 1e-11 * max(1.0, np.max(np.abs(a["grad"])))
        np.testing.assert_allclose(b["F2"], a["F2"], rtol=1e-12, atol=1e-12)


def test_solvers_agree_on_converging_instances():
    """Reference-order and laned solver are the same algorithm: where both converge they end at
    the same minimiser (1e-4, the north_star tolerance) with the same cost."""
    dims = Dims()
    P = instances.generate(dims, 24, seed=3)
    cfg = SolverSettings()
    Ua, SA = oracle.solve_batch(dims, RobotSpec(), cfg, P, None, threads=8)
    Ub, SB = oracle.solve_batch(dims, RobotSpec(), cfg, P, None, threads=8, laned=True)
    both = (SA[:, 9] == 0) & (SB[:, 9] == 0)
    assert both.sum() >= 3
    assert (np.max(np.abs(Ua - Ub), axis=1)[both] <= 1e-4).all()
    np.testing.assert_allclose(SA[both, 0], SB[both, 0], rtol=1e-6)


def test_panoc_amplifies_roundoff():
    """Why parity is tested bit-exactly against the laned oracle: a 1e-14 perturbation of the
    initial guess changes the reference-order oracle's own answer by orders of magnitude more
    within tens of iterations (the Lipschitz estimate divides a gradient difference by
    |h| = 6e-12, then line-search branches diverge)."""
    dims = Dims()
    P = instances.generate(dims, 16, seed=3)
    cfg = SolverSettings(max_inner=40, max_outer=2)
    U0 = np.zeros((16, dims.nu_total))
    Ua, _ = oracle.solve_batch(dims, RobotSpec(), cfg, P, U0, threads=8)
    Ub, _ = oracle.solve_batch(dims, RobotSpec(), cfg, P, U0 + 1e-14, threads=8)
    du = np.max(np.abs(Ua - Ub), axis=1)
    assert np.median(du) > 1e-10        # >= 4 orders of magnitude amplification


def test_iteration_budget_is_the_out_of_time_status():
    """cfg.max_inner_total restates the reference's wall-clock cap (mpc_fast.yaml:45) as a budget on
    inner iterations: both oracles stop at the budget, report NotConvergedOutOfTime (2) exactly for
    the instances that needed more, and are untouched when it is off or generous."""
    from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances
    from oracle import oracle
    import numpy as np
    d = Dims()
    P = instances.generate(d, 12, seed=3)
    for laned in (False, True):
        U0, S0 = oracle.solve_batch(d, RobotSpec(), SolverSettings(), P, None, threads=8, laned=laned)
        U1, S1 = oracle.solve_batch(d, RobotSpec(), SolverSettings(max_inner_total=700), P, None, threads=8, laned=laned)
        U2, S2 = oracle.solve_batch(d, RobotSpec(), SolverSettings(max_inner_total=10**6), P, None, threads=8, laned=laned)
        np.testing.assert_array_equal(U0, U2)
        np.testing.assert_array_equal(S0, S2)
        need_more = S0[:, 6] > 700
        assert need_more.any() and (~need_more).any()
        assert (S1[need_more, 9] == 2).all() and (S1[need_more, 6] == 700).all()
        np.testing.assert_array_equal(S1[~need_more], S0[~need_more])
        np.testing.assert_array_equal(U1[~need_more], U0[~need_more])
