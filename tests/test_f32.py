"""The f32 boundary mode (include/mpcb.h "f32 twins"): float32 buffers in and out, f64
arithmetic.  Stated tolerance: bit-identical to the f64 entry points on the widened inputs, rounded to
float32; against the f64 solve of the UNROUNDED parameters the difference is what rounding p to float32
(1e-6 m in the positions) does to the solve: <= 1e-3 in u on instances that converge in both."""
import numpy as np
import pytest

from dyobav_mpcnwta_warehouse_b200 import Dims, RobotSpec, SolverSettings, instances

pytestmark = pytest.mark.gpu
TOL_U_F32 = 1e-3


@pytest.mark.parametrize("dims,starts,modes,peds", [(Dims(), 4, 3, 2), (Dims(Ndyn=40), 1, 20, 2),
                                                    (Dims(N=20, Nother=2, Nstc=3, Ndyn=64), 1, 4, 16)])
def test_f32_twins_equal_f64_on_widened_inputs(dims, starts, modes, peds):
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    assert torch.cuda.is_available()
    n_p = 12
    P = instances.generate(dims, n_p, seed=61, modes=modes, pedestrians=peds)
    U0 = instances.multistart_guesses(dims, P, starts, 61)
    s = BatchedSolver(dims, RobotSpec(), SolverSettings(max_inner=150, max_outer=4))
    P32 = torch.as_tensor(P, dtype=torch.float32, device="cuda").contiguous()
    U32 = torch.as_tensor(U0, dtype=torch.float32, device="cuda").contiguous()
    o32 = s.run_batch_f32(P32, U32, starts=starts)
    o64 = s.run_batch(P32.double().contiguous(), U32.double().contiguous(), starts=starts)
    torch.cuda.synchronize()
    for k in ("u", "cost", "fpr", "f1_infeas", "f2_norm", "penalty", "y"):
        assert o32[k].dtype == torch.float32
        assert torch.equal(o32[k], o64[k].float()), k
    for k in ("exit_status", "n_outer", "n_inner", "evals"):
        assert torch.equal(o32[k], o64[k]), k
    # evaluation twin
    e32 = s.evaluate_f32(P32, U32, starts=starts)
    e64 = s.evaluate(P32.double().contiguous(), U32.double().contiguous(), starts=starts)
    torch.cuda.synchronize()
    for k in ("f", "psi", "grad", "F1", "F2"):
        assert torch.equal(e32[k], e64[k].float()), k


def test_f32_mode_against_the_f64_solve_of_unrounded_parameters():
    """What the f32 boundary costs: rounding p to float32, carried through PANOC."""
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    dims, n_p = Dims(), 96
    P = instances.generate(dims, n_p, seed=62)
    s = BatchedSolver(dims, RobotSpec(), SolverSettings())
    o64 = s.run_batch(torch.as_tensor(P, device="cuda").contiguous())
    o32 = s.run_batch_f32(torch.as_tensor(P, dtype=torch.float32, device="cuda").contiguous())
    torch.cuda.synchronize()
    both = ((o64["exit_status"] == 0) & (o32["exit_status"] == 0)).cpu().numpy()
    assert both.sum() >= 10
    du = (o32["u"].double() - o64["u"]).abs().max(1).values.cpu().numpy()
    assert (du[both] <= TOL_U_F32).all(), np.sort(du[both])[-5:]
    assert (o64["exit_status"] == o32["exit_status"]).double().mean().item() >= 0.9


def test_f32_argument_errors():
    import torch
    from dyobav_mpcnwta_warehouse_b200.solver import BatchedSolver
    s = BatchedSolver(Dims(), RobotSpec(), SolverSettings())
    with pytest.raises(RuntimeError):
        s.run_batch_f32(torch.zeros(2, 2778, dtype=torch.float64, device="cuda"))
    with pytest.raises(RuntimeError):
        s.run_batch(torch.zeros(2, 2778, dtype=torch.float32, device="cuda"))
