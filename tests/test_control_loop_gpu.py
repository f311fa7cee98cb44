"""SimpleControlLoop (control_loop.py:343-431) with the loop on the device: batched CSTR plant + NMPC
(+ EKF beside it).  Checks: every solve optimal, the closed loop converges to the set point, the plant (RK4, 10 sub-steps)
lands where the controller's collocation model predicted, and the one-instance mode reproduces instance 0 of the batched
run while filling ``plant.solution`` like the reference does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _loop(N=10):
    from hilo_mpc_b200.control_loop import SimpleControlLoop
    from hilo_mpc_b200.examples import cstr_nmpc
    from hilo_mpc_b200.library import cstr_model
    plant = cstr_model(0.005)
    nmpc = cstr_nmpc(N=N)
    return SimpleControlLoop(plant, nmpc), plant, nmpc


def test_batched_closed_loop_converges_and_matches_prediction(cuda_device):
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states
    from hilo_mpc_b200.library import cstr_steady_state
    loop, plant, nmpc = _loop()
    B, steps = 16, 25
    x0 = cstr_initial_states(B, seed=5)
    out = loop.run(steps, x0=torch.tensor(x0, device=cuda_device))
    assert out['x'].shape == (steps + 1, B, 4) and out['u'].shape == (steps, B, 2)
    assert (out['status'] == 1).all()
    x_ss, u_ss = cstr_steady_state()
    scale = torch.tensor([1., 1., 100., 100.], device=cuda_device, dtype=torch.float64)
    xs = torch.tensor(x_ss, device=cuda_device)
    d0 = ((out['x'][0] - xs) / scale).abs().max(dim=1).values
    d1 = ((out['x'][-1] - xs) / scale).abs().max(dim=1).values
    assert (d1 < d0).all() and d1.mean().item() < 0.3 * d0.mean().item(), (d0, d1)   # 25 steps = 0.125 h of a slow plant
    # inputs stay inside their box (cstr bounds)
    assert (out['u'][..., 0] >= 5. - 1e-9).all() and (out['u'][..., 0] <= 100. + 1e-9).all()
    assert (out['u'][..., 1] >= -8500. - 1e-6).all() and (out['u'][..., 1] <= 1e-6).all()
    # one more step by hand: the plant must land on the controller's predicted next state (same ODE, Radau-3 vs RK4/10)
    x = out['x'][-1].clone()
    nmpc.optimize(x)
    x_pred, _, _ = nmpc.return_prediction()
    nxt = loop.run(1, x0=x)['x'][1]
    assert ((nxt - x_pred[:, 1, :4]) / scale).abs().max().item() < 1e-5


def test_single_instance_mode_fills_plant_solution(cuda_device):
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states
    x0 = cstr_initial_states(3, seed=9)
    loop_b, _, _ = _loop()
    ref = loop_b.run(4, x0=torch.tensor(x0, device=cuda_device))
    loop_s, plant, nmpc = _loop()
    plant.set_initial_conditions(x0[0])
    loop_s.run(4)
    xs = np.asarray(plant.solution['x'])
    assert xs.shape[1] == 5   # initial condition + 4 steps
    np.testing.assert_allclose(xs.T, ref['x'][:, 0].cpu().numpy(), rtol=0, atol=1e-7)
    assert nmpc.solver_status_code == 1


def test_observer_runs_beside_the_loop(cuda_device):
    """EKF fed with the plant's measurements (c_B, T_R) and the applied inputs tracks the plant state."""
    import torch
    from hilo_mpc_b200 import EKF
    from hilo_mpc_b200.control_loop import SimpleControlLoop
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    from hilo_mpc_b200.library import cstr_model
    plant = cstr_model(0.005)
    ekf = EKF(cstr_model(0.005))
    ekf.Q = np.diag([1e-4, 1e-4, 1e-2, 1e-2])
    ekf.R = np.diag([1e-6, 1e-4])
    ekf.setup()
    x0 = cstr_initial_states(1, seed=2)[0]
    ekf.set_initial_guess(x0 * np.array([1.05, 0.95, 1.01, 0.99]), P0=np.diag([1e-2, 1e-2, 4., 4.]))
    plant.set_initial_conditions(x0)
    loop = SimpleControlLoop(plant, cstr_nmpc(N=10), observer=ekf)
    out = loop.run(20)
    err0 = np.abs(x0 * np.array([.05, -.05, .01, -.01]))
    err1 = (out['x_estimate'][-1, 0] - out['x'][-1, 0]).abs().cpu().numpy()
    assert (err1[1:3] < 0.2 * err0[1:3] + 1e-6).all(), (err0, err1)   # measured states are pulled in
    assert np.isfinite(err1).all()
