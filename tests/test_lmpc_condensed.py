"""LMPC as the north star states it: condensing + batched dense QP, with the reference's own equality block available
through ``bug_compat=True`` (mpc.py:2243).  CPU part: the condensed matrices against the oracle's
solve of the reference's FLAT QP (oracle/lmpc.py).  GPU part: ``hb_lmpc_solve`` through ``LMPC.optimize`` for both modes."""
import numpy as np
import pytest

from test_hostsim_lmpc_mhe import double_integrator


def _mpc(bug_compat, N=10):
    from hilo_mpc_b200 import LMPC
    mpc = LMPC(double_integrator(), bug_compat=bug_compat)
    mpc.Q, mpc.R, mpc.horizon = np.eye(2), 1, N
    mpc.set_box_constraints(x_lb=[-5, -5], x_ub=[5, 5], u_lb=[-1], u_ub=[1])
    mpc.setup()
    return mpc


def _numpy_condensed_solution(c, x0):
    """Exact solve of the condensed QP by enumerating the active set with SciPy (small n): reference for the CPU test."""
    from scipy.optimize import minimize, Bounds, LinearConstraint
    H, g = c['H'], c['F'] @ x0
    d = c['Phi'] @ x0
    res = minimize(lambda u: .5 * u @ H @ u + g @ u, np.zeros(c['n']), jac=lambda u: H @ u + g, hess=lambda u: H,
                   method='trust-constr', bounds=Bounds(c['ulb'], c['uub']),
                   constraints=[LinearConstraint(c['Gam'], c['xlb'] - d, c['xub'] - d)],
                   options=dict(gtol=1e-12, xtol=1e-14, barrier_tol=1e-13, maxiter=3000))
    return res.x


@pytest.mark.parametrize('bug_compat', [False, True])
def test_condensed_matrices_reproduce_the_flat_qp_of_the_reference(bug_compat):
    from oracle.lmpc import solve_lmpc
    mpc = _mpc(bug_compat)
    c = mpc._condensed
    assert mpc.engine is None and c['n'] == 10 and c['mx'] == 20
    A, B = np.array([[1, .5], [0, 1]]), np.array([[.125], [.5]])
    np.testing.assert_allclose(c['A'], A)
    np.testing.assert_allclose(c['B'], B)
    # (the reference's kron(B, I_N) QP is INFEASIBLE for e.g. x0 = [3.5, 1.5], which the intended QP handles: the
    # acausal input block cannot brake that state inside the box)
    for x0 in ([1., 1.], [3.5, 1.5] if not bug_compat else [2., .5], [-4.5, -.5]):
        x0 = np.array(x0)
        xo, uo = solve_lmpc(A, B, np.eye(2), np.eye(1), np.eye(2), 10, x0, [-5, -5], [5, 5], [-1], [1], bug_compat=bug_compat)
        u = _numpy_condensed_solution(c, x0)
        np.testing.assert_allclose(u, uo[:, 0], atol=2e-6)
        np.testing.assert_allclose((c['Phi'] @ x0 + c['Gam'] @ u).reshape(10, 2), xo[1:], atol=2e-6)
    # the two input blocks give different QPs for nx = 2 (they coincide only for nx = nu = 1)
    other = _mpc(not bug_compat)._condensed
    assert np.abs(other['Gam'] - c['Gam']).max() > .1


@pytest.mark.gpu
@pytest.mark.parametrize('bug_compat', [False, True])
def test_condensed_dense_qp_kernel_matches_the_oracle(cuda_device, bug_compat):
    """hb_lmpc_solve (batched interior-point method on the condensed QP) vs the oracle's solve of the reference's flat QP;
    closed loop of tests/test_LMPC.py:21-44 through the single-instance API."""
    import torch
    from oracle.lmpc import solve_lmpc
    A, B = np.array([[1, .5], [0, 1]]), np.array([[.125], [.5]])
    mpc = _mpc(bug_compat)
    rng = np.random.default_rng(0)
    X0 = np.vstack([[1., 1.], [3.5, 1.5] if not bug_compat else [2., .5], [-4.5, -.5],
                    rng.uniform(-1.5, 1.5, size=(61, 2))])   # (states from which the box can be kept: feasible QPs)
    u0 = mpc.optimize(torch.tensor(X0, device=cuda_device))
    assert u0.shape == (64, 1)
    assert (mpc.solver_status_code == 1).all()
    x_pred, u_pred = mpc.return_prediction()
    for b in (0, 1, 2, 17, 63):
        xo, uo = solve_lmpc(A, B, np.eye(2), np.eye(1), np.eye(2), 10, X0[b], [-5, -5], [5, 5], [-1], [1], bug_compat=bug_compat)
        np.testing.assert_allclose(u_pred[b].cpu().numpy(), uo, atol=1e-6)
        np.testing.assert_allclose(x_pred[b].cpu().numpy(), xo, atol=1e-6)
    if not bug_compat:
        model = double_integrator()
        model.set_initial_conditions(x0=[1, 1])
        x0 = np.array([1., 1.])
        for i in range(60):
            u = mpc.optimize(x0=x0)
            assert u.shape == (1, 1) and mpc.solver_status_code == 1
            model.simulate(u=u)
            x0 = model.solution['x:f'].reshape(-1)
        assert np.abs(x0).max() < 1e-3
