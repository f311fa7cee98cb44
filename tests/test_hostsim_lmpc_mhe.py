"""LMPC and MHE (a4) kernel logic on the CPU: host build of the kernel bodies vs the oracle's flat NLP of
the same transcription; LMPC additionally against an independent dense QP solve (SciPy) and the LQR golden."""
import numpy as np
import pytest
import scipy.optimize as sopt

from hilo_mpc_b200 import Model
from hilo_mpc_b200.codegen import OcpGenerator
from hilo_mpc_b200.lmpc import LMPC
from hilo_mpc_b200.mhe import MHE
from oracle.nmpc import FlatNlp
from runner import HostOcp


def double_integrator(Ts=.5):
    # tests/test_LMPC.py:8-19
    model = Model(discrete=True)
    model.A = np.array([[1, model.dt], [0, 1]])
    model.B = np.array([[model.dt ** 2 / 2], [model.dt]])
    model.setup(dt=Ts)
    return model


def _lmpc(N=10):
    mpc = LMPC(double_integrator())
    mpc.Q = np.eye(2)
    mpc.R = 1
    mpc.horizon = N
    mpc.set_box_constraints(x_lb=[-5, -5], x_ub=[5, 5], u_lb=[-1], u_ub=[1])
    mpc._build_module = False
    mpc.setup(solver='riccati')   # the stage-structured path (the condensed dense-QP path needs no generated module)
    return mpc


def _dense_qp_reference(x0, N, Ts=.5):
    """Independent check: condensed QP  min 1/2 sum x_k'Qx_k + 1/2 x_N'P x_N + 1/2 sum u_k'R u_k  solved with SciPy's
    trust-constr on the input sequence (bounds on u, linear inequality for the state box)."""
    A = np.array([[1, Ts], [0, 1]])
    Bm = np.array([[Ts ** 2 / 2], [Ts]])
    Phi = np.vstack([np.linalg.matrix_power(A, k) for k in range(N + 1)])
    Gam = np.zeros((2 * (N + 1), N))
    for k in range(1, N + 1):
        for j in range(k):
            Gam[2 * k:2 * k + 2, j] = (np.linalg.matrix_power(A, k - 1 - j) @ Bm)[:, 0]
    H = Gam.T @ Gam + np.eye(N)
    g = Gam.T @ Phi @ x0
    cons = sopt.LinearConstraint(Gam[2:], -5 - (Phi @ x0)[2:], 5 - (Phi @ x0)[2:])
    res = sopt.minimize(lambda u: .5 * u @ H @ u + g @ u, np.zeros(N), jac=lambda u: H @ u + g, hess=lambda u: H,
                        method='trust-constr', bounds=sopt.Bounds(-1, 1), constraints=[cons],
                        options=dict(gtol=1e-12, xtol=1e-14, barrier_tol=1e-12, maxiter=2000))
    return res.x


def test_lmpc_double_integrator_matches_oracle_and_dense_qp():
    mpc = _lmpc(10)
    spec = mpc.spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    x0 = np.array([[1., 1.], [3.5, 1.5], [-3., -1.]])
    B, N = 3, 10
    x = np.tile(spec.x_guess, (B, N + 1, 1))
    u = np.tile(spec.u_guess, (B, N, 1))
    out = host.solve(x0, x, u, None, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub, dt=.5, tol=1e-10)
    assert (out['status'] == 1).all()
    nlp = FlatNlp(spec, exact_hessian=True)
    for b in range(B):
        res, xo, uo = nlp.solve(x0[b], dt=.5, finish_barrier=True, tol=1e-10)
        assert res.status == 1
        np.testing.assert_allclose(out['u'][b], uo.T, atol=1e-7)
        np.testing.assert_allclose(out['x'][b], xo.T, atol=1e-7)
        np.testing.assert_allclose(out['u'][b, :, 0], _dense_qp_reference(x0[b], N), atol=2e-6)


def test_unconstrained_lq_problem_reproduces_lqr_golden():
    """tests/test_LQR.py:314-342 system: with no active bounds the first LMPC move equals -K x0 with the finite-horizon
    Riccati gain K (golden) when the terminal weight is Q and the horizon is 6 (= 5 Riccati iterations + terminal)."""
    from oracle.lqr import lqr_gain
    model = Model(discrete=True)
    model.set_dynamical_states('x', 'y', 'z')
    model.set_inputs('u', 'w')
    model.set_dynamical_equations(['x + dt*(2*y + u)', 'y - dt*x', 'z + dt*w'])
    model.setup(dt=1.)
    mpc = LMPC(model)
    mpc.Q, mpc.R, mpc.horizon = np.eye(3), np.eye(2), 6
    mpc._build_module = False
    mpc.setup(solver='riccati')
    spec = mpc.spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    x0 = np.array([[1., 0., 1.]])
    out = host.solve(x0, np.zeros((1, 7, 3)), np.zeros((1, 6, 2)), None, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub,
                     dt=1., tol=1e-10)
    assert out['status'][0] == 1
    A = np.array([[1., 2., 0.], [-1., 1., 0.], [0., 0., 1.]])
    Bm = np.array([[1., 0.], [0., 0.], [0., 1.]])
    K, _ = lqr_gain(A, Bm, np.eye(3), np.eye(2), horizon=5)
    np.testing.assert_allclose(K, [[1.39207671, 1.35221712, 0.], [0., 0., .61802575]], rtol=1e-7)
    np.testing.assert_allclose(out['u'][0, 0], -K @ x0[0], atol=1e-8)


def reaction_model(dt=.25):
    # two-state reaction A -> B with unknown rate constant k (parameter), measurement of B
    model = Model(name='reaction')
    x = model.set_dynamical_states(['ca', 'cb'])
    u = model.set_inputs(['f'])
    p = model.set_parameters(['k'])
    model.set_measurements(['y'])
    model.set_measurement_equations([x[1]])
    model.set_dynamical_equations([u[0] * (1. - x[0]) - p[0] * x[0], p[0] * x[0] - u[0] * x[1]])
    model.setup(dt=dt)
    return model


def _mhe(N=6):
    mhe = MHE(reaction_model())
    mhe.horizon = N
    mhe.quad_arrival_cost.add_states(weights=[10., 10.], guess=[.5, .2])
    mhe.quad_arrival_cost.add_parameters(weights=[1.], guess=[1.])
    mhe.quad_stage_cost.add_measurements(weights=[100.])
    mhe.quad_stage_cost.add_state_noise(weights=[20., 20.])
    mhe.set_box_constraints(x_lb=[0., 0.], x_ub=[1., 1.], p_lb=[.1], p_ub=[5.])
    mhe._build_module = False
    mhe.setup()
    return mhe


def test_mhe_matches_oracle():
    N = 6
    mhe = _mhe(N)
    spec = mhe.spec
    assert spec.nx == 3 and spec.nu == 2 and not spec.fixed_x0 and spec.lag_from == 1
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    rng = np.random.default_rng(0)
    B = 2
    # synthetic measurement history: true k = 1.7
    x_true = np.array([.6, .25])
    ys, us = [], []
    from oracle.estimators import rk4
    f = lambda x, u, p: np.array([u[0] * (1 - x[0]) - p[0] * x[0], p[0] * x[0] - u[0] * x[1]])
    for k in range(N):
        us.append([.4])
        ys.append([x_true[1] + rng.normal() * 5e-3])
        x_true = rk4(f, x_true, [.4], [1.7], .25, 20)
    sp = np.concatenate([np.array(us), np.array(ys)], axis=1)  # [N, nu + ny]
    sp = np.stack([sp, sp + rng.normal(size=sp.shape) * 1e-3])
    ap = np.array([[.5, .2, 1.], [.55, .25, 1.2]])
    x = np.tile(spec.x_guess, (B, N + 1, 1))
    u = np.tile(spec.u_guess, (B, N, 1))
    xc = np.tile(spec.x_guess, (B, N, spec.degree))
    out = host.solve(np.zeros((B, 3)), x, u, xc, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub, dt=.25, sp=sp, ap=ap,
                     tol=1e-10)
    assert (out['status'] == 1).all()
    nlp = FlatNlp(spec, exact_hessian=True)
    for b in range(B):
        res, xo, uo = nlp.solve(None, dt=.25, sps=sp[b], ap=ap[b], finish_barrier=True, tol=1e-10)
        assert res.status == 1
        np.testing.assert_allclose(out['x'][b], xo.T, atol=2e-7)
        np.testing.assert_allclose(out['u'][b], uo.T, atol=2e-7)
    # the parameter state is constant over the horizon and the estimate moved from the guess towards the truth
    k_est = out['x'][0, :, 2]
    assert np.ptp(k_est) < 1e-9 and abs(k_est[0] - 1.7) < abs(1. - 1.7)
