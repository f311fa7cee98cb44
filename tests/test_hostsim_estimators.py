"""Estimator kernel logic on the CPU (host build of the same kernel bodies, test infrastructure): the reference's golden
numbers (tests/test_KFs.py) and oracle parity for UKF / PF on a nonlinear model."""
import numpy as np
import pytest

from hilo_mpc_b200 import Model
from hilo_mpc_b200.estgen import build_est_module
from oracle import estimators as est
from est_runner import HostEst


def kf_model():
    model = Model()
    model.set_equations(equations="""
        dx_1/dt = -k_1*x_1(t) + u(k)
        dx_2/dt = k_1*x_1(t) - k_2*x_2(t)
        y(k) = x_2(t)
        """)
    model.discretize('erk', order=1, inplace=True)
    model.setup(dt=1.)
    return model


def scalar_model():
    model = Model(discrete=True)
    model.set_dynamical_states('x')
    model.set_measurements('y')
    model.set_dynamical_equations('x/2 + 25*dt*x/(1 + x^2)')
    model.set_measurement_equations('x^2/20')
    model.setup(dt=1.)
    return model


def test_parser_builds_the_kf_model():
    m = kf_model()
    assert m.dynamical_state_names == ['x_1', 'x_2'] and m.input_names == ['u']
    assert m.parameter_names == ['k_1', 'k_2'] and m.measurement_names == ['y'] and m.discrete and m.is_linear()


def test_kf_predict_update_golden():  # tests/test_KFs.py:489-522
    h = HostEst(build_est_module(kf_model(), hostsim=True))
    out = h.call('ekf', [.8, 0.], P=np.eye(2), y=[0.], u=[.8], p=[.5, .4], Q=.01 * np.eye(2), R=[[1.]],
                 do_update=False)
    np.testing.assert_allclose(np.column_stack([out['x'][0], out['P'][0]]), [[1.2, .26, .25], [.4, .25, .62]], rtol=1e-12)
    out = h.call('ekf', [1.2, .4], P=[[.26, .25], [.25, .62]], y=[.322052], u=[.8], p=[.5, .4], Q=np.eye(2),
                 R=[[.064]], do_predict=False)
    np.testing.assert_allclose(np.column_stack([out['x'][0], out['P'][0]]),
                               [[1.17151023, .16862573, .023391813], [.32934538, .023391813, .0580117]], rtol=2e-7)
    np.testing.assert_allclose(out['y_pred'], [[.4]])


def test_kf_one_step_golden():  # tests/test_KFs.py:298-335
    h = HostEst(build_est_module(kf_model(), hostsim=True))
    out = h.call('ekf', [.8, 0.], P=np.eye(2), y=[.3894626], u=[.8], p=[.5, .4], Q=.01 * np.eye(2), R=[[.064]])
    np.testing.assert_allclose(out['x'][0], [1.19614861, .39044856], rtol=1e-7)


def test_ekf_and_ukf_one_step_golden():  # tests/test_KFs.py:548-566, 606-624
    h = HostEst(build_est_module(scalar_model(), hostsim=True))
    out = h.call('ekf', [9.], P=[[1.]], y=[2.59109], Q=[[10.]], R=[[1.]])
    np.testing.assert_allclose(out['x'][0], [7.206059], rtol=1e-7)
    out = h.call('ukf', [9.], P=[[1.]], y=[2.59109], Q=[[10.]], R=[[1.]])
    np.testing.assert_allclose(out['x'][0], [7.2739647], rtol=1e-7)


def _cstr():
    from hilo_mpc_b200.library import cstr_model, cstr_rhs, cstr_steady_state
    m = cstr_model(dt=0.005)
    f = lambda x, u, p: np.array(cstr_rhs(x, u, exp=np.exp))
    hfun = lambda x, u, p: np.array([x[1], x[2]])
    x_ss, u_ss = cstr_steady_state()
    return m, f, hfun, x_ss, u_ss


def test_ukf_and_ekf_cstr_match_oracle():
    """Continuous 4-state CSTR, non-diagonal P (exercises the upper-factor column convention)."""
    m, f, hfun, x_ss, u_ss = _cstr()
    h = HostEst(build_est_module(m, hostsim=True))
    rng = np.random.default_rng(0)
    A = rng.normal(size=(4, 4)) * np.array([.05, .05, .5, .5])
    P0 = A @ A.T + np.diag([1e-3, 1e-3, 1e-1, 1e-1])
    Q, R = np.diag([1e-4, 1e-4, 1e-2, 1e-2]), np.diag([1e-3, 1e-1])
    x0 = x_ss + np.array([.05, -.03, 1., -1.])
    y = np.array([x_ss[1], x_ss[2]])
    for kind in ('ukf', 'ekf'):
        out = h.call(kind, x0, P=P0, y=y, u=u_ss, Q=Q, R=R, dt=0.005, substeps=10, alpha=.5)
        if kind == 'ukf':
            xp, Pp, X = est.ukf_predict(f, x0, P0, u_ss, None, Q, dt=0.005, discrete=False, substeps=10, alpha=.5)
            xo, Po, _ = est.ukf_update(hfun, xp, Pp, X, y, u_ss, None, R, alpha=.5)
        else:
            xp, Pp = est.ekf_predict(f, x0, P0, u_ss, None, Q, dt=0.005, discrete=False, substeps=10)
            xo, Po, _ = est.ekf_update(hfun, xp, Pp, y, u_ss, None, R)
        assert out['status'][0] == 0
        np.testing.assert_allclose(out['x'][0], xo, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out['P'][0], Po, rtol=1e-7, atol=1e-12)


def test_pf_step_indices_bit_exact_and_moments():
    m = scalar_model()
    h = HostEst(build_est_module(m, hostsim=True))
    f, hfun = (lambda x, u, p: x / 2 + 25 * x / (1 + x ** 2)), (lambda x, u, p: x ** 2 / 20)
    rng = np.random.default_rng(5)
    B, ns = 3, 64
    X = rng.normal(9., 1., size=(B, ns, 1))
    w = rng.normal(0., np.sqrt(10.), size=(B, ns, 1))
    v = rng.normal(0., 1., size=(B, ns, 1))
    unif = rng.random((B, ns))
    y = np.array([[2.5], [4.], [7.]])
    out = h.call('pf', np.zeros((B, 1)), P=np.zeros((B, 1, 1)), X=X, y=y, Q=[[10.]], R=[[1.]], w=w, v=v, unif=unif)
    for b in range(B):
        xm, P, Xr, q, idx = est.pf_step(f, hfun, X[b].T, y[b], None, None, w[b].T, v[b].T, np.array([[1.]]), unif[b])
        np.testing.assert_allclose(out['q'][b], q, rtol=1e-12)
        # indices are integer work: bit-exact given the kernel's own weights and the same uniforms
        np.testing.assert_array_equal(out['idx'][b], est.pf_resample_indices(out['q'][b], unif[b]))
        np.testing.assert_array_equal(out['idx'][b], idx)
        np.testing.assert_allclose(out['x'][b], xm, rtol=1e-12)
        np.testing.assert_allclose(out['P'][b], np.atleast_2d(P), rtol=1e-10)


def test_pf_prediction_only_pass_and_prior_editing_loop():
    """``do_update = 0`` (the pass the prior-editing loop repeats, pf.py:378-398) leaves particles and estimate alone and
    delivers the predicted noisy measurements; the editing loop driven with it equals the oracle's restatement."""
    m = scalar_model()
    h = HostEst(build_est_module(m, hostsim=True))
    f, hfun = (lambda x, u, p: x / 2 + 25 * x / (1 + x ** 2)), (lambda x, u, p: x ** 2 / 20)
    rng = np.random.default_rng(5)
    ns, R, K = 40, .04, .5
    X0 = rng.normal(9., 2.5, size=(1, ns, 1))
    y = np.array([[3.6]])
    w = rng.normal(0., 1., size=(1, ns, 1))
    v = rng.normal(0., np.sqrt(R), size=(1, ns, 1))
    x_in = np.full((1, 1), 9.)
    out = h.call('pf', x_in, P=np.ones((1, 1, 1)), X=X0, y=y, Q=[[1.]], R=[[R]], w=w, v=v, unif=np.zeros((1, ns)),
                 do_update=False)
    np.testing.assert_array_equal(out['X'], X0)                    # particles untouched
    np.testing.assert_array_equal(out['x'], x_in)                  # estimate untouched
    Ys = out['work'][ns + ns:ns + 2 * ns]                          # Xs | cdf | Ys  (csrc/est_module.cuh)
    Yo = np.array([hfun(f(X0[0, k, 0], None, None), None, None) for k in range(ns)]) + v[0, :, 0]
    np.testing.assert_allclose(Ys, Yo, rtol=1e-13)
    # the loop of ParticleFilter._edit_prior on the host kernels, jitter drawn like the oracle draws it
    draws = np.random.RandomState(3)
    pdf = lambda mu, cov, n: draws.multivariate_normal(mu, cov, size=n)
    X = X0.copy()
    rounds = 0
    while True:
        o = h.call('pf', x_in, P=np.ones((1, 1, 1)), X=X, y=y, Q=[[1.]], R=[[R]], w=w, v=v, unif=np.zeros((1, ns)),
                   do_update=False)
        need = np.abs(y[0, 0] - o['work'][2 * ns:3 * ns]) > 6. * np.sqrt(R)
        if not need.any():
            break
        rounds += 1
        spread = X[0].max(axis=0) - X[0].min(axis=0)
        X[0, need] += pdf(np.zeros(1), K * np.diag(spread) * float(need.sum()) ** -1., int(need.sum()))
    draws = np.random.RandomState(3)
    Xe = est.pf_prior_editing(f, hfun, X0[0].T, y[0], None, None, w[0].T, v[0].T, np.array([[R]]), K, pdf)
    assert rounds >= 1
    np.testing.assert_allclose(X[0].T, Xe, rtol=1e-13)


def bioreactor_model():
    """tests/test_KFs.py:695-712 (equation-string front end; parameters in order of appearance)."""
    m = Model()
    m.set_equations(equations="""
        dT/dt = alpha*(T_amb - T(t))
        dcB/dt = r - D(k)*cB(t)
        dcS/dt = -r/Y - D(k)*cS(t)

        y_0(k) = T(t)
        y_1(k) = cB(t)

        r = (mu_0 + mu_1*T(t))*cS(t)*cB(t)/(K + cS(t))
        """)
    m.setup(dt=.1)
    return m


def test_ukf_bioreactor_sigma_point_golden_through_the_kernel_bodies():
    """tests/test_KFs.py:716-756 (atol 1e-3 there; the reference integrates with CVODES, the kernel with 10 RK4 steps)."""
    from test_oracle_golden import UKF_BIO_P, UKF_BIO_PRED, UKF_BIO_UPD
    m = bioreactor_model()
    assert m.parameter_names == ['alpha', 'T_amb', 'mu_0', 'mu_1', 'K', 'Y']
    h = HostEst(build_est_module(m, hostsim=True))
    out = h.call('ukf', [299.876, .217108, 20.], P=np.eye(3), y=[0., 0.], u=[.01], p=UKF_BIO_P, Q=np.zeros((3, 3)),
                 R=np.eye(2), dt=.1, substeps=10, alpha=1., do_update=False)
    got = np.concatenate([out['x'][0][:, None], out['P'][0], out['X'][0].T], axis=1)
    np.testing.assert_allclose(got, UKF_BIO_PRED, atol=1e-3)
    out = h.call('ukf', UKF_BIO_PRED[:, 0], P=UKF_BIO_PRED[:, 1:4], X=UKF_BIO_PRED[:, 4:].T, y=[300.941, .245805],
                 u=[.01], p=UKF_BIO_P, Q=np.zeros((3, 3)), R=np.diag([.25, .01]), dt=.1, alpha=1., do_predict=False)
    np.testing.assert_allclose(np.concatenate([out['x'][0][:, None], out['P'][0]], axis=1), UKF_BIO_UPD, atol=1e-3)
    np.testing.assert_allclose(out['y_pred'][0], [299.925, .219434], atol=1e-3)
