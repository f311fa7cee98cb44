"""GPU parity for the estimator rows of SURVEY.md §8a (a5 KF/EKF, a6 UKF, a7 PF) through the public API and the C ABI.
Golden numbers are the reference's own (tests/test_KFs.py); batched cases are compared with the NumPy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_hostsim_estimators import _cstr, kf_model, scalar_model  # model builders shared with the CPU tests


def test_kalman_filter_one_step_and_passthroughs(cuda_device):
    from hilo_mpc_b200 import KF
    model = kf_model()
    kf = KF(model)
    kf.setup()
    model.set_initial_conditions([.8, 0.])
    model.set_initial_parameter_values([.5, .4])
    kf.R = .064
    kf.Q = [.01, .01]
    kf.set_initial_guess([.8, 0.])
    kf.set_initial_parameter_values([.5, .4])
    model.simulate(u=.8)
    np.testing.assert_allclose(model.solution['x:f'], [[1.2], [.4]], rtol=1e-12)  # ERK1 of the linear model
    kf.estimate(y=.3894626, u=.8)
    np.testing.assert_allclose(kf.solution.get_by_id('x:f'), np.array([[1.19614861], [.39044856]]), rtol=1e-7)
    # tests/test_KFs.py:489-522
    kf2 = KF(model)
    kf2.setup()
    pred = kf2.predict(np.append(np.array([[.8], [0.]]), np.eye(2), axis=1), np.append([.8], [.5, .4]), .01 * np.eye(2))
    np.testing.assert_allclose(pred, np.array([[1.2, .26, .25], [.4, .25, .62]]))
    upd, y_pred = kf2.update(np.array([[1.2, .26, .25], [.4, .25, .62]]), .322052, np.append([.8], [.5, .4]), .064)
    np.testing.assert_allclose(upd, np.array([[1.17151023, .16862573, .023391813], [.32934538, .023391813, .0580117]]),
                               rtol=2e-7)
    np.testing.assert_allclose(y_pred, np.array([[.4]]))


def test_ekf_ukf_one_step_golden(cuda_device):
    from hilo_mpc_b200 import EKF, UKF
    for cls, ref in ((EKF, 7.206059), (UKF, 7.2739647)):
        est_ = cls(scalar_model())
        est_.setup()
        est_.Q = 10.
        est_.R = 1.
        est_.set_initial_guess(9.)
        est_.estimate(y=2.59109)
        np.testing.assert_allclose(est_.solution.get_by_id('x:f'), np.array([[ref]]), rtol=1e-7)


def test_kf_rejects_nonlinear_and_ekf_warns_on_linear():
    from hilo_mpc_b200 import EKF, KF, UKF
    with pytest.raises(ValueError):
        KF(scalar_model())
    with pytest.warns(UserWarning):
        EKF(kf_model())
    with pytest.warns(UserWarning):
        UKF(kf_model())


def test_batched_ukf_cstr_matches_oracle(cuda_device):
    import torch
    from hilo_mpc_b200 import UKF
    from oracle import estimators as est
    m, f, hfun, x_ss, u_ss = _cstr()
    B = 512
    rng = np.random.default_rng(2)
    x0 = x_ss + rng.normal(size=(B, 4)) * np.array([.05, .05, 1., 1.])
    ys = np.column_stack([x_ss[1] + rng.normal(size=B) * .02, x_ss[2] + rng.normal(size=B) * .3])
    Q, R = np.diag([1e-4, 1e-4, 1e-2, 1e-2]), np.diag([1e-3, 1e-1])
    ukf = UKF(m, alpha=.5)
    ukf.setup()
    ukf.Q, ukf.R = Q, R
    P0 = np.diag([1e-2, 1e-2, 1., 1.])
    ukf.set_initial_guess(x0, P0=P0)
    x, P = ukf.estimate(y=ys, u=u_ss)
    x, P = x.cpu().numpy(), P.cpu().numpy()
    for b in (0, 100, 511):
        xp, Pp, X = est.ukf_predict(f, x0[b], P0, u_ss, None, Q, dt=0.005, discrete=False, substeps=10, alpha=.5)
        xo, Po, _ = est.ukf_update(hfun, xp, Pp, X, ys[b], u_ss, None, R, alpha=.5)
        np.testing.assert_allclose(x[b], xo, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(P[b], Po, rtol=1e-7, atol=1e-12)


def test_particle_filter_indices_bit_exact(cuda_device):
    import torch
    from hilo_mpc_b200 import PF
    from oracle import estimators as est
    pf = PF(scalar_model())
    pf.setup(n_samples=200)
    pf.Q, pf.R = 10., 1.
    B, ns = 256, 200
    rng = np.random.default_rng(7)
    pf.set_initial_guess(np.full((B, 1), 9.))
    X = rng.normal(9., 1., size=(B, ns, 1))
    pf.set_particles(X)
    w = rng.normal(0., np.sqrt(10.), size=(B, ns, 1))
    v = rng.normal(0., 1., size=(B, ns, 1))
    unif = rng.random((B, ns))
    y = rng.uniform(1., 8., size=(B, 1))
    x, P = pf.estimate(y=y, w=w, v=v, uniforms=unif)
    q = pf.last_weights.cpu().numpy()
    idx = pf.last_indices.cpu().numpy()
    f, hfun = (lambda x_, u, p: x_ / 2 + 25 * x_ / (1 + x_ ** 2)), (lambda x_, u, p: x_ ** 2 / 20)
    for b in range(B):
        np.testing.assert_array_equal(idx[b], est.pf_resample_indices(q[b], unif[b]))
    for b in (0, 255):
        xm, Pm, _, qo, io = est.pf_step(f, hfun, X[b].T, y[b], None, None, w[b].T, v[b].T, np.array([[1.]]), unif[b])
        np.testing.assert_allclose(q[b], qo, rtol=1e-12)
        np.testing.assert_array_equal(idx[b], io)
        np.testing.assert_allclose(x[b].cpu().numpy(), xm, rtol=1e-12)


def test_particle_filter_warp_kernel_is_bit_identical_to_thread_per_filter(cuda_device):
    """One warp per filter with the particles on its lanes (est_lanes.cuh::PfLanes) keeps every running sum of the
    reference in particle order, so weights, indices, particles and moments equal the one-thread-per-filter kernel bit
    for bit - also with a particle count that is not a multiple of 32 and on a model with several states."""
    import torch
    from hilo_mpc_b200 import PF
    from hilo_mpc_b200.library import cstr_model
    rng = np.random.default_rng(11)
    for model, ns, B in ((scalar_model(), 200, 300), (cstr_model(dt=0.005), 48, 130)):
        nx, nu = model.n_x, model.n_u
        x0 = np.array([9.]) if nx == 1 else np.array([2.14, 1.09, 114.2, 112.9])
        Xp = x0 + (1. if nx == 1 else .05) * rng.standard_normal((B, ns, nx))
        w = rng.normal(0., .1, size=(B, ns, nx))
        v = rng.normal(0., .1, size=(B, ns, model.n_y))
        un = rng.random((B, ns))
        outs = []
        for variant in (0, 1):
            pf = PF(model)
            pf.setup(n_samples=ns)
            pf.Q, pf.R = (10., 1.) if nx == 1 else (np.eye(nx) * 1e-2, np.eye(model.n_y))
            pf.kernel_variant = variant
            pf.set_initial_guess(np.tile(x0, (B, 1)))
            pf.set_particles(Xp)
            y = np.full((B, 1), 3.2) if nx == 1 else np.tile([1.09, 114.2], (B, 1))
            kw = {} if nu == 0 else dict(u=np.tile([14.19, -1113.5], (B, 1)))
            x, P = pf.estimate(y=y, w=w, v=v, uniforms=un, **kw)
            torch.cuda.synchronize()
            outs.append([t.cpu().numpy().copy() for t in (x, P, pf.last_weights, pf.last_indices, pf.particles)])
        for a_, b_ in zip(*outs):
            np.testing.assert_array_equal(a_, b_)


def test_pf_resample_entry_matches_numpy_choice(cuda_device):
    """hb_pf_resample vs np.random.choice on the legacy global RNG (pf.py:406), including a large ragged case."""
    import ctypes as C
    import torch
    from hilo_mpc_b200 import _lib
    lib = _lib.load()
    for n in (15, 1000, 4097):
        B = 7
        rngq = np.random.default_rng(n)
        q = rngq.random((B, n))
        q /= q.sum(axis=1, keepdims=True)
        ref, unif = [], []
        for b in range(B):
            np.random.seed(100 + b)
            ref.append(np.random.choice(n, size=n, replace=True, p=q[b]))
            np.random.seed(100 + b)
            unif.append(np.random.random_sample(n))
        qt = torch.tensor(q, device=cuda_device)
        ut = torch.tensor(np.stack(unif), device=cuda_device)
        cdf = torch.empty_like(qt)
        idx = torch.zeros(B, n, dtype=torch.int64, device=cuda_device)
        lib.hb_pf_resample.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = lib.hb_pf_resample(B, n, qt.data_ptr(), ut.data_ptr(), cdf.data_ptr(), idx.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        np.testing.assert_array_equal(idx.cpu().numpy(), np.stack(ref))


def test_pf_resampling_indices_equal_the_executed_reference_block(cuda_device):
    """tests/golden/reference_numeric.json: ``ind`` of pf.py:406 obtained by executing the reference's own source with the
    legacy RNG seeded; hb_pf_resample fed with the uniforms that seed produces must return the same integers."""
    import ctypes as C
    import json
    import os
    import torch
    from hilo_mpc_b200 import _lib
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_numeric.json')))
    lib = _lib.load()
    for c in gold['pf_resample_block']['cases']:
        n = c['n_s']
        np.random.seed(c['seed'])
        unif = np.random.random_sample(n)
        qt = torch.tensor(np.array(c['q'])[None], device=cuda_device)
        ut = torch.tensor(unif[None], device=cuda_device)
        cdf = torch.empty_like(qt)
        idx = torch.zeros(1, n, dtype=torch.int64, device=cuda_device)
        rc = lib.hb_pf_resample(1, n, qt.data_ptr(), ut.data_ptr(), cdf.data_ptr(), idx.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        np.testing.assert_array_equal(idx.cpu().numpy()[0], np.array(c['ind']))


def test_particle_filter_roughening_matches_oracle(cuda_device):
    """pf.py:411-421: roughening after the resampling.  With explicit draws the engine must equal the oracle; with the
    legacy global RNG seeded it must draw in the reference's order (process noise, measurement noise, resampling
    uniforms, roughening jitter), i.e. equal the oracle fed with the numbers the same seed produces."""
    from hilo_mpc_b200 import PF
    from oracle import estimators as est
    f, hfun = (lambda x_, u, p: x_ / 2 + 25 * x_ / (1 + x_ ** 2)), (lambda x_, u, p: x_ ** 2 / 20)
    ns = 50
    rng = np.random.default_rng(3)
    X0 = rng.normal(9., 1., size=(1, ns, 1))
    y = np.array([[3.2]])
    # explicit draws, batch of 2 identical filters
    pf = PF(scalar_model(), roughening=True, K=.3)
    pf.setup(n_samples=ns)
    pf.Q, pf.R = 10., 1.
    pf.set_initial_guess(np.full((2, 1), 9.))
    pf.set_particles(np.repeat(X0, 2, axis=0))
    w = rng.normal(0., np.sqrt(10.), size=(1, ns, 1))
    v = rng.normal(0., 1., size=(1, ns, 1))
    unif = rng.random((1, ns))
    jit = rng.normal(0., .1, size=(1, ns, 1))
    rep = lambda a: np.repeat(a, 2, axis=0)
    x, P = pf.estimate(y=rep(y), w=rep(w), v=rep(v), uniforms=rep(unif), roughening_noise=rep(jit))
    xm, Pm, Xr, _, _ = est.pf_step(f, hfun, X0[0].T, y[0], None, None, w[0].T, v[0].T, np.array([[1.]]), unif[0],
                                   roughening_noise=jit[0].T)
    np.testing.assert_allclose(x[0].cpu().numpy(), xm, rtol=1e-12)
    np.testing.assert_allclose(P[1].cpu().numpy().reshape(-1), np.atleast_1d(Pm).reshape(-1), rtol=1e-12)
    np.testing.assert_allclose(pf.particles[0].cpu().numpy(), Xr.T, rtol=1e-12)
    # legacy global RNG, single filter: same seed -> same numbers in the reference's order
    pf1 = PF(scalar_model(), roughening=True, K=.2)
    pf1.setup(n_samples=ns)
    pf1.Q, pf1.R = 10., 1.
    pf1.set_initial_guess(np.array([9.]))
    pf1.set_particles(X0[0].T)
    np.random.seed(77)
    pf1.estimate(y=y[0])
    np.random.seed(77)
    w1 = est.lhsnorm(np.zeros(1), np.array([[10.]]), ns)      # the filter's default density (pf.py:80, :363)
    v1 = (np.sqrt(np.array([[1.]])) @ np.random.randn(1, ns)).T
    u1 = np.random.random_sample(ns)
    _, _, Xres, _, _ = est.pf_step(f, hfun, X0[0].T, y[0], None, None, w1.T, v1.T, np.array([[1.]]), u1)
    j1 = est.lhsnorm(np.zeros(1), est.pf_roughening_covariance(Xres, .2), ns)
    xm1, Pm1, Xr1, _, _ = est.pf_step(f, hfun, X0[0].T, y[0], None, None, w1.T, v1.T, np.array([[1.]]), u1,
                                      roughening_noise=j1.T)
    np.testing.assert_allclose(pf1.particles[0].cpu().numpy(), Xr1.T, rtol=1e-11)
    np.testing.assert_allclose(np.asarray(pf1.solution['x'])[:, -1], xm1, rtol=1e-11)


def test_particle_filter_prior_editing_matches_oracle(cuda_device):
    """pf.py:378-398: particles whose predicted measurement is more than 6 sigma off are jittered before the propagation
    until none is left.  Legacy global RNG seeded: the engine must draw in the reference's order (process noise,
    measurement noise, editing jitter round by round, resampling uniforms) and equal the oracle's restatement."""
    from hilo_mpc_b200 import PF
    from oracle import estimators as est
    f, hfun = (lambda x_, u, p: x_ / 2 + 25 * x_ / (1 + x_ ** 2)), (lambda x_, u, p: x_ ** 2 / 20)
    ns, R, K = 40, .04, .5
    rng = np.random.default_rng(5)
    X0 = rng.normal(9., 2.5, size=(ns, 1))
    y = np.array([3.6])
    for variant in (1, 0):
        pf = PF(scalar_model(), prior_editing=True, K=K)
        pf.kernel_variant = variant
        pf.setup(n_samples=ns)
        pf.Q, pf.R = 1., R
        pf.set_initial_guess(np.array([9.]))
        pf.set_particles(X0.T)
        np.random.seed(123)
        pf.estimate(y=y)
        assert pf.last_prior_editing_rounds >= 1          # the case does exercise the loop
        np.random.seed(123)
        w1 = est.lhsnorm(np.zeros(1), np.array([[1.]]), ns)
        v1 = (np.sqrt(np.array([[R]])) @ np.random.randn(1, ns)).T
        Xe = est.pf_prior_editing(f, hfun, X0.T, y, None, None, w1.T, v1.T, np.array([[R]]), K, est.lhsnorm)
        u1 = np.random.random_sample(ns)
        xm, Pm, Xr, q1, i1 = est.pf_step(f, hfun, Xe, y, None, None, w1.T, v1.T, np.array([[R]]), u1)
        np.testing.assert_array_equal(pf.last_indices[0].cpu().numpy(), i1)
        np.testing.assert_allclose(pf.particles[0].cpu().numpy(), Xr.T, rtol=1e-11)
        np.testing.assert_allclose(np.asarray(pf.solution['x'])[:, -1], xm, rtol=1e-11)


def test_ukf_bioreactor_sigma_point_golden(cuda_device):
    """tests/test_KFs.py:716-756 through the reference's positional predict / update pass-throughs (kf.py:309-325)."""
    from hilo_mpc_b200 import UKF
    from test_hostsim_estimators import bioreactor_model
    from test_oracle_golden import UKF_BIO_P, UKF_BIO_PRED, UKF_BIO_UPD
    ukf = UKF(bioreactor_model(), alpha=1.)
    ukf.setup()
    x = np.array([[299.876], [.217108], [20.]])
    prediction = ukf.predict(np.append(x, np.eye(3), axis=1), np.append([.01], UKF_BIO_P), np.zeros((3, 3)))
    np.testing.assert_allclose(prediction, UKF_BIO_PRED, atol=1e-3)
    update, y_pred = ukf.update(UKF_BIO_PRED, np.array([[300.941], [.245805]]), np.append([.01], UKF_BIO_P),
                                np.diag([.25, .01]))
    np.testing.assert_allclose(update, UKF_BIO_UPD, atol=1e-3)
    np.testing.assert_allclose(np.asarray(y_pred).reshape(-1), [299.925, .219434], atol=1e-3)
