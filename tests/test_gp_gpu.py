"""GPU parity for the GP row: golden LML of the reference (tests/test_GPs.py:354-363), SE Gram golden
(tests/test_kernels.py), predictions vs the NumPy/SciPy oracle at the BASELINE size (n=512, d=6), self-consistency
properties the reference tests (noise-free variance < noisy variance, tests/test_GPs.py:654-699)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gp_log_marginal_likelihood_golden(cuda_device):
    from hilo_mpc_b200.gp import GP
    gp = GP(['x', 'y'], 'z')
    X = np.array([[0., .5, 1. / np.sqrt(2.), np.sqrt(3.) / 2., 1., 0.],
                  [1., np.sqrt(3.) / 2., 1. / np.sqrt(2.), .5, 0., -1.]])
    y = np.array([[0., np.pi / 6., np.pi / 4., np.pi / 3., np.pi / 2., np.pi]])
    gp.set_training_data(X, y)
    gp.setup()
    np.testing.assert_approx_equal(gp.log_marginal_likelihood(), -9.82229944, significant=9)


@pytest.mark.parametrize('n,d', [(512, 6), (100, 3), (37, 1)])
def test_gp_fit_and_predict_match_oracle(cuda_device, n, d):
    from hilo_mpc_b200.gp import GP, Kernel
    from oracle.gp import GpOracle
    rng = np.random.default_rng(n)
    X = rng.uniform(-1., 1., size=(d, n))
    ls = rng.uniform(.5, 1.5, size=d)
    ytrue = np.sin(X.sum(0)) + .1 * X[0] ** 2
    y = ytrue + rng.normal(size=n) * 1e-2
    gp = GP([f'f{k}' for k in range(d)], 'y', kernel=Kernel.squared_exponential(signal_variance=1.3, length_scales=ls,
                                                                                 ard=True), noise_variance=1e-4)
    gp.set_training_data(X, y[None])
    gp.setup()
    orc = GpOracle(X, y, noise_variance=1e-4, signal_variance=1.3, length_scales=ls)
    np.testing.assert_allclose(gp.log_marginal_likelihood(), orc.log_marginal_likelihood(), rtol=1e-9)
    L = gp._fit['L'].cpu().numpy()
    np.testing.assert_allclose(L, orc.L, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(gp._fit['alpha_host'], orc.alpha, rtol=1e-5, atol=1e-6)
    Xq = rng.uniform(-1., 1., size=(d, 1000 + 3))
    mean, var = gp.predict(Xq)
    mo, vo = orc.predict(Xq)
    np.testing.assert_allclose(mean[0], mo, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(var[0], vo, rtol=1e-6, atol=1e-9)
    _, var_nf = gp.predict(Xq, noise_free=True)
    assert (var_nf < var).all()
    # symbolic prediction (how a GP enters a Model) evaluates to the same mean
    from hilo_mpc_b200 import symbolic as sx
    xs = sx.sym('x', d)
    ms, _ = gp.predict(xs)
    F = sx.Function('m', [('x', xs)], [('m', ms)])
    np.testing.assert_allclose(F(Xq[:, :5])[0], mo[:5], rtol=1e-7, atol=1e-8)


def test_gp_inside_nmpc_matches_oracle(cuda_device):
    """BASELINE config 3 (GP-output-feedback NMPC) at a size the oracle finishes in seconds: the CSTR's side-reaction rate
    is replaced by the posterior mean of a GP fitted on the GPU (`hb_gp_fit`), embedded symbolically like the reference
    does (dynamic_model.py:3040-3080), compiled into the node kernels, and the batched solve is compared with the oracle
    on the same NLP.  The learned model must also be a good surrogate: first moves within 2 % of the true-model NMPC."""
    import torch
    from hilo_mpc_b200.examples import cstr_gp_nmpc, cstr_initial_states, cstr_nmpc, cstr_side_reaction_gp
    from hilo_mpc_b200.library import cstr_side_reaction
    from oracle.nmpc import FlatNlp
    gp = cstr_side_reaction_gp(n_train=24)
    rng = np.random.default_rng(0)
    Xq = np.vstack([rng.uniform(.8, 1.8, 200), rng.uniform(115., 135., 200)])
    mean, var = gp.predict(Xq)
    truth = cstr_side_reaction(Xq[0], Xq[1])
    assert np.abs(mean[0] - truth).max() < 0.05 * np.abs(truth).max()
    assert (var > 0).all()
    nmpc = cstr_gp_nmpc(gp, N=10)
    x0 = cstr_initial_states(16, seed=3)
    u0 = nmpc.optimize(torch.tensor(x0, device=cuda_device))
    assert (nmpc.solver_status_code == 1).all()
    x_pred, u_pred, _ = nmpc.return_prediction()
    xs, us = np.asarray(nmpc._xs_full), np.asarray(nmpc._us_full)
    nlp = FlatNlp(nmpc.spec, exact_hessian=True)
    for b in (0, 9):
        res, xo, uo = nlp.solve(x0[b] / xs, dt=nmpc.sampling_interval, finish_barrier=True)
        assert res.status == 1
        np.testing.assert_allclose(x_pred[b].cpu().numpy() / xs, xo.T, atol=2e-6)
        np.testing.assert_allclose(u_pred[b].cpu().numpy() / us, uo.T, atol=2e-5)
    true_nmpc = cstr_nmpc(N=10)
    u_true = true_nmpc.optimize(torch.tensor(x0, device=cuda_device))
    rel = ((u0 - u_true).abs() / torch.tensor(us, device=cuda_device)).max().item()
    assert rel < 0.02 * (u_true.abs() / torch.tensor(us, device=cuda_device)).max().item() + 0.02


def test_config3_gp_with_512_training_points_as_device_function(cuda_device):
    """BASELINE config 3 at its stated GP size: 512 training points.  The posterior mean enters the model as ONE call
    node; the generated node kernels evaluate it (value, gradient, Hessian) with a loop over the training table in
    constant memory (csrc/gp_device.cuh) instead of 512 unrolled terms.  The device function must equal hb_gp_predict on
    the same points, and the batched NMPC solve must equal the oracle's solve of the same flat NLP."""
    import torch
    from hilo_mpc_b200 import symbolic as sx
    from hilo_mpc_b200.examples import cstr_gp_nmpc, cstr_initial_states, cstr_side_reaction_gp
    from hilo_mpc_b200.library import cstr_side_reaction
    from oracle.nmpc import FlatNlp
    gp = cstr_side_reaction_gp(n_train=512)
    xq = sx.sym('q', 2)
    mean_sym, var_sym = gp.predict(xq)
    assert var_sym is None and any(n.op.startswith('gp:') for n in sx.topo_sort(list(mean_sym)))
    rng = np.random.default_rng(0)
    Xq = np.vstack([rng.uniform(.8, 1.8, 100), rng.uniform(115., 135., 100)])
    F = sx.Function('m', [('q', xq)], [('m', mean_sym)])
    mean_dev, _ = gp.predict(Xq, mean_only=True)
    np.testing.assert_allclose(np.asarray(F(Xq)).reshape(-1), mean_dev.reshape(-1), rtol=1e-9, atol=1e-9)
    truth = cstr_side_reaction(Xq[0], Xq[1])
    assert np.abs(mean_dev[0] - truth).max() < 0.02 * np.abs(truth).max()
    nmpc = cstr_gp_nmpc(gp, N=10)
    x0 = cstr_initial_states(32, seed=3)
    nmpc.optimize(torch.tensor(x0, device=cuda_device))
    assert (nmpc.solver_status_code == 1).all()
    x_pred, u_pred, _ = nmpc.return_prediction()
    xs, us = np.asarray(nmpc._xs_full), np.asarray(nmpc._us_full)
    nlp = FlatNlp(nmpc.spec, exact_hessian=True)
    for b in (0, 9, 31):
        res, xo, uo = nlp.solve(x0[b] / xs, dt=nmpc.sampling_interval, finish_barrier=True)
        assert res.status == 1
        np.testing.assert_allclose(x_pred[b].cpu().numpy() / xs, xo.T, atol=2e-6)
        np.testing.assert_allclose(u_pred[b].cpu().numpy() / us, uo.T, atol=2e-5)


def test_gp_fit_model_reproduces_gpml_golden(cuda_device):
    """Hyper-parameter training: tests/test_GPs.py:833-849 - 20 points from the reference's LCG
    generator, default SE kernel and start values, fitted [noise_variance, length_scale, signal_variance] must equal the
    gpml toolbox values to rtol 1e-5; the same through the oracle, and a fixed hyper-parameter stays put."""
    from hilo_mpc_b200.gp import GP
    from oracle.gp import fit_hyperparameters, lcg_randn
    x = lcg_randn(.8, (20, 1))
    y = np.sin(3 * x) + .1 * lcg_randn(.9, (20, 1))
    gp = GP(['x'], ['y'])
    gp.set_training_data(x.T, y.T)
    gp.setup()
    lml0 = gp.log_marginal_likelihood()
    gp.fit_model()
    vals = [p.value for p in gp.hyperparameters]
    np.testing.assert_allclose(vals, [.0085251, .5298217, .8114553], rtol=1e-5)
    sn2, ls, sf2, res = fit_hyperparameters(x.T, y.T)
    np.testing.assert_allclose(vals, [sn2, ls[0], sf2], rtol=1e-6)
    assert gp.log_marginal_likelihood() > lml0
    np.testing.assert_allclose(gp.log_marginal_likelihood(), -res.fun, rtol=1e-10)
    # predictions use the fitted values
    mean, var = gp.predict(np.linspace(-3, 3, 61).reshape(1, -1))
    assert np.isfinite(mean).all() and (var > 0).all()
    gp2 = GP(['x'], ['y'], noise_variance=.1 ** 2)
    gp2.set_training_data(x.T, y.T)
    gp2.setup()
    gp2.fit_model(fixed=('noise_variance',))
    assert gp2.noise_variance == .1 ** 2
    assert gp2.log_marginal_likelihood() <= gp.log_marginal_likelihood() + 1e-9


@pytest.mark.parametrize('name,gold', [('matern_32', [.0088262, .8329355, .9398366]),
                                       ('matern_52', [.0086694, .7180206, .9571137])])
def test_gp_matern_kernels_reproduce_goldens(cuda_device, name, gold):
    """Matern kernels on the device: Gram goldens of tests/test_kernels.py (through the fit's Cholesky factor), fitted
    hyper-parameters = gpml values of tests/test_GPs.py:862-867 (rtol 1e-5), predictions = oracle."""
    from hilo_mpc_b200.gp import GP, Kernel
    from oracle.gp import GpOracle, lcg_randn, stationary_kernel
    kind = {'matern_32': 'M32', 'matern_52': 'M52'}[name]
    # Gram: L L^T - noise I must equal the reference's covariance matrix
    x5 = np.array([[1., 2., 3., 4., 5.]])
    g5 = GP(['x'], ['y'], kernel=getattr(Kernel, name)(), noise_variance=.5)
    g5.set_training_data(x5, np.zeros((1, 5)))
    g5.setup()
    L = g5._fit['L'].cpu().numpy()
    np.testing.assert_allclose(L @ L.T - .5 * np.eye(5), stationary_kernel(x5, kind=kind), atol=1e-14)
    np.testing.assert_allclose(getattr(Kernel, name)()(x5), stationary_kernel(x5, kind=kind), atol=1e-15)
    if kind == 'M32':
        np.testing.assert_allclose((L @ L.T)[0, 1:], [.48335772, .13973135, .03431324, .007767734], rtol=2e-7)
    # training
    x = lcg_randn(.8, (20, 1))
    y = np.sin(3 * x) + .1 * lcg_randn(.9, (20, 1))
    gp = GP(['x'], ['y'], kernel=getattr(Kernel, name)())
    gp.set_training_data(x.T, y.T)
    gp.setup()
    gp.fit_model()
    vals = [p.value for p in gp.hyperparameters]
    np.testing.assert_allclose(vals, gold, rtol=1e-5)
    # prediction against the oracle at the fitted values
    xq = np.linspace(-3, 3, 61).reshape(1, -1)
    mean, var = gp.predict(xq)
    orc = GpOracle(x.T, y.T, noise_variance=vals[0], length_scales=vals[1], signal_variance=vals[2], kind=kind)
    mo, vo = orc.predict(xq)
    np.testing.assert_allclose(mean[0], mo, atol=1e-9)
    np.testing.assert_allclose(var[0], vo, atol=1e-9)


def test_gp_exponential_kernel_training(cuda_device):
    """tests/test_GPs.py:854-860: the noise variance collapses (< 1e-5), [length_scale, signal_variance] are pinned."""
    from hilo_mpc_b200.gp import GP, Kernel
    from oracle.gp import lcg_randn
    x = lcg_randn(.8, (20, 1))
    y = np.sin(3 * x) + .1 * lcg_randn(.9, (20, 1))
    gp = GP(['x'], ['y'], kernel=Kernel.exponential())
    gp.set_training_data(x.T, y.T)
    gp.setup()
    gp.fit_model()
    vals = [p.value for p in gp.hyperparameters]
    assert vals[0] < 1e-5
    np.testing.assert_allclose(vals[1:], [.9769135, .5935212], rtol=1e-5)


def test_gp_rasmussen_regression_goldens(cuda_device):
    """tests/test_GPs.py:1095-1124: LML of the Matern-3/2 GP with the (here pre-subtracted) mean 0.5 x + 1 = -11.9706317,
    fitted hyper-parameters of the default SE model = gpml values to rtol 1e-6."""
    from hilo_mpc_b200.gp import GP, Kernel
    from test_oracle_golden import rasmussen_data
    x, y, mu = rasmussen_data()
    gp = GP(['x'], ['y'], kernel=Kernel.matern_32(length_scales=.25), noise_variance=.1 ** 2)
    gp.set_training_data(x.T, (y - mu).T)
    gp.setup()
    np.testing.assert_approx_equal(gp.log_marginal_likelihood(), -11.9706317, significant=9)
    gp1 = GP(['x'], ['y'], noise_variance=.1 ** 2)
    gp1.set_training_data(x.T, y.T)
    gp1.setup()
    gp1.fit_model()
    np.testing.assert_allclose([p.value for p in gp1.hyperparameters], [.0222571, .3703166, 3.9427837], rtol=1e-6)


def test_gp_mean_functions(cuda_device):
    """Mean objects (gp/mean.py): LML golden of tests/test_GPs.py:1120 with the mean 0.5 x + 1 given as
    Mean.linear(fixed) + Mean.one(); all five hyper-parameters of the linear + constant mean model fitted to the gpml
    values (:1125-1126, rtol 1e-6); predictions carry the mean."""
    from hilo_mpc_b200 import Mean
    from hilo_mpc_b200.gp import GP, Kernel
    from oracle.gp import GpOracle
    from test_oracle_golden import rasmussen_data
    x, y, mu = rasmussen_data()
    mean = Mean.linear(coefficient=.5, bounds={'coefficient': 'fixed'}) + Mean.one()
    np.testing.assert_allclose(mean(x.T), mu.T, rtol=1e-15)
    gp = GP(['x'], ['y'], mean=mean, kernel=Kernel.matern_32(length_scales=.25), noise_variance=.1 ** 2)
    gp.set_training_data(x.T, y.T)
    gp.setup()
    np.testing.assert_approx_equal(gp.log_marginal_likelihood(), -11.9706317, significant=9)
    xq = np.linspace(-1.9, 1.9, 101).reshape(1, -1)
    m, v = gp.predict(xq)
    mo, vo = GpOracle(x.T, (y - mu).T, noise_variance=.01, length_scales=.25, kind='M32').predict(xq)
    np.testing.assert_allclose(m[0], mo + (.5 * xq[0] + 1.), atol=1e-10)
    np.testing.assert_allclose(v[0], vo, atol=1e-10)
    gp2 = GP(['x'], ['y'], mean=Mean.linear(coefficient=0.) + Mean.constant(bias=0.), noise_variance=.1 ** 2)
    gp2.set_training_data(x.T, y.T)
    gp2.setup()
    gp2.fit_model()
    np.testing.assert_allclose([p.value for p in gp2.hyperparameters],
                               [.02183, 1.1918832, 1.4624534, .2201449, .4017997], rtol=1e-6)


@pytest.mark.parametrize('kernel', ['squared_exponential', 'exponential', 'matern_32', 'matern_52'])
def test_lml_gradient_kernel_matches_finite_differences(cuda_device, kernel):
    """``hb_gp_lml_grad`` (own GEMM + one reduction pass over the pairs) against central differences of the negative log
    marginal likelihood in the log hyper-parameters, ARD length scales, every stationary kernel family."""
    from hilo_mpc_b200.gp import GP, Kernel
    rng = np.random.default_rng(17)
    n, d = 70, 3
    X = rng.uniform(-1., 1., size=(d, n))
    y = (np.sin(2. * X[0]) + X[1] * X[2] + .05 * rng.normal(size=n))[None, :]
    gp = GP([f'f{k}' for k in range(d)], 'y', noise_variance=.04,
            kernel=getattr(Kernel, kernel)(signal_variance=1.7, length_scales=[.6, 1.1, .9], ard=True))
    gp.set_training_data(X, y)
    gp.setup()
    nl = d
    base = np.concatenate([[np.log(gp.noise_variance)], np.log(gp.kernel.length_scales), [np.log(gp.kernel.signal_variance)]])
    free = np.ones(base.size, dtype=bool)
    f0, g = gp._nlml_and_grad(base.copy(), free, base)
    assert np.isfinite(f0) and g.shape == (nl + 2,)
    h = 1e-5
    fd = np.empty_like(g)
    for j in range(base.size):
        tp, tm = base.copy(), base.copy()
        tp[j] += h
        tm[j] -= h
        fd[j] = (gp._nlml_and_grad(tp, free, base)[0] - gp._nlml_and_grad(tm, free, base)[0]) / (2. * h)
    np.testing.assert_allclose(g, fd, rtol=2e-6, atol=1e-7)
