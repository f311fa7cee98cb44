"""Pins the oracle restatements on the golden numbers held by the reference's own tests."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle.gp import GpOracle, se_kernel
from oracle.lqr import lqr_gain


def _kf_model():
    # tests/test_KFs.py:252-262: dx1 = -k1 x1 + u, dx2 = k1 x1 - k2 x2, y = x2, ERK1 with dt = 1
    def f(x, u, p):
        return np.array([x[0] + (-p[0] * x[0] + u[0]), x[1] + (p[0] * x[0] - p[1] * x[1])])

    def h(x, u, p):
        return np.array([x[1]])
    return f, h


def test_kf_predict_golden():  # tests/test_KFs.py:489-503
    f, _ = _kf_model()
    x, P = est.ekf_predict(f, np.array([.8, 0.]), np.eye(2), np.array([.8]), np.array([.5, .4]), .01 * np.eye(2))
    np.testing.assert_allclose(np.column_stack([x, P]), [[1.2, .26, .25], [.4, .25, .62]], rtol=1e-12)


def test_kf_update_golden():  # tests/test_KFs.py:505-522
    _, h = _kf_model()
    pred = np.array([[1.2, .26, .25], [.4, .25, .62]])
    x, P, yp = est.ekf_update(h, pred[:, 0], pred[:, 1:], .322052, np.array([.8]), np.array([.5, .4]), np.array([[.064]]))
    np.testing.assert_allclose(np.column_stack([x, P]), [[1.17151023, .16862573, .023391813],
                                                         [.32934538, .023391813, .0580117]], rtol=2e-7)
    np.testing.assert_allclose(yp, [.4])


def test_kf_one_step_golden():  # tests/test_KFs.py:298-335
    f, h = _kf_model()
    u, p = np.array([.8]), np.array([.5, .4])
    x, P = est.ekf_predict(f, np.array([.8, 0.]), np.eye(2), u, p, .01 * np.eye(2))
    x, P, _ = est.ekf_update(h, x, P, .3894626, u, p, np.array([[.064]]))
    np.testing.assert_allclose(x, [1.19614861, .39044856], rtol=1e-7)


def _scalar_model():
    # tests/test_KFs.py:553-557: x+ = x/2 + 25 dt x/(1+x^2), y = x^2/20, dt = 1
    return (lambda x, u, p: x / 2 + 25 * x / (1 + x ** 2)), (lambda x, u, p: x ** 2 / 20)


def test_ekf_one_step_golden():  # tests/test_KFs.py:548-566
    f, h = _scalar_model()
    x, P = est.ekf_predict(f, np.array([9.]), np.eye(1), None, None, 10. * np.eye(1))
    x, P, _ = est.ekf_update(h, x, P, 2.59109, None, None, np.eye(1))
    np.testing.assert_allclose(x, [7.206059], rtol=1e-7)


def test_ukf_one_step_golden():  # tests/test_KFs.py:606-624 (alpha=1e-3, beta=2, kappa=0)
    f, h = _scalar_model()
    x, P, X = est.ukf_predict(f, np.array([9.]), np.eye(1), None, None, 10. * np.eye(1))
    x, P, _ = est.ukf_update(h, x, P, X, 2.59109, None, None, np.eye(1))
    np.testing.assert_allclose(x, [7.2739647], rtol=1e-7)


def test_lqr_riccati_golden():  # tests/test_LQR.py:314-342: x+ = x + dt(2y + p u), y+ = y - dt x, z+ = z + dt w
    for p, K_ref in ((1., [[1.39207671, 1.35221712, 0.], [0., 0., .61802575]]),
                     (0., [[0., 0., 0.], [0., 0., .61802575]])):
        A = np.array([[1., 2., 0.], [-1., 1., 0.], [0., 0., 1.]])
        B = np.array([[p, 0.], [0., 0.], [0., 1.]])
        K, _ = lqr_gain(A, B, np.eye(3), np.eye(2), horizon=5)
        np.testing.assert_allclose(K, K_ref, rtol=1e-7, atol=1e-12)


def test_se_gram_golden():  # tests/test_kernels.py:262-277, 370-388
    x = np.array([[1., 2., 3., 4., 5.]])
    np.testing.assert_allclose(se_kernel(x)[0], [1., .60653066, .135335283, .0111089965, .000335462628], rtol=1e-7)
    cov = se_kernel(x, x / 2.)
    np.testing.assert_allclose(cov[0], [.8824969, 1., .8824969, .60653066, .32465247], rtol=1e-7)
    np.testing.assert_allclose(cov[4], [.0000400652974, .000335462628, .0021874911, .0111089965, .04393693], rtol=1e-7)


def test_gp_log_marginal_likelihood_golden():  # tests/test_GPs.py:327-363
    X = np.array([[0., .5, 1. / np.sqrt(2.), np.sqrt(3.) / 2., 1., 0.],
                  [1., np.sqrt(3.) / 2., 1. / np.sqrt(2.), .5, 0., -1.]])
    y = np.array([0., np.pi / 6., np.pi / 4., np.pi / 3., np.pi / 2., np.pi])
    gp = GpOracle(X, y, noise_variance=1., signal_variance=1., length_scales=1.)
    np.testing.assert_approx_equal(gp.log_marginal_likelihood(), -9.82229944, significant=9)


def test_pf_resampling_matches_numpy_choice():
    """pf.py:406: indices must be bit-identical to ``np.random.choice(n, n, p=q)`` on the legacy global RNG."""
    rng_q = np.random.default_rng(3)
    for n in (15, 64, 1000):
        q = rng_q.random(n)
        q /= q.sum()
        np.random.seed(1234)
        ref = np.random.choice(n, size=n, replace=True, p=q)
        np.random.seed(1234)
        u = np.random.random_sample(n)
        np.testing.assert_array_equal(est.pf_resample_indices(q, u), ref)


def test_gp_fitted_hyperparameters_golden():  # tests/test_GPs.py:833-849: gpml values for the SE kernel, rtol 1e-5
    from oracle.gp import fit_hyperparameters, lcg_randn
    x = lcg_randn(.8, (20, 1))
    y = np.sin(3 * x) + .1 * lcg_randn(.9, (20, 1))
    sn2, ls, sf2, _ = fit_hyperparameters(x.T, y.T)
    np.testing.assert_allclose([sn2, ls[0], sf2], [.0085251, .5298217, .8114553], rtol=1e-5)


def test_matern_gram_golden():  # tests/test_kernels.py:1232-1247 (M32), :1826-1845 (M52), :856-875 (E)
    from oracle.gp import stationary_kernel
    x = np.array([[1., 2., 3., 4., 5.]])
    np.testing.assert_allclose(stationary_kernel(x, kind='M32')[0], [1., .48335772, .13973135, .03431324, .007767734],
                               rtol=2e-7)
    np.testing.assert_allclose(stationary_kernel(x, x / 2., kind='M52')[0],
                               [.82864914, 1., .82864914, .52399411, .28316327], rtol=2e-7)
    np.testing.assert_allclose(stationary_kernel(x, x / 2., kind='E')[0],
                               [.60653066, 1., .60653066, .36787944, .22313016], rtol=2e-7)


@pytest.mark.parametrize('kind,gold', [('M32', [.0088262, .8329355, .9398366]), ('M52', [.0086694, .7180206, .9571137])])
def test_gp_fitted_hyperparameters_matern_golden(kind, gold):  # tests/test_GPs.py:862-867: gpml values, rtol 1e-5
    from oracle.gp import fit_hyperparameters, lcg_randn
    x = lcg_randn(.8, (20, 1))
    y = np.sin(3 * x) + .1 * lcg_randn(.9, (20, 1))
    sn2, ls, sf2, _ = fit_hyperparameters(x.T, y.T, kind=kind)
    np.testing.assert_allclose([sn2, ls[0], sf2], gold, rtol=1e-5)


def test_gp_fitted_hyperparameters_exponential_golden():  # tests/test_GPs.py:854-860: noise -> 0, [l, sf2] pinned
    from oracle.gp import fit_hyperparameters, lcg_randn
    x = lcg_randn(.8, (20, 1))
    y = np.sin(3 * x) + .1 * lcg_randn(.9, (20, 1))
    sn2, ls, sf2, _ = fit_hyperparameters(x.T, y.T, kind='E')
    assert sn2 < 1e-5
    np.testing.assert_allclose([ls[0], sf2], [.9769135, .5935212], rtol=1e-5)


# tests/test_KFs.py:690-756: UKF sigma-point golden on the 3-state bioreactor (continuous model, dt = 0.1, alpha = 1);
# the reference integrates with CVODES and asserts atol = 1e-3
UKF_BIO_P = [.15, 303.15, .13, .00025, 15., .14]     # alpha, T_amb, mu_0, mu_1, K, Y (order of appearance)
UKF_BIO_PRED = np.array([
    [299.925, 0.970453, 1.08254e-06, -2.11206e-05, 299.925, 301.631, 299.925, 299.925, 298.218, 299.925, 299.925],
    [0.219433, 1.08254e-06, 1.02165, -0.0848792, 0.219446, 0.219451, 1.97011, 0.219537, 0.21944, -1.53128, 0.219345],
    [19.9619, -2.11206e-05, -0.0848792, 1.00427, 19.9618, 19.9617, 19.8164, 21.6914, 19.9618, 20.1075, 18.2322]])
UKF_BIO_UPD = np.array([[300.733, 0.198539, -2.03814e-06, 1.56415e-06],
                        [0.245549, -2.03814e-06, 0.00990874, -0.000819447],
                        [19.9597, 1.56415e-06, -0.000819447, 0.997286]])


def bioreactor_rhs(x, u, p):
    T, cB, cS = x
    D = np.atleast_1d(u)[0]
    alpha, T_amb, mu0, mu1, K, Y = p
    r = (mu0 + mu1 * T) * cS * cB / (K + cS)
    return np.array([alpha * (T_amb - T), r - D * cB, -r / Y - D * cS])


def test_ukf_bioreactor_sigma_point_golden():
    h = lambda x, u, p: np.array([x[0], x[1]])
    xp, Pp, X = est.ukf_predict(bioreactor_rhs, np.array([299.876, .217108, 20.]), np.eye(3), [.01], UKF_BIO_P,
                                np.zeros((3, 3)), dt=.1, discrete=False, substeps=10, alpha=1.)
    np.testing.assert_allclose(np.concatenate([xp[:, None], Pp, X], axis=1), UKF_BIO_PRED, atol=1e-3)
    xu, Pu, yp = est.ukf_update(h, UKF_BIO_PRED[:, 0], UKF_BIO_PRED[:, 1:4], UKF_BIO_PRED[:, 4:],
                                np.array([300.941, .245805]), [.01], UKF_BIO_P, np.diag([.25, .01]), alpha=1.)
    np.testing.assert_allclose(np.concatenate([xu[:, None], Pu], axis=1), UKF_BIO_UPD, atol=1e-3)
    np.testing.assert_allclose(yp, [299.925, .219434], atol=1e-3)


def rasmussen_data():
    """tests/test_GPs.py:1102-1113: 20 points drawn from a Matern-3/2 GP (l = 0.25) around the mean 0.5 x + 1."""
    from oracle.gp import lcg_randn, stationary_kernel
    x = lcg_randn(.3, (20, 1))
    K = stationary_kernel(x.T, kind='M32', length_scales=.25)
    mu = .5 * x + 1.
    y = np.linalg.cholesky(K) @ lcg_randn(.15, (20, 1)) + mu + .1 * lcg_randn(.2, (20, 1))
    return x, y, mu


def test_gp_rasmussen_regression_goldens():  # tests/test_GPs.py:1120-1124
    from oracle.gp import fit_hyperparameters
    x, y, mu = rasmussen_data()
    o = GpOracle(x.T, (y - mu).T, noise_variance=.01, signal_variance=1., length_scales=.25, kind='M32')
    np.testing.assert_approx_equal(o.log_marginal_likelihood(), -11.9706317, significant=9)
    sn2, ls, sf2, _ = fit_hyperparameters(x.T, y.T, noise_variance=.01)
    np.testing.assert_allclose([sn2, ls[0], sf2], [.0222571, .3703166, 3.9427837], rtol=1e-6)


def test_gp_trainable_mean_golden():  # tests/test_GPs.py:1125-1126, 1150-1155: linear + constant mean, all five fitted
    from oracle.gp import fit_hyperparameters
    x, y, _ = rasmussen_data()
    Phi = np.column_stack([x[:, 0], np.ones(20)])
    sn2, ls, sf2, res = fit_hyperparameters(x.T, y.T, noise_variance=.01, Phi=Phi)
    np.testing.assert_allclose([sn2, res.beta[0], res.beta[1], ls[0], sf2],
                               [.02183, 1.1918832, 1.4624534, .2201449, .4017997], rtol=1e-6)
