"""Oracle (test infrastructure): exact Gaussian-process regression with the squared-exponential kernel.

Restates ``hilo_mpc/modules/machine_learning/gp/inference.py:197-217`` (``L = chol(sigma_n^2 I + K)``,
``alpha = L^-T L^-1 (y - m)``, ``LML = -1/2 (y-m)^T alpha - sum log diag L - D/2 log 2 pi``, ``mu* = m(x*) + K*^T alpha``,
``var* = k** - v^T v`` with ``v = L^-1 K*``) and the SE covariance ``hilo_mpc/modules/machine_learning/gp/kernel.py:650-734``
(``k(x, x') = sigma_f^2 exp(-1/2 (x-x')^T M (x-x'))``, ``M = diag(1/l^2)``; hyper-parameters are stored as logs: signal
variance as ``log sigma_f``, noise as ``log sigma_n``, gp.py:302-303).  Data layout as in the reference: ``X`` is
``(n_features, n_observations)`` (gp.py:347-354).  Pinned by tests/test_kernels.py:262-277, 370-388 and
tests/test_GPs.py:354-363 (LML = -9.82229944).
"""
import numpy as np
import scipy.linalg as sla


# Stationary kernels as functions of r^2 = sum_k ((x_k - x'_k) / l_k)^2 (kernel.py:538-555): value g(r) and the factor
# h(r) of the length-scale derivative, dK/dlog l_k = sigma_f^2 h(r) D2_k with D2_k = ((x_k - x'_k) / l_k)^2.
#   'SE'  (kernel.py:650-734)    g = exp(-r^2/2)                                 h = g
#   'E'   (kernel.py:783-826, p=0, nu=1/2)  g = exp(-r)                          h = exp(-r) / r
#   'M32' (p=1, nu=3/2, d = sqrt(3) r)      g = (1 + d) exp(-d)                  h = 3 exp(-d)
#   'M52' (p=2, nu=5/2, d = sqrt(5) r)      g = (1 + d + d^2/3) exp(-d)          h = 5/3 (1 + d) exp(-d)
KINDS = ('SE', 'E', 'M32', 'M52')


def stationary_g_h(r2, kind):
    r = np.sqrt(np.maximum(r2, 0.))
    if kind == 'SE':
        g = np.exp(-.5 * r2)
        return g, g
    if kind == 'E':
        g = np.exp(-r)
        with np.errstate(divide='ignore', invalid='ignore'):
            h = np.where(r > 0., g / np.where(r > 0., r, 1.), 0.)
        return g, h
    if kind == 'M32':
        d = np.sqrt(3.) * r
        return (1. + d) * np.exp(-d), 3. * np.exp(-d)
    if kind == 'M52':
        d = np.sqrt(5.) * r
        return (1. + d + d * d / 3.) * np.exp(-d), 5. / 3. * (1. + d) * np.exp(-d)
    raise ValueError(kind)


def stationary_kernel(X, X_bar=None, signal_variance=1., length_scales=1., kind='SE'):
    X = np.atleast_2d(np.asarray(X, dtype=float))
    X_bar = X if X_bar is None else np.atleast_2d(np.asarray(X_bar, dtype=float))
    ls = np.broadcast_to(np.asarray(length_scales, dtype=float).reshape(-1), (X.shape[0],))
    r2 = (((X[:, :, None] - X_bar[:, None, :]) / ls[:, None, None]) ** 2).sum(0)
    return signal_variance * stationary_g_h(r2, kind)[0]


def se_kernel(X, X_bar=None, signal_variance=1., length_scales=1.):
    X = np.atleast_2d(np.asarray(X, dtype=float))
    X_bar = X if X_bar is None else np.atleast_2d(np.asarray(X_bar, dtype=float))
    ls = np.broadcast_to(np.asarray(length_scales, dtype=float).reshape(-1), (X.shape[0],))
    A = X / ls[:, None]
    Bm = X_bar / ls[:, None]
    d2 = (A * A).sum(0)[:, None] + (Bm * Bm).sum(0)[None, :] - 2. * A.T @ Bm
    return signal_variance * np.exp(-.5 * np.maximum(d2, 0.))


class GpOracle:
    def __init__(self, X, y, noise_variance=1., signal_variance=1., length_scales=1., mean=0., kind='SE'):
        self.X = np.atleast_2d(np.asarray(X, dtype=float))
        self.y = np.asarray(y, dtype=float).reshape(-1)
        self.hyp = dict(noise_variance=noise_variance, signal_variance=signal_variance, length_scales=length_scales)
        self.mean = mean
        self.kind = kind
        n = self.X.shape[1]
        K = stationary_kernel(self.X, None, signal_variance, length_scales, kind)
        self.L = np.linalg.cholesky(noise_variance * np.eye(n) + K)
        self.alpha = sla.cho_solve((self.L, True), self.y - mean)

    def log_marginal_likelihood(self):
        n = self.X.shape[1]
        return float(-.5 * (self.y - self.mean) @ self.alpha - np.log(np.diag(self.L)).sum() - n / 2 * np.log(2 * np.pi))

    def predict(self, Xq, noise_free=False):
        Xq = np.atleast_2d(np.asarray(Xq, dtype=float))
        Ks = stationary_kernel(self.X, Xq, self.hyp['signal_variance'], self.hyp['length_scales'], self.kind)
        mu = self.mean + Ks.T @ self.alpha
        v = sla.solve_triangular(self.L, Ks, lower=True)
        var = self.hyp['signal_variance'] - (v * v).sum(0)
        if not noise_free:
            var = var + self.hyp['noise_variance']  # gp.py:709-712
        return mu, var


# ---------------------------------------------------------------------------------------------------------------------
# Hyper-parameter training (gp.py:660-697: minimise the negative log marginal likelihood over the logs of the
# hyper-parameters; inference.py:211-217).  Pinned by the gpml values of tests/test_GPs.py:846-849 (SE kernel, 20 points
# from the LCG generator of tests/test_GPs.py:1173-1199, restated below as a data generator).
def lcg_randn(seed, shape):
    """Normal variates of the reference's test-data generator (tests/test_GPs.py:1173-1199): Lehmer LCG + Box-Muller."""
    n = int(np.prod(shape))
    N = int(np.ceil(n / 2) * 2)
    a, m = 7 ** 5, 2 ** 31 - 1
    q, r = np.fix(m / a), np.remainder(m, a)
    u = np.zeros(N + 1)
    u[0] = np.fix(seed * 2 ** 31)
    for k in range(1, N + 1):
        u[k] = a * np.remainder(u[k - 1], q) - r * np.fix(u[k - 1] / q)
        if u[k] < 0.:
            u[k] += m
    u = u[1:N + 1] / (2 ** 31)
    N2 = N // 2
    w = np.sqrt(-2 * np.log(u[:N2]))
    x = np.concatenate([w * np.cos(2 * np.pi * u[N2:N]), w * np.sin(2 * np.pi * u[N2:N])])
    return np.reshape(x[:n], shape)


def nlml_and_grad(theta, X, y, mean=0., ard=False, kind='SE', Phi=None):
    """Negative LML and its gradient w.r.t. theta = [log noise_variance, log length_scale(s), log signal_variance, beta]:
    dNLML/dtheta_j = 1/2 tr((K^-1 - alpha alpha^T) dK/dtheta_j) (Rasmussen & Williams eq. 5.9); ``Phi`` (n x m) is the
    basis of a linear-in-parameters mean function m(X) = Phi beta (gp/mean.py: linear + constant terms, coefficients not
    log-transformed, gp.py:197), dNLML/dbeta = -Phi^T alpha."""
    X = np.atleast_2d(np.asarray(X, dtype=float))
    y = np.asarray(y, dtype=float).reshape(-1)
    d, n = X.shape
    nl = d if ard else 1
    sn2, ls, sf2 = np.exp(theta[0]), np.exp(theta[1:1 + nl]), np.exp(theta[1 + nl])
    if Phi is not None:
        mean = mean + np.asarray(Phi, dtype=float) @ theta[2 + nl:]
    lsv = np.broadcast_to(ls, (d,))
    D2 = ((X[:, :, None] - X[:, None, :]) / lsv[:, None, None]) ** 2        # (d, n, n)
    gk, hk = stationary_g_h(D2.sum(0), kind)
    Kse, Kh = sf2 * gk, sf2 * hk
    L = np.linalg.cholesky(Kse + sn2 * np.eye(n))
    alpha = sla.cho_solve((L, True), y - mean)
    nlml = .5 * (y - mean) @ alpha + np.log(np.diag(L)).sum() + n / 2 * np.log(2 * np.pi)
    W = sla.cho_solve((L, True), np.eye(n)) - np.outer(alpha, alpha)
    g = np.empty_like(theta)
    g[0] = .5 * sn2 * np.trace(W)
    if ard:
        for k in range(d):
            g[1 + k] = .5 * (W * Kh * D2[k]).sum()
    else:
        g[1] = .5 * (W * Kh * D2.sum(0)).sum()
    g[1 + nl] = .5 * (W * Kse).sum()
    if Phi is not None:
        g[2 + nl:] = -np.asarray(Phi, dtype=float).T @ alpha
    return float(nlml), g


def fit_hyperparameters(X, y, noise_variance=1., length_scales=1., signal_variance=1., mean=0., ard=False, kind='SE',
                        Phi=None, beta0=None):
    """Returns (noise_variance, length_scales, signal_variance) at the minimiser of the negative LML (L-BFGS on the logs,
    started from the given values)."""
    import scipy.optimize as sopt
    d = np.atleast_2d(X).shape[0]
    ls0 = np.broadcast_to(np.asarray(length_scales, dtype=float).reshape(-1), (d if ard else 1,))
    th0 = np.concatenate([[np.log(noise_variance)], np.log(ls0), [np.log(signal_variance)]])
    nk = th0.size
    if Phi is not None:
        th0 = np.concatenate([th0, np.zeros(np.asarray(Phi).shape[1]) if beta0 is None else np.asarray(beta0, float)])
    res = sopt.minimize(nlml_and_grad, th0, args=(X, y, mean, ard, kind, Phi), jac=True, method='L-BFGS-B',
                        options=dict(maxiter=500, ftol=1e-15, gtol=1e-10))
    th = res.x
    res.beta = th[nk:]   # coefficients of the mean function (empty without one)
    return float(np.exp(th[0])), np.exp(th[1:nk - 1]), float(np.exp(th[nk - 1])), res
