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


def se_kernel(X, X_bar=None, signal_variance=1., length_scales=1.):
    X = np.atleast_2d(np.asarray(X, dtype=float))
    X_bar = X if X_bar is None else np.atleast_2d(np.asarray(X_bar, dtype=float))
    ls = np.broadcast_to(np.asarray(length_scales, dtype=float).reshape(-1), (X.shape[0],))
    A = X / ls[:, None]
    Bm = X_bar / ls[:, None]
    d2 = (A * A).sum(0)[:, None] + (Bm * Bm).sum(0)[None, :] - 2. * A.T @ Bm
    return signal_variance * np.exp(-.5 * np.maximum(d2, 0.))


class GpOracle:
    def __init__(self, X, y, noise_variance=1., signal_variance=1., length_scales=1., mean=0.):
        self.X = np.atleast_2d(np.asarray(X, dtype=float))
        self.y = np.asarray(y, dtype=float).reshape(-1)
        self.hyp = dict(noise_variance=noise_variance, signal_variance=signal_variance, length_scales=length_scales)
        self.mean = mean
        n = self.X.shape[1]
        K = se_kernel(self.X, None, signal_variance, length_scales)
        self.L = np.linalg.cholesky(noise_variance * np.eye(n) + K)
        self.alpha = sla.cho_solve((self.L, True), self.y - mean)

    def log_marginal_likelihood(self):
        n = self.X.shape[1]
        return float(-.5 * (self.y - self.mean) @ self.alpha - np.log(np.diag(self.L)).sum() - n / 2 * np.log(2 * np.pi))

    def predict(self, Xq, noise_free=False):
        Xq = np.atleast_2d(np.asarray(Xq, dtype=float))
        Ks = se_kernel(self.X, Xq, self.hyp['signal_variance'], self.hyp['length_scales'])
        mu = self.mean + Ks.T @ self.alpha
        v = sla.solve_triangular(self.L, Ks, lower=True)
        var = self.hyp['signal_variance'] - (v * v).sum(0)
        if not noise_free:
            var = var + self.hyp['noise_variance']  # gp.py:709-712
        return mu, var
