"""Oracle (test infrastructure): Kalman-filter family and particle filter, restated in NumPy.

Follows ``hilo_mpc/modules/estimator/kf.py`` — predict :71-133 (``P+ = F P F^T + Q`` discrete, ``Pdot = F P + P F^T + Q``
continuous, integrated together with the state), update :135-186 (``K = P_xy P_yy^-1`` through ``ca.solve``,
``P+ = P - K P_yy K^T``), UKF parameters :486-503 (``lambda = alpha^2 (n+kappa) - n``, ``gamma = sqrt(n+lambda)``, weights,
``ca.chol`` = *upper* factor), UKF predict :505-554 (sigma points use the COLUMNS of the upper factor), UKF update
:556-604 (re-uses the propagated sigma points) — and ``hilo_mpc/modules/estimator/pf.py`` — propagate :103-148, likelihood
:150-166 (univariate normal pdf with ``sqrt(R)``, normalised), resampling :406 (``np.random.choice(n, n, p=q)`` =
``cdf = q.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, side='right')`` on legacy-RNG uniforms), estimate
:419-421 (``mean``, ``np.cov``).

Pinned by the reference's own golden numbers (tests/test_KFs.py:298-335, 489-522, 548-566, 606-624) in
``tests/test_oracle_golden.py``.  Model functions are plain callables ``f(x, u, p) -> x_next`` (discrete) or
``-> dx/dt`` (continuous, integrated here with classical RK4 sub-steps), ``h(x, u, p) -> y``.
"""
import numpy as np


def rk4(f, x, u, p, dt, substeps=1):
    h = dt / substeps
    for _ in range(substeps):
        k1 = f(x, u, p)
        k2 = f(x + .5 * h * k1, u, p)
        k3 = f(x + .5 * h * k2, u, p)
        k4 = f(x + h * k3, u, p)
        x = x + h / 6. * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def jac_cs(fun, x, *rest):
    """Complex-step Jacobian of ``fun(x, *rest)`` with respect to ``x``."""
    x = np.asarray(x, dtype=float)
    cols = []
    for j in range(x.size):
        xp = x.astype(complex)
        xp[j] += 1e-30j
        cols.append(np.atleast_1d(fun(xp, *rest)).imag / 1e-30)
    return np.stack(cols, axis=1)


def ekf_predict(f, x, P, u, p, Q, dt=1., discrete=True, substeps=10):
    """kf.py:87-133. Continuous: joint RK4 integration of ``[x; vec P]`` with ``Pdot = F(x) P + P F(x)^T + Q``."""
    n = x.size
    if discrete:
        F = jac_cs(f, x, u, p)
        return f(x, u, p), F @ P @ F.T + Q

    def aug(z, u_, p_):
        xx, PP = z[:n], z[n:].reshape(n, n)
        F = jac_cs(lambda a, b, c: f(a, b, c), np.real(xx), u_, p_)
        return np.concatenate([f(xx, u_, p_), (F @ PP + PP @ F.T + Q).reshape(-1)])
    z = rk4(aug, np.concatenate([x, P.reshape(-1)]), u, p, dt, substeps)
    return z[:n], z[n:].reshape(n, n)


def ekf_update(h, x, P, y, u, p, R):
    """kf.py:169-180."""
    H = jac_cs(h, x, u, p)
    y_pred = np.atleast_1d(h(x, u, p))
    P_xy = P @ H.T
    P_yy = H @ P @ H.T + R
    K = np.linalg.solve(P_yy.T, P_xy.T).T
    return x + K @ (np.atleast_1d(y) - y_pred), P - K @ P_yy @ K.T, y_pred


def ukf_weights(n, alpha=1e-3, beta=2., kappa=0.):
    lam = alpha ** 2 * (n + kappa) - n
    gamma = np.sqrt(n + lam)
    w = np.zeros((2, 2 * n + 1))
    w[0, 0] = lam / (n + lam)
    w[1, 0] = lam / (n + lam) + 1 - alpha ** 2 + beta
    w[:, 1:] = 1 / (2 * (n + lam))
    return lam, gamma, w


def ukf_predict(f, x, P, u, p, Q, dt=1., discrete=True, substeps=10, alpha=1e-3, beta=2., kappa=0.):
    """kf.py:505-554. Returns ``x_pred, P_pred, X`` (propagated sigma points, n x (2n+1))."""
    n = x.size
    _, gamma, w = ukf_weights(n, alpha, beta, kappa)
    Rup = np.linalg.cholesky(P).T  # ca.chol: upper factor, R^T R = P
    cols = [x] + [x + gamma * Rup[:, k] for k in range(n)] + [x - gamma * Rup[:, k] for k in range(n)]
    X = np.stack([f(c, u, p) if discrete else rk4(f, c, u, p, dt, substeps) for c in cols], axis=1)
    x_pred = X @ w[0]
    dX = X - x_pred[:, None]
    P_pred = Q + (w[1] * dX) @ dX.T
    return x_pred, P_pred, X


def ukf_update(h, x, P, X, y, u, p, R, alpha=1e-3, beta=2., kappa=0.):
    """kf.py:556-604."""
    n = x.size
    _, _, w = ukf_weights(n, alpha, beta, kappa)
    Y = np.stack([np.atleast_1d(h(X[:, k], u, p)) for k in range(2 * n + 1)], axis=1)
    y_pred = Y @ w[0]
    dY = Y - y_pred[:, None]
    dX = X - x[:, None]
    P_xy = (w[1] * dX) @ dY.T
    P_yy = R + (w[1] * dY) @ dY.T
    K = np.linalg.solve(P_yy.T, P_xy.T).T
    return x + K @ (np.atleast_1d(y) - y_pred), P - K @ P_yy @ K.T, y_pred


def pf_resample_indices(q, uniforms):
    """``np.random.choice(n, n, p=q)`` given the uniforms the legacy RNG would draw (numpy/random/mtrand.pyx
    ``choice``): sequential cumulative sum, normalisation by the last entry, right-sided search."""
    cdf = np.cumsum(np.asarray(q, dtype=float))
    cdf /= cdf[-1]
    return cdf.searchsorted(np.asarray(uniforms, dtype=float), side='right')


def pf_roughening_covariance(Xr, K):
    """pf.py:411-414: the resampled particles are jittered with N(0, K diag(max - min) n_s^(-1/n_x)) (the range enters
    as a VARIANCE, as in the reference)."""
    n, ns = Xr.shape
    return K * np.diag(Xr.max(axis=1) - Xr.min(axis=1)) * ns ** (-1. / n)


def pf_step(f, h, X, y, u, p, w, v, R, uniforms, dt=1., discrete=True, substeps=10, roughening_noise=None):
    """pf.py:363-421 without prior editing. ``X``: n x ns particles, ``w``: n x ns process noise, ``v``: ny x ns
    measurement noise, ``roughening_noise``: n x ns jitter added after the resampling (pf.py:411-416; drawn by the caller
    from ``pf_roughening_covariance``) or None. Returns ``x_mean, P, X_resampled, q, idx``."""
    ns = X.shape[1]
    Xp = np.stack([f(X[:, k], u, p) if discrete else rk4(f, X[:, k], u, p, dt, substeps) for k in range(ns)], axis=1)
    Y = np.stack([np.atleast_1d(h(Xp[:, k], u, p)) for k in range(ns)], axis=1)
    Xp = Xp + w
    Y = Y + v
    sigma = np.sqrt(R)
    q = np.exp(-.5 * ((Y - np.atleast_1d(y)[:, None]) / sigma) ** 2) / (np.sqrt(2 * np.pi) * sigma)
    q = (q / q.sum(axis=1, keepdims=True))[0]
    idx = pf_resample_indices(q, uniforms)
    Xr = Xp[:, idx]
    if roughening_noise is not None:
        Xr = Xr + roughening_noise
    return Xr.sum(axis=1) / ns, np.cov(Xr), Xr, q, idx


def pf_prior_editing(f, h, X, y, u, p, w, v, R, K, pdf, dt=1., discrete=True, substeps=10, max_rounds=1000):
    """pf.py:378-398, restated: particles whose predicted noisy measurement lies more than six standard deviations from
    ``y`` are jittered BEFORE the propagation (range of all particles x K x n_r^(-1/n), density ``pdf(mu, cov, n) ->
    (n, n_x)``) and the prediction is repeated until none is left.  ``X``: n x ns.  Returns the edited particle set."""
    X = np.array(X, dtype=float)
    n = X.shape[0]
    sigma = np.sqrt(R)

    def predicted(Xc):
        Xp = np.stack([f(Xc[:, k], u, p) if discrete else rk4(f, Xc[:, k], u, p, dt, substeps) for k in range(Xc.shape[1])],
                      axis=1)
        return np.stack([np.atleast_1d(h(Xp[:, k], u, p)) for k in range(Xc.shape[1])], axis=1) + v

    for _ in range(max_rounds):
        need = (np.abs(np.atleast_1d(y)[:, None] - predicted(X)) > 6 * sigma)[0]
        n_r = int(need.sum())
        if n_r == 0:
            return X
        rng_x = X.max(axis=1) - X.min(axis=1)
        dx = pdf(np.zeros(n), K * np.diag(rng_x) * n_r ** (-1. / n), n_r)
        X[:, np.where(need)[0]] += np.asarray(dx).T
    raise RuntimeError("prior editing did not terminate")


def lhsnorm(mu, sigma, n, rng=np.random):
    """pf.py:425-447, restated: Latin-hypercube normal sample, ``(n, n_mu)``; ``rng`` = the legacy global RNG module."""
    from scipy.stats import norm
    mu = np.asarray(mu, dtype=float)
    z = rng.multivariate_normal(mu, sigma, size=n)
    x = np.zeros_like(z)
    for k in range(mu.size):
        x[np.argsort(z[:, k]), k] = np.arange(1, n + 1)
    x = (x - rng.rand(*x.shape)) / n
    for k in range(mu.size):
        x[:, k] = norm.ppf(x[:, k], loc=mu[k], scale=np.sqrt(sigma[k][k]))
    return x
