"""Kalman-filter family and particle filter with the reference's API, batched on the GPU.

Mirrors ``hilo_mpc/modules/estimator/base.py`` (``_Estimator``: ``set_initial_guess`` :337, ``set_initial_parameter_values``
:357, ``Q``/``R`` setters :99-125, 253-279, input coercion :164-239), ``kf.py`` (``KalmanFilter`` :340, ``ExtendedKalmanFilter`` :383,
``UnscentedKalmanFilter`` :428 with ``alpha/beta/kappa`` :606-640, ``setup`` :207, ``estimate`` :279, ``predict``/``update``
:309-325) and ``pf.py`` (``ParticleFilter`` :51, ``sample_size`` :263, ``estimate`` :340).  A filter holds either one
instance (reference semantics, results in ``solution``) or a batch: ``set_initial_guess`` with ``x0`` of shape ``(B, nx)``
runs ``B`` independent filters on identical models; ``estimate(y=...)`` then takes ``(B, ny)`` measurements.
"""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np

from .estgen import EstEngine
from .series import TimeSeries


def _cov(arg, n: int, what: str) -> np.ndarray:
    """Scalar -> s*I, vector -> diag, matrix as is (estimator/base.py:99-125)."""
    a = np.asarray(arg, dtype=float)
    if a.ndim == 0:
        return float(a) * np.eye(n)
    if a.ndim == 1:
        if a.size != n:
            raise ValueError(f"Dimension mismatch. Supplied dimension is {a.size}x{a.size}, but required dimension is "
                             f"{n}x{n}")
        return np.diag(a)
    if a.shape != (n, n):
        raise ValueError(f"Dimension mismatch. Supplied dimension is {a.shape[0]}x{a.shape[1]}, but required dimension"
                         f" is {n}x{n}")
    return a.copy()


class _Estimator:
    _label = 'Estimator'

    def __init__(self, model, id=None, name=None, plot_backend=None):
        if not model.is_setup():
            raise RuntimeError(f"Model is not set up. Run Model.setup() before passing it to the {self._label}.")
        self._model = model.copy(setup=True)
        self.id, self.name = id, name or self._label
        self._n_x, self._n_u, self._n_p = model.n_x, model.n_u, model.n_p
        self._n_y = model.n_y if model.n_y > 0 else model.n_x
        self._Q = None
        self._R = None
        self._x = None  # device tensor [B, nx]
        self._P = None  # device tensor [B, nx, nx]
        self._p_values = None
        self._single = True
        self._t = 0.
        self._is_setup = False
        self._engine: Optional[EstEngine] = None
        self._solution = TimeSeries()
        self._substeps = int(model._solver_options.get('substeps', 10))

    # ------------------------------------------------------------------ covariances
    @property
    def Q(self):
        return self._Q

    @Q.setter
    def Q(self, arg):
        self._Q = _cov(arg, self._n_x, 'Q')

    process_noise_covariance = Q

    @property
    def R(self):
        return self._R

    @R.setter
    def R(self, arg):
        self._R = _cov(arg, self._n_y, 'R')

    measurement_noise_covariance = R

    @property
    def P(self):
        return None if self._P is None else (self._P[0].cpu().numpy() if self._single else self._P)

    error_covariance = P

    @property
    def solution(self):
        return self._solution

    def is_setup(self) -> bool:
        return self._is_setup

    # ------------------------------------------------------------------ set-up / initial values
    def setup(self, **kwargs) -> None:
        m = self._model
        self._engine = EstEngine(m, build=getattr(self, '_build_module', True))
        self._solution = TimeSeries(t=['t'], x=m.dynamical_state_names, y=m.measurement_names or
                                    m.dynamical_state_names, P=[f'P_{k}' for k in range(self._n_x ** 2)])
        self._is_setup = True

    def set_initial_guess(self, x0, t0: float = 0., z0=None, P0=None) -> None:
        x0 = np.asarray(x0.cpu() if hasattr(x0, 'cpu') else x0, dtype=float)
        if x0.ndim <= 1 or (x0.ndim == 2 and x0.shape[1] == 1 and self._n_x > 1):
            x0 = x0.reshape(1, -1)
            self._single = True
        else:
            self._single = False
        if x0.shape[1] != self._n_x:
            raise ValueError(f"Dimension mismatch. Supplied {x0.shape[1]} initial states, required {self._n_x}")
        B = x0.shape[0]
        P0 = np.eye(self._n_x) if P0 is None else _cov(P0, self._n_x, 'P0')
        self._x_host, self._P_host = x0, np.broadcast_to(P0, (B, self._n_x, self._n_x)).copy()
        self._x = self._P = None  # moved to the device lazily (no GPU needed to configure a filter)
        self._t = float(t0)
        if self._single:
            self._solution.set('t', [t0])
            self._solution.set('x', x0[0])
            self._solution.set('P', self._P_host[0].reshape(-1))

    def set_initial_parameter_values(self, p) -> None:
        p = np.asarray(p, dtype=float).reshape(-1)
        if p.size != self._n_p:
            raise ValueError(f"Dimension mismatch. Supplied {p.size} parameter values, required {self._n_p}")
        self._p_values = p

    def _to_device(self):
        if self._x is None:
            eng = self._engine
            self._x = eng.tensor(self._x_host)
            self._P = eng.tensor(self._P_host)

    def _check(self, y):
        if not self._is_setup:
            raise RuntimeError(f"{self._label} is not set up. Run {type(self).__name__}.setup() before running "
                               f"simulations.")
        if getattr(self, '_x_host', None) is None:
            raise RuntimeError(f"No initial guess for the states found. Please set initial guess before running the "
                               f"{self._label}!")
        if y is None:
            raise RuntimeError("No measurement data supplied.")

    def _inputs(self, y, u, p, B):
        eng = self._engine
        yt = eng.tensor(y).reshape(-1, self._n_y)
        if yt.shape[0] != B:
            raise ValueError(f"Dimension mismatch in y: expected ({B}, {self._n_y}), got {tuple(yt.shape)}")
        if self._n_u and u is None:
            raise RuntimeError(f"The model has {self._n_u} input(s) but no input was supplied")
        if self._n_p and p is None:
            p = self._p_values
            if p is None:
                raise RuntimeError(f"The model has {self._n_p} parameter(s) but no values were supplied")
        return yt, (None if not self._n_u else eng.tensor(u)), (None if not self._n_p else eng.tensor(p))

    def _store(self, y_pred):
        self._t += self._model.dt_value
        if self._single:
            self._solution.update(t=[self._t], x=self._x[0].cpu().numpy(), P=self._P[0].reshape(-1).cpu().numpy(),
                                  y=y_pred[0].cpu().numpy())


class _KalmanFilter(_Estimator):
    _label = 'Kalman filter'
    _kernel = 'ekf'

    def _step(self, y, u, p, do_predict=True, do_update=True, x=None, P=None):
        from . import _lib
        torch = _lib.require_cuda()
        eng = self._engine
        eng.open()
        if self._Q is None or self._R is None:
            raise RuntimeError("The noise covariances Q and R must be set before estimating")
        self._to_device()
        B = self._x.shape[0]
        yt, ut, pt = self._inputs(y, u, p, B)
        ypred = torch.zeros(B, self._n_y, dtype=torch.float64, device=eng.device)
        status = torch.zeros(B, dtype=torch.int32, device=eng.device)
        extra = self._extra(B, torch, eng)
        eng.call(self._kernel, B, self._model.dt_value, self._x, P=self._P, y=yt, u=ut, p=pt, Q=eng.tensor(self._Q),
                 R=eng.tensor(self._R), y_pred=ypred, status=status, substeps=self._substeps, do_predict=do_predict,
                 do_update=do_update, t0=self._t, **extra)
        self._last_status = status
        if self._single:   # one filter: a failed factorisation is an error (the batched path exposes `last_status`)
            if int(status[0].item()) != 0:
                raise RuntimeError(f"{self._label}: numerical failure (covariance not positive definite or singular "
                                   f"innovation covariance); the estimate was left unchanged")
        return ypred

    @property
    def last_status(self):
        """Per-instance status of the last step (device int32 tensor): 0 ok, -1 numerical failure (estimate unchanged)."""
        return getattr(self, '_last_status', None)

    def _extra(self, B, torch, eng):
        return {}

    def estimate(self, y=None, u=None, p=None, steps=1, **kwargs):
        self._check(y)
        ypred = self._step(y, u, p)
        self._store(ypred)
        if self._single:
            return None
        return self._x, self._P

    def predict(self, x0, up=None, Q=None):
        """``predict([x|P], [u;p], Q)`` -> ``[x|P]`` pass-through (kf.py:309-316, tests/test_KFs.py:489-503)."""
        x0 = np.asarray(x0, dtype=float)
        n = self._n_x
        self.set_initial_guess(x0[:, 0], P0=x0[:, 1:1 + n])
        if Q is not None:
            self.Q = Q
        if self._R is None:
            self.R = 1.
        up = np.zeros(0) if up is None else np.asarray(up, dtype=float).reshape(-1)
        u, p = up[:self._n_u], up[self._n_u:]
        self._step(np.zeros(self._n_y), u if self._n_u else None, p if self._n_p else None, do_update=False)
        return self._result_matrix()

    def update(self, pred, y, up=None, R=None):
        pred = np.asarray(pred, dtype=float)
        n = self._n_x
        self.set_initial_guess(pred[:, 0], P0=pred[:, 1:1 + n])
        self._restore_sigma(pred)
        if R is not None:
            self.R = R
        if self._Q is None:
            self.Q = 1.
        up = np.zeros(0) if up is None else np.asarray(up, dtype=float).reshape(-1)
        u, p = up[:self._n_u], up[self._n_u:]
        ypred = self._step(y, u if self._n_u else None, p if self._n_p else None, do_predict=False)
        out = np.column_stack([self._x[0].cpu().numpy(), self._P[0].cpu().numpy()])
        return out, ypred[0].cpu().numpy().reshape(-1, 1)

    def _restore_sigma(self, pred):
        pass

    def _result_matrix(self):
        return np.column_stack([self._x[0].cpu().numpy(), self._P[0].cpu().numpy()])


class KalmanFilter(_KalmanFilter):
    def __init__(self, model, id=None, name=None, plot_backend=None, square_root_form=True):
        if not model.is_linear():
            raise ValueError("The supplied model is nonlinear. Please use an estimator targeted at the estimation of "
                             "nonlinear systems.")  # kf.py:349-351
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)


class ExtendedKalmanFilter(_KalmanFilter):
    _label = 'Extended Kalman filter'

    def __init__(self, model, id=None, name=None, plot_backend=None, square_root_form=True):
        if model.is_linear():
            warnings.warn("The supplied model is linear. For better efficiency use an observer targeted at the "
                          "estimation of linear systems.")  # kf.py:392-394
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)


class UnscentedKalmanFilter(_KalmanFilter):
    _label = 'Unscented Kalman filter'
    _kernel = 'ukf'

    def __init__(self, model, id=None, name=None, plot_backend=None, alpha=.001, beta=2., kappa=0.):
        if model.is_linear():
            warnings.warn("The supplied model is linear. For better efficiency use an observer targeted at the "
                          "estimation of linear systems.")  # kf.py:440-442
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)
        self._alpha = self._beta = self._kappa = None
        self.alpha, self.beta, self.kappa = alpha, beta, kappa
        self._X = None
        self.kernel_variant = 1   # 1: one lane per sigma point (csrc/est_lanes.cuh), 0: one thread per filter

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, arg):
        if not 1e-4 <= arg <= 1.:  # kf.py:606-616
            raise ValueError("The parameter alpha needs to lie in the interval [1e-4, 1]")
        self._alpha = float(arg)

    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, arg):
        self._beta = float(arg)

    @property
    def kappa(self):
        return self._kappa

    @kappa.setter
    def kappa(self, arg):
        if arg < 0.:
            raise ValueError("The parameter kappa needs to be greater or equal to 0")
        self._kappa = float(arg)

    def _extra(self, B, torch, eng):
        ns = 2 * self._n_x + 1
        if self._X is None or self._X.shape[0] != B:
            self._X = torch.zeros(B, ns, self._n_x, dtype=torch.float64, device=eng.device)
        return dict(X=self._X, alpha=self._alpha, beta=self._beta, kappa=self._kappa, variant=self.kernel_variant)

    def _result_matrix(self):  # kf.py:549-553: [x_pred | P_pred | X]
        return np.column_stack([self._x[0].cpu().numpy(), self._P[0].cpu().numpy(), self._X[0].cpu().numpy().T])

    def _restore_sigma(self, pred):
        n = self._n_x
        if pred.shape[1] >= 1 + n + 2 * n + 1:
            self._engine.open()
            self._X = self._engine.tensor(pred[:, 1 + n:].T.copy()).reshape(1, 2 * n + 1, n)
        elif self._X is None:
            raise ValueError("UKF.update needs the propagated sigma points: pass the full matrix returned by predict() "
                             "([x | P | X], kf.py:549-553)")


def lhsnorm(mu, sigma, n):
    """Latin-hypercube sample of a normal distribution, the particle filter's default density (pf.py:425-447): ``n`` draws
    of ``N(mu, sigma)`` fix the rank order in every coordinate, the values are the normal quantiles of one uniformly
    jittered point per stratum ``((rank - U) / n)`` with the marginal standard deviations ``sqrt(sigma_kk)``.  Draws from
    the legacy NumPy global RNG in the reference's order (one ``multivariate_normal``, one ``rand``).  Returns ``(n, n_mu)``."""
    from scipy.stats import norm
    mu = np.asarray(mu, dtype=float).reshape(-1)
    sigma = np.asarray(sigma, dtype=float)
    z = np.random.multivariate_normal(mu, sigma, size=n)
    rank = np.empty_like(z)
    order = np.argsort(z, axis=0)
    for k in range(mu.size):
        rank[order[:, k], k] = np.arange(1, n + 1)
    strata = (rank - np.random.rand(*z.shape)) / n
    return norm.ppf(strata, loc=mu[None, :], scale=np.sqrt(np.diag(sigma))[None, :])


class ParticleFilter(_Estimator):
    _label = 'Particle filter'

    def __init__(self, model, id=None, name=None, plot_backend=None, variant=None, roughening=False,
                 prior_editing=False, K=.2):
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)
        self._prior_editing = bool(prior_editing)
        self.last_prior_editing_rounds = 0
        self._roughening = bool(roughening)
        self._roughening_tuning_param = float(K)   # pf.py:72-76
        self.last_roughening_noise = None
        self._sample_size = 15  # pf.py:79
        self._pdf = lhsnorm      # pf.py:80: initial sample, process noise and roughening all draw from it
        self.kernel_variant = 1  # 1: one warp per filter, particles on the lanes (csrc/est_lanes.cuh), 0: one thread per filter
        self._X = None
        self.last_indices = None
        self.last_weights = None

    @property
    def sample_size(self):
        return self._sample_size

    @sample_size.setter
    def sample_size(self, n):
        if not isinstance(n, (int, np.integer)) or n < 2:
            raise ValueError("The sample size must be an integer greater than 1")
        self._sample_size = int(n)

    def setup(self, n_samples: Optional[int] = None, **kwargs) -> None:
        if n_samples is not None:
            self.sample_size = n_samples
        super().setup(**kwargs)

    def set_particles(self, X) -> None:
        """Particles ``(n_x, n_s)`` (reference layout, pf.py:352) or batched ``(B, n_s, n_x)``."""
        X = np.asarray(X, dtype=float)
        if X.ndim == 2:
            X = X.T[None]
        self._X_host = X.copy()
        self._X = None
        self._sample_size = X.shape[1]

    @property
    def probability_density_function(self):
        return self._pdf

    @probability_density_function.setter
    def probability_density_function(self, pdf):
        """``pdf(mu, sigma, n) -> (n, n_mu)`` or ``(n_mu, n)`` samples (pf.py:193-247)."""
        if not callable(pdf):
            raise TypeError("The probability density function must be callable as pdf(mu, sigma, n)")
        self._pdf = pdf

    def _draw(self, mu, sigma, n):
        X = np.asarray(self._pdf(np.asarray(mu, dtype=float), np.asarray(sigma, dtype=float), n), dtype=float)
        nm = np.asarray(mu).size
        if X.shape == (n, nm):
            return X
        if X.shape == (nm, n):    # the reference transposes user densities that return (n_mu, n) (pf.py:175-185)
            return X.T
        raise ValueError(f"Dimension mismatch. Expected dimension {nm}x{n}, got {X.shape[1]}x{X.shape[0]}.")

    def _initial_sample(self):
        """Default initial sample around the initial guess with the filter's density (pf.py:168-187), legacy global RNG."""
        x, P = self._x_host, self._P_host
        out = np.stack([self._draw(x[b], P[b], self._sample_size) for b in range(x.shape[0])])
        self._X_host = out
        self._X = None

    def _roughen(self, torch, roughening_noise):
        """pf.py:411-421: jitter the resampled particles with N(0, K diag(max - min) n_s^(-1/n_x)) (drawn like the reference
        draws it: legacy global RNG, after the resampling draw) and take the moments of the jittered set."""
        X = self._X                                                    # (B, n_s, n_x), resampled by the kernel
        B, ns, nx = X.shape
        if roughening_noise is None:
            rng_range = (X.amax(dim=1) - X.amin(dim=1)).cpu().numpy()  # (B, n_x)
            scale = self._roughening_tuning_param * ns ** (-1. / nx)
            roughening_noise = np.stack([self._draw(np.zeros(nx), scale * np.diag(rng_range[b]), ns) for b in range(B)])
        dx = self._engine.tensor(roughening_noise).reshape(B, ns, nx)
        self.last_roughening_noise = dx
        X += dx
        mean = X.mean(dim=1)
        Xc = X - mean[:, None, :]
        self._x.copy_(mean)
        self._P.copy_(Xc.transpose(1, 2) @ Xc / (ns - 1))              # np.cov: unbiased

    def _edit_prior(self, torch, B, ns, nx, yt, work, common, max_rounds: int = 1000):
        """Prior editing (pf.py:378-398): particles whose predicted noisy measurement is more than six standard
        deviations away from the measurement are jittered before the propagation - range of the particle set x K x
        n_r^(-1/n_x), drawn from the filter's density in the reference's order - and the prediction pass (kernel call with
        ``do_update=0``: nothing but the predicted measurements is produced) is repeated until none is left.  The check
        runs on the device; only the count per filter and the jittered particles cross the bus."""
        eng = self._engine
        sig6 = 6. * float(np.sqrt(np.asarray(self._R, dtype=float).reshape(-1)[0]))
        unif0 = torch.zeros(B, ns, dtype=torch.float64, device=eng.device)
        Ys = work[B * ns * nx + B * ns:B * ns * nx + 2 * B * ns].view(B, ns)   # csrc/est_module.cuh: Xs | cdf | Ys
        self.last_prior_editing_rounds = 0
        for _ in range(max_rounds):
            eng.call('pf', B, self._model.dt_value, self._x, X=self._X, unif=unif0, do_update=False, **common)
            need = (yt[:, :1] - Ys).abs() > sig6
            n_r = need.sum(dim=1).cpu().numpy()
            if not n_r.any():
                return
            self.last_prior_editing_rounds += 1
            Xh = self._X.cpu().numpy()
            mask = need.cpu().numpy()
            for b in np.nonzero(n_r)[0]:
                spread = Xh[b].max(axis=0) - Xh[b].min(axis=0)
                cov = self._roughening_tuning_param * np.diag(spread) * float(n_r[b]) ** (-1. / nx)
                Xh[b, mask[b]] += self._draw(np.zeros(nx), cov, int(n_r[b]))
            self._X = eng.tensor(Xh)
        raise RuntimeError("Prior editing did not terminate: the measurement is out of reach of the particle set")

    def estimate(self, y=None, u=None, p=None, w=None, v=None, uniforms=None, roughening_noise=None, **kwargs):
        """One filter step. The random draws follow the reference (legacy NumPy global RNG, pf.py:363-365, 406, 413) unless
        ``w`` (process noise ``(B, n_s, n_x)``), ``v`` (measurement noise), ``uniforms`` and (with roughening)
        ``roughening_noise`` ``(B, n_s, n_x)`` are passed explicitly."""
        from . import _lib
        torch = _lib.require_cuda()
        self._check(y)
        if self._Q is None or self._R is None:
            raise RuntimeError("The noise covariances Q and R must be set before estimating")
        eng = self._engine
        eng.open()
        self._to_device()
        if getattr(self, '_X_host', None) is None:
            self._initial_sample()
        if self._X is None:
            self._X = eng.tensor(self._X_host)
        B, ns, nx, ny = self._X.shape[0], self._sample_size, self._n_x, self._n_y
        if w is None:
            w = np.stack([self._draw(np.zeros(nx), self._Q, ns) for _ in range(B)])   # pf.py:363
        if v is None:
            v = np.stack([(np.sqrt(self._R) @ np.random.randn(ny, ns)).T for _ in range(B)])  # pf.py:365
        yt, ut, pt = self._inputs(y, u, p, B)
        q = torch.zeros(B, ns, dtype=torch.float64, device=eng.device)
        idx = torch.zeros(B, ns, dtype=torch.int64, device=eng.device)
        ypred = torch.zeros(B, ny, dtype=torch.float64, device=eng.device)
        status = torch.zeros(B, dtype=torch.int32, device=eng.device)
        work = torch.empty(eng.work_doubles(B, ns), dtype=torch.float64, device=eng.device)
        wt, vt = eng.tensor(w), eng.tensor(v)
        common = dict(P=self._P, y=yt, u=ut, p=pt, Q=eng.tensor(self._Q), R=eng.tensor(self._R), y_pred=ypred, w=wt, v=vt,
                      q=q, idx=idx, work=work, status=status, substeps=self._substeps, n_particles=ns, t0=self._t,
                      variant=self.kernel_variant)
        if self._prior_editing:
            self._edit_prior(torch, B, ns, nx, yt, work, common)
        if uniforms is None:   # drawn after the prior editing, like the reference (pf.py:406)
            uniforms = np.stack([np.random.random_sample(ns) for _ in range(B)])  # what np.random.choice draws
        eng.call('pf', B, self._model.dt_value, self._x, X=self._X, unif=eng.tensor(uniforms), **common)
        self.last_indices, self.last_weights = idx, q
        if self._roughening:
            self._roughen(torch, roughening_noise)
        self._store(ypred)
        if self._single:
            return None
        return self._x, self._P

    @property
    def particles(self):
        return self._X


KF = KalmanFilter
EKF = ExtendedKalmanFilter
UKF = UnscentedKalmanFilter
PF = ParticleFilter
