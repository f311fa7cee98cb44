"""Gaussian-process regression with the reference's API (squared-exponential and Matern kernels, exact inference) on the
GPU.

Mirrors ``hilo_mpc/modules/machine_learning/gp/gp.py`` (``GaussianProcess`` :112, ``set_training_data`` :378, ``setup``
:522, ``predict`` :699, ``log_marginal_likelihood`` :653, ``noise_variance`` :302) and ``kernel.py``
(``Kernel.squared_exponential`` :229-260, ``SquaredExponentialKernel`` :704-734).  Numeric queries run the
``hb_gp_fit`` / ``hb_gp_predict`` kernels (Gram, blocked Cholesky with fp64 tensor-core trailing updates, ``L^-1 K*`` on
the tensor pipe); symbolic queries (``predict(x_sym)``, which is how the reference embeds a GP into a ``Model``,
``dynamic_model.py:3040-3080``) return the posterior mean/variance as expressions of the fitted numbers.
Hyper-parameter training (``fit_model``, gp.py:660-697): the negative log marginal likelihood and its
analytic gradient are evaluated on the device (``hb_gp_fit`` for the factorisation, the trace terms as device matrix
products), the quasi-Newton iteration over the handful of log-hyper-parameters runs on the host.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Optional, Sequence

import numpy as np

from . import _lib
from . import symbolic as sx


class StationaryKernel:
    """Covariance functions of ``r^2 = sum_k ((x_k - x'_k) / l_k)^2`` (kernel.py:488-563): squared exponential
    (:704-734) and the Matern family with half-integer ``nu`` (:737-917).  ``kind`` is the index the device kernels use
    (``hb_gp_fit_kernel``); ``g_h`` returns the value ``g(r)`` and the factor ``h(r)`` of the length-scale derivative
    ``dK/dlog l_k = sigma_f^2 h(r) D2_k`` (both for NumPy arrays and device tensors)."""
    acronym = 'SE'
    kind = 0

    def __init__(self, active_dims=None, signal_variance=1., length_scales=1., ard=False, bounds=None):
        self.active_dims = None if active_dims is None else list(active_dims)
        self.signal_variance = float(signal_variance)
        self.length_scales = np.atleast_1d(np.asarray(length_scales, dtype=float))
        self.ard = bool(ard)

    def length_scale_vector(self, d: int) -> np.ndarray:
        ls = self.length_scales
        if ls.size == 1:
            return np.full(d, float(ls[0]))
        if ls.size != d:
            raise ValueError("Length scales vector dimension does not equal input space dimension.")
        return ls.copy()

    def g_h(self, r2, xp=np):
        sq = xp.sqrt
        if self.kind == 0:
            g = xp.exp(-.5 * r2)
            return g, g
        r = sq(r2)
        if self.kind == 1:
            g = xp.exp(-r)
            return g, xp.where(r > 0., g / xp.where(r > 0., r, xp.ones_like(r)), xp.zeros_like(r))
        if self.kind == 2:
            d = 3. ** .5 * r
            e = xp.exp(-d)
            return (1. + d) * e, 3. * e
        d = 5. ** .5 * r
        e = xp.exp(-d)
        return (1. + d + d * d / 3.) * e, 5. / 3. * (1. + d) * e

    def __call__(self, X, X_bar=None):
        """Numeric covariance matrix (kernel.py:97-140) - small host-side helper used by tests / set-up code."""
        X = np.atleast_2d(np.asarray(X, dtype=float))
        Xb = X if X_bar is None else np.atleast_2d(np.asarray(X_bar, dtype=float))
        if X.shape[0] != Xb.shape[0]:
            raise AssertionError("X and X_bar do not have the same input space dimensions")
        ls = self.length_scale_vector(X.shape[0])
        d2 = (((X[:, :, None] - Xb[:, None, :]) / ls[:, None, None]) ** 2).sum(0)
        return self.signal_variance * self.g_h(d2)[0]


class SquaredExponentialKernel(StationaryKernel):
    acronym, kind = 'SE', 0


class ExponentialKernel(StationaryKernel):      # kernel.py:829-856 (Matern with p = 0)
    acronym, kind = 'E', 1


class Matern32Kernel(StationaryKernel):         # kernel.py:859-886
    acronym, kind = 'M32', 2


class Matern52Kernel(StationaryKernel):         # kernel.py:889-916
    acronym, kind = 'M52', 3


class Mean:
    """Prior mean functions (gp/mean.py:45-499): sums of a linear term ``c^T x`` and constants.  ``m(X) = Phi(X) beta``
    with the basis ``Phi`` (rows of X for the linear coefficients, ones for a bias); ``free`` marks the entries of ``beta``
    that hyper-parameter training may change (mean hyper-parameters are not log-transformed, gp.py:197)."""

    def __init__(self, terms):
        self.terms = list(terms)   # (kind, value(s), fixed) with kind in 'linear', 'constant'

    # factories of the reference (mean.py:173-241)
    @staticmethod
    def constant(bias=1., hyperprior=None, bounds=None):
        if hyperprior is not None:
            raise NotImplementedError("Hyper-priors are not supported by the B200 engine yet")
        return Mean([('constant', np.atleast_1d(float(bias)), _is_fixed(bounds, 'bias'))])

    @staticmethod
    def zero():
        return Mean([('constant', np.zeros(1), True)])

    @staticmethod
    def one():
        return Mean([('constant', np.ones(1), True)])

    @staticmethod
    def linear(coefficient=1., hyperprior=None, bounds=None):
        if hyperprior is not None:
            raise NotImplementedError("Hyper-priors are not supported by the B200 engine yet")
        return Mean([('linear', np.atleast_1d(np.asarray(coefficient, dtype=float)), _is_fixed(bounds, 'coefficient'))])

    def __add__(self, other):
        if not isinstance(other, Mean):
            return NotImplemented
        return Mean(self.terms + other.terms)

    def _expanded(self, d):
        out = []
        for kind, val, fixed in self.terms:
            if kind == 'linear':
                val = np.full(d, float(val[0])) if val.size == 1 else val
                if val.size != d:
                    raise ValueError("Dimension of the linear mean coefficient does not match the number of features")
            out.append((kind, np.array(val, dtype=float), fixed))
        return out

    def basis(self, X, xp=np):
        """``Phi`` as a list of rows (one per entry of beta) for X of shape (d, m)."""
        rows = []
        for kind, val, _ in self._expanded(X.shape[0]):
            rows += [X[k] for k in range(val.size)] if kind == 'linear' else [xp.ones_like(X[0])]
        return rows

    def beta(self, d):
        return np.concatenate([val for _, val, _ in self._expanded(d)])

    def free(self, d):
        return np.concatenate([np.full(val.size, not fixed) for _, val, fixed in self._expanded(d)])

    def set_beta(self, beta, d):
        terms, o = [], 0
        for kind, val, fixed in self._expanded(d):
            terms.append((kind, np.array(beta[o:o + val.size], dtype=float), fixed))
            o += val.size
        self.terms = terms

    def names(self, d):
        out = []
        for kind, val, _ in self._expanded(d):
            out += [f'Lin.coefficient'] * val.size if kind == 'linear' else ['Const.bias']
        return out

    def __call__(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        rows = self.basis(X)
        return (np.stack(rows).T @ self.beta(X.shape[0])).reshape(1, -1) if rows else np.zeros((1, X.shape[1]))


def _is_fixed(bounds, name):
    return isinstance(bounds, dict) and bounds.get(name) == 'fixed'


class Kernel:
    @staticmethod
    def squared_exponential(active_dims=None, signal_variance=1., length_scales=1., ard=False, bounds=None):
        return SquaredExponentialKernel(active_dims, signal_variance, length_scales, ard, bounds)

    @staticmethod
    def exponential(active_dims=None, signal_variance=1., length_scales=1., ard=False, bounds=None):
        return ExponentialKernel(active_dims, signal_variance, length_scales, ard, bounds)

    @staticmethod
    def matern_32(active_dims=None, signal_variance=1., length_scales=1., ard=False, bounds=None):
        return Matern32Kernel(active_dims, signal_variance, length_scales, ard, bounds)

    @staticmethod
    def matern_52(active_dims=None, signal_variance=1., length_scales=1., ard=False, bounds=None):
        return Matern52Kernel(active_dims, signal_variance, length_scales, ard, bounds)

    @staticmethod
    def _unsupported(name):
        def factory(*args, **kwargs):
            raise NotImplementedError(f"Kernel '{name}' is not supported by the B200 engine yet (SURVEY.md §2 row 7: "
                                      f"only the squared-exponential and Matern kernels are on the hot path)")
        return factory


for _name in ('rational_quadratic', 'periodic', 'linear', 'polynomial',
              'constant', 'neural_network', 'gamma_exponential', 'piecewise_polynomial', 'dot_product'):
    setattr(Kernel, _name, staticmethod(Kernel._unsupported(_name)))


class GaussianProcess:
    def __init__(self, features, labels, inference=None, likelihood=None, mean=None, kernel=None, noise_variance=1.,
                 hyperprior=None, id=None, name=None, solver=None, solver_options=None, **kwargs):
        self.features = [features] if isinstance(features, str) else list(features)
        labels = [labels] if isinstance(labels, str) else list(labels)
        if len(labels) > 1:
            raise ValueError("Training a GP on multiple labels is not supported. Please use 'GPArray'.")  # gp.py:133-135
        self.labels = labels
        self.kernel = kernel if kernel is not None else Kernel.squared_exponential()
        if not isinstance(self.kernel, StationaryKernel):
            raise NotImplementedError("Only the squared-exponential and Matern kernels are supported by the B200 engine")
        # a number = constant prior mean handled by the device kernels; a Mean object = m(X) subtracted from the targets
        self.mean = mean if isinstance(mean, Mean) else None
        self.prior_mean = 0. if (mean is None or self.mean is not None) else float(mean)
        self.noise_variance = float(noise_variance)
        self.id, self.name = id, name or 'GP'
        self._X = self._y = None
        self._is_setup = False
        self._fit = None  # device tensors: X, L, Linv, alpha, ls ; host copies: alpha, lml

    @property
    def n_features(self):
        return len(self.features)

    @property
    def X_train(self):
        return self._X

    @property
    def y_train(self):
        return self._y

    def set_training_data(self, X, y) -> None:
        X = np.atleast_2d(np.asarray(X, dtype=float))
        y = np.atleast_2d(np.asarray(y, dtype=float))
        if X.shape[0] != self.n_features:
            raise ValueError(f"Dimension of training inputs does not match the number of features: {X.shape[0]} vs "
                             f"{self.n_features}")  # gp.py:347-354
        if y.shape[0] != 1 or y.shape[1] != X.shape[1]:
            raise ValueError("Number of observations in training matrix and target vector do not match!")
        if self._is_setup:
            warnings.warn("Dimensions of training data set changed. Please run setup() method again.")
        self._X, self._y = X, y
        self._is_setup = False

    def set_training_shards(self, X_shard, y_shard) -> None:
        """Training data distributed over the ranks of the process group by observations: all-gather the shards (NCCL
        over NVLink on GPUs; the one collective of the whole path) and keep the full set on every rank,
        so that the fit and the predictions stay replicated and local."""
        from .distributed import all_gather_training_shards
        X, y = all_gather_training_shards(X_shard, y_shard)
        self.set_training_data(X.cpu().numpy(), y.cpu().numpy())

    def is_setup(self) -> bool:
        return self._is_setup

    def setup(self, **kwargs) -> None:
        """Fit on the device: Gram + Cholesky + alpha + log marginal likelihood (gp.py:522-641, inference.py:197-210)."""
        if self._X is None or self._y is None:
            raise RuntimeError("The training data has not been set. Please run the method set_training_data() to "
                               "proceed.")
        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = torch.device('cuda', torch.cuda.current_device())
        f8 = dict(dtype=torch.float64, device=dev)
        d, n = self._X.shape
        ls = self.kernel.length_scale_vector(d)
        Xt = torch.tensor(self._X, **f8).contiguous()
        y_eff = self._y.reshape(-1) - (self.mean(self._X).reshape(-1) if self.mean is not None else 0.)
        yt = torch.tensor(y_eff, **f8).contiguous()
        lst = torch.tensor(ls, **f8)
        L = torch.empty(n, n, **f8)
        Linv = torch.empty(n, n, **f8)
        alpha = torch.empty(n, **f8)
        lml = torch.zeros(1, **f8)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        lib.hb_gp_fit_work_doubles.restype = C.c_size_t
        work = torch.empty(lib.hb_gp_fit_work_doubles(n), **f8)
        vp, i, dd = C.c_void_p, C.c_int, C.c_double
        lib.hb_gp_fit_kernel.argtypes = [i, i, i, vp, vp, vp, dd, dd, dd, vp, vp, vp, vp, vp, vp, vp]
        rc = lib.hb_gp_fit_kernel(self.kernel.kind, n, d, Xt.data_ptr(), yt.data_ptr(), lst.data_ptr(), self.kernel.signal_variance,
                           self.noise_variance, self.prior_mean, L.data_ptr(), Linv.data_ptr(), alpha.data_ptr(),
                           lml.data_ptr(), work.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'hb_gp_fit_kernel')
        if int(info.item()) != 0:
            raise RuntimeError(f"Covariance matrix is not positive definite (pivot {int(info.item()) - 1}); increase "
                               f"the noise variance")
        self._fit = dict(X=Xt, ls=lst, L=L, Linv=Linv, alpha=alpha, lml=float(lml.item()),
                         alpha_host=alpha.cpu().numpy(), dev=dev)
        self._is_setup = True

    @property
    def hyperparameters(self):
        """``[noise_variance, length_scales, signal_variance]`` as objects with ``name`` / ``value`` (the order of the
        reference's list, gp.py:215-240; tests/test_GPs.py:846-849 reads ``[p.value for p in gp.hyperparameters]``)."""
        from types import SimpleNamespace as NS
        ls = self.kernel.length_scales
        mean_par = []
        if self.mean is not None:
            d = len(self.features)
            mean_par = [NS(name=n, value=float(v)) for n, v in zip(self.mean.names(d), self.mean.beta(d))]
        return [NS(name='GP.noise_variance', value=self.noise_variance)] + mean_par + [
                NS(name=f'{self.kernel.acronym}.length_scales', value=float(ls[0]) if ls.size == 1 else ls.copy()),
                NS(name=f'{self.kernel.acronym}.signal_variance', value=self.kernel.signal_variance)]

    def _nlml_and_grad(self, theta, free, base):
        """Negative LML and its gradient w.r.t. the free log-hyper-parameters: factorisation by ``hb_gp_fit`` (setup),
        ``dNLML/dtheta_j = 1/2 tr((K^-1 - alpha alpha^T) dK/dtheta_j)`` by ``hb_gp_lml_grad`` (csrc/gp_kernels.cu)."""
        torch = _lib.require_cuda()
        full = base.copy()
        full[free] = theta
        d = self._X.shape[0]
        nl = self.kernel.length_scales.size
        if free[0]:                                      # fixed values are not passed through log/exp
            self.noise_variance = float(np.exp(full[0]))
        if free[1:1 + nl].any():
            self.kernel.length_scales = np.where(free[1:1 + nl], np.exp(full[1:1 + nl]), self.kernel.length_scales)
        if free[1 + nl]:
            self.kernel.signal_variance = float(np.exp(full[1 + nl]))
        nm = full.size - (2 + nl)                        # coefficients of the mean function (not log-transformed)
        if nm and free[2 + nl:].any():
            self.mean.set_beta(full[2 + nl:], d)
        try:
            self.setup()
        except RuntimeError:   # not positive definite at this trial point: reject the step
            return 1e300, np.zeros_like(theta)
        fit = self._fit
        X, Linv, alpha = fit['X'], fit['Linv'], fit['alpha']
        n = X.shape[1]
        lib = _lib.load()
        work = torch.empty(lib.hb_gp_lml_grad_work_doubles(n), dtype=torch.float64, device=X.device)
        gd = torch.empty(d + 2, dtype=torch.float64, device=X.device)
        st = torch.cuda.current_stream(X.device)
        _lib.check(lib.hb_gp_lml_grad(self.kernel.kind, n, d, X.data_ptr(), fit['ls'].data_ptr(),
                                      float(self.kernel.signal_variance), float(self.noise_variance), Linv.data_ptr(),
                                      alpha.data_ptr(), work.data_ptr(), gd.data_ptr(), st.cuda_stream), 'hb_gp_lml_grad')
        gh = gd.cpu().numpy()                            # [noise | length scales (d) | signal variance], log parameters
        g = np.empty(full.size)
        g[0] = gh[0]
        g[1:1 + nl] = gh[1:1 + d].sum() if nl == 1 else gh[1:1 + d]
        g[1 + nl] = gh[1 + d]
        if nm:                                           # dNLML/dbeta_j = -Phi_j^T alpha
            Phi = torch.stack(self.mean.basis(X, xp=torch))
            g[2 + nl:] = -(Phi @ alpha).cpu().numpy()
        return -fit['lml'], g[free]

    def fit_model(self, fixed: Sequence[str] = (), maxiter: int = 500) -> None:
        """Optimise the hyper-parameters by minimising the negative log marginal likelihood (gp.py:660-697), starting from
        the current values.  ``fixed``: names out of ``'noise_variance'``, ``'length_scales'``, ``'signal_variance'`` that
        keep their values (the reference's ``prior = 'Delta'`` / ``'fixed'`` bounds); coefficients of a ``Mean`` object are
        trained too unless created with ``bounds={'coefficient': 'fixed'}`` / ``Mean.one()`` / ``Mean.zero()``."""
        if not self._is_setup:
            raise RuntimeError("The GP has not been set up yet. Please run the setup() method before fitting.")
        import scipy.optimize as sopt
        nl = self.kernel.length_scales.size
        d = self._X.shape[0]
        beta = self.mean.beta(d) if self.mean is not None else np.zeros(0)
        base = np.concatenate([[np.log(self.noise_variance)], np.log(self.kernel.length_scales),
                               [np.log(self.kernel.signal_variance)], beta])
        names = ['noise_variance'] + ['length_scales'] * nl + ['signal_variance']
        unknown = [f for f in fixed if f not in names]
        if unknown:
            raise ValueError(f"Unknown hyper-parameter name(s) {unknown}")
        free = np.concatenate([np.array([n not in fixed for n in names]),
                               self.mean.free(d) if self.mean is not None else np.zeros(0, dtype=bool)])
        if not free.any():
            return
        res = sopt.minimize(self._nlml_and_grad, base[free], args=(free, base), jac=True, method='L-BFGS-B',
                            options=dict(maxiter=maxiter, ftol=1e-15, gtol=1e-10))
        self._nlml_and_grad(res.x, free, base)   # leaves the fitted hyper-parameters and the factorisation in place
        self._optimization_stats = dict(success=bool(res.success) or res.status == 2, message=str(res.message),
                                        iterations=int(res.nit), nlml=float(res.fun))

    def log_marginal_likelihood(self) -> float:
        if not self._is_setup:
            raise RuntimeError("The GP has not been set up yet. Please run the setup() method before.")
        return self._fit['lml']

    def predict(self, X_query, noise_free: bool = False, mean_only: bool = False):
        """``(mean[1, m], var[1, m])`` for numeric ``X_query`` of shape ``(d, m)`` (NumPy in -> NumPy out, CUDA tensor in
        -> CUDA tensors out); symbolic queries return expressions (gp.py:699-718)."""
        if not self._is_setup:
            raise RuntimeError("The GP has not been set up yet. Please run the setup() method before.")
        if sx.is_symbolic(X_query):
            return self._predict_symbolic(sx.as_vector(X_query), noise_free)
        torch = _lib.require_cuda()
        lib = _lib.load()
        fit = self._fit
        is_tensor = torch.is_tensor(X_query)
        Xq = X_query if is_tensor else torch.as_tensor(np.atleast_2d(np.asarray(X_query, dtype=float)))
        Xq = Xq.to(device=fit['dev'], dtype=torch.float64)
        if Xq.dim() == 1:
            Xq = Xq.reshape(-1, 1)
        d, n = self._X.shape
        if Xq.shape[0] != d:
            raise ValueError(f"Dimension of the query points does not match the number of features: {Xq.shape[0]} vs {d}")
        Xq = Xq.contiguous()
        M = Xq.shape[1]
        f8 = dict(dtype=torch.float64, device=fit['dev'])
        mean = torch.empty(M, **f8)
        var = None if mean_only else torch.empty(M, **f8)
        lib.hb_gp_predict_work_doubles.restype = C.c_size_t
        lib.hb_gp_predict_work_doubles.argtypes = [C.c_int, C.c_int]
        work = torch.empty(lib.hb_gp_predict_work_doubles(n, M), **f8)
        vp, i, dd = C.c_void_p, C.c_int, C.c_double
        lib.hb_gp_predict_kernel.argtypes = [i, i, i, i, vp, vp, vp, dd, dd, dd, vp, vp, vp, vp, vp, vp]
        rc = lib.hb_gp_predict_kernel(self.kernel.kind, n, d, M, fit['X'].data_ptr(), Xq.data_ptr(), fit['ls'].data_ptr(),
                               self.kernel.signal_variance, 0. if noise_free else self.noise_variance, self.prior_mean,
                               fit['Linv'].data_ptr(), fit['alpha'].data_ptr(), mean.data_ptr(),
                               None if var is None else var.data_ptr(), work.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'hb_gp_predict_kernel')
        if self.mean is not None:   # m(x*) = Phi(x*) beta, on the device
            beta = self.mean.beta(d)
            for bj, row in zip(beta, self.mean.basis(Xq, xp=torch)):
                mean += float(bj) * row
        if is_tensor:
            return mean.reshape(1, M), (None if var is None else var.reshape(1, M))
        return mean.cpu().numpy().reshape(1, M), (None if var is None else var.cpu().numpy().reshape(1, M))

    def _predict_symbolic(self, xq, noise_free):
        if self.kernel.kind != 0:
            raise NotImplementedError("Symbolic queries (embedding into a Model) are implemented for the "
                                      "squared-exponential kernel only")
        d, n = self._X.shape
        if len(xq) != d:
            raise ValueError(f"Dimension of the query point does not match the number of features: {len(xq)} vs {d}")
        ls = self.kernel.length_scale_vector(d)
        alpha = self._fit['alpha_host']
        mean = sx.Expr.const(self.prior_mean)
        if self.mean is not None:
            beta = self.mean.beta(d)
            rows = self.mean.basis(np.array([[xi] for xi in xq], dtype=object))
            for bj, row in zip(beta, rows):
                mean = mean + float(bj) * (row[0] if isinstance(row, np.ndarray) else row)
        if n > self.symbolic_unroll_limit:
            # large training set: the mean enters the model as ONE call node (a loop over the training table inside the
            # generated device function, csrc/gp_device.cuh) instead of n unrolled terms; no symbolic variance
            Xs = (self._X / np.asarray(ls, dtype=float)[:, None]).T
            gid = sx.register_gp_table(Xs, np.asarray(alpha, dtype=float) * self.kernel.signal_variance,
                                       1. / np.asarray(ls, dtype=float))
            return np.array([mean + sx.gp_mean_node(gid, xq)], dtype=object), None
        ks = []
        for i in range(n):
            d2 = sx.Expr.const(0.)
            for k in range(d):
                r = (xq[k] - self._X[k, i]) / ls[k]
                d2 = d2 + r * r
            ki = self.kernel.signal_variance * sx.exp(-.5 * d2)
            ks.append(ki)
            mean = mean + alpha[i] * ki
        Linv = self._fit['Linv'].cpu().numpy()
        var = sx.Expr.const(self.kernel.signal_variance + (0. if noise_free else self.noise_variance))
        if n <= 64:  # the variance expression is O(n^2); only built for small training sets
            for i in range(n):
                vi = sx.Expr.const(0.)
                for j in range(i + 1):
                    vi = vi + Linv[i, j] * ks[j]
                var = var - vi * vi
        else:
            var = None
        return np.array([mean], dtype=object), (None if var is None else np.array([var], dtype=object))


GaussianProcess.symbolic_unroll_limit = 32   # training sets up to this size are unrolled into the expression graph

GP = GaussianProcess
