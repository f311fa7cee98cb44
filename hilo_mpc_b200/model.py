"""``Model``: symbolic ODE / difference-equation model with the reference's front-end API.

Mirrors the subset of ``hilo_mpc/modules/dynamic_model/dynamic_model.py`` that the hot-path classes consume:
``set_dynamical_states/inputs/parameters/measurements`` (:737, :1085, :1193, :875), ``set_dynamical_equations`` /
``set_measurement_equations`` / ``set_equations`` (:1293, :1352, :1508), state-space setters ``A,B,C,D`` (:1555-1668),
``setup(dt=)`` (:3727), ``discretize`` (:3600), ``linearize``/``is_linear`` (:3670, :2465), ``scale`` (:3127 /
``base.py:1562-1591``), ``set_initial_conditions`` (:3360), ``simulate`` (:3911), ``copy`` (:3562), ``solution``.

Differences, by design: expressions are ``hilo_mpc_b200.symbolic`` DAGs (no CasADi); ``simulate`` integrates on the
GPU with a generated fixed-step kernel (explicit RK4 sub-stepping instead of CVODES); DAE models are not supported.
"""
from __future__ import annotations

import copy as _copy
import warnings
from typing import Optional, Sequence

import numpy as np

from . import symbolic as sx
from .parsing import parse_dynamic_equations, parse_expression
from .series import TimeSeries


def _names_from_args(args, default_prefix: str, n: Optional[int] = None):
    if len(args) == 1 and isinstance(args[0], (list, tuple, np.ndarray)):
        return [str(k) for k in args[0]]
    if len(args) == 1 and isinstance(args[0], str) and n is not None and n > 1:
        return [f'{args[0]}_{k}' for k in range(n)]
    return [str(k) for k in args]


class Model:
    def __init__(self, id=None, name=None, discrete: bool = False, solver=None, solver_options=None,
                 time_unit: str = 'h', plot_backend=None, use_sx: bool = True):
        self.id = id
        self.name = name or 'model'
        self._discrete = bool(discrete)
        self._solver = solver or ('rk4' if not discrete else None)
        self._solver_options = dict(solver_options or {})
        self.time_unit = time_unit
        self._t = sx.sym('t')
        self._dt = sx.sym('dt')
        self._x = np.empty((0,), dtype=object)
        self._u = np.empty((0,), dtype=object)
        self._p = np.empty((0,), dtype=object)
        self._y = np.empty((0,), dtype=object)
        self._ode = np.empty((0,), dtype=object)
        self._meas = np.empty((0,), dtype=object)
        self._A = self._B = self._C = self._D = None
        self._is_setup = False
        self._dt_value: Optional[float] = None
        self._x0 = None
        self._t0 = 0.
        self._p_values = None
        self._solution = TimeSeries()
        self._x_eq = None
        self._u_eq = None
        self._is_linearized = False
        self._sim_module = None
        self._learned = []  # learned components (GP / ANN) substituted into the equations

    # ------------------------------------------------------------------ symbols
    @property
    def t(self): return self._t
    @property
    def time(self): return self._t
    @property
    def dt(self): return self._dt
    @property
    def sampling_time(self): return self._dt
    @property
    def x(self): return self._x
    @property
    def dynamical_states(self): return self._x
    @property
    def u(self): return self._u
    @property
    def inputs(self): return self._u
    @property
    def p(self): return self._p
    @property
    def parameters(self): return self._p
    @property
    def y(self): return self._y
    @property
    def measurements(self): return self._y
    @property
    def z(self): return np.empty((0,), dtype=object)
    @property
    def n_x(self): return len(self._x)
    @property
    def n_u(self): return len(self._u)
    @property
    def n_p(self): return len(self._p)
    @property
    def n_y(self): return len(self._y)
    @property
    def n_z(self): return 0
    @property
    def n_q(self): return 0
    @property
    def dynamical_state_names(self): return [s.name for s in self._x]
    @property
    def input_names(self): return [s.name for s in self._u]
    @property
    def parameter_names(self): return [s.name for s in self._p]
    @property
    def measurement_names(self): return [s.name for s in self._y]
    @property
    def algebraic_state_names(self): return []
    @property
    def discrete(self): return self._discrete
    @property
    def continuous(self): return not self._discrete
    @property
    def solution(self): return self._solution
    @property
    def ode(self): return self._ode
    @property
    def dynamical_equations(self): return self._ode
    @property
    def meas(self): return self._meas
    @property
    def measurement_equations(self): return self._meas

    def _set_symbols(self, attr, args, prefix, n=None):
        names = _names_from_args(args, prefix, n)
        vec = sx.named_syms(names)
        setattr(self, attr, vec)
        self._is_setup = False
        return vec if len(vec) != 1 else vec  # always indexable (x[0]) like the reference's SX vectors

    def set_dynamical_states(self, *args, **kwargs):
        return self._set_symbols('_x', args, 'x')

    def set_inputs(self, *args, **kwargs):
        return self._set_symbols('_u', args, 'u')

    def set_parameters(self, *args, **kwargs):
        return self._set_symbols('_p', args, 'p')

    def set_measurements(self, *args, **kwargs):
        return self._set_symbols('_y', args, 'y')

    def set_algebraic_states(self, *args, **kwargs):
        raise NotImplementedError("DAE models (algebraic states) are not supported by the B200 engine yet")

    set_algebraic_variables = set_algebraic_states

    def add_parameters(self, parameters, position=None, **kwargs):
        self._p = np.concatenate([self._p, sx.as_vector(parameters)])
        self._is_setup = False

    # ------------------------------------------------------------------ equations
    def _known(self):
        known = {s.name: s for s in np.concatenate([self._x, self._u, self._p, self._y])}
        known['dt'] = self._dt
        known['t'] = self._t
        known['k'] = self._t
        return known

    def _coerce_equations(self, arg):
        if isinstance(arg, str):
            arg = [arg]
        out = []
        for e in (arg if isinstance(arg, (list, tuple, np.ndarray)) else [arg]):
            if isinstance(e, str):
                free = []
                expr, free = parse_expression(e, self._known(), free)
                if free:
                    self._p = np.concatenate([self._p, np.array(free, dtype=object)])
                out.append(expr)
            else:
                out.append(sx.Expr.wrap(e))
        return np.array(out, dtype=object)

    def set_dynamical_equations(self, arg, delimiter=None, check_properties=False):
        self._ode = self._coerce_equations(arg)
        self._A = self._B = None
        self._is_setup = False

    def set_measurement_equations(self, arg, delimiter=None, check_properties=False):
        self._meas = self._coerce_equations(arg)
        self._C = self._D = None
        self._is_setup = False

    def set_algebraic_equations(self, arg, **kwargs):
        raise NotImplementedError("DAE models (algebraic equations) are not supported by the B200 engine yet")

    def set_equations(self, equations=None, **kwargs):
        if equations is not None:
            parsed = parse_dynamic_equations(equations, discrete=self._discrete, dt=self._dt, t=self._t,
                                             x=list(self._x) or None, y=list(self._y) or None,
                                             u=list(self._u) or None, p=list(self._p) or None)
            self._x = np.array(parsed['x'], dtype=object)
            self._y = np.array(parsed['y'], dtype=object)
            self._u = np.array(parsed['u'], dtype=object)
            self._p = np.array(parsed['p'], dtype=object)
            self._ode = np.array(parsed['ode'], dtype=object)
            self._meas = np.array(parsed['meas'], dtype=object)
        ode = kwargs.get('ode')
        if ode is not None:
            self.set_dynamical_equations(ode)
        meas = kwargs.get('meas')
        if meas is not None:
            self.set_measurement_equations(meas)
        if kwargs.get('alg') is not None:
            raise NotImplementedError("DAE models (algebraic equations) are not supported by the B200 engine yet")
        self._is_setup = False

    # ------------------------------------------------------------------ state-space form
    def _ensure_ss_symbols(self, a=None, b=None, c=None):
        if a is not None and self.n_x == 0:
            n = np.asarray(a, dtype=object).shape[0]
            self._x = sx.named_syms(['x'] if n == 1 else [f'x_{k}' for k in range(n)])
        if b is not None and self.n_u == 0:
            m = np.atleast_2d(np.asarray(b, dtype=object)).shape[1]
            self._u = sx.named_syms(['u'] if m == 1 else [f'u_{k}' for k in range(m)])
        if c is not None and self.n_y == 0:
            q = np.atleast_2d(np.asarray(c, dtype=object)).shape[0]
            self._y = sx.named_syms(['y'] if q == 1 else [f'y_{k}' for k in range(q)])

    def _refresh_ss(self):
        if self._A is not None:
            ode = np.asarray(self._A, dtype=object) @ self._x
            if self._B is not None:
                ode = ode + np.asarray(self._B, dtype=object) @ self._u
            self._ode = sx.as_vector(ode)
        if self._C is not None:
            meas = np.asarray(self._C, dtype=object) @ self._x
            if self._D is not None:
                meas = meas + np.asarray(self._D, dtype=object) @ self._u
            self._meas = sx.as_vector(meas)
        self._is_setup = False

    @property
    def A(self): return self._A
    @A.setter
    def A(self, a):
        self._ensure_ss_symbols(a=a)
        self._A = np.atleast_2d(np.asarray(a, dtype=object))
        self._refresh_ss()
    state_matrix = A

    @property
    def B(self): return self._B
    @B.setter
    def B(self, b):
        b = np.asarray(b, dtype=object)
        if b.ndim == 1:
            b = b.reshape(-1, 1)
        self._ensure_ss_symbols(b=b)
        self._B = b
        self._refresh_ss()
    input_matrix = B

    @property
    def C(self): return self._C
    @C.setter
    def C(self, c):
        c = np.atleast_2d(np.asarray(c, dtype=object))
        self._ensure_ss_symbols(c=c)
        self._C = c
        self._refresh_ss()
    output_matrix = C

    @property
    def D(self): return self._D
    @D.setter
    def D(self, d):
        self._D = np.atleast_2d(np.asarray(d, dtype=object))
        self._refresh_ss()
    feedthrough_matrix = D

    # ------------------------------------------------------------------ structure queries
    def is_setup(self) -> bool:
        return self._is_setup

    def is_linear(self) -> bool:
        xu = np.concatenate([self._x, self._u])
        if len(self._ode) and sx.depends_on(sx.jacobian(self._ode, xu).reshape(-1), xu):
            return False
        if len(self._meas) and sx.depends_on(sx.jacobian(self._meas, xu).reshape(-1), xu):
            return False
        return True

    def is_linearized(self) -> bool:
        return self._is_linearized

    def is_time_variant(self) -> bool:
        return sx.depends_on(np.concatenate([self._ode, self._meas]), [self._t])

    def is_autonomous(self) -> bool:
        return self.n_u == 0

    def check_consistency(self):
        if len(self._ode) != self.n_x:
            raise RuntimeError(f"Dimension mismatch between the dynamical states ({self.n_x}) and the dynamical "
                               f"equations ({len(self._ode)})")
        if len(self._meas) != self.n_y:
            raise RuntimeError(f"Dimension mismatch between the measurements ({self.n_y}) and the measurement "
                               f"equations ({len(self._meas)})")

    # ------------------------------------------------------------------ setup / transforms
    def setup(self, dt: Optional[float] = None, **kwargs) -> None:
        self.check_consistency()
        if dt is not None:
            self._dt_value = float(dt)
        elif self._dt_value is None:
            self._dt_value = 1.
        names = {}
        for key, nm in (('x', self.dynamical_state_names), ('u', self.input_names), ('p', self.parameter_names),
                        ('y', self.measurement_names)):
            names[key] = nm
        self._solution = TimeSeries(t=['t'], **names)
        self._sim_module = None
        self._is_setup = True

    @property
    def dt_value(self) -> Optional[float]:
        return self._dt_value

    def copy(self, name=None, setup: bool = False, **kwargs) -> 'Model':
        new = Model.__new__(Model)
        new.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        new._solution = TimeSeries()
        new._sim_module = None
        new._learned = list(self._learned)
        if name is not None:
            new.name = name
        new._is_setup = False
        if setup:
            new.setup(dt=self._dt_value)
        return new

    def discretize(self, method: str, inplace: bool = False, **kwargs) -> Optional['Model']:
        """Explicit Runge-Kutta discretisation ``x_{k+1} = x_k + dt*sum b_i k_i`` of a continuous model
        (``dynamic_model.py:2113-2457`` with ``modeling.py:1214-1281``). ``method``: 'erk' (order=1..4) or 'rk4'."""
        if self._discrete:
            warnings.warn("The model is already discrete; nothing to be done")
            return None if inplace else self.copy()
        from .transcription import tableau_for_order
        if method == 'rk4':
            order = 4
        elif method in ('erk', 'rk'):
            order = int(kwargs.get('order', 1))
        else:
            raise NotImplementedError(f"Discretization method '{method}' is not supported by the B200 engine "
                                      f"(use 'erk' or 'rk4'; collocation is applied inside NMPC/MHE)")
        tab = tableau_for_order(order)
        A, b, c = tab['A'], tab['b'], tab['c']
        ks = []
        for i in range(order):
            xi = self._x
            for j in range(i):
                if A[i, j] != 0.:
                    xi = xi + self._dt * A[i, j] * ks[j]
            ks.append(sx.substitute(self._ode, np.concatenate([[self._t], self._x]),
                                    np.concatenate([[self._t + self._dt * c[i]], xi])))
        xf = self._x
        for i in range(order):
            if b[i] != 0.:
                xf = xf + self._dt * b[i] * ks[i]
        target = self if inplace else self.copy()
        target._ode = sx.as_vector(xf)
        target._discrete = True
        target._A = target._B = None
        target._solver = None
        target._is_setup = False
        return None if inplace else target

    def linearize(self, name: str = None, trajectory=None) -> 'Model':
        """First-order Taylor model around the equilibrium set by ``set_equilibrium_point`` (or kept symbolic in the
        operating point, as the reference does, ``dynamic_model.py:2488-2731``): states/inputs become deviations."""
        lin = self.copy(name=name)
        xu = np.concatenate([self._x, self._u])
        A = sx.jacobian(self._ode, self._x) if self.n_x else np.zeros((0, 0), dtype=object)
        B = sx.jacobian(self._ode, self._u) if self.n_u else np.zeros((self.n_x, 0), dtype=object)
        C = sx.jacobian(self._meas, self._x) if self.n_y else None
        D = sx.jacobian(self._meas, self._u) if self.n_y and self.n_u else None
        if sx.depends_on(np.concatenate([m.reshape(-1) for m in (A, B) if m is not None]), xu):
            if self._x_eq is None:
                raise RuntimeError("Model is nonlinear. Set an equilibrium point (Model.set_equilibrium_point) "
                                   "before linearizing.")
            eq = np.concatenate([self._x_eq, self._u_eq if self._u_eq is not None else np.zeros(self.n_u)])
            A = sx.substitute(A, xu, eq)
            B = sx.substitute(B, xu, eq)
            if C is not None:
                C = sx.substitute(C, xu, eq)
            if D is not None:
                D = sx.substitute(D, xu, eq)
        lin._A, lin._B, lin._C, lin._D = A, (B if self.n_u else None), C, D
        lin._refresh_ss()
        lin._is_linearized = True
        return lin

    def set_equilibrium_point(self, x_eq=None, z_eq=None, u_eq=None):
        self._x_eq = None if x_eq is None else np.asarray(x_eq, dtype=float).reshape(-1)
        self._u_eq = None if u_eq is None else np.asarray(u_eq, dtype=float).reshape(-1)

    def scaled(self, x_scale=None, u_scale=None):
        """Return ``(ode, meas)`` in scaled coordinates ``x = x~ * s`` with ``ode_i /= s_i`` for continuous models
        (``base.py:1562-1591``); for discrete models the successor state is divided by the scale as well."""
        ode, meas = self._ode, self._meas
        olds, news = [], []
        if x_scale is not None:
            olds.append(self._x)
            news.append(self._x * np.asarray(x_scale, dtype=float))
        if u_scale is not None:
            olds.append(self._u)
            news.append(self._u * np.asarray(u_scale, dtype=float))
        if olds:
            ode = sx.substitute(ode, np.concatenate(olds), np.concatenate(news))
            if len(meas):
                meas = sx.substitute(meas, np.concatenate(olds), np.concatenate(news))
        if x_scale is not None:
            ode = ode / np.asarray(x_scale, dtype=float)
        return sx.as_vector(ode), sx.as_vector(meas)

    # ------------------------------------------------------------------ learned components
    def substitute_from(self, obj) -> None:
        """Replace parameters named like the labels of a learned model (GP / ANN) by its symbolic prediction
        (``dynamic_model.py:3040-3125``)."""
        labels = [lbl for lbl in obj.labels]
        names = self.parameter_names
        features = []
        known = self._known()
        for f in obj.features:
            if f not in known:
                raise ValueError(f"Feature '{f}' of the learned model is not a variable of this model")
            features.append(known[f])
        pred = obj.predict(np.array(features, dtype=object))
        mean = pred[0] if isinstance(pred, tuple) else pred
        mean = sx.as_vector(mean)
        olds, news, keep = [], [], []
        for k, nm in enumerate(names):
            if nm in labels:
                olds.append(self._p[k])
                news.append(mean[labels.index(nm)])
            else:
                keep.append(self._p[k])
        if not olds:
            raise ValueError("None of the labels of the learned model matches a parameter of this model")
        self._ode = sx.substitute(self._ode, olds, news)
        if len(self._meas):
            self._meas = sx.substitute(self._meas, olds, news)
        self._p = np.array(keep, dtype=object)
        self._learned.append(obj)
        self._is_setup = False

    # ------------------------------------------------------------------ initial conditions / simulation
    def set_initial_conditions(self, x0, t0: float = 0., z0=None) -> None:
        x0 = np.asarray(x0, dtype=float).reshape(-1)
        if x0.size != self.n_x:
            raise ValueError(f"Dimension mismatch. Supplied {x0.size} initial states, required {self.n_x}")
        self._x0 = x0
        self._t0 = float(t0)
        self._solution.set('t', [t0])
        self._solution.set('x', x0)

    def set_initial_parameter_values(self, p) -> None:
        p = np.asarray(p, dtype=float).reshape(-1)
        if p.size != self.n_p:
            raise ValueError(f"Dimension mismatch. Supplied {p.size} parameter values, required {self.n_p}")
        self._p_values = p

    def reset_solution(self, keep_initial_conditions: bool = True, keep_equilibrium_point: bool = True) -> None:
        self._solution.reset(keep_first=False)
        if keep_initial_conditions and self._x0 is not None:
            self._solution.set('t', [self._t0])
            self._solution.set('x', self._x0)

    def simulate(self, u=None, p=None, steps: int = 1, **kwargs):
        """Advance the stored state by ``steps`` sampling intervals on the GPU (``dynamic_model.py:3911-4212``)."""
        if not self._is_setup:
            raise RuntimeError("Model is not set up. Run Model.setup() before running simulations.")
        if self._solution.is_empty('x'):
            raise RuntimeError("No initial dynamical states found. Please set initial conditions before simulating "
                               "the model!")
        from .simulate import simulate_model
        return simulate_model(self, u=u, p=p, steps=steps, **kwargs)
