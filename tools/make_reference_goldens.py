#!/usr/bin/env python
"""Golden vectors produced by EXECUTING the reference's own source (read from /root/reference at generation time only).

CasADi is not installable here, so the reference package cannot be imported.  Some routines on the hot path are pure
NumPy/SciPy though and touch CasADi only through a couple of trivial calls; this script extracts their source text from
the reference tree, executes it unmodified in a namespace with a *stub* ``ca`` (only ``collocation_points``, ``sum2`` and
``DM``-free helpers, restated below) and stores inputs and outputs as fixtures under ``tests/golden/``:

* ``RungeKutta._construct_polynomial_basis`` (hilo_mpc/util/modeling.py:1091-1127): the collocation constants B, C, D, tau
  every collocation transcription is built from (NMPC, MHE).
* ``lhsnorm`` (hilo_mpc/modules/estimator/pf.py:425-447): Latin-hypercube normal sampling of the particle filter.
* the resampling / roughening / moment block of ``ParticleFilter.estimate`` (pf.py:405-421), with the legacy global RNG.
* the equation-string parser (hilo_mpc/util/parsing.py, the whole module) and ``Activation`` / ``net_to_casadi_graph``
  (hilo_mpc/util/machine_learning.py) on symbols that are closures over a name -> value environment.
* ``RungeKutta._collocation`` / ``RungeKutta._explicit`` (hilo_mpc/util/modeling.py:1129-1281): one interval of the
  transcription on a toy problem, ``ca.Function`` calls with symbolic arguments emulated by re-binding names.
* ``GenericCost`` / ``QuadraticCost`` / ``MHEQuadraticCost`` (hilo_mpc/util/modeling.py:38-818) from their class sources on
  stub models, and the matrix-construction lines of ``LMPC.setup`` (mpc.py:2179-2266) with a numeric stub.
* the horizon loop of ``NMPC._setup`` (mpc.py: ``for ii in range(self._prediction_horizon)`` ... ``time += dt_ii``) around
  the executed ``RungeKutta._collocation``: objective and constraint vector of the whole NLP; the loop of ``MHE.setup``
  (mhe.py) the same way on the reaction model of the MHE tests.
* ``UnscentedKalmanFilter._setup_parameters / _setup_predict / _setup_update`` (kf.py:486-604) on numeric matrices.
* ``NMPC._get_tvp_parameters_values`` (mpc.py:293-333): the horizon of stored time-varying parameter values iteration by
  iteration, including what it does once the horizon runs past the end of the data.

Run:  python tools/make_reference_goldens.py   (needs /root/reference; the tests only read the committed JSON).
"""
import ast
import json
import os
import sys
import textwrap

import numpy as np
from scipy.stats import norm

REF = os.environ.get('HILO_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden',
                   'reference_numeric.json')


class _CaStub:
    """The two CasADi calls the extracted sources make."""

    @staticmethod
    def collocation_points(degree, method):
        # casadi.collocation_points: Gauss-Legendre points / Radau IIA points on (0, 1] (roots of the shifted Legendre /
        # of P_{n-1} - P_n), computed here from the Legendre polynomials
        from numpy.polynomial import legendre as L
        if method == 'legendre':
            r = L.legroots([0] * degree + [1])
        elif method == 'radau':
            c = np.zeros(degree + 1)
            c[degree - 1], c[degree] = 1., -1.          # P_{n-1} - P_n has the root x = 1 (right end point)
            r = L.legroots(c)
        else:
            raise ValueError(method)
        return sorted(float(v) for v in (np.real(r) + 1.) / 2.)

    @staticmethod
    def sum2(a):
        return np.sum(np.asarray(a), axis=1)


class _H(float):
    """Stands in for the symbolic step size ``h``: ``h.zeros(n)`` and ``h * tau`` are all the source needs."""

    def zeros(self, n):
        return np.zeros(n)


def _function_source(path, name, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    nodes = tree.body
    if cls is not None:
        nodes = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == name)
    lines = src.split('\n')[fn.lineno - 1:fn.end_lineno]
    return textwrap.dedent('\n'.join(lines)), (fn.lineno, fn.end_lineno)


class _SX:
    """Stub of ``casadi.SX`` for the parser: a named symbol or a closure over an environment ``{name: value}``."""

    def __init__(self, fn, name=None):
        self.fn, self._name = fn, name

    @staticmethod
    def sym(name):
        return _SX(lambda env, n=name: env[n], name)

    def name(self):
        return self._name

    def elements(self):
        return [self]

    @staticmethod
    def _f(v):
        return v.fn if isinstance(v, _SX) else (lambda env, c=float(v): c)

    @staticmethod
    def _bin(a, b, op):
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return NotImplemented          # let numpy broadcast over vectors of symbols
        fa, fb = _SX._f(a), _SX._f(b)
        return _SX(lambda env: op(fa(env), fb(env)))

    def zeros(self, n):                    # h.zeros(n) in RungeKutta._construct_polynomial_basis
        return np.zeros(n, dtype=object)

    def __add__(self, o): return _SX._bin(self, o, lambda a, b: a + b)
    def __radd__(self, o): return _SX._bin(o, self, lambda a, b: a + b)
    def __sub__(self, o): return _SX._bin(self, o, lambda a, b: a - b)
    def __rsub__(self, o): return _SX._bin(o, self, lambda a, b: a - b)
    def __mul__(self, o): return _SX._bin(self, o, lambda a, b: a * b)
    def __rmul__(self, o): return _SX._bin(o, self, lambda a, b: a * b)
    def __truediv__(self, o): return _SX._bin(self, o, lambda a, b: a / b)
    def __rtruediv__(self, o): return _SX._bin(o, self, lambda a, b: a / b)
    def __pow__(self, o): return _SX._bin(self, o, lambda a, b: a ** b)
    def __rpow__(self, o): return _SX._bin(o, self, lambda a, b: a ** b)

    def __neg__(self):
        f = self.fn
        return _SX(lambda env: -f(env))


def _stub_casadi_module():
    import math
    import types
    m = types.ModuleType('casadi')
    m.SX = _SX
    m.MX = _SX

    def lift1(fun):
        def wrapped(a):
            if isinstance(a, np.ndarray):      # vectors are object arrays of scalar symbols
                return np.array([wrapped(v) for v in a.flat], dtype=object).reshape(a.shape)
            if isinstance(a, _SX):
                f = a.fn
                return _SX(lambda env: fun(f(env)))
            return fun(a)
        return wrapped

    def lift2(fun):
        def wrapped(a, b):
            if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
                a2, b2 = np.broadcast_arrays(np.asarray(a, dtype=object), np.asarray(b, dtype=object))
                return np.array([wrapped(u, v) for u, v in zip(a2.flat, b2.flat)], dtype=object).reshape(a2.shape)
            return _SX._bin(a, b, fun) if isinstance(a, _SX) or isinstance(b, _SX) else fun(a, b)
        return wrapped
    m.sum1 = lambda a: np.sum(a, axis=0)
    for name, fun in (('sqrt', math.sqrt), ('exp', math.exp), ('log', math.log), ('log10', math.log10),
                      ('sign', lambda v: float((v > 0) - (v < 0))), ('fabs', abs), ('sin', math.sin), ('cos', math.cos),
                      ('tan', math.tan), ('asin', math.asin), ('acos', math.acos), ('atan', math.atan),
                      ('sinh', math.sinh), ('cosh', math.cosh), ('tanh', math.tanh), ('asinh', math.asinh),
                      ('acosh', math.acosh), ('atanh', math.atanh)):
        setattr(m, name, lift1(fun))
    for name, fun in (('fmin', min), ('fmax', max), ('atan2', math.atan2), ('plus', lambda a, b: a + b),
                      ('minus', lambda a, b: a - b), ('times', lambda a, b: a * b), ('rdivide', lambda a, b: a / b),
                      ('power', lambda a, b: a ** b)):
        setattr(m, name, lift2(fun))
    return m


PARSER_CASES = [
    # (name, discrete, equations) - strings of the reference's tests (tests/test_parsing.py, test_KFs.py, test_MHE.py,
    # test_NMPC.py) restricted to ODE models, plus the features the front end has to get right: constants, auxiliary
    # expressions that are inlined, inputs v(k), parameters in order of appearance, '^', line continuation '...'
    ('double_replace', False, """
        dx/dt = u(k)
        dy/dt = ar + ar

        ar = sin(2*x(t))
        """),
    ('hill_oscillator', False, """
        # ODE's
        d/dt(x_1(t)) = x_2(t)
        d/dt(x_2(t)) = -p*x_1(t)

        # Algebraic equations
        p = r - 2*q*cos(omega*t)

        # Measurement equations
        y(k) = x_1(t) + x_2(t)*sin(omega*t)
        """),
    ('hill_oscillator_discrete', True, """
        # ODE's
        x_1(k + 1) = x_1(k) + x_2(k)*dt
        x_2(k + 1) = x_2(k) - p*x_1(k)*dt

        # Algebraic equations
        p = r - 2*q*cos(omega*t(k))

        # Measurement equations
        y(k) = x_1(k) + x_2(k)*sin(omega*t(k))
        """),
    ('cart_pole_nested_derivative', False, """
        # DAE's
        dx(t)/dt = v(t)
        dv(t)/dt = (3/4*m*g*sin(theta(t))*cos(theta(t)) - 1/2*m*l*sin(theta(t))*omega(t)^2 + F(k))/...
        (M + m - 3/4*m*cos(theta(t))^2)
        d/dt(theta(t)) = omega(t)
        d/dt(omega(t)) = 3/(2*l)*(dv(t)/dt*cos(theta(t)) + g*sin(theta(t)))
        y(k) = h + l*cos(theta(t))

        # Constants
        g = 9.81
        """),
    ('bioreactor', False, """
        dT/dt = alpha*(T_amb - T(t))
        dcB/dt = r - D(k)*cB(t)
        dcS/dt = -r/Y - D(k)*cS(t)

        y_0(k) = T(t)
        y_1(k) = cB(t)

        r = (mu_0 + mu_1*T(t))*cS(t)*cB(t)/(K + cS(t))
        """),
    ('kf_two_tank', False, """
        dx_1/dt = -k_1*x_1(t) + u(k)
        dx_2/dt = k_1*x_1(t) - k_2*x_2(t)
        y(k) = x_2(t)
        """),
    ('functions', False, """
        da/dt = -sqrt(abs(a(t))) + exp(-b(t)^2)*u_1(k) + log(1 + c^2) + tanh(a(t))*arctan(b(t))
        db/dt = min(a(t), b(t)) - max(u_2(k), 0.5) + arctan2(a(t), 1 + b(t)^2) + sign(u_1(k))*log10(10 + a(t)^2)
        z_out(k) = sinh(a(t)/10) + cosh(b(t)/10) + arsinh(a(t)) - cos(c)*tan(b(t)/4)
        c = 0.75
        """),
]


def _parser_goldens():
    import importlib.util
    import warnings
    path = os.path.join(REF, 'hilo_mpc', 'util', 'parsing.py')
    sys.modules['casadi'] = _stub_casadi_module()
    try:
        spec = importlib.util.spec_from_file_location('_ref_parsing', path)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            spec.loader.exec_module(mod)      # the reference file itself, unmodified
            cases = []
            for name, discrete, text in PARSER_CASES:
                t, dt = _SX.sym('t'), _SX.sym('dt')
                model = mod.parse_dynamic_equations(text.split('\n'), discrete=discrete, t=[t], dt=[dt])
                names = {k: [v.name() for v in model[k]] for k in ('x', 'y', 'z', 'u', 'p')}
                rng = np.random.RandomState(abs(hash(name)) % 1000 if False else len(name))
                allnames = sorted(set(sum(names.values(), [])) | {'t', 'dt'})
                points = []
                for _ in range(3):
                    env = {n: float(rng.uniform(.2, 1.5)) for n in allnames}
                    vals = {k: [float(_SX._f(e)(env)) for e in model[k]] for k in ('ode', 'alg', 'meas', 'quad')}
                    points.append({'env': env, 'values': vals})
                cases.append({'name': name, 'discrete': discrete, 'equations': text, 'names': names,
                              'const': {k: float(v) for k, v in model['const'].items()}, 'points': points})
    finally:
        del sys.modules['casadi']
    return {'source': 'hilo_mpc/util/parsing.py (whole module, executed with a stub casadi)', 'cases': cases}


class _LazyEnv:
    """Environment with re-bound names: a binding is evaluated (in the outer environment) only when it is looked up, so
    a substitution of variables the expression does not depend on needs no value for them."""

    def __init__(self, outer, binds):
        self.outer, self.binds = outer, binds

    def __getitem__(self, name):
        b = self.binds.get(name)
        return self.outer[name] if b is None else b(self.outer)


def _rebind(e, names, vals):
    f = _SX._f(e)
    binds = dict(zip(names, vals))
    return _SX(lambda env: f(_LazyEnv(env, binds)))


class _Vec(np.ndarray):
    """Vector / matrix of ``_SX`` scalars with the few ``casadi.SX`` methods the discretisation routines call."""

    def __new__(cls, items=(), shape=None):
        a = np.empty(len(items), dtype=object)
        for k, v in enumerate(items):
            a[k] = v
        if shape is not None:
            a = a.reshape(shape)
        return a.view(cls)

    def size1(self):
        return self.shape[0] if self.ndim else 1

    def is_empty(self):
        return self.size == 0

    @staticmethod
    def sym(name, n=1):
        if isinstance(n, tuple):
            return _Vec([_SX.sym(f'{name}_{i}_{j}') for i in range(n[0]) for j in range(n[1])], shape=n)
        return _Vec([_SX.sym(f'{name}_{i}') for i in range(n)])


class _SymFunction:
    """``ca.Function(name, ins, outs)`` on closures: a call with new (symbolic or numeric) arguments rebinds the names of
    the input symbols in the environment the output closures are evaluated in."""

    def __init__(self, name, ins, outs, *opts):
        self.names = [[s.name() for s in np.asarray(v, dtype=object).reshape(-1)] for v in ins]
        self.outs = [np.asarray(o, dtype=object).reshape(-1) for o in outs]

    def __call__(self, *args):
        binds = []
        for names, a in zip(self.names, args):
            vals = list(np.asarray(a, dtype=object).reshape(-1)) if not isinstance(a, _SX) else [a]
            assert len(vals) == len(names), (names, vals)
            binds += [(n, _SX._f(v)) for n, v in zip(names, vals)]

        def rebound(e):
            return _rebind(e, [n for n, _ in binds], [b for _, b in binds])
        res = [_Vec([rebound(e) for e in o]) for o in self.outs]
        return res[0] if len(res) == 1 else res


def _discretisation_goldens():
    """``RungeKutta._collocation`` and ``RungeKutta._explicit`` (hilo_mpc/util/modeling.py:1128-1281) executed from source
    on a toy problem with explicit time dependence and a Lagrange term: collocation equations, end state, quadrature /
    one explicit Runge-Kutta step, evaluated at random points."""
    path = os.path.join(REF, 'hilo_mpc', 'util', 'modeling.py')
    ca = _stub_casadi_module()
    ca.SX = _Vec
    ca.MX = _Vec
    ca.Function = _SymFunction
    ca.vertcat = lambda *a: _Vec([e for v in a for e in np.asarray(v, dtype=object).reshape(-1)])
    ca.collocation_points = _CaStub.collocation_points
    src = open(path).read()
    tree = ast.parse(src)
    lines = src.split('\n')
    node = next(n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], 'id', '') == 'EXPLICIT_METHODS')
    ns = {'np': np, 'ca': ca, 'Symbolic': object}
    exec('\n'.join(lines[node.lineno - 1:node.end_lineno]), ns)
    spans = {}
    body = []
    for name, deco in (('_construct_polynomial_basis', '@staticmethod'), ('_collocation', '@classmethod'),
                       ('_explicit', '@classmethod')):   # the decorators the methods carry in the reference
        fsrc, spans[name] = _function_source(path, name, cls='RungeKutta')
        body.append(textwrap.indent(deco + '\n' + fsrc.replace(deco + '\n', ''), '    '))
    exec('class RungeKutta:\n' + '\n'.join(body), ns)
    RK = ns['RungeKutta']

    def problem():
        t, dt = _Vec.sym('t'), _SX.sym('dt_0')
        x, u, p = _Vec.sym('x', 2), _Vec.sym('u', 1), _Vec.sym('p', 1)
        z = _Vec.sym('z', 0)
        sin, cos = ca.sin, ca.cos
        ode = _Vec([x[1] + p[0] * sin(t[0]) * u[0], -x[0] * x[0] * x[1] + cos(x[0]) * u[0] - .3 * t[0]])
        quad = _Vec([x[0] * x[0] + .1 * u[0] * u[0] + x[1] * sin(t[0])])
        return dict(ode=ode, alg=_Vec([]), quad=quad, meas=None, t=t, x=x, z=z, u=u, p=p, dt=dt)
    out = {'source': f"hilo_mpc/util/modeling.py:{spans['_collocation'][0]}-{spans['_explicit'][1]}",
           'toy': 'x0\' = x1 + p sin(t) u,  x1\' = -x0^2 x1 + cos(x0) u - 0.3 t,  l = x0^2 + 0.1 u^2 + x1 sin(t)',
           'collocation': [], 'explicit': []}
    rng = np.random.RandomState(12)

    def ev(e, env):
        return [float(_SX._f(v)(env)) for v in np.asarray(e, dtype=object).reshape(-1)]
    for points in ('radau', 'legendre'):
        for degree in (2, 3, 4):
            pr = problem()
            RK._collocation(pr, degree=degree, collocation_points=points)
            pts = []
            for _ in range(2):
                env = {'t_0': rng.uniform(0., 2.), 'dt_0': rng.uniform(.05, .5), 'x_0': rng.normal(), 'x_1': rng.normal(),
                       'u_0': rng.normal(), 'p_0': rng.normal()}
                for k in range(degree):
                    env[f'x_{k}_0'], env[f'x_{k}_1'] = rng.normal(), rng.normal()
                pts.append({'env': env, 'ce': ev(pr['collocation_equations'], env), 'xf': ev(pr['ode'], env),
                            'J': ev(pr['quad'], env)})
            out['collocation'].append({'points': points, 'degree': degree, 'evals': pts})
    for order in (1, 2, 3, 4):
        pr = problem()
        RK._explicit(pr, order=order)
        pts = []
        for _ in range(2):
            env = {'t_0': rng.uniform(0., 2.), 'dt_0': rng.uniform(.05, .5), 'x_0': rng.normal(), 'x_1': rng.normal(),
                   'u_0': rng.normal(), 'p_0': rng.normal()}
            pts.append({'env': env, 'xf': ev(pr['ode'], env), 'J': ev(pr['quad'], env)})
        out['explicit'].append({'order': order, 'evals': pts})
    return out


def _quadratic_cost_goldens():
    """``GenericCost`` + ``QuadraticCost`` (hilo_mpc/util/modeling.py:38-531) executed from their class sources (helpers
    from hilo_mpc/util/util.py likewise) on a stub model: reference tracking of states and inputs with scaling, a
    stabilisation term, a full weight matrix and the input-change term, exactly as ``NMPC._define_cost_terms`` sets them
    up (mpc.py:207-242).  Stored: the cost as a function of the (scaled) model variables and the ``u_old`` placeholder."""
    import collections
    path = os.path.join(REF, 'hilo_mpc', 'util', 'modeling.py')
    upath = os.path.join(REF, 'hilo_mpc', 'util', 'util.py')
    ca = _stub_casadi_module()

    class _DM(np.ndarray):
        def __new__(cls, a=()):
            return np.asarray(a, dtype=float).view(cls)
    ca.DM = _DM
    ca.SX = _Vec
    ca.MX = type('_MX', (), {})          # nothing is an MX here
    ca.vertcat = lambda *a: _Vec([e for v in a for e in np.asarray(v, dtype=object).reshape(-1)])
    ca.mtimes = lambda a, b: np.dot(a, b)

    def substitute(expr, old, new):
        names = [sym.name() for sym in np.asarray(old, dtype=object).reshape(-1)]
        vals = [_SX._f(v) for v in np.asarray(new, dtype=object).reshape(-1)]
        assert len(names) == len(vals)

        def rebound(e):
            return _rebind(e, names, vals)
        if isinstance(expr, np.ndarray):
            return _Vec([rebound(e) for e in expr.reshape(-1)])
        return rebound(expr)
    ca.substitute = substitute
    ns = {'np': np, 'ca': ca, 'sys': sys, 'warnings': __import__('warnings')}
    for name in ('check_and_wrap_to_list', 'check_if_has_duplicates', 'check_if_list_of_type', 'check_if_list_of_none',
                 'check_if_square', 'who_am_i'):
        exec(_function_source(upath, name)[0], ns)
    spans = {}
    for name in ('GenericCost', 'QuadraticCost'):
        src, spans[name] = _class_source(path, name)
        exec(src, ns)

    class _Names(_Vec):
        pass

    def named(names):
        v = _Vec([_SX.sym(n) for n in names]).view(_Names)
        v._names = list(names)
        return v

    class _Model:
        def __init__(self):
            self._x, self._u, self._y = named(['c_A', 'c_B', 'T']), named(['F', 'Q']), named([])
            self.dynamical_state_names, self.input_names, self.measurement_names = self._x._names, self._u._names, []
            self.n_y, self.meas = 0, _Vec([])
    cases = []
    rng = np.random.RandomState(21)
    x_scale, u_scale = [1., 2., 100.], [10., 1000.]
    setups = collections.OrderedDict([
        ('reference_tracking_scaled', lambda q: (q.add_states(names=['c_A', 'T'], weights=[1., 10.], ref=[2.1, 110.]),
                                                 q.add_inputs(names=['F', 'Q'], weights=[.1, .2], ref=[14., -1100.]))),
        ('stabilisation_and_matrix_weight', lambda q: (q.add_states(names=['c_B'], weights=3.),
                                                       q.add_inputs(names=['Q', 'F'], weights=np.array([[2., .5], [.5, 1.]])))),
        ('input_change', lambda q: (q.add_states(names=['c_A', 'c_B', 'T'], weights=[1., 2., 3.], ref=[1., 1., 100.]),
                                    q.add_inputs_change(names=['F'], weights=[5.]),
                                    q.add_inputs_change(names=['Q'], weights=[7.]))),
    ])
    for label, build in setups.items():
        model = _Model()
        q = ns['QuadraticCost'](model)
        build(q)
        q._setup(x_scale=x_scale, u_scale=u_scale, y_scale=[], path_variables=[], time_variable=None)
        old_names = [sym.name() for sym in np.asarray(q.input_change_placeholder, dtype=object).reshape(-1)] \
            if len(q.input_change_placeholder) else []
        pts = []
        for _ in range(3):
            env = {n: float(rng.normal()) for n in model._x._names + model._u._names + old_names}
            pts.append({'env': env, 'cost': float(_SX._f(q.cost)(env))})
        cases.append({'label': label, 'x_scale': x_scale, 'u_scale': u_scale, 'u_old_names': old_names,
                      'ind_input_changes': [int(i) for i in q._ind_input_changes], 'points': pts})
    # ---- the MHE flavour: costs on PHYSICAL quantities (x -> x * scale is substituted in _setup), guesses and
    # measurements as placeholders
    src, spans['MHEQuadraticCost'] = _class_source(path, 'MHEQuadraticCost')
    exec(src, ns)

    class _MheModel:
        def __init__(self):
            self.x, self.p, self.u, self.y = named(['a', 'b']), named(['k']), named(['f']), named(['y1', 'y2'])
            self._x, self._p, self._u, self._y = self.x, self.p, self.u, self.y
            self.dynamical_state_names, self.parameter_names = self.x._names, self.p._names
            self.input_names, self.measurement_names = self.u._names, self.y._names
            self.n_x, self.n_p, self.n_u, self.n_y = 2, 1, 1, 2
            self.meas = _Vec([self.x[0] * self.x[1] + self.p[0], self.x[1] * self.u[0]])
    xs, ps, ws = [2., 5.], [3.], [.5, .25]
    mm = _MheModel()
    arr = ns['MHEQuadraticCost'](mm)
    arr.add_states(weights=[10., 5.], guess=[.5, .2])
    arr.add_parameters(weights=[2.], guess=[1.])
    arr._setup(x_scale=xs, w_scale=ws, p_scale=ps, u_scale=[1.])
    stg = ns['MHEQuadraticCost'](mm)
    stg.add_measurements(weights=[100., 10.])
    stg.add_state_noise(weights=[20., 30.])
    stg._setup(x_scale=xs, w_scale=ws, p_scale=ps, u_scale=[1.])

    def placeholders(q):
        return [sym.name() for lst in (q._iter_ref_list, q._tv_ref_list) for r in lst
                for sym in np.asarray(r['placeholder'], dtype=object).reshape(-1)]
    mhe_pts = []
    w_names = [sym.name() for sym in np.asarray(stg._w, dtype=object).reshape(-1)]
    for _ in range(3):
        env = {n: float(rng.normal()) for n in ['a', 'b', 'k', 'f'] + w_names + placeholders(arr) + placeholders(stg)}
        mhe_pts.append({'env': env, 'arrival': float(_SX._f(arr.cost)(env)), 'stage': float(_SX._f(stg.cost)(env))})
    mhe = {'x_scale': xs, 'p_scale': ps, 'w_scale': ws, 'w_names': w_names, 'arrival_placeholders': placeholders(arr),
           'stage_placeholders': placeholders(stg), 'points': mhe_pts,
           'model': "x = [a, b], p = [k], u = [f], y1 = a b + k, y2 = b f"}
    return {'source': f"hilo_mpc/util/modeling.py:{spans['GenericCost'][0]}-{spans['MHEQuadraticCost'][1]}", 'cases': cases,
            'mhe': mhe}


def _lmpc_goldens():
    """The matrix construction of ``LMPC.setup`` (hilo_mpc/modules/controller/mpc.py, from ``n_x = self._n_x`` to
    ``v_ub = ca.vertcat(...)``) executed from its source lines with a numeric stub ``ca`` (DM = NumPy arrays): the QP the
    reference really hands to qpOASES - in particular the input block of the equality constraints."""
    import scipy.linalg
    path = os.path.join(REF, 'hilo_mpc', 'modules', 'controller', 'mpc.py')
    lines = open(path).read().split('\n')
    tree = ast.parse('\n'.join(lines))
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'LMPC')
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'setup')
    body = lines[fn.lineno - 1:fn.end_lineno]
    start = next(i for i, ln in enumerate(body) if ln.strip() == 'n_x = self._n_x')
    end = next(i for i, ln in enumerate(body) if ln.strip().startswith('v_ub = ca.vertcat('))
    block = textwrap.dedent('\n'.join(body[start:end + 1]))

    class _DMns:
        @staticmethod
        def zeros(*shape):
            shape = shape[0] if len(shape) == 1 and isinstance(shape[0], tuple) else shape
            return np.zeros(shape if len(shape) > 1 else (shape[0], 1))

        @staticmethod
        def ones(*shape):
            shape = shape[0] if len(shape) == 1 and isinstance(shape[0], tuple) else shape
            return np.ones(shape if len(shape) > 1 else (shape[0], 1))

        @staticmethod
        def eye(n):
            return np.eye(n)
    ca = type('ca', (), {})()
    ca.DM = _DMns
    ca.kron = lambda a, b: np.kron(np.atleast_2d(a), np.atleast_2d(b))
    ca.horzcat = lambda *a: np.hstack([np.atleast_2d(v) for v in a])
    ca.vertcat = lambda *a: np.vstack([np.atleast_2d(v) for v in a])
    ca.diagcat = lambda *a: scipy.linalg.block_diag(*a)
    cases = []
    rng = np.random.RandomState(8)
    for nx, nu, N in ((2, 1, 10), (2, 1, 3), (3, 2, 4), (1, 1, 5)):
        if (nx, nu) == (2, 1):
            A, B = np.array([[1., .5], [0., 1.]]), np.array([[.125], [.5]])      # double integrator, tests/test_LMPC.py:8-33
        else:
            A, B = rng.normal(size=(nx, nx)), rng.normal(size=(nx, nu))
        Q, R, P = np.diag(rng.uniform(.5, 2., nx)), np.diag(rng.uniform(.5, 2., nu)), np.diag(rng.uniform(.5, 2., nx))

        class _Mdl:
            state_matrix, input_matrix, n_x = A, B, nx

        class _Self:
            _n_x, _n_u, _horizon, horizon, _n_tvp = nx, nu, N, N, 0
            _model = _Mdl()
            _x_lb, _x_ub = -np.arange(1, nx + 1.).reshape(-1, 1), np.arange(1, nx + 1.).reshape(-1, 1) * 2
            _u_lb, _u_ub = -np.ones((nu, 1)), np.ones((nu, 1)) * 3
        obj = _Self()
        obj.Q, obj.R, obj.P = Q, R, P
        env = {'np': np, 'ca': ca, 'self': obj}
        exec(block, env)
        cases.append({'nx': nx, 'nu': nu, 'N': N, 'A': A.tolist(), 'B': B.tolist(), 'Q': Q.tolist(), 'R': R.tolist(),
                      'P': P.tolist(), 'Aeq': np.asarray(env['Aeq']).tolist(), 'beq': np.asarray(env['beq']).reshape(-1).tolist(),
                      'H': np.asarray(env['H']).tolist(), 'v_lb': np.asarray(env['v_lb']).reshape(-1).tolist(),
                      'v_ub': np.asarray(env['v_ub']).reshape(-1).tolist(), 'x_ind': env['x_ind'], 'u_ind': env['u_ind']})
    return {'source': f'hilo_mpc/modules/controller/mpc.py:{fn.lineno + start}-{fn.lineno + end}', 'cases': cases}


def _nmpc_assembly_goldens():
    """The horizon loop of ``NMPC._setup`` (hilo_mpc/modules/controller/mpc.py, ``for ii in range(self._prediction_horizon)``
    ... ``time += dt_ii``) executed from its source lines: how the per-interval pieces are assembled into the NLP - the
    collocation and continuity constraints interval by interval, the sum of the quadratures, the Mayer term at
    ``time + dt`` of the last interval, the running time, and ``u_old``: the parameter for the first interval, the
    interval's OWN input afterwards.  The per-interval function is the executed
    ``RungeKutta._collocation`` (above); the glue that hands it to the loop is this script's."""
    mpath = os.path.join(REF, 'hilo_mpc', 'util', 'modeling.py')
    cpath = os.path.join(REF, 'hilo_mpc', 'modules', 'controller', 'mpc.py')
    opath = os.path.join(REF, 'hilo_mpc', 'modules', 'optimizer.py')
    ca = _stub_casadi_module()
    ca.SX = _Vec
    ca.MX = _Vec
    ca.Function = _SymFunction
    ca.vertcat = lambda *a: _Vec([e for v in a for e in np.asarray(v, dtype=object).reshape(-1)])
    ca.collocation_points = _CaStub.collocation_points
    ns = {'np': np, 'ca': ca, 'Symbolic': object}
    body = []
    for name, deco in (('_construct_polynomial_basis', '@staticmethod'), ('_collocation', '@classmethod')):
        fsrc, _ = _function_source(mpath, name, cls='RungeKutta')
        body.append(textwrap.indent(deco + '\n' + fsrc.replace(deco + '\n', ''), '    '))
    exec('class RungeKutta:\n' + '\n'.join(body), ns)
    N, deg, nx = 3, 2, 2
    t, dt = _Vec.sym('t'), _SX.sym('dt_0')
    x, u, p = _Vec.sym('x', nx), _Vec.sym('u', 1), _Vec.sym('p', 2)        # p = [model parameter, u_old]
    sin, cos = ca.sin, ca.cos
    pr = dict(ode=_Vec([x[1] + p[0] * sin(t[0]) * u[0], -x[0] * x[0] * x[1] + cos(x[0]) * u[0] - .3 * t[0]]),
              alg=_Vec([]), meas=None, t=t, x=x, z=_Vec.sym('z', 0), u=u, p=p, dt=dt,
              quad=_Vec([x[0] * x[0] + .1 * u[0] * u[0] + x[1] * sin(t[0]) + 5. * (u[0] - p[1]) * (u[0] - p[1])]))
    ns['RungeKutta']._collocation(pr, degree=deg, collocation_points='radau')
    ipsyms = _Vec([e for v in pr['collocation_points_ode'] for e in v])
    stage = _SymFunction('stage', [t, _Vec([dt]), ipsyms, x, u, _Vec([p[0]]), _Vec([p[1]])],
                         [pr['collocation_equations'], pr['ode'], pr['quad']])

    def int_dynamics_fun(time, dt_ii, ip_ii, x_ii, u_ii, zp_ii, p_ii, e_soft, tv_ref_ii, u_old0):
        return stage(time, _Vec([dt_ii]), ip_ii, x_ii, u_ii, p_ii, u_old0)
    tm, xm = _Vec.sym('tm'), _Vec.sym('xm', nx)
    mayer = _SymFunction('may', [tm, xm], [_Vec([2. * xm[0] * xm[0] + xm[1] * xm[1] + .1 * tm[0] * xm[0]])])
    # the loop, from the reference
    lines = open(cpath).read().split('\n')
    start = next(i for i, ln in enumerate(lines) if ln.strip() == 'for ii in range(self._prediction_horizon):'
                 and 'x_ii = x[ii, 0]' in lines[i + 1])
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == 'time += dt_ii')
    loop = textwrap.dedent('\n'.join(lines[start:end + 1]))
    rsrc, _ = _function_source(opath, '_rearrange_parameters', cls='DynamicOptimization')
    exec(rsrc, ns)

    class _Flag:
        is_set = False
        is_soft = False

    class _QS:
        _ind_input_changes = [0]

    class _Self:
        _prediction_horizon = _control_horizon = N
        quad_stage_cost = _QS()
        _nlp_options = {'integration_method': 'collocation', 'objective_function': 'continuous', 'ipopt_debugger': False}
        _lag_term_flag = _may_term_flag = True
        _lag_term_fun = None
        terminal_constraint = stage_constraint = _Flag()
        _n_tvp = 0
        _g_indices = {}

        def _may_term_fun(self, tt, xx, tv):
            return mayer(tt, xx)

        def _rearrange_parameters(self, tvp, cp):
            return ns['_rearrange_parameters'](self, tvp, cp)
    xs = np.empty((N + 1, 1), dtype=object)
    us = np.empty((N, 1), dtype=object)
    ips = np.empty((N, 1), dtype=object)
    zps = np.empty((N, 1), dtype=object)
    for k in range(N + 1):
        xs[k, 0] = _Vec.sym(f'X{k}', nx)
    for k in range(N):
        us[k, 0], ips[k, 0], zps[k, 0] = _Vec.sym(f'U{k}', 1), _Vec.sym(f'C{k}', deg * nx), _Vec([])

    class _Mdl:
        n_x = nx
    env = {'np': np, 'ca': ca, 'self': _Self(), 'x': xs, 'u': us, 'ip': ips, 'zp': zps, 'model': _Mdl(),
           'tv_p': np.empty((0, N), dtype=object), 'c_p': _Vec.sym('CP', 1), 'u_old': _Vec.sym('UOLD', 1),
           'tv_ref_sc': np.empty((0, N), dtype=object), 'tv_ref_tc': _Vec([]), '_dt': _Vec.sym('PDT', N),
           'time': _Vec.sym('PTIME'), 'g': [], 'g_lb': [], 'g_ub': [], 'J': 0, 'ind_g': 0,
           'int_dynamics_fun': int_dynamics_fun, 'e_soft_stage': _Vec([]), 'e_soft_term': _Vec([]),
           'gk_col_lb': [np.zeros(deg * nx)], 'gk_col_ub': [np.zeros(deg * nx)]}
    exec(loop, env)
    gvec = [e for blk in env['g'] for e in np.asarray(blk, dtype=object).reshape(-1)]
    names = [f'X{k}_{i}' for k in range(N + 1) for i in range(nx)] + [f'U{k}_0' for k in range(N)] + \
        [f'C{k}_{i}' for k in range(N) for i in range(deg * nx)] + ['CP_0', 'UOLD_0', 'PTIME_0'] + [f'PDT_{k}' for k in range(N)]
    rng = np.random.RandomState(33)
    pts = []
    for _ in range(3):
        ev = {n: float(rng.normal()) for n in names}
        ev['PTIME_0'] = float(rng.uniform(0., 2.))
        for k in range(N):
            ev[f'PDT_{k}'] = float(rng.uniform(.05, .4))
        ev0 = dict(ev)
        pts.append({'env': ev0, 'g': [float(_SX._f(e)(dict(ev0))) for e in gvec],
                    'J': float(_SX._f(np.asarray(env['J'], dtype=object).reshape(-1)[0])(dict(ev0)))})
    return {'source': f'hilo_mpc/modules/controller/mpc.py:{start + 1}-{end + 1}', 'N': N, 'degree': deg, 'points': pts,
            'g_lb_ub_all_zero': bool(all(not np.any(b) for b in env['g_lb']) and all(not np.any(b) for b in env['g_ub'])),
            'toy': "x0' = x1 + p sin(t) u, x1' = -x0^2 x1 + cos(x0) u - 0.3 t; l = x0^2 + 0.1 u^2 + x1 sin(t) + 5 (u - u_old)^2; "
                   "Mayer 2 x0^2 + x1^2 + 0.1 t x0"}


def _mhe_assembly_goldens():
    """The horizon loop of ``MHE.setup`` (hilo_mpc/modules/estimator/mhe.py, ``for ii in range(self._horizon)`` ...
    ``time += dt_ii``) executed from its source lines on the reaction model of the MHE tests
    (ca' = f (1 - ca) - k ca, cb' = k ca - f cb, y = cb), Radau-3, with the arrival and stage cost functions taken from
    the executed ``MHEQuadraticCost``: continuity with the additive state noise, arrival cost on interval 0 only, stage
    cost from interval 1 on, one parameter variable for the whole horizon."""
    mpath = os.path.join(REF, 'hilo_mpc', 'util', 'modeling.py')
    upath = os.path.join(REF, 'hilo_mpc', 'util', 'util.py')
    epath = os.path.join(REF, 'hilo_mpc', 'modules', 'estimator', 'mhe.py')
    ca = _stub_casadi_module()

    class _DM(np.ndarray):
        def __new__(cls, a=()):
            return np.asarray(a, dtype=float).view(cls)
    ca.DM = _DM
    ca.SX = _Vec
    ca.MX = _Vec
    ca.Function = _SymFunction
    ca.vertcat = lambda *a: _Vec([e for v in a for e in np.asarray(v, dtype=object).reshape(-1)])
    ca.mtimes = lambda a, b: np.dot(a, b)
    ca.collocation_points = _CaStub.collocation_points
    ca.sum1 = lambda a: np.sum(np.asarray(a, dtype=object).reshape(-1))

    def substitute(expr, old, new):
        names = [sym.name() for sym in np.asarray(old, dtype=object).reshape(-1)]
        vals = [_SX._f(v) for v in np.asarray(new, dtype=object).reshape(-1)]

        def rebound(e):
            return _rebind(e, names, vals)
        if isinstance(expr, np.ndarray):
            return _Vec([rebound(e) for e in expr.reshape(-1)])
        return rebound(expr)
    ca.substitute = substitute
    ns = {'np': np, 'ca': ca, 'sys': sys, 'Symbolic': object, 'warnings': __import__('warnings')}
    for name in ('check_and_wrap_to_list', 'check_if_has_duplicates', 'check_if_list_of_type', 'check_if_list_of_none',
                 'check_if_square', 'who_am_i'):
        exec(_function_source(upath, name)[0], ns)
    for name in ('GenericCost', 'MHEQuadraticCost'):
        exec(_class_source(mpath, name)[0], ns)
    body = []
    for name, deco in (('_construct_polynomial_basis', '@staticmethod'), ('_collocation', '@classmethod')):
        fsrc, _ = _function_source(mpath, name, cls='RungeKutta')
        body.append(textwrap.indent(deco + '\n' + fsrc.replace(deco + '\n', ''), '    '))
    exec('class RungeKutta:\n' + '\n'.join(body), ns)
    N, deg, nx = 4, 3, 2

    class _Names(_Vec):
        pass

    def named(names):
        v = _Vec([_SX.sym(n) for n in names]).view(_Names)
        v._names = list(names)
        return v

    class _Model:
        def __init__(self):
            self.x, self.p, self.u, self.y = named(['ca', 'cb']), named(['k']), named(['f']), named(['y'])
            self._x, self._p, self._u, self._y = self.x, self.p, self.u, self.y
            self.dynamical_state_names, self.parameter_names = self.x._names, self.p._names
            self.input_names, self.measurement_names = self.u._names, self.y._names
            self.n_x, self.n_p, self.n_u, self.n_y = 2, 1, 1, 1
            self.meas = _Vec([self.x[1]])
    mdl = _Model()
    arr = ns['MHEQuadraticCost'](mdl)
    arr.add_states(weights=[10., 10.], guess=[.5, .2])
    arr.add_parameters(weights=[1.], guess=[1.])
    arr._setup(x_scale=[1., 1.], w_scale=[1., 1.], p_scale=[1.], u_scale=[1.])
    stg = ns['MHEQuadraticCost'](mdl)
    stg.add_measurements(weights=[100.])
    stg.add_state_noise(weights=[20., 20.])
    stg._setup(x_scale=[1., 1.], w_scale=[1., 1.], p_scale=[1.], u_scale=[1.])

    def syms_of(lst):
        return _Vec([sym for r in lst for sym in np.asarray(r['placeholder'], dtype=object).reshape(-1)])
    xarr_ph = _Vec([sym for r in arr._iter_ref_list if r['type'] == 'states' for sym in r['placeholder']])
    parr_ph = _Vec([sym for r in arr._iter_ref_list if r['type'] == 'parameters' for sym in r['placeholder']])
    arrival_fun = _SymFunction('arr', [mdl.x, mdl.p, xarr_ph, parr_ph], [_Vec([arr.cost])])
    stage_fun = _SymFunction('stg', [stg._w, mdl.x, syms_of(stg._tv_ref_list)], [_Vec([stg.cost])])
    # per-interval collocation of the model (executed reference routine)
    t, dt = _Vec.sym('t'), _SX.sym('dt_0')
    xx, uu, pp = _Vec.sym('x', nx), _Vec.sym('u', 1), _Vec.sym('p', 1)
    pr = dict(ode=_Vec([uu[0] * (1. - xx[0]) - pp[0] * xx[0], pp[0] * xx[0] - uu[0] * xx[1]]), alg=_Vec([]),
              quad=_Vec([0. * xx[0]]), meas=None, t=t, x=xx, z=_Vec.sym('z', 0), u=uu, p=pp, dt=dt)
    ns['RungeKutta']._collocation(pr, degree=deg, collocation_points='radau')
    ipsyms = _Vec([e for v in pr['collocation_points_ode'] for e in v])
    stage = _SymFunction('stage', [t, _Vec([dt]), ipsyms, xx, uu, pp], [pr['collocation_equations'], pr['ode']])

    def int_dynamics_fun(time, dt_ii, ip_ii, x_ii, u_ii, zp_ii, p_, e_soft, y_ii):
        return stage(time, _Vec([dt_ii]), ip_ii, x_ii, u_ii, p_)
    lines = open(epath).read().split('\n')
    start = next(i for i, ln in enumerate(lines) if ln.strip() == 'for ii in range(self._horizon):'
                 and 'x_ii = x[ii, 0]' in lines[i + 1])
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == 'time += dt_ii')
    loop = textwrap.dedent('\n'.join(lines[start:end + 1]))

    class _Flag:
        is_set = False

    class _Self:
        _horizon = N
        _state_noise_flag = True
        _nlp_options = {'integration_method': 'collocation'}
        _arrival_term_flag = _stage_term_flag = True
        stage_constraint = _Flag()

        def _arrival_term_fun(self, x_ii, p_, xa, pa):
            return arrival_fun(x_ii, p_, xa, pa)

        def _stage_term_fun(self, w_ii, x_ii, y_ii):
            return stage_fun(w_ii, x_ii, y_ii)
    xs, ws = np.empty((N + 1, 1), dtype=object), np.empty((N, 1), dtype=object)
    ips, zps = np.empty((N, 1), dtype=object), np.empty((N, 1), dtype=object)
    for k in range(N + 1):
        xs[k, 0] = _Vec.sym(f'X{k}', nx)
    for k in range(N):
        ws[k, 0], ips[k, 0], zps[k, 0] = _Vec.sym(f'W{k}', nx), _Vec.sym(f'C{k}', deg * nx), _Vec([])
    um = np.empty((1, N), dtype=object)
    ym = np.empty((1, N), dtype=object)
    for k in range(N):
        um[0, k], ym[0, k] = _SX.sym(f'UM{k}'), _SX.sym(f'YM{k}')

    class _Mdl2:
        n_x = nx
    dtv, time_now = _Vec.sym('PDT', N), _Vec.sym('PTIME')
    env = {'np': np, 'ca': ca, 'self': _Self(), 'x': xs, 'w': ws, 'ip': ips, 'zp': zps, 'u_meas': um.view(_Vec),
           'y_meas': ym.view(_Vec), 'p': _Vec.sym('P', 1), 'model': _Mdl2(), '_dt': dtv, 'x_arrival': _Vec.sym('XA', nx),
           'p_arrival': _Vec.sym('PA', 1), 'time': time_now - ca.sum1(dtv), 'g': [], 'g_lb': [], 'g_ub': [], 'J': 0,
           'int_dynamics_fun': int_dynamics_fun, 'e_soft_stage': _Vec([]), 'gk_col_lb': [np.zeros(deg * nx)],
           'gk_col_ub': [np.zeros(deg * nx)]}
    exec(loop, env)
    gvec = [e for blk in env['g'] for e in np.asarray(blk, dtype=object).reshape(-1)]
    names = ['P_0'] + [f'X{k}_{i}' for k in range(N + 1) for i in range(nx)] + [f'W{k}_{i}' for k in range(N) for i in range(nx)] + \
        [f'C{k}_{i}' for k in range(N) for i in range(deg * nx)] + [f'UM{k}' for k in range(N)] + [f'YM{k}' for k in range(N)] + \
        ['XA_0', 'XA_1', 'PA_0', 'PTIME_0']
    rng = np.random.RandomState(44)
    pts = []
    for _ in range(3):
        ev = {n: float(rng.normal()) for n in names}
        for k in range(N):
            ev[f'PDT_{k}'] = .25          # the independent restatement has one dt
        pts.append({'env': dict(ev), 'g': [float(_SX._f(e)(dict(ev))) for e in gvec],
                    'J': float(_SX._f(np.asarray(env['J'], dtype=object).reshape(-1)[0])(dict(ev)))})
    return {'source': f'hilo_mpc/modules/estimator/mhe.py:{start + 1}-{end + 1}', 'N': N, 'degree': deg, 'dt': .25,
            'points': pts}


def _ukf_goldens():
    """``UnscentedKalmanFilter._setup_parameters / _setup_predict / _setup_update`` (hilo_mpc/modules/estimator/kf.py:486-604)
    executed from their source text.  The routines are straight-line matrix code on their "symbols", so the stub hands
    them NUMERIC matrices with CasADi's shape semantics (always 2-D, ``X[:, k]`` a column) under the symbols' names and
    records what they pass to ``ca.Function`` as outputs: the numbers one predict + update step produces.  ``ca.chol`` is
    the upper factor (R^T R = A), ``ca.solve`` a dense solve.  Model: the 3-state bioreactor of tests/test_KFs.py:695-712
    as an explicit-Euler discrete map, two measurements."""
    path = os.path.join(REF, 'hilo_mpc', 'modules', 'estimator', 'kf.py')
    import warnings

    class _Mx(np.matrix):
        def name(self):
            return 'P'

        def numel(self):
            return self.size

    values = {}

    class _Sym:
        @staticmethod
        def sym(name, *shape):
            return _Mx(np.array(values[name], dtype=float))

        def __new__(cls, v=0):
            return _Zero()                  # ca.MX(0): the start value of the running sums

    class _Zero:
        def __add__(self, o):
            return _Mx(np.array(o, dtype=float))
        __radd__ = __iadd__ = __add__
    ca = type('ca', (), {})()
    ca.SX = ca.MX = _Sym
    ca.DM = float
    ca.vertcat = lambda *a: _Mx(np.vstack([np.asarray(v).reshape(-1, 1) for v in a if np.size(v)] or [np.zeros((0, 1))]))
    ca.horzcat = lambda *a: _Mx(np.hstack([np.asarray(v) for v in a]))
    ca.chol = lambda a: _Mx(np.linalg.cholesky(np.asarray(a)).T)
    ca.solve = lambda A, B: _Mx(np.linalg.solve(np.asarray(A), np.asarray(B)))
    captured = {}

    class _Function:
        def __init__(self, name, ins, outs, *names):
            captured[name] = [np.asarray(o).copy() for o in outs]
    ca.Function = _Function
    import typing
    ns = {'np': np, 'ca': ca, 'warnings': warnings, 'Optional': typing.Optional, 'Union': typing.Union}
    body = []
    for name in ('_setup_parameters', '_setup_predict', '_setup_update'):
        fsrc, span = _function_source(path, name, cls='UnscentedKalmanFilter')
        body.append(textwrap.indent(fsrc, '    '))
    exec('class UKF:\n' + '\n'.join(body), ns)
    par = dict(alpha_=.1, T_amb=25., mu_0=.2, mu_1=.01, K=.5, Y=.6)
    dt = .1

    def f(xc, uc):
        T, cB, cS = xc
        r = (par['mu_0'] + par['mu_1'] * T) * cS * cB / (par['K'] + cS)
        return np.array([T + dt * par['alpha_'] * (par['T_amb'] - T), cB + dt * (r - uc[0] * cB), cS + dt * (-r / par['Y'] - uc[0] * cS)])

    class _Model:
        n_x, n_u, n_p, n_y = 3, 1, 0, 2
        dynamical_state_names = ['T', 'cB', 'cS']

        def __call__(self, x0=None, p=None, which=None):
            X = np.asarray(x0)
            uc = np.asarray(p).reshape(-1)
            if which == 'meas_function':
                return {'yf': _Mx(X[:2, :])}
            return {'xf': _Mx(np.stack([f(X[:, k], uc) for k in range(X.shape[1])], axis=1))}

    class _Sol:
        def setup(self, *a, **k):
            pass
    cases = []
    rng = np.random.RandomState(55)
    for alpha, beta, kappa in ((.5, 2., 0.), (1e-3, 2., 0.), (1., 0., 1.)):
        obj = ns['UKF']()
        obj._model, obj._solution = _Model(), _Sol()
        obj._alpha, obj._beta, obj._kappa = alpha, beta, kappa
        A = rng.normal(size=(3, 3))
        x = np.array([30., 1.2, 4.]) + rng.normal(size=3) * .1
        P = A @ A.T * .01 + np.eye(3) * .02
        Q, R = np.diag(rng.uniform(1e-4, 1e-3, 3)), np.diag(rng.uniform(1e-3, 1e-2, 2))
        u, y = np.array([.05 + .01 * rng.normal()]), np.array([30.2, 1.25]) + rng.normal(size=2) * .05
        values.update({'a': np.eye(3), 'x': x.reshape(-1, 1), 'u': u.reshape(-1, 1), 'p': np.zeros((0, 1)), 'P': P, 'Q': Q})
        obj._setup_parameters()
        obj._sqrt = lambda M: ca.chol(M)          # the Function('sqrt', [a], [ca.chol(a)]) built from the placeholder
        obj._setup_predict()
        pred = captured['prediction_step'][0]
        x_pred, P_pred, Xp = pred[:, :1], pred[:, 1:4], pred[:, 4:]
        values.update({'x': x_pred, 'P': P_pred, 'X': Xp, 'y': y.reshape(-1, 1), 'R': R})
        obj._setup_update()
        up, y_pred = captured['update_step']
        cases.append({'alpha': alpha, 'beta': beta, 'kappa': kappa, 'x': x.tolist(), 'P': P.tolist(), 'Q': Q.tolist(),
                      'R': R.tolist(), 'u': u.tolist(), 'y': y.tolist(), 'x_pred': x_pred.reshape(-1).tolist(),
                      'P_pred': P_pred.tolist(), 'X_prop': Xp.tolist(), 'x_up': up[:, 0].tolist(), 'P_up': up[:, 1:].tolist(),
                      'y_pred': np.asarray(y_pred).reshape(-1).tolist(), 'weights': obj._weights.tolist(),
                      'gamma': float(obj._gamma)})
    return {'source': 'hilo_mpc/modules/estimator/kf.py:486-604', 'dt': dt, 'parameters': par, 'cases': cases}


def _class_source(path, name):
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    return '\n'.join(src.split('\n')[node.lineno - 1:node.end_lineno]), (node.lineno, node.end_lineno)


def _network_goldens():
    """``Activation`` and ``net_to_casadi_graph`` (hilo_mpc/util/machine_learning.py:44-118, :521-578) executed from their
    source text with the stub ``ca``: forward pass of trained weights incl. z-score scaling of inputs and outputs."""
    path = os.path.join(REF, 'hilo_mpc', 'util', 'machine_learning.py')
    ca = _stub_casadi_module()

    class _Function:
        def __init__(self, name, ins, outs, in_names, out_names):
            self.ins, self.outs, self.out_names = ins, outs, out_names

        def __call__(self, xv):
            env = {s.name(): float(v) for s, v in zip(self.ins[0], xv)}
            return {n: [float(_SX._f(e)(env)) for e in np.asarray(o, dtype=object).reshape(-1)]
                    for n, o in zip(self.out_names, self.outs)}
    ca.Function = _Function
    src_a, span_a = _class_source(path, 'Activation')
    src_f, span_f = _function_source(path, 'net_to_casadi_graph')
    ns = {'ca': ca, 'np': np}
    exec(src_a, ns)
    exec(src_f, ns)

    class _Layer:
        def __init__(self, nodes, activation):
            self.type, self.nodes, self.activation = 'dense', nodes, activation

    class _Scaler:
        def __init__(self, mean, scale):
            self.mean_, self.scale_ = np.asarray(mean, dtype=float), np.asarray(scale, dtype=float)
    cases = []
    for seed, n_in, hidden, n_out, scaled in ((1, 3, [('tanh', 7)], 2, False), (2, 5, [('sigmoid', 6)], 4, True),
                                              (3, 2, [('tanh', 5), ('softplus', 4)], 1, True),
                                              (4, 4, [('linear', 3), ('tanh', 8)], 2, False)):
        rng = np.random.RandomState(seed)
        sizes = [n_in] + [h for _, h in hidden] + [n_out]
        weights = [rng.normal(size=(sizes[k + 1], sizes[k])) for k in range(len(sizes) - 1)]
        bias = [rng.normal(size=sizes[k + 1]) for k in range(len(sizes) - 1)]
        kw = {}
        if scaled:
            kw = dict(input_scaling=_Scaler(rng.normal(size=n_in), rng.uniform(.5, 2., n_in)),
                      output_scaling=_Scaler(rng.normal(size=n_out), rng.uniform(.5, 2., n_out)))
        x = np.array([_SX.sym(f'x{k}') for k in range(n_in)], dtype=object)
        fun = ns['net_to_casadi_graph']({'weights': weights, 'bias': bias}, x.copy(), [_Layer(h, a) for a, h in hidden], **kw)
        fun.ins = [list(x)]      # the routine scales `h` in place; the inputs of the function are the plain symbols
        pts = []
        for _ in range(3):
            xv = rng.normal(size=n_in)
            pts.append({'x': xv.tolist(), 'out': fun(xv)})
        cases.append({'n_in': n_in, 'hidden': [[a, h] for a, h in hidden], 'n_out': n_out,
                      'weights': [w.tolist() for w in weights], 'bias': [b.tolist() for b in bias],
                      'input_scaling': None if not scaled else [kw['input_scaling'].mean_.tolist(), kw['input_scaling'].scale_.tolist()],
                      'output_scaling': None if not scaled else [kw['output_scaling'].mean_.tolist(), kw['output_scaling'].scale_.tolist()],
                      'points': pts})
    return {'source': f'hilo_mpc/util/machine_learning.py:{span_a[0]}-{span_a[1]}, {span_f[0]}-{span_f[1]}', 'cases': cases}


def main():
    out = {'reference': 'hilo-mpc/hilo-mpc @ /root/reference', 'generator': 'tools/make_reference_goldens.py'}
    ca = _CaStub()

    # ---- collocation constants
    path = os.path.join(REF, 'hilo_mpc', 'util', 'modeling.py')
    src, span = _function_source(path, '_construct_polynomial_basis', cls='RungeKutta')
    ns = {'np': np, 'ca': ca, 'Symbolic': object}
    exec(src.replace('@staticmethod\n', ''), ns)
    basis = {}
    for method in ('radau', 'legendre'):
        for deg in (1, 2, 3, 4, 5):
            B, C, D, T = ns['_construct_polynomial_basis'](deg, method, _H(1.))
            basis[f'{method}{deg}'] = {'B': B.tolist(), 'C': C.tolist(), 'D': D.tolist(), 'T': np.asarray(T).tolist()}
    out['polynomial_basis'] = {'source': f'hilo_mpc/util/modeling.py:{span[0]}-{span[1]}', 'values': basis}

    # ---- lhsnorm
    path = os.path.join(REF, 'hilo_mpc', 'modules', 'estimator', 'pf.py')
    src, span = _function_source(path, 'lhsnorm')
    ns = {'np': np, 'norm': norm}
    exec(src, ns)
    cases = []
    for seed, mu, sigma, n in ((0, [0., 0.], [[1., 0.], [0., 1.]], 15), (7, [1., -2., .5], [[.5, .1, 0.], [.1, 2., .3], [0., .3, 1.]], 40)):
        np.random.seed(seed)
        x = ns['lhsnorm'](np.array(mu), np.array(sigma), n)
        cases.append({'seed': seed, 'mu': mu, 'sigma': sigma, 'n': n, 'x': x.tolist()})
    out['lhsnorm'] = {'source': f'hilo_mpc/modules/estimator/pf.py:{span[0]}-{span[1]}', 'cases': cases}

    # ---- resampling / roughening / moments block of ParticleFilter.estimate
    lines = open(path).read().split('\n')
    start = next(i for i, ln in enumerate(lines) if '# Resample (Survival of the fittest)' in ln)
    end = next(i for i, ln in enumerate(lines) if 'P = convert(np.cov(X), ca.DM)' in ln)
    block = textwrap.dedent('\n'.join(lines[start:end + 1]))

    class _Q:   # q.full().flatten()
        def __init__(self, a):
            self.a = np.asarray(a, dtype=float)

        def full(self):
            return self.a.reshape(1, -1)

    pf_cases = []
    for seed, n_x, n_s, rough in ((3, 2, 15, False), (11, 3, 64, True), (5, 4, 200, True)):
        rng = np.random.RandomState(1000 + seed)
        X = rng.standard_normal((n_x, n_s)) * np.arange(1, n_x + 1)[:, None] + np.arange(n_x)[:, None]
        Y = X[:1] ** 2
        q = rng.uniform(.01, 1., n_s)
        q /= q.sum()

        class _Self:
            _sample_size = n_s
            _roughening = rough
            _n_x = n_x
            _roughening_tuning_param = .2
            _transpose_pdf = True
            _pdf = staticmethod(ns['lhsnorm'])
        env = {'np': np, 'ca': ca, 'self': _Self(), 'X': X.copy(), 'Y': Y.copy(), 'q': _Q(q),
               'convert': lambda a, _t: np.asarray(a)}
        ca.DM = None
        np.random.seed(seed)
        exec(block, env)
        pf_cases.append({'seed': seed, 'n_x': n_x, 'n_s': n_s, 'roughening': rough, 'K': .2, 'X_in': X.tolist(),
                         'q': q.tolist(), 'ind': np.asarray(env['ind']).tolist(), 'X_out': np.asarray(env['X']).tolist(),
                         'x_mean': np.asarray(env['x']).tolist(), 'P': np.asarray(env['P']).tolist()})
    out['pf_resample_block'] = {'source': f'hilo_mpc/modules/estimator/pf.py:{start + 1}-{end + 1}', 'cases': pf_cases}

    # ---- horizon of the time-varying parameters, iteration by iteration (NMPC._get_tvp_parameters_values)
    path = os.path.join(REF, 'hilo_mpc', 'modules', 'controller', 'mpc.py')
    src, span = _function_source(path, '_get_tvp_parameters_values', cls='NMPC')

    class _DM:
        @staticmethod
        def zeros(n, m):
            return np.zeros((n, m))
    ca.DM = _DM
    import warnings
    ns = {'np': np, 'ca': ca, 'warnings': warnings}
    exec(src, ns)
    tvp_cases = []
    for N, lengths, iters in ((4, (4, 7), 12), (5, (5,), 11), (3, (10, 3), 9), (6, (8,), 20)):
        values = {f'p{k}': [float(10 * (k + 1) + i) for i in range(L)] for k, L in enumerate(lengths)}

        class _Self:
            _n_iterations = 0
            _time_varying_parameters = list(values)
            _time_varying_parameters_values = values
            _time_varying_parameters_horizon = None
            prediction_horizon = N
            _prediction_horizon = N
            _n_tvp = len(values)
        obj = _Self()
        horizons = []
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for ci in range(iters):
                obj._n_iterations = ci
                ns['_get_tvp_parameters_values'](obj, values)
                horizons.append(np.asarray(obj._time_varying_parameters_horizon).tolist())
        tvp_cases.append({'N': N, 'values': values, 'horizons': horizons})
    out['tvp_horizon'] = {'source': f'hilo_mpc/modules/controller/mpc.py:{span[0]}-{span[1]}', 'cases': tvp_cases}

    # ---- equation-string parser (hilo_mpc/util/parsing.py), executed as a module with a stub `casadi`
    out['parser'] = _parser_goldens()

    # ---- export of a trained feed-forward network to an expression graph (util/machine_learning.py)
    out['net_to_graph'] = _network_goldens()

    # ---- one interval of the collocation / explicit Runge-Kutta transcription (util/modeling.py)
    out['discretisation'] = _discretisation_goldens()

    # ---- quadratic stage cost incl. scaled references and the input-change term (util/modeling.py)
    out['quadratic_cost'] = _quadratic_cost_goldens()

    # ---- the QP matrices LMPC.setup builds (mpc.py)
    out['lmpc_matrices'] = _lmpc_goldens()

    # ---- the horizon loop of NMPC._setup (mpc.py)
    out['nmpc_assembly'] = _nmpc_assembly_goldens()

    # ---- the horizon loop of MHE.setup (mhe.py)
    out['mhe_assembly'] = _mhe_assembly_goldens()

    # ---- one predict + update step of the unscented Kalman filter (kf.py)
    out['ukf_step'] = _ukf_goldens()

    with open(OUT, 'w') as f:
        json.dump(out, f)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    sys.exit(main())
