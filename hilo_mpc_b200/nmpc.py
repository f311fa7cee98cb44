"""``NMPC``: nonlinear model predictive controller with the reference's API, solved in batches on the GPU.

API mirror of ``hilo_mpc/modules/controller/mpc.py:42-1931`` and the ``DynamicOptimization`` option surface
(``hilo_mpc/modules/optimizer.py:847-1763``).  ``setup()`` performs the symbolic part of the reference's
``_setup`` (costs, scaling, path variables, discretisation choice; mpc.py:1133-1437) and hands an ``OcpSpec`` to the CUDA
generator; ``optimize()`` packs the per-call numbers (mpc.py:744-857) and runs one batched solve.  ``x0`` may be a
single state (``(nx,)`` -> returns ``(nu, 1)`` like the reference) or a batch ``(B, nx)`` (-> ``(B, nu)``).
"""
from __future__ import annotations

import time as _time
import warnings
from typing import Dict, List, Optional

import numpy as np

from . import symbolic as sx
from .costs import GenericCost, QuadraticCost
from .ocp import OcpSpec
from .series import TimeSeries

_STATUS_TEXT = {1: 'solve_succeeded', 2: 'solved_to_acceptable_level', 3: 'infeasible_problem_detected',
                4: 'restoration_failed', 5: 'maximum_iterations_exceeded', -1: 'error'}


def _vec(v, n, name, default):
    if v is None:
        return np.full(n, default, dtype=float)
    v = np.asarray(v, dtype=float).reshape(-1)
    if v.size != n:
        raise ValueError(f"{name} has {v.size} elements but {n} are required")
    return v.copy()


class DynamicOptimization:
    """Option / bound / scaling surface shared by NMPC and MHE (optimizer.py:847-1763)."""
    _type = 'DynamicOptimization'

    def __init__(self, model, id=None, name=None, plot_backend=None, stats: bool = False, use_sx: bool = True):
        if not model.is_setup():
            raise RuntimeError(f"Model is not set up. Run Model.setup() before passing it to the {self._type}.")
        self._model_orig = model
        self._model = model.copy(setup=True)
        self.id = id
        self.name = name or self._type
        self._stats = stats
        self._prediction_horizon = None
        self._control_horizon = None
        self._sampling_interval = model.dt_value
        self._x_ub = self._x_lb = self._u_ub = self._u_lb = None
        self._x_guess = self._u_guess = None
        self._x_scaling = self._u_scaling = self._y_scaling = None
        self._tvp_names: List[str] = []
        self._tvp_values: Dict[str, np.ndarray] = {}
        self._nlp_options = None
        self._nlp_opts: Dict[str, object] = {}
        self._solver_name = 'ipopt'
        self._nlp_setup_done = False
        self._n_iterations = 0
        self._time = 0.
        self._solver_status_code = None
        self._solver_status = None
        self._engine = None
        self._spec: Optional[OcpSpec] = None
        self._solution = TimeSeries()

    # ------------------------------------------------------------------ horizons
    @property
    def horizon(self):
        return self._prediction_horizon

    @horizon.setter
    def horizon(self, arg):
        if not isinstance(arg, (int, np.integer)) or arg <= 0:
            raise TypeError("The horizon must be a positive integer")
        self._prediction_horizon = int(arg)
        self._control_horizon = int(arg)

    @property
    def prediction_horizon(self):
        return self._prediction_horizon

    @prediction_horizon.setter
    def prediction_horizon(self, arg):
        if not isinstance(arg, (int, np.integer)) or arg <= 0:
            raise TypeError("The prediction horizon must be a positive integer")
        self._prediction_horizon = int(arg)

    @property
    def control_horizon(self):
        return self._control_horizon

    @control_horizon.setter
    def control_horizon(self, arg):
        if not isinstance(arg, (int, np.integer)) or arg <= 0:
            raise TypeError("The control horizon must be a positive integer")
        self._control_horizon = int(arg)

    @property
    def sampling_interval(self):
        return self._sampling_interval

    def set_sampling_interval(self, dt=None):
        self._sampling_interval = float(dt) if dt is not None else self._model.dt_value

    # ------------------------------------------------------------------ bounds, guesses, scaling
    def set_box_constraints(self, x_ub=None, x_lb=None, u_ub=None, u_lb=None, y_ub=None, y_lb=None, z_ub=None,
                            z_lb=None):
        m = self._model
        self._x_ub = _vec(x_ub, m.n_x, 'x_ub', np.inf)
        self._x_lb = _vec(x_lb, m.n_x, 'x_lb', -np.inf)
        self._u_ub = _vec(u_ub, m.n_u, 'u_ub', np.inf)
        self._u_lb = _vec(u_lb, m.n_u, 'u_lb', -np.inf)
        for lo, hi, nm in ((self._x_lb, self._x_ub, 'x'), (self._u_lb, self._u_ub, 'u')):
            flip = lo > hi  # optimizer.py:123-136 flips inverted bounds
            if flip.any():
                warnings.warn(f"Some lower bounds of {nm} are larger than the upper bounds. I am flipping them.")
                lo[flip], hi[flip] = hi[flip], lo[flip].copy()
        if y_ub is not None or y_lb is not None:
            raise NotImplementedError("Measurement box constraints are not supported by the B200 engine yet")

    def set_initial_guess(self, x_guess=None, u_guess=None, z_guess=None):
        m = self._model
        self._x_guess = _vec(x_guess, m.n_x, 'x_guess', 0.)
        self._u_guess = _vec(u_guess, m.n_u, 'u_guess', 0.)

    def set_scaling(self, x_scaling=None, u_scaling=None, y_scaling=None):
        m = self._model
        self._x_scaling = _vec(x_scaling, m.n_x, 'x_scaling', 1.)
        self._u_scaling = _vec(u_scaling, m.n_u, 'u_scaling', 1.)
        self._y_scaling = _vec(y_scaling, m.n_y, 'y_scaling', 1.)

    def set_time_varying_parameters(self, names=None, values=None):
        names = [] if names is None else ([names] if isinstance(names, str) else list(names))
        for n in names:
            if n not in self._model.parameter_names:
                raise ValueError(f"The time-varying parameter {n} is not a parameter of the model. The model "
                                 f"parameters are {self._model.parameter_names}")
        self._tvp_names = names
        self._tvp_values = {}
        if values is not None:
            for n in names:
                if n not in values:
                    raise ValueError(f"No values supplied for the time-varying parameter {n}")
                self._tvp_values[n] = np.asarray(values[n], dtype=float).reshape(-1)

    # ------------------------------------------------------------------ options
    def set_nlp_options(self, *args, **kwargs):
        possible = {'integration_method': ['collocation', 'rk4', 'erk', 'discrete', 'idas', 'cvodes'],
                    'solver': ['ipopt'], 'collocation_points': ['radau', 'legendre'],
                    'objective_function': ['discrete', 'continuous'], 'warm_start': [True, False], 'degree': None,
                    'order': None, 'print_level': [0, 1], 'ipopt_debugger': [True, False]}
        opts = {'integration_method': 'collocation', 'collocation_points': 'radau', 'degree': 3, 'print_level': 1,
                'warm_start': True, 'solver': 'ipopt', 'ipopt_debugger': False, 'order': None,
                'objective_function': 'discrete' if self._model.discrete else 'continuous'}
        given = args[0] if args and isinstance(args[0], dict) else kwargs
        for key, value in (given or {}).items():
            if key not in opts:
                raise ValueError(f"The option named {key} does not exist. Possible options are {list(opts.keys())}.")
            if possible[key] is not None and value not in possible[key]:
                raise ValueError(f"The option {key} is set to value {value} but the only allowed values are "
                                 f"{possible[key]}.")
            opts[key] = value
        if opts['integration_method'] != 'discrete' and self._model.discrete:
            if given and 'integration_method' in given:
                warnings.warn(f"The integration method is set to {opts['integration_method']} but I notice that the "
                              f"model is in discrete time. I am overwriting and using discrete mode.")
            opts['integration_method'] = 'discrete'
        if opts['integration_method'] in ('idas', 'cvodes'):
            raise NotImplementedError("SUNDIALS-in-the-loop shooting ('idas'/'cvodes') is not available in the B200 "
                                      "engine; use 'collocation', 'rk4', 'erk' or 'discrete'.")
        if opts['integration_method'] == 'rk4':
            opts['order'] = 4
        elif opts['integration_method'] == 'erk' and opts['order'] is None:
            opts['order'] = 1
        self._nlp_options = opts

    def set_solver_opts(self, *args, **kwargs):
        """IPOPT-style option names are accepted (``ipopt.tol``, ``ipopt.max_iter``, ``ipopt.mu_init``,
        ``ipopt.acceptable_tol``); engine-specific keys: ``b200.line_search_candidates``, ``b200.check_every``."""
        given = args[0] if args and isinstance(args[0], dict) else kwargs
        self._nlp_opts = dict(given or {})

    def set_nlp_solver(self, solver):
        if solver != 'ipopt':
            raise ValueError(f"The solver {solver} does no exist. The possible solver are ['ipopt'].")
        self._solver_name = solver

    def _solver_numbers(self):
        o = self._nlp_opts
        tol = float(o.get('ipopt.tol', o.get('tol', 1e-8)))
        return dict(tol=tol, acc_tol=float(o.get('ipopt.acceptable_tol', 1e-6)),
                    max_iter=int(o.get('ipopt.max_iter', o.get('max_iter', 3000))),
                    mu_init=float(o.get('ipopt.mu_init', .1)),
                    n_alpha=int(o.get('b200.line_search_candidates', 8)),
                    check_every=int(o.get('b200.check_every', 4)),
                    riccati_variant=int(o.get('b200.riccati_variant', 2)),
                    warm_start=int(o.get('b200.warm_start', 0)),
                    mu_init_warm=float(o.get('b200.warm_start_mu', 1e-4)),
                    eval_variant=int(o.get('b200.eval_variant', 1)),
                    step_variant=int(o.get('b200.step_variant', 2)),
                    compaction=bool(int(o.get('b200.compaction', 1))),
                    adaptive_poll=bool(o.get('b200.adaptive_poll', True)),
                    regularization={'inertia': 0, 'mirror': 1}[str(o.get('b200.regularization', 'inertia'))],
                    regularization_eps=float(o.get('b200.regularization_eps', 1e-4)),
                    regularization_mu=float(o.get('b200.regularization_mu', 0.)),
                    shift_warm_start=int(o.get('b200.shift_warm_start', 0)))

    def is_setup(self) -> bool:
        return self._nlp_setup_done

    @property
    def _time_varying_parameters(self):
        """Name the reference's control loop reads (control_loop.py:357-362)."""
        return self._tvp_names

    @property
    def solver_status_code(self):
        return self._solver_status_code

    def _exact_hessian(self) -> bool:
        h = str(self._nlp_opts.get('b200.hessian', 'exact')).lower()
        if h not in ('exact', 'gauss-newton'):
            raise ValueError("b200.hessian must be 'exact' or 'gauss-newton'")
        return h == 'exact'

    @property
    def solution(self):
        return self._solution

    @property
    def type(self):
        return self._type


class NMPC(DynamicOptimization):
    _type = 'NMPC'

    def __init__(self, model, id=None, name=None, plot_backend=None, use_sx: bool = True, stats: bool = False):
        super().__init__(model, id=id, name=name, plot_backend=plot_backend, stats=stats, use_sx=use_sx)
        m = self._model
        self.quad_stage_cost = QuadraticCost(m)
        self.quad_terminal_cost = QuadraticCost(m)
        self.stage_cost = GenericCost(m)
        self.terminal_cost = GenericCost(m)
        self._paths_var_list: List[dict] = []
        self._time_var = None
        self._uold = None
        self._last = None
        self._batch = None

    # ------------------------------------------------------------------ extra builders
    @property
    def n_of_path_vars(self):
        return len(self._paths_var_list)

    def create_path_variable(self, name='theta', u_pf_lb=0.0001, u_pf_ub=1., u_pf_ref=None, u_pf_weight=10.,
                             theta_guess=0., theta_lb=0., theta_ub=np.inf):
        """Virtual path parameter with dynamics ``theta' = u_theta`` (mpc.py:1025-1053, 1173-1204)."""
        theta = sx.sym(name)
        self._paths_var_list.append(dict(theta=theta, name=name, u_pf_lb=u_pf_lb, u_pf_ub=u_pf_ub, u_pf_ref=u_pf_ref,
                                         u_pf_weight=u_pf_weight, theta_guess=theta_guess, theta_lb=theta_lb,
                                         theta_ub=theta_ub))
        return theta

    def get_time_variable(self):
        if self._time_var is None:
            self._time_var = self._model.t
        return self._time_var

    def set_quadratic_stage_cost(self, states=None, cost_states=None, states_references=None, inputs=None,
                                 cost_inputs=None, inputs_references=None):
        if states is not None:
            self.quad_stage_cost.add_states(states, cost_states, states_references)
        if inputs is not None:
            self.quad_stage_cost.add_inputs(inputs, cost_inputs, inputs_references)

    def set_quadratic_terminal_cost(self, states=None, cost=None, references=None):
        if states is not None:
            self.quad_terminal_cost.add_states(states, cost, references)

    def set_stage_constraints(self, *args, **kwargs):
        raise NotImplementedError("General nonlinear stage constraints are not supported by the B200 engine yet "
                                  "(SURVEY.md §8f-4); box constraints are.")

    set_terminal_constraints = set_stage_constraints
    set_custom_constraints_function = set_stage_constraints

    def minimize_final_time(self, weight=1.):
        raise NotImplementedError("Minimum-time problems are not supported by the B200 engine yet")

    # ------------------------------------------------------------------ setup
    def _quad_to_expr(self, qc: QuadraticCost, xs, us, x_scale, u_scale, y_scale, meas_scaled, ref_slots, kind):
        """Return (expr, ls_terms, du_expr, du_syms) for one QuadraticCost in scaled coordinates."""
        expr = sx.Expr.const(0.)
        ls = []
        du_expr = sx.Expr.const(0.)
        du_syms = []
        for term in qc.terms:
            ind, W = term['ind'], term['W']
            if term['kind'] == 'states':
                var, scale = xs[ind], x_scale[ind]
            elif term['kind'] in ('inputs', 'inputs_change'):
                var, scale = us[ind], u_scale[ind]
            else:
                var, scale = meas_scaled[ind], y_scale[ind]
            if term['kind'] == 'inputs_change':
                uo = sx.named_syms([f"{n}_old" for n in term['names']])
                r = var - uo
                du_expr = du_expr + sx.quad_form(r, W)
                du_syms.append((uo, ind))
                continue
            ref = term['ref']
            if ref is None and not term['trajectory_tracking']:
                r = var  # stabilisation (modeling.py:255)
            elif term['path_following']:
                r = var - ref  # path expression of theta is substituted as is (modeling.py:258-266)
            elif term['trajectory_tracking']:
                if ref is not None:
                    r = var - ref / scale  # function of time supplied symbolically
                else:
                    slots = sx.named_syms([f"{n}_ref_{kind}" for n in term['names']])
                    ref_slots.append((slots, term['names'], scale))
                    r = var - slots / scale
            else:
                r = var - np.asarray(ref, dtype=float) / scale  # modeling.py:310
            r = sx.as_vector(r)
            expr = expr + sx.quad_form(r, W)
            ls.append((r, W))
        return expr, ls, du_expr, du_syms

    def setup(self, options=None, solver_options=None) -> None:
        m = self._model
        if self._prediction_horizon is None:
            raise ValueError("The prediction horizon must be set before setting up the NMPC (nmpc.horizon = N).")
        if self._control_horizon != self._prediction_horizon:
            raise NotImplementedError("control_horizon != prediction_horizon is not supported by the B200 engine yet")
        if options is not None:
            self.set_nlp_options(options)
        elif self._nlp_options is None:
            self.set_nlp_options()
        if solver_options is not None:
            self.set_solver_opts(solver_options)
        if self._x_scaling is None:
            self.set_scaling()
        if self._x_lb is None:
            self.set_box_constraints()
        if self._x_guess is None:
            self.set_initial_guess()
        if self._sampling_interval is None:
            self.set_sampling_interval()
        opts = self._nlp_options
        nx0, nu0 = m.n_x, m.n_u
        x_scale = list(self._x_scaling)
        u_scale = list(self._u_scaling)
        x_lb, x_ub, x_guess = list(self._x_lb), list(self._x_ub), list(self._x_guess)
        u_lb, u_ub, u_guess = list(self._u_lb), list(self._u_ub), list(self._u_guess)
        x_syms = list(m.x)
        u_syms = list(m.u)
        ode = list(m.ode)
        extra_stage = sx.Expr.const(0.)
        # ---- path variables (mpc.py:1173-1204)
        for pv in self._paths_var_list:
            u_theta = sx.sym('u_' + pv['name'])
            x_syms.append(pv['theta'])
            u_syms.append(u_theta)
            ode.append(pv['theta'] + self._sampling_interval * u_theta if m.discrete else u_theta)
            x_guess.append(pv['theta_guess'])
            u_guess.append(pv['u_pf_lb'] + 0.0001)
            u_lb.append(pv['u_pf_lb'])
            u_ub.append(pv['u_pf_ub'])
            x_lb.append(pv['theta_lb'])
            x_ub.append(pv['theta_ub'])
            u_scale.append(1.)
            x_scale.append(1.)
            if pv['u_pf_ref'] is not None:
                extra_stage = extra_stage + (u_theta - pv['u_pf_ref']) ** 2 * pv['u_pf_weight']
        x_syms = np.array(x_syms, dtype=object)
        u_syms = np.array(u_syms, dtype=object)
        x_scale, u_scale = np.asarray(x_scale, dtype=float), np.asarray(u_scale, dtype=float)
        nx, nu = len(x_syms), len(u_syms)
        # ---- parameters: time-varying ones first, then constant ones (mpc.py:1574-1580)
        p_names = m.parameter_names
        tvp_idx = [p_names.index(n) for n in self._tvp_names]
        cp_idx = [k for k in range(len(p_names)) if k not in tvp_idx]
        self._tvp_idx, self._cp_idx = tvp_idx, cp_idx
        # ---- scaled symbols
        xs = sx.named_syms([s.name for s in x_syms])
        us = sx.named_syms([s.name for s in u_syms])
        p_slots = sx.named_syms([f'p_{p_names[k]}' for k in tvp_idx + cp_idx])
        t, dt = sx.sym('t'), sx.sym('dt')
        old = np.concatenate([x_syms, u_syms, m.p[tvp_idx + cp_idx] if len(p_names) else [], [m.t], [m.dt]])
        new = np.concatenate([xs * x_scale, us * u_scale, p_slots, [t], [dt]])
        ode_s = sx.substitute(np.array(ode, dtype=object), old, new) / x_scale
        meas_s = sx.substitute(m.meas, old, new) / self._y_scaling if m.n_y else np.zeros(0, dtype=object)
        # ---- costs: expressions use the model symbols as *scaled* variables (mpc.py:207-263 defines the cost before
        #      the model is scaled, so the same symbols then denote x/x_scaling)
        raw_old = np.concatenate([x_syms, u_syms, m.p[tvp_idx + cp_idx] if len(p_names) else [], [m.t], [m.dt]])
        raw_new = np.concatenate([xs, us, p_slots, [t], [dt]])
        ref_slots_sc, ref_slots_tc = [], []
        q_expr, q_ls, du_expr, du_syms = self._quad_to_expr(self.quad_stage_cost, xs, us, x_scale, u_scale,
                                                            self._y_scaling, meas_s, ref_slots_sc, 'sr')
        t_expr, t_ls, _, _ = self._quad_to_expr(self.quad_terminal_cost, xs, us, x_scale, u_scale, self._y_scaling,
                                                meas_s, ref_slots_tc, 'tr')

        def to_scaled(e):
            return sx.substitute(e, raw_old, raw_new)
        q_ls = [(to_scaled(r), W) for r, W in q_ls]
        t_ls = [(to_scaled(r), W) for r, W in t_ls]
        lag_rest = sx.Expr.const(0.)
        if self.stage_cost.is_set:
            lag_rest = lag_rest + to_scaled(self.stage_cost.cost)
        extra_stage = sx.substitute(extra_stage, raw_old, raw_new) if self._paths_var_list else extra_stage
        for pv in self._paths_var_list:
            pass
        # path-variable velocity terms are least-squares too
        for pv, uth in zip(self._paths_var_list, us[nu0:]):
            if pv['u_pf_ref'] is not None:
                q_ls.append((np.array([uth - pv['u_pf_ref']], dtype=object), np.array([[pv['u_pf_weight']]])))
        lag = to_scaled(q_expr) + lag_rest + extra_stage
        may_rest = to_scaled(self.terminal_cost.cost) if self.terminal_cost.is_set else sx.Expr.const(0.)
        may = to_scaled(t_expr) + may_rest
        du_expr = to_scaled(du_expr)
        uold_syms = np.concatenate([s for s, _ in du_syms]) if du_syms else np.zeros(0, dtype=object)
        self._du_ind = [i for _, ind in du_syms for i in ind]
        # ---- stage / terminal parameter vectors
        sr = np.concatenate([s for s, _, _ in ref_slots_sc]) if ref_slots_sc else np.zeros(0, dtype=object)
        tr = np.concatenate([s for s, _, _ in ref_slots_tc]) if ref_slots_tc else np.zeros(0, dtype=object)
        self._ref_slots_sc, self._ref_slots_tc = ref_slots_sc, ref_slots_tc
        sp = np.concatenate([p_slots, sr])
        method = {'collocation': 'collocation', 'rk4': 'erk', 'erk': 'erk', 'discrete': 'discrete'}[
            opts['integration_method']]
        name = f"nmpc_{m.name}_{method}_N{self._prediction_horizon}"
        self._spec = OcpSpec(
            nx=nx, nu=nu, N=self._prediction_horizon, t=t, dt=dt, x=xs, u=us, sp=sp, tp=tr, uold=uold_syms,
            ode=sx.as_vector(ode_s), lag=lag, lag_du=du_expr, may=may, lag_ls=q_ls, lag_rest=lag_rest,
            may_ls=t_ls, may_rest=may_rest, method=method, degree=int(opts['degree']),
            points=opts['collocation_points'], order=int(opts['order'] or 1),
            objective=opts['objective_function'], x_lb=np.asarray(x_lb, dtype=float) / x_scale,
            x_ub=np.asarray(x_ub, dtype=float) / x_scale, u_lb=np.asarray(u_lb, dtype=float) / u_scale,
            u_ub=np.asarray(u_ub, dtype=float) / u_scale, x_guess=np.asarray(x_guess, dtype=float) / x_scale,
            u_guess=np.asarray(u_guess, dtype=float) / u_scale, exact_hessian=self._exact_hessian(), name=name)
        self._xs_full, self._us_full = x_scale, u_scale
        self._scaling_cache = None
        self._nx0, self._nu0 = nx0, nu0
        from .engine import OcpEngine
        # `_build_module = False` (set by CPU-only tooling) skips nvcc; such an object cannot optimize()
        self._engine = OcpEngine(self._spec, build=getattr(self, '_build_module', True))
        names = dict(t=['t'], x=[s.name for s in x_syms], u=[s.name for s in u_syms],
                     extime=['extime'], niterations=['niterations'], solvstatus=['solvstatus'])
        self._solution = TimeSeries(**names)
        self._n_iterations = 0
        self._uold = None
        self._nlp_setup_done = True

    @property
    def spec(self) -> OcpSpec:
        return self._spec

    @property
    def engine(self):
        return self._engine

    # ------------------------------------------------------------------ per-call parameter packing (mpc.py:474-515)
    def _stage_parameters(self, cp, tvp, ref_sc, torch, device, B):
        s = self._spec
        N = s.N
        n_tvp, n_cp = len(self._tvp_idx), len(self._cp_idx)
        if s.nsp == 0:
            return None
        cols = []
        if n_tvp:
            block = np.zeros((N, n_tvp))
            for j, name in enumerate(self._tvp_names):
                passed = tvp is not None and name in tvp
                vals = tvp[name] if passed else self._tvp_values.get(name)
                if vals is None:
                    raise ValueError(f"No values for the time-varying parameter {name} were supplied")
                vals = np.asarray(vals, dtype=float).reshape(-1)
                if len(vals) < N:
                    raise TypeError(f"When passing time-varying parameters, you need to pass a number of values at "
                                    f"least as long as the prediction horizon. The parameter {name} has {len(vals)} "
                                    f"values but the MPC has a prediction horizon length of {N}.")
                if passed:   # values handed to optimize() ARE the horizon of this step (mpc.py:342-355)
                    block[:, j] = vals[:N]
                    continue
                # Stored values (mpc.py:293-333, optimizer.py:1106-1143): the horizon moves on by one value per
                # iteration.  Past the end of the data the reference "loops back" with the index ci - N floor(ci / N) of
                # the iteration ci that appends the value - i.e. position (ci + j + 1) mod N of the data, not a wrap
                # modulo its length; reproduced as it is (tests/golden/reference_numeric.json, 'tvp_horizon').
                pos = self._n_iterations + np.arange(N)
                beyond = pos >= len(vals)
                if beyond.any() and self._n_iterations > 0:
                    warnings.warn("The prediction horizon is predicting outside the values of the time varying "
                                  "parameters. I am now taking looping back the values and start from there. "
                                  "See documentation for info.")
                block[:, j] = vals[np.where(beyond, (pos + 1) % N, np.minimum(pos, len(vals) - 1))]
            cols.append(torch.tensor(block, dtype=torch.float64, device=device).unsqueeze(0))
        if n_cp:
            if cp is None:
                raise ValueError(f"The model has {n_cp} constant parameter(s): pass them to optimize with cp=...")
            cpt = torch.as_tensor(np.asarray(cp, dtype=float) if not torch.is_tensor(cp) else cp,
                                  dtype=torch.float64, device=device)
            if cpt.numel() not in (n_cp, n_cp * B):
                raise ValueError(f"cp has {cpt.numel()} elements, expected {n_cp} (shared) or {n_cp * B} (one row "
                                 f"per instance)")
            cpt = cpt.reshape(-1, n_cp)
            cols.append(cpt.unsqueeze(1).expand(cpt.shape[0], N, n_cp))
        for slots, names, scale in self._ref_slots_sc:
            vals = []
            for n in names:
                if ref_sc is None or n not in ref_sc:
                    raise ValueError(f"No stage reference trajectory supplied for '{n}' (pass ref_sc={{'{n}': ...}})")
                v = np.asarray(ref_sc[n], dtype=float).reshape(-1)
                vals.append(v[:N] if v.size >= N else np.resize(v, N))
            cols.append(torch.tensor(np.stack(vals, axis=1), dtype=torch.float64, device=device).unsqueeze(0))
        nb = max(c.shape[0] for c in cols)
        cols = [c.expand(nb, N, c.shape[2]) for c in cols]
        return torch.cat(cols, dim=2).contiguous()

    def _scaling_tensors(self, torch, dev):
        """Scaling vectors on the device, created once (two pageable host-to-device copies per call otherwise)."""
        cache = getattr(self, '_scaling_cache', None)
        if cache is None or cache[0] != dev:
            cache = (dev, torch.tensor(self._xs_full, dtype=torch.float64, device=dev),
                     torch.tensor(self._us_full, dtype=torch.float64, device=dev))
            self._scaling_cache = cache
        return cache[1], cache[2]

    def optimize(self, x0, cp=None, tvp=None, v0=None, runs=0, fix_x0=True, **kwargs):
        if not self._nlp_setup_done:
            raise RuntimeError("The NMPC is not set up. Run NMPC.setup() before optimizing.")
        if not isinstance(runs, (int, np.integer)) or runs < 0:
            raise ValueError("The number of multi-start runs must be a non-negative integer")
        if not fix_x0 and ('x0_lb' in kwargs or 'x0_ub' in kwargs):
            raise NotImplementedError("fix_x0=False with bounds on the initial state other than the state bounds "
                                      "(x0_lb / x0_ub) is not supported by the B200 engine yet")
        from . import _lib
        torch = _lib.require_cuda()
        eng = self._engine if fix_x0 else self._free_x0_engine()
        eng._open()
        dev = eng._device
        s = self._spec
        t_start = _time.time()
        is_tensor = torch.is_tensor(x0)
        x0t = x0.to(device=dev, dtype=torch.float64) if is_tensor else \
            torch.as_tensor(np.asarray(x0, dtype=float), dtype=torch.float64, device=dev)
        if x0t.dim() <= 1 or (x0t.dim() == 2 and x0t.shape == (self._nx0, 1) and self._nx0 > 1):
            if x0t.numel() != self._nx0:
                raise ValueError(f"x0 must have {self._nx0} rows (single instance) or shape (B, {self._nx0}); got "
                                 f"{tuple(x0t.shape)}")
            single = True
            x0t = x0t.reshape(1, self._nx0)
        elif x0t.dim() == 2 and x0t.shape[1] == self._nx0:
            single = False
        else:
            raise ValueError(f"x0 must have {self._nx0} rows (single instance) or shape (B, {self._nx0}); got "
                             f"{tuple(x0t.shape)}")
        B = x0t.shape[0]
        xs_t, us_t = self._scaling_tensors(torch, dev)
        # augmented state: path variables start from their previous predicted value (or the guess)
        if s.nx > self._nx0:
            if self._last is not None and self._last['x'].shape[0] == B:
                extra = self._last['x'][:, 1, self._nx0:] * xs_t[self._nx0:]
            else:
                extra = torch.tensor([pv['theta_guess'] for pv in self._paths_var_list], dtype=torch.float64,
                                     device=dev).expand(B, -1)
            x0t = torch.cat([x0t, extra], dim=1)
        x0s = x0t / xs_t  # mpc.py:801
        sp = self._stage_parameters(cp, tvp, kwargs.get('ref_sc'), torch, dev, B)
        tp = None
        if self._ref_slots_tc:
            ref_tc = kwargs.get('ref_tc') or {}
            vals = [float(np.asarray(ref_tc[n]).reshape(-1)[0]) for _, names, _ in self._ref_slots_tc for n in names]
            tp = torch.tensor(vals, dtype=torch.float64, device=dev)
        uold = None
        if s.nuold:
            if self._uold is None or self._uold.shape[0] != B:
                self._uold = torch.tensor(s.u_guess[self._du_ind], dtype=torch.float64, device=dev).expand(B, -1)
            uold = self._uold.contiguous()
        if self._batch != B or not self._nlp_options.get('warm_start', True):
            eng.reset_guess()
            self._batch = B
        if v0 is not None:
            eng._alloc(B, self._solver_numbers()['n_alpha'])
            eng.reset_guess(v0)
        num = self._solver_numbers()
        if runs or not fix_x0:
            eng._alloc(B, num['n_alpha'])
        if not fix_x0:   # x_0 is a decision variable inside the state bounds; the measured state is only its initial guess
            eng._t['x'][:, 0, :] = x0s
        solve = lambda: eng.solve(x0s, dt=self._sampling_interval, time0=self._time, sp=sp, tp=tp, uold=uold, **num)
        out = solve() if not runs else self._multi_start(eng, solve, int(runs), float(kwargs.get('pert_factor', .1)), torch)
        self._last = out
        if s.nuold:
            self._uold = out['u'][:, 0, self._du_ind].clone()  # next call's u_old (mpc.py:488-493)
        u0 = out['u'][:, 0, :self._nu0] * us_t[:self._nu0]  # mpc.py:856
        self._time += self._sampling_interval
        self._n_iterations += 1
        if single:
            st = int(out['status'][0].item())
            self._solver_status_code = st
            self._solver_status = _STATUS_TEXT.get(st, 'error')
            u_np = u0[0].detach().cpu().numpy().reshape(-1, 1)
            self._solution.update(t=[self._time - self._sampling_interval],
                                  x=(x0t[0].detach().cpu().numpy()), u=(out['u'][0, 0] * us_t).detach().cpu().numpy())
            if self._stats:
                self._solution.update(extime=[_time.time() - t_start], niterations=[int(out['iters'][0].item())],
                                      solvstatus=[st])
            return u_np
        self._solver_status_code = out['status']
        return u0 if is_tensor else u0.detach().cpu().numpy()

    def _free_x0_engine(self):
        """Second generated module of the same problem with a free initial state (``optimize(..., fix_x0=False)``,
        mpc.py:803-807), built on first use."""
        if getattr(self, '_engine_free', None) is None:
            import copy
            from .engine import OcpEngine
            spec = copy.copy(self._spec)
            spec.fixed_x0 = False
            spec.name = self._spec.name + '_freex0'
            self._engine_free = OcpEngine(spec, build=getattr(self, '_build_module', True))
        return self._engine_free

    def _multi_start(self, eng, solve, runs, pert_factor, torch):
        """``runs`` solves from randomly perturbed starting points, the best converged one per instance is kept
        (mpc.py:728-740): first from the current guess ``v0``, then from ``clip(v0 + v0 (1 - 2 rand) pert_factor, lb, ub)``
        with ``rand`` drawn from the legacy NumPy RNG over the flat decision vector ``[x_0..x_N | u | X]`` of each instance."""
        t = eng._t
        B = t['x'].shape[0]
        keys = ('x', 'u', 'xc') if eng.dims.nc else ('x', 'u')
        v0 = {k: t[k].clone() for k in keys}
        s = self._spec
        f8 = dict(dtype=torch.float64, device=t['x'].device)
        from .engine import _finite
        lb = {'x': torch.tensor(_finite(s.x_lb), **f8), 'u': torch.tensor(_finite(s.u_lb), **f8)}
        ub = {'x': torch.tensor(_finite(s.x_ub), **f8), 'u': torch.tensor(_finite(s.u_ub), **f8)}
        if 'xc' in keys:
            lb['xc'], ub['xc'] = lb['x'].repeat(s.degree), ub['x'].repeat(s.degree)
        best, best_f = None, torch.full((B,), float('inf'), **f8)
        sizes = [int(np.prod(v0[k].shape[1:])) for k in keys]
        for r in range(runs):
            eng._solved_once = False          # every run is a fresh start of the solver from the given point
            out = solve()
            f = out['stats'][:, 13]           # ST_OBJ in csrc/ocp_kernels.cuh
            ok = ((out['status'] == 1) | (out['status'] == 2)) & (f < best_f)
            if best is None:
                best = {k: out[k].clone() for k in ('x', 'u', 'lam', 'status', 'iters', 'stats')}
                if out.get('xc') is not None:
                    best['xc'] = out['xc'].clone()
                best_f = torch.where(ok, f, best_f)
            elif ok.any():
                for k, v in best.items():
                    m = ok.reshape((-1,) + (1,) * (v.dim() - 1))
                    best[k] = torch.where(m, out[k], v)
                best_f = torch.where(ok, f, best_f)
            if r + 1 < runs:
                rnd = np.stack([np.random.rand(sum(sizes)) for _ in range(B)])
                pos = 0
                for k, n in zip(keys, sizes):
                    fac = torch.tensor(1. - 2. * rnd[:, pos:pos + n], **f8).reshape(v0[k].shape)
                    pos += n
                    t[k].copy_(torch.minimum(torch.maximum(v0[k] + v0[k] * fac * pert_factor, lb[k]), ub[k]))
        for k in keys:                        # the kept solution is the next warm start (mpc.py:737-738)
            t[k].copy_(best[k])
        t['lam'].copy_(best['lam'])
        t['status'].copy_(best['status'])
        best.setdefault('xc', None)
        return best

    def return_prediction(self):
        """``(x_pred [nx, N+1], u_pred [nu, N], None)`` for a single instance, batched tensors otherwise
        (mpc.py:1803-1823)."""
        if self._last is None:
            raise RuntimeError("No prediction available: call optimize() first")
        torch = __import__('torch')
        out = self._last
        xs_t, us_t = self._scaling_tensors(torch, out['x'].device)
        x = out['x'] * xs_t
        u = out['u'] * us_t
        if x.shape[0] == 1:
            return x[0].T.cpu().numpy(), u[0].T.cpu().numpy(), None
        return x, u, None

    @property
    def solver_status_code(self):
        return self._solver_status_code

    @property
    def obj_fun(self):
        if self._last is None:
            return None
        return self._last['stats'][:, 13]  # ST_OBJ in csrc/ocp_kernels.cuh
