"""``MHE``: moving-horizon estimator with the reference's API, solved in batches on the GPU.

Mirrors ``hilo_mpc/modules/estimator/mhe.py`` (``MovingHorizonEstimator`` :63, ``setup`` :418-790, ``add_measurements``
:274-309, ``estimate`` :311-416, arrival update :248-270) and the MHE cost builder ``hilo_mpc/util/modeling.py:533-777``
(``quad_arrival_cost.add_states/add_parameters``, ``quad_stage_cost.add_measurements/add_state_noise``).

NLP restated (mhe.py:596-743): variables ``[p | x_0..x_N | w_0..w_{N-1} | collocation]``; dynamics
``x_{k+1} = F(x_k, u_meas_k, p) + w_k``; cost ``arrival(x_0, p)`` at ``k = 0`` and ``stage(w_k, h(x_k), y_meas_k)`` for
``k >= 1`` only.  For the stage-structured engine the constant parameters become extra
states with trivial dynamics and the state noise is the stage input entering additively after the integration.
``estimate`` returns ``(x_N, p)`` (one-step-ahead, mhe.py:381-384) once ``horizon`` measurements are stored, and
``(None, None)`` before.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import symbolic as sx
from .costs import create_weight_matrix
from .nmpc import DynamicOptimization, _STATUS_TEXT, _vec
from .ocp import OcpSpec
from .series import TimeSeries


class MHEQuadraticCost:
    def __init__(self, model):
        self._model = model
        self.terms: List[dict] = []
        self.x_guess = None
        self.p_guess = None
        self.has_state_noise = False

    def _W(self, weights, n):
        W = create_weight_matrix(np.asarray(weights, dtype=float) if isinstance(weights, (list, tuple)) else weights)
        if W.shape[0] != n:
            raise ValueError(f"Weights dimension {W.shape} does not match the number of variables {n}")
        return W

    def add_measurements(self, weights, names=None):
        m = self._model
        names = m.measurement_names if names is None else ([names] if isinstance(names, str) else list(names))
        for n in names:
            if n not in m.measurement_names:
                raise ValueError(f"The measurement {n} does not exist. The available states are {m.measurement_names}")
        ind = [m.measurement_names.index(n) for n in names]
        self.terms.append(dict(kind='measurements', ind=ind, W=self._W(weights, len(ind))))

    def add_state_noise(self, weights):
        self.has_state_noise = True
        self.terms.append(dict(kind='state_noise', ind=list(range(self._model.n_x)),
                               W=self._W(weights, self._model.n_x)))

    def add_inputs(self, names, weights):
        raise NotImplementedError("Input-measurement costs are not supported by the B200 engine yet")

    def add_states(self, weights, guess):
        guess = np.asarray(guess, dtype=float).reshape(-1)
        if guess.size != self._model.n_x:
            raise ValueError(f"The guess must have the same dimension of the model states.There are "
                             f"{self._model.n_x} states but guess has {guess.size} values.")
        self.x_guess = guess
        self.terms.append(dict(kind='states', ind=list(range(self._model.n_x)), W=self._W(weights, self._model.n_x)))

    def add_parameters(self, weights, guess):
        guess = np.asarray(guess, dtype=float).reshape(-1)
        if guess.size != self._model.n_p:
            raise ValueError(f"The guess must have the same dimension of the model parameters.There are "
                             f"{self._model.n_p} states but guess has {guess.size} values.")
        self.p_guess = guess
        self.terms.append(dict(kind='parameters', ind=list(range(self._model.n_p)),
                               W=self._W(weights, self._model.n_p)))


class MovingHorizonEstimator(DynamicOptimization):
    _type = 'MHE'

    def __init__(self, model, id=None, name=None, plot_backend=None, time=0):
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)
        m = self._model
        self.quad_arrival_cost = MHEQuadraticCost(m)
        self.quad_stage_cost = MHEQuadraticCost(m)
        self._time = float(time)
        self._w_lb = self._w_ub = self._p_lb = self._p_ub = None
        self._w_guess = self._p_guess = None
        self._w_scaling = self._p_scaling = None
        self._window = None   # [B, N, nu + ny] device tensor: the measurements inside the horizon
        self._meas_counter = 0
        self._horizon_is_reached = False
        self._last = None
        self._x_arrival = self._p_arrival = None

    # ------------------------------------------------------------------ option surface (mhe.py:792-1060)
    def set_box_constraints(self, x_ub=None, x_lb=None, w_ub=None, w_lb=None, p_lb=None, p_ub=None, z_ub=None,
                            z_lb=None):
        m = self._model
        self._x_ub, self._x_lb = _vec(x_ub, m.n_x, 'x_ub', np.inf), _vec(x_lb, m.n_x, 'x_lb', -np.inf)
        self._w_ub, self._w_lb = _vec(w_ub, m.n_x, 'w_ub', np.inf), _vec(w_lb, m.n_x, 'w_lb', -np.inf)
        self._p_ub, self._p_lb = _vec(p_ub, m.n_p, 'p_ub', np.inf), _vec(p_lb, m.n_p, 'p_lb', -np.inf)

    def set_initial_guess(self, x_guess=None, w_guess=None, p_guess=None, z_guess=None):
        m = self._model
        self._x_guess = _vec(x_guess, m.n_x, 'x_guess', 0.)
        self._w_guess = _vec(w_guess, m.n_x, 'w_guess', 0.)
        self._p_guess = _vec(p_guess, m.n_p, 'p_guess', 0.)

    def set_scaling(self, x_scaling=None, w_scaling=None, p_scaling=None, u_scaling=None):
        m = self._model
        self._x_scaling = _vec(x_scaling, m.n_x, 'x_scaling', 1.)
        self._w_scaling = _vec(w_scaling, m.n_x, 'w_scaling', 1.)
        self._p_scaling = _vec(p_scaling, m.n_p, 'p_scaling', 1.)
        self._u_scaling = _vec(u_scaling, m.n_u, 'u_scaling', 1.)

    def set_nlp_options(self, *args, **kwargs):
        given = dict(args[0]) if args and isinstance(args[0], dict) else dict(kwargs)
        self._arrival_update = given.pop('arrival_guess_update', 'smoothing')
        if self._arrival_update not in ('smoothing',):
            raise NotImplementedError("The filtering update is not yet implemented.")
        given.pop('objective_function', None)
        super().set_nlp_options(given)

    # ------------------------------------------------------------------ setup
    def setup(self, options=None, nlp_opts=None, solver='ipopt', **kwargs) -> None:
        m = self._model
        if self._prediction_horizon is None:
            raise ValueError("The horizon must be set before setting up the MHE (mhe.horizon = N).")
        if options is not None:
            self.set_nlp_options(options)
        elif self._nlp_options is None:
            self.set_nlp_options()
        if nlp_opts is not None:
            self.set_solver_opts(nlp_opts)
        if self._x_scaling is None:
            self.set_scaling()
        if self._x_lb is None:
            self.set_box_constraints()
        arr, stg = self.quad_arrival_cost, self.quad_stage_cost
        if self._x_guess is None:
            self.set_initial_guess(x_guess=arr.x_guess, p_guess=arr.p_guess)
        nx, nu_m, npar, ny = m.n_x, m.n_u, m.n_p, m.n_y
        N = self._prediction_horizon
        has_noise = stg.has_state_noise
        nw = nx if has_noise else 1  # a cost-free dummy input keeps the stage structure when there is no state noise
        xs_sc, ps_sc, ws_sc = self._x_scaling, self._p_scaling, self._w_scaling
        # augmented scaled state [x; p], stage input w, stage parameters [u_meas; y_meas]
        xa = sx.named_syms(m.dynamical_state_names + [f'p_{n}' for n in m.parameter_names])
        w = sx.named_syms([f'w_{n}' for n in m.dynamical_state_names] if has_noise else ['w_dummy'])
        um = sx.named_syms([f'{n}_meas' for n in m.input_names])
        ym = sx.named_syms([f'{n}_meas' for n in m.measurement_names])
        t, dt = sx.sym('t'), sx.sym('dt')
        x_phys = xa[:nx] * xs_sc
        p_phys = xa[nx:] * ps_sc
        old = np.concatenate([m.x, m.u, m.p, [m.t], [m.dt]])
        new = np.concatenate([x_phys, um, p_phys, [t], [dt]])
        f = sx.substitute(m.ode, old, new) / xs_sc
        zero = sx.Expr.const(0.)
        ode = np.concatenate([f, xa[nx:] if m.discrete else np.array([zero] * npar, dtype=object)])
        h = sx.substitute(m.meas, old, new) if ny else np.zeros(0, dtype=object)
        # costs act on physical quantities (modeling.py:678-684 substitutes x -> x*scale into the MHE cost)
        lag, lag_ls = zero, []
        for term in stg.terms:
            if term['kind'] == 'measurements':
                r = h[term['ind']] - ym[term['ind']]
            elif term['kind'] == 'state_noise':
                r = w * ws_sc
            else:
                raise NotImplementedError(f"Stage cost on {term['kind']} is not supported")
            r = sx.as_vector(r)
            lag = lag + sx.quad_form(r, term['W'])
            lag_ls.append((r, term['W']))
        ap = sx.named_syms([f'{n}_arr' for n in m.dynamical_state_names + m.parameter_names])
        arrival, arr_ls = zero, []
        for term in arr.terms:
            if term['kind'] == 'states':
                r = x_phys - ap[:nx]
            elif term['kind'] == 'parameters':
                r = p_phys - ap[nx:]
            else:
                raise NotImplementedError(f"Arrival cost on {term['kind']} is not supported")
            r = sx.as_vector(r)
            arrival = arrival + sx.quad_form(r, term['W'])
            arr_ls.append((r, term['W']))
        opts = self._nlp_options
        method = {'collocation': 'collocation', 'rk4': 'erk', 'erk': 'erk', 'discrete': 'discrete'}[
            opts['integration_method']]
        gadd = np.zeros((nx + npar, nw))
        if has_noise:
            gadd[:nx, :nx] = np.eye(nx)
        w_lb = self._w_lb / ws_sc if has_noise else np.array([-np.inf])
        w_ub = self._w_ub / ws_sc if has_noise else np.array([np.inf])
        w_guess = self._w_guess / ws_sc if has_noise else np.zeros(1)
        self._spec = OcpSpec(
            nx=nx + npar, nu=nw, N=N, t=t, dt=dt, x=xa, u=w, sp=np.concatenate([um, ym]),
            tp=np.zeros(0, dtype=object), uold=np.zeros(0, dtype=object), ode=sx.as_vector(ode), lag=lag,
            lag_du=zero, may=zero, lag_ls=lag_ls, lag_rest=None, may_ls=[], may_rest=None, method=method,
            degree=int(opts['degree']), points=opts['collocation_points'], order=int(opts['order'] or 1),
            objective='discrete',
            x_lb=np.concatenate([self._x_lb / xs_sc, self._p_lb / ps_sc]),
            x_ub=np.concatenate([self._x_ub / xs_sc, self._p_ub / ps_sc]), u_lb=w_lb, u_ub=w_ub,
            x_guess=np.concatenate([self._x_guess / xs_sc, self._p_guess / ps_sc]), u_guess=w_guess, gadd=gadd,
            fixed_x0=False, lag_from=1, arr=arrival, arr_ls=arr_ls, ap=ap, exact_hessian=self._exact_hessian(),
            name=f'mhe_{m.name}_{method}_N{N}')
        from .engine import OcpEngine
        self._engine = OcpEngine(self._spec, build=getattr(self, '_build_module', True))
        self._solution = TimeSeries(t=['t'], x=m.dynamical_state_names, p=m.parameter_names)
        self._nlp_setup_done = True
        self._x_arrival = None if arr.x_guess is None else arr.x_guess.copy()
        self._p_arrival = None if arr.p_guess is None else arr.p_guess.copy()

    @property
    def spec(self):
        return self._spec

    @property
    def engine(self):
        return self._engine

    @property
    def has_state_noise(self):
        return self.quad_stage_cost.has_state_noise

    # ------------------------------------------------------------------ measurements / estimate
    def add_measurements(self, y_meas, u_meas=None) -> None:
        """Append one sample; ``y_meas`` ``(ny,)`` or batched ``(B, ny)``, NumPy or device tensor (mhe.py:274-309).
        The horizon window ``[u_meas | y_meas]`` of every instance lives on the device (``[B, N, nu + ny]``, the stage
        parameters of the solve): a sample costs one copy of its own bytes, not of the window."""
        if not self._nlp_setup_done:
            raise RuntimeError("You need to setup the MHE by running .setup() before running add_measurements")
        from . import _lib
        torch = _lib.require_cuda()
        eng = self._engine
        eng._open()
        dev = eng._device
        m = self._model
        N = self._prediction_horizon

        def as_dev(v, n, what):
            t = v.to(device=dev, dtype=torch.float64) if torch.is_tensor(v) else \
                torch.as_tensor(np.asarray(v, dtype=float), dtype=torch.float64, device=dev)
            if t.numel() == n:
                return t.reshape(1, n)
            if t.numel() % n != 0:
                raise ValueError(f"{what} has {t.numel()} elements, expected a multiple of {n}")
            return t.reshape(-1, n)
        y = as_dev(y_meas, m.n_y, 'y_meas')
        if m.n_u:
            if u_meas is None:
                raise ValueError(f"The model has {m.n_u} input(s): pass their measured values with u_meas=")
            u = as_dev(u_meas, m.n_u, 'u_meas')
        else:
            u = None
        B = max(y.shape[0], 1 if u is None else u.shape[0])
        win = getattr(self, '_window', None)
        if win is None or win.shape[0] != B:
            if win is not None and self._meas_counter > 0 and win.shape[0] != 1:
                raise ValueError(f"The batch size changed from {win.shape[0]} to {B} while the horizon window is filling")
            new = torch.zeros(B, N, m.n_u + m.n_y, dtype=torch.float64, device=dev)
            if win is not None and self._meas_counter > 0:
                new.copy_(win.expand(B, -1, -1))
            self._window = win = new
        if self._meas_counter >= N:   # window full: drop the oldest sample
            win[:, :-1] = win[:, 1:].clone()
            pos = N - 1
        else:
            pos = self._meas_counter
        if u is not None:
            win[:, pos, :m.n_u] = u
        win[:, pos, m.n_u:] = y
        self._meas_counter += 1
        if self._meas_counter >= N:
            self._horizon_is_reached = True

    def estimate(self, x_arrival=None, p_arrival=None, v0=None, runs=0, **kwargs):
        if not self._nlp_setup_done:
            raise ValueError("You need to setup the nlp before optimizing. Type *mheObject*.setup()")
        self._time += self._sampling_interval
        if not self._horizon_is_reached:
            return None, None
        from . import _lib
        torch = _lib.require_cuda()
        eng = self._engine
        eng._open()
        dev = eng._device
        m, s = self._model, self._spec
        nx, npar, N = m.n_x, m.n_p, s.N
        sp = self._window
        B = sp.shape[0]
        f8 = dict(dtype=torch.float64, device=dev)

        def arrival(given, stored, n, what):
            v = stored if given is None else given
            if v is None:
                raise ValueError(f"No arrival guess for the {what}: use quad_arrival_cost.add_{what}(weights, guess)")
            t = v.to(**f8) if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=float), **f8)
            return t.reshape(-1, n).expand(B, n)
        xa = arrival(x_arrival, self._x_arrival, nx, 'states')
        apt = xa if npar == 0 else torch.cat([xa, arrival(p_arrival, self._p_arrival, npar, 'parameters')], dim=1)
        apt = apt.contiguous()
        if getattr(self, '_x0_dummy', None) is None or self._x0_dummy.shape[0] != B:
            self._x0_dummy = torch.zeros(B, s.nx, **f8)
        if getattr(self, '_batch', None) != B:
            eng.reset_guess()
            self._batch = B
        num = self._solver_numbers()
        out = eng.solve(self._x0_dummy, dt=self._sampling_interval, time0=self._time - N * self._sampling_interval,
                        sp=sp, ap=apt, **num)
        self._last = out
        if getattr(self, '_xs_dev', None) is None or self._xs_dev.device != dev:
            self._xs_dev = torch.tensor(np.concatenate([self._x_scaling, self._p_scaling]), **f8)
        xs_t = self._xs_dev
        xN = out['x'][:, N, :] * xs_t
        # arrival update 'smoothing' (mhe.py:254-256 reads node 2 of the previous solution); kept on the device
        node = min(2, N)
        self._x_arrival = out['x'][:, node, :nx] * xs_t[:nx]
        if npar:
            self._p_arrival = out['x'][:, 0, nx:] * xs_t[nx:]
        st = out['status']
        if B == 1:
            code = int(st[0].item())
            self._solver_status_code, self._solver_status = code, _STATUS_TEXT.get(code, 'error')
            x_opt = xN[0, :nx].cpu().numpy().reshape(-1, 1)
            p_opt = xN[0, nx:].cpu().numpy().reshape(-1, 1) if npar else None
            self._solution.update(t=[self._time], x=x_opt[:, 0], p=None if p_opt is None else p_opt[:, 0])
            return x_opt, p_opt
        self._solver_status_code = st
        return xN[:, :nx], (xN[:, nx:] if npar else None)

    def return_mhe_estimation(self):
        if self._last is None:
            raise RuntimeError("No estimation available: call estimate() first")
        return self._last['x'], self._last['u']


MHE = MovingHorizonEstimator
