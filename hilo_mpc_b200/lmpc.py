"""``LMPC``: linear MPC with the reference's API.

Mirrors ``hilo_mpc/modules/controller/mpc.py:1934-2459``: ``Q, R, P`` setters (:2437-2459), ``horizon``,
``set_box_constraints``, ``set_scaling``, ``setup`` (:2143), ``optimize(x0)`` -> ``u_0`` (:2307), ``return_prediction`` (:2408);
``set_stage_constraints`` / ``set_custom_constraints_function`` / ``set_initial_guess`` raise ``NotImplementedError``
(:2396-2406).  The reference assembles the sparse QP ``min 1/2 v^T H v  s.t.  A x_k - x_{k+1} + B u_k = 0`` with
``v = [x_0..x_N | u_0..u_{N-1}]``, ``H = blkdiag(I_N (x) Q, P, I_N (x) R)`` and hands it to qpOASES each step (:2195-2276,
:2374).  Here the same QP keeps its stage structure and is solved for a whole batch by the Riccati interior-point
engine (one Newton system = one Riccati sweep; the QP is strictly convex, so the unique minimiser is the qpOASES one).

Input block: the reference builds it as ``kron(B, I_N)`` (mpc.py:2243), which matches the stage-major variable order only for
``nx = nu = 1``.  This class implements the mathematically intended ``I_N (x) B``.
"""
from __future__ import annotations

import numpy as np

from . import symbolic as sx
from .costs import create_weight_matrix
from .nmpc import DynamicOptimization, _STATUS_TEXT
from .ocp import OcpSpec
from .series import TimeSeries


class LMPC(DynamicOptimization):
    _type = 'LMPC'

    def __init__(self, model, id=None, name=None, plot_backend=None, use_sx: bool = True):
        if not model.discrete:
            raise TypeError("The model used for the LMPC must be discrete. Use model.discretize()")  # mpc.py:1937-1947
        if not model.is_linear():
            raise TypeError("The model used for the LMPC must be linear. Use model.linearize() to obtain a linear "
                            "model or use the nonlinear MPC.")
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)
        self._Q = self._R = self._P = None
        self._last = None
        self._batch = None

    # ------------------------------------------------------------------ weights
    @property
    def Q(self):
        return self._Q

    @Q.setter
    def Q(self, arg):
        self._Q = self._weight(arg, self._model.n_x, 'Q')

    @property
    def R(self):
        return self._R

    @R.setter
    def R(self, arg):
        self._R = self._weight(arg, self._model.n_u, 'R')

    @property
    def P(self):
        return self._P

    @P.setter
    def P(self, arg):
        self._P = self._weight(arg, self._model.n_x, 'P')

    @staticmethod
    def _weight(arg, n, name):
        W = create_weight_matrix(np.asarray(arg, dtype=float) if isinstance(arg, (list, tuple)) else arg, name)
        if W.shape != (n, n):
            raise ValueError(f"The matrix {name} must be {n}x{n}, got {W.shape[0]}x{W.shape[1]}")
        return W

    def set_stage_constraints(self, *args, **kwargs):
        raise NotImplementedError("Stage constraints are not yet implemented for the linear MPC")

    def set_custom_constraints_function(self, *args, **kwargs):
        raise NotImplementedError("Custom constraints are not yet implemented for the linear MPC")

    def set_initial_guess(self, *args, **kwargs):
        raise NotImplementedError("Initial guess is not necessary for the linear MPC")

    # ------------------------------------------------------------------ setup / optimize
    def setup(self, options=None, solver_options=None, solver='qpoases') -> None:
        m = self._model
        if self._prediction_horizon is None:
            raise ValueError("The horizon must be set before setting up the LMPC (lmpc.horizon = N).")
        if self._Q is None or self._R is None:
            raise ValueError("The weighting matrices Q and R must be set before setting up the LMPC")
        if self._P is None:
            self._P = self._Q.copy()
        if solver_options is not None:
            self.set_solver_opts(solver_options)
        if self._x_scaling is None:
            self.set_scaling()
        if self._x_lb is None:
            self.set_box_constraints()
        if self._sampling_interval is None:
            self.set_sampling_interval()
        nx, nu = m.n_x, m.n_u
        xs, us = sx.named_syms(m.dynamical_state_names), sx.named_syms(m.input_names)
        p_names = m.parameter_names
        tvp_idx = [p_names.index(n) for n in self._tvp_names]
        cp_idx = [k for k in range(len(p_names)) if k not in tvp_idx]
        self._tvp_idx, self._cp_idx = tvp_idx, cp_idx
        p_slots = sx.named_syms([f'p_{p_names[k]}' for k in tvp_idx + cp_idx])
        t, dt = sx.sym('t'), sx.sym('dt')
        x_scale, u_scale = self._x_scaling, self._u_scaling
        old = np.concatenate([m.x, m.u, m.p[tvp_idx + cp_idx] if len(p_names) else [], [m.t], [m.dt]])
        new = np.concatenate([xs * x_scale, us * u_scale, p_slots, [t], [dt]])
        ode = sx.substitute(m.ode, old, new) / x_scale
        # weights are scaled like the reference does for Q, P, R (mpc.py:2082-2089): cost on physical variables
        Qs = self._Q * np.outer(x_scale, x_scale)
        Ps = self._P * np.outer(x_scale, x_scale)
        Rs = self._R * np.outer(u_scale, u_scale)
        lag = .5 * sx.quad_form(xs, Qs) + .5 * sx.quad_form(us, Rs)
        may = .5 * sx.quad_form(xs, Ps)
        N = self._prediction_horizon
        self._spec = OcpSpec(
            nx=nx, nu=nu, N=N, t=t, dt=dt, x=xs, u=us, sp=p_slots, tp=np.zeros(0, dtype=object),
            uold=np.zeros(0, dtype=object), ode=sx.as_vector(ode), lag=lag, lag_du=sx.Expr.const(0.), may=may,
            lag_ls=[(xs, .5 * Qs), (us, .5 * Rs)], lag_rest=None, may_ls=[(xs, .5 * Ps)], may_rest=None,
            method='discrete', objective='discrete', x_lb=self._x_lb / x_scale, x_ub=self._x_ub / x_scale,
            u_lb=self._u_lb / u_scale, u_ub=self._u_ub / u_scale, x_guess=np.zeros(nx), u_guess=np.zeros(nu),
            name=f'lmpc_{m.name}_N{N}')
        g = self._spec
        # start inside the bounds: the guess is projected on the box midpoint where zero is infeasible
        for lo, hi, gs in ((g.x_lb, g.x_ub, g.x_guess), (g.u_lb, g.u_ub, g.u_guess)):
            bad = (gs <= lo) | (gs >= hi)
            mid = np.where(np.isfinite(lo) & np.isfinite(hi), .5 * (lo + hi), np.where(np.isfinite(lo), lo + 1., hi - 1.))
            gs[bad] = mid[bad]
        from .engine import OcpEngine
        self._engine = OcpEngine(self._spec, build=getattr(self, '_build_module', True))
        self._solution = TimeSeries(t=['t'], x=m.dynamical_state_names, u=m.input_names)
        self._nlp_setup_done = True

    @property
    def spec(self):
        return self._spec

    @property
    def engine(self):
        return self._engine

    def optimize(self, x0, cp=None, tvp=None, **kwargs):
        if not self._nlp_setup_done:
            raise RuntimeError("The LMPC is not set up. Run LMPC.setup() before optimizing.")
        from . import _lib
        torch = _lib.require_cuda()
        eng = self._engine
        eng._open()
        dev = eng._device
        nx, nu = self._spec.nx, self._spec.nu
        is_tensor = torch.is_tensor(x0)
        x0t = x0.to(device=dev, dtype=torch.float64) if is_tensor else \
            torch.as_tensor(np.asarray(x0, dtype=float), dtype=torch.float64, device=dev)
        single = x0t.dim() <= 1 or (x0t.dim() == 2 and x0t.shape == (nx, 1) and nx > 1)
        x0t = x0t.reshape(1, nx) if single else x0t
        if x0t.shape[1] != nx:
            raise ValueError(f"x0 must have {nx} rows (single instance) or shape (B, {nx}); got {tuple(x0t.shape)}")
        B = x0t.shape[0]
        xs_t = torch.tensor(self._x_scaling, dtype=torch.float64, device=dev)
        us_t = torch.tensor(self._u_scaling, dtype=torch.float64, device=dev)
        sp = None
        if self._spec.nsp:
            from .nmpc import NMPC
            self._ref_slots_sc = []
            sp = NMPC._stage_parameters(self, cp, tvp, None, torch, dev, B)
        if self._batch != B:
            eng.reset_guess()
            self._batch = B
        num = self._solver_numbers()
        out = eng.solve(x0t / xs_t, dt=self._sampling_interval, time0=self._time, sp=sp, **num)
        self._last = out
        self._time += self._sampling_interval
        self._n_iterations += 1
        u0 = out['u'][:, 0, :] * us_t
        if single:
            st = int(out['status'][0].item())
            self._solver_status_code, self._solver_status = st, _STATUS_TEXT.get(st, 'error')
            self._solution.update(t=[self._time - self._sampling_interval], x=x0t[0].cpu().numpy(),
                                  u=u0[0].cpu().numpy())
            return u0[0].cpu().numpy().reshape(-1, 1)
        self._solver_status_code = out['status']
        return u0 if is_tensor else u0.cpu().numpy()

    def return_prediction(self):
        if self._last is None:
            raise RuntimeError("No prediction available: call optimize() first")
        torch = __import__('torch')
        out = self._last
        x = out['x'] * torch.tensor(self._x_scaling, dtype=torch.float64, device=out['x'].device)
        u = out['u'] * torch.tensor(self._u_scaling, dtype=torch.float64, device=out['x'].device)
        if x.shape[0] == 1:
            return x[0].T.cpu().numpy(), u[0].T.cpu().numpy()
        return x, u

    @property
    def solver_status_code(self):
        return self._solver_status_code
