"""``LMPC``: linear MPC with the reference's API.

Mirrors ``hilo_mpc/modules/controller/mpc.py:1934-2459``: ``Q, R, P`` setters (:2437-2459), ``horizon``,
``set_box_constraints``, ``set_scaling``, ``setup`` (:2143), ``optimize(x0)`` -> ``u_0`` (:2307), ``return_prediction`` (:2408);
``set_stage_constraints`` / ``set_custom_constraints_function`` / ``set_initial_guess`` raise ``NotImplementedError``
(:2396-2406).  The reference assembles the sparse QP ``min 1/2 v^T H v  s.t.  A x_k - x_{k+1} + B u_k = 0`` with
``v = [x_0..x_N | u_0..u_{N-1}]``, ``H = blkdiag(I_N (x) Q, P, I_N (x) R)`` and hands it to qpOASES each step (:2195-2276,
:2374).  Here the same QP keeps its stage structure and is solved for a whole batch by the Riccati interior-point
engine (one Newton system = one Riccati sweep; the QP is strictly convex, so the unique minimiser is the qpOASES one).

Two solvers (``setup(solver=...)``):
  * ``'qpoases'`` / ``'condensed'`` (default): the horizon is condensed on the host once (``x = Phi x0 + Gam u``) and every
    instance solves the dense QP in the ``N nu`` inputs with the batched interior-point kernel ``hb_lmpc_solve``
    (csrc/lmpc_qp.cuh) - the "condensing + dense QP" path of the north star;
  * ``'riccati'``: the same QP in its stage structure through the Riccati interior-point engine (also used when the model
    has time-varying parameters, whose condensed matrices would change every step).

Input block: the reference builds it as ``kron(B, I_N)`` (mpc.py:2243), which matches the stage-major variable order only for
``nx = nu = 1``.  Default here is the mathematically intended ``I_N (x) B``;
``LMPC(model, bug_compat=True)`` reproduces the reference's equality block entry by entry (condensed solver only), i.e.
the QP the reference actually hands to qpOASES.
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

    def __init__(self, model, id=None, name=None, plot_backend=None, use_sx: bool = True, bug_compat: bool = False):
        if not model.discrete:
            raise TypeError("The model used for the LMPC must be discrete. Use model.discretize()")  # mpc.py:1937-1947
        if not model.is_linear():
            raise TypeError("The model used for the LMPC must be linear. Use model.linearize() to obtain a linear "
                            "model or use the nonlinear MPC.")
        super().__init__(model, id=id, name=name, plot_backend=plot_backend)
        self._Q = self._R = self._P = None
        self._last = None
        self._batch = None
        self._bug_compat = bool(bug_compat)
        self._condensed = None

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
        if solver not in ('qpoases', 'condensed', 'riccati'):
            raise ValueError(f"The solver {solver} does no exist. The possible solver are ['qpoases', 'condensed', "
                             f"'riccati'].")
        use_condensed = solver in ('qpoases', 'condensed') and not tvp_idx and not cp_idx
        if self._bug_compat and not use_condensed:
            raise NotImplementedError("bug_compat=True reproduces the reference's kron(B, I_N) equality block, which has no "
                                      "stage structure: it needs the condensed solver (no model parameters)")
        self._condensed = self._condense(ode, xs, us, Qs, Ps, Rs) if use_condensed else None
        from .engine import OcpEngine
        self._engine = None if use_condensed else OcpEngine(self._spec, build=getattr(self, '_build_module', True))
        self._solution = TimeSeries(t=['t'], x=m.dynamical_state_names, u=m.input_names)
        self._nlp_setup_done = True

    def _condense(self, ode, xs, us, Qs, Ps, Rs):
        """Condensed matrices of the reference's QP (mpc.py:2195-2266; scaled coordinates): with ``v = [x_0..x_N | u]``,
        ``H = blkdiag(I_N (x) Q, P, I_N (x) R)`` and the equality block ``Aeq v = 0`` whose row block k reads
        ``A x_k - x_{k+1} + (input block) u = 0``, the states are ``x_{1..N} = Phi x_0 + Gam u`` and the QP in ``u`` has
        ``Hc = Gam^T Qbar Gam + Rbar`` and linear term ``(Gam^T Qbar Phi) x_0``, ``Qbar = blkdiag(I_{N-1} (x) Q, P)``.
        The cost of ``x_0`` itself is constant.  Input block: ``I_N (x) B`` or, with ``bug_compat``, the reference's
        ``kron(B, I_N)`` - entry ``B[i, j]`` at row ``i N + a``, column ``j N + a`` of the stage-major system (mpc.py:2243)."""
        nx, nu, N = len(xs), len(us), self._prediction_horizon
        free = [v for v in sx.symbols_of(sx.as_vector(ode)) if v not in set(xs) | set(us)]   # t, dt of the discrete map
        vals = [self._sampling_interval if v.name == 'dt' else self._time for v in free]
        point = np.concatenate([np.zeros(nx + nu), vals])
        syms = np.concatenate([xs, us, np.array(free, dtype=object)])
        A = np.array(sx.evaluate(sx.jacobian(ode, xs), syms, point), dtype=float).reshape(nx, nx)
        Bm = np.array(sx.evaluate(sx.jacobian(ode, us), syms, point), dtype=float).reshape(nx, nu)
        E = np.kron(Bm, np.eye(N)) if self._bug_compat else np.kron(np.eye(N), Bm)     # (N nx) x (N nu)
        Phi = np.zeros((N * nx, nx))
        Gam = np.zeros((N * nx, N * nu))
        xk_x0, xk_u = np.eye(nx), np.zeros((nx, N * nu))
        for k in range(N):       # x_{k+1} = A x_k + E[rows of stage k] u
            xk_x0 = A @ xk_x0
            xk_u = A @ xk_u + E[k * nx:(k + 1) * nx]
            Phi[k * nx:(k + 1) * nx] = xk_x0
            Gam[k * nx:(k + 1) * nx] = xk_u
        # x_1..x_{N-1} carry Q, x_N carries P; the Hessian of the reference's objective 1/2 v^T H v with H built from the
        # SCALED weight matrices (mpc.py:2082-2089; the Lagrange term of the stage-structured form is 1/2 x^T Q x as well)
        Qbar = np.kron(np.eye(N), Qs)
        Qbar[(N - 1) * nx:, (N - 1) * nx:] = Ps
        Rbar = np.kron(np.eye(N), Rs)
        Hc = Gam.T @ Qbar @ Gam + Rbar
        Hc = .5 * (Hc + Hc.T)
        F = Gam.T @ Qbar @ Phi
        g = self._spec
        return dict(H=Hc, F=F, Gam=Gam, Phi=Phi, ulb=np.tile(g.u_lb, N), uub=np.tile(g.u_ub, N),
                    xlb=np.tile(g.x_lb, N), xub=np.tile(g.x_ub, N), n=N * nu, mx=N * nx, A=A, B=Bm)

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
        if self._condensed is not None:
            return self._optimize_condensed(x0, torch)
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

    def _optimize_condensed(self, x0, torch):
        from . import _lib
        lib = _lib.load()
        c = self._condensed
        nx, nu, N = self._spec.nx, self._spec.nu, self._spec.N
        dev = torch.device('cuda', torch.cuda.current_device())
        f8 = dict(dtype=torch.float64, device=dev)
        if 'dev' not in c or c['dev'] != dev:
            from .engine import _finite
            for k in ('H', 'F', 'Gam', 'Phi'):
                c[k + '_t'] = torch.tensor(c[k], **f8).contiguous()
            for k in ('ulb', 'uub', 'xlb', 'xub'):
                c[k + '_t'] = torch.tensor(_finite(c[k]), **f8)
            c['xs_t'] = torch.tensor(self._x_scaling, **f8)
            c['us_t'] = torch.tensor(self._u_scaling, **f8)
            c['dev'] = dev
        is_tensor = torch.is_tensor(x0)
        x0t = x0.to(**f8) if is_tensor else torch.as_tensor(np.asarray(x0, dtype=float), **f8)
        single = x0t.dim() <= 1 or (x0t.dim() == 2 and x0t.shape == (nx, 1) and nx > 1)
        x0t = x0t.reshape(1, nx) if single else x0t
        if x0t.shape[1] != nx:
            raise ValueError(f"x0 must have {nx} rows (single instance) or shape (B, {nx}); got {tuple(x0t.shape)}")
        B = x0t.shape[0]
        x0s = (x0t / c['xs_t']).contiguous()
        n, mx = c['n'], c['mx']
        if c.get('batch') != B:
            c['u'] = torch.empty(B, n, **f8)
            c['xp'] = torch.empty(B, mx, **f8)
            c['status'] = torch.zeros(B, dtype=torch.int32, device=dev)
            c['iters'] = torch.zeros(B, dtype=torch.int32, device=dev)
            c['work'] = torch.empty(lib.hb_lmpc_work_doubles(B, n, mx), **f8)
            c['batch'] = B
        o = self._nlp_opts
        tol = float(o.get('ipopt.tol', o.get('tol', 1e-12)))   # relative KKT tolerance of the dense-QP kernel
        rc = lib.hb_lmpc_solve(B, n, mx, nx, c['H_t'].data_ptr(), c['F_t'].data_ptr(), c['Gam_t'].data_ptr(),
                               c['Phi_t'].data_ptr(), c['ulb_t'].data_ptr(), c['uub_t'].data_ptr(), c['xlb_t'].data_ptr(),
                               c['xub_t'].data_ptr(), x0s.data_ptr(), c['u'].data_ptr(), c['xp'].data_ptr(),
                               c['status'].data_ptr(), c['iters'].data_ptr(), tol, int(o.get('max_iter', 100)),
                               c['work'].data_ptr(), c['work'].numel(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, 'hb_lmpc_solve')
        c['_keep'] = x0s
        u = c['u'].reshape(B, N, nu)
        x = torch.cat([x0s[:, None, :], c['xp'].reshape(B, N, nx)], dim=1)
        self._last = dict(x=x, u=u, status=c['status'], iters=c['iters'])
        self._time += self._sampling_interval
        self._n_iterations += 1
        u0 = u[:, 0, :] * c['us_t']
        if single:
            st = int(c['status'][0].item())
            self._solver_status_code, self._solver_status = st, _STATUS_TEXT.get(st, 'error')
            self._solution.update(t=[self._time - self._sampling_interval], x=x0t[0].cpu().numpy(), u=u0[0].cpu().numpy())
            return u0[0].cpu().numpy().reshape(-1, 1)
        self._solver_status_code = c['status']
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
