"""Oracle (test infrastructure): the reference's flat NMPC/MHE NLP, restated, and solved on the CPU.

Restates ``NMPC._setup`` (``hilo_mpc/modules/controller/mpc.py:1440-1728``) for the feature set the engine
supports (ODE models; collocation / explicit RK / discrete transcription; quadratic or generic stage & terminal
costs; box bounds; input-change cost; time-varying and constant parameters; references):

decision vector ``v`` (mpc.py:1455-1527):   ``x_0..x_N | u_0..u_{N-1} | for each stage: X_{k,1..deg}``
constraints ``g`` per stage (mpc.py:1627-1669): ``collocation equations (deg*nx)`` then ``x_{k+1} - xf`` (nx)
objective (mpc.py:1672-1682): ``sum_k quad_k`` (continuous) or ``sum_k lag(x_k, u_k)`` (discrete), ``+ may(xf_{N-1})``
collocation (modeling.py:1180-1195): ``dt*f(t+T[i-1], X_i, u, p) - sum_j C[j,i] X_j = 0``, ``xf = sum_j D[j] X_j``,
``quad = sum_i B[i]*lag(t+T[i-1], X_i, u)*dt``;   ERK (modeling.py:1255-1281).

Bounds: ``x_lb/x_ub`` on every node and every collocation state (mpc.py:1325-1326, 1465-1473, 1503-1506); ``x_0`` pinned via
``lb = ub = x0`` (mpc.py:797-807).  Solver: ``oracle.ipm.solve_nlp`` (IPOPT restatement).

The per-stage expressions are differentiated with ``hilo_mpc_b200.symbolic`` (validated against complex-step
differentiation in ``tests/test_symbolic.py``) and evaluated for all stages at once.
"""
import numpy as np
import scipy.sparse as sp

from hilo_mpc_b200 import symbolic as sx
from .collocation import polynomial_basis, tableau_for_order
from .ipm import solve_nlp


class FlatNlp:
    """Flat NLP of one OCP instance. ``spec`` is an ``hilo_mpc_b200.ocp.OcpSpec``."""

    def __init__(self, spec, exact_hessian=True):
        self.spec = s = spec
        self.exact_hessian = exact_hessian
        nx, nu, N = s.nx, s.nu, s.N
        self.ncol = ncol = s.ncol
        # ---- index maps into v (mpc.py:1462-1527)
        off = 0
        self.x_ind = [np.arange(off + k * nx, off + (k + 1) * nx) for k in range(N + 1)]
        off += (N + 1) * nx
        self.u_ind = [np.arange(off + k * nu, off + (k + 1) * nu) for k in range(N)]
        off += N * nu
        self.c_ind = [np.arange(off + k * ncol, off + (k + 1) * ncol) for k in range(N)]
        off += N * ncol
        self.n_v = off
        self.n_gs = ncol + nx
        self.n_g = N * self.n_gs
        self._build_stage_functions()
        # local variable order w = [x_k, u_k, X_k, x_{k+1}] -> global indices, per stage
        self.w_ind = np.stack([np.concatenate([self.x_ind[k], self.u_ind[k], self.c_ind[k], self.x_ind[k + 1]])
                               for k in range(N)])
        self.nw = self.w_ind.shape[1]

    # ------------------------------------------------------------------ symbolic stage functions
    def _stage_exprs(self, with_lag, with_du, with_may):
        s = self.spec
        nx, nu = s.nx, s.nu
        xk = sx.sym('xk', nx)
        uk = sx.sym('uk', nu)
        xn = sx.sym('xn', nx)
        t, dt = sx.sym('t'), sx.sym('dt')
        sp_, tp_ = sx.sym('sp', s.nsp), sx.sym('tp', s.ntp)
        uo = sx.sym('uo', s.nuold + s.nap)  # stage-0 parameters: previous inputs, then arrival-cost parameters

        def f_at(tt, xx, uu):
            return sx.substitute(s.ode, np.concatenate([[s.t], s.x, s.u, s.sp, [s.dt]]),
                                 np.concatenate([[tt], xx, uu, sp_, [dt]]))

        def lag_at(tt, xx, uu):
            e = sx.substitute(s.lag, np.concatenate([[s.t], s.x, s.u, s.sp, [s.dt]]),
                              np.concatenate([[tt], xx, uu, sp_, [dt]]))
            if with_du and s.nuold > 0:
                # the input-change cost is part of the Lagrange term (modeling.py:243-249); for stages >= 1 the
                # reference passes u_old := u_k itself, so the term vanishes identically there
                e = e + sx.substitute(s.lag_du, np.concatenate([s.u, s.uold]), np.concatenate([uu, uo[:s.nuold]]))
            return e

        zero = sx.Expr.const(0.)
        if s.method == 'collocation':
            deg = s.degree
            B, C, D, tau = polynomial_basis(deg, s.points)
            Xc = [sx.sym(f'X{i}', nx) for i in range(deg)]
            ce = []
            xf = D[0] * xk
            quad = zero
            for i in range(1, deg + 1):
                xp = C[0, i] * xk
                for j in range(deg):
                    xp = xp + C[j + 1, i] * Xc[j]
                ti = t + dt * tau[i - 1]  # modeling.py:1187 evaluates at t + T[i-1]
                fi = f_at(ti, Xc[i - 1], uk)
                ce.append(dt * fi - xp)
                xf = xf + D[i] * Xc[i - 1]
                quad = quad + B[i] * lag_at(ti, Xc[i - 1], uk) * dt
            w = np.concatenate([xk, uk] + Xc + [xn])
            g = np.concatenate(ce + [xn - xf])
        elif s.method == 'erk':
            tab = tableau_for_order(s.order)
            A, b, c = tab['A'], tab['b'], tab['c']
            ks, kq = [], []
            for i in range(s.order):
                ki = np.array([zero] * nx, dtype=object)
                for j in range(i):
                    ki = ki + A[i, j] * ks[j]
                ks.append(f_at(t + dt * c[i], xk + dt * ki, uk))
                kq.append(lag_at(t + dt * c[i], xk + dt * ki, uk))
            xf = xk
            quad = zero
            for i in range(s.order):
                xf = xf + dt * b[i] * ks[i]
                quad = quad + dt * b[i] * kq[i]
            w = np.concatenate([xk, uk, xn])
            g = xn - xf
        elif s.method == 'discrete':
            xf = f_at(t, xk, uk)
            quad = zero
            w = np.concatenate([xk, uk, xn])
            g = xn - xf
        else:
            raise ValueError(s.method)
        if s.gadd is not None:
            g = g - np.concatenate([np.zeros(len(g) - nx), np.asarray(s.gadd) @ uk])
            xf = xf + np.asarray(s.gadd) @ uk
        J = zero
        if with_lag:
            if s.method == 'discrete' or s.objective == 'discrete':
                J = J + lag_at(t, xk, uk)  # mpc.py:1676-1679
            else:
                J = J + quad
        if with_du and s.arr is not None:  # arrival cost on the first node (mhe.py:738-740)
            J = J + sx.substitute(s.arr, np.concatenate([s.x, s.ap]), np.concatenate([xk, uo[s.nuold:]]))
        if with_may:
            J = J + sx.substitute(s.may, np.concatenate([[s.t], s.x, s.tp]), np.concatenate([[t + dt], xf, tp_]))
        return dict(w=w, g=g, J=J, t=t, dt=dt, sp=sp_, tp=tp_, uo=uo)

    def _make_fun(self, e):
        w, g, J = e['w'], e['g'], e['J']
        lam = sx.sym('lam', len(g))
        Jg = sx.jacobian(g, w)
        gJ = sx.gradient(J, w)
        L = J + sx.dot(lam, g) if self.exact_hessian else J
        H = sx.hessian(L, w)
        ins = [('w', w), ('t', e['t']), ('dt', e['dt']), ('sp', e['sp']), ('tp', e['tp']), ('uo', e['uo']),
               ('lam', lam)]
        return sx.Function('stage', ins, [('J', J), ('gJ', gJ), ('g', g), ('Jg', Jg), ('H', H)])

    def _build_stage_functions(self):
        s = self.spec
        # three stage kinds: first (input-change cost), middle, last (carries the Mayer term)
        self._funs = {}
        for kind, (du, may) in {'first': (True, False), 'mid': (False, False), 'last': (False, True),
                                'only': (True, True)}.items():
            for lag in (True, False):
                self._funs[(kind, lag)] = None
        self._fun_args = {'first': (True, False), 'mid': (False, False), 'last': (False, True), 'only': (True, True)}

    def _fun(self, kind, lag):
        if self._funs[(kind, lag)] is None:
            du, may = self._fun_args[kind]
            self._funs[(kind, lag)] = self._make_fun(self._stage_exprs(lag, du, may))
        return self._funs[(kind, lag)]

    def _kinds(self):
        N = self.spec.N
        kinds = []
        for k in range(N):
            if N == 1:
                kind = 'only'
            elif k == 0:
                kind = 'first'
            elif k == N - 1:
                kind = 'last'
            else:
                kind = 'mid'
            kinds.append((kind, k >= self.spec.lag_from))
        return kinds

    # ------------------------------------------------------------------ numeric callback
    def callback(self, time, dts, sps, tp, uold):
        """Returns ``callback(v, lam)`` for ``oracle.ipm.solve_nlp``. ``sps``: (N, nsp); ``dts``: (N,)."""
        s = self.spec
        N = s.N
        kinds = self._kinds()
        times = time + np.concatenate([[0.], np.cumsum(dts)[:-1]])
        groups = {}
        for k, kd in enumerate(kinds):
            groups.setdefault(kd, []).append(k)
        tp = np.asarray(tp, dtype=float).reshape(-1)
        uold = np.asarray(uold, dtype=float).reshape(-1)
        sps = np.asarray(sps, dtype=float).reshape(N, s.nsp)
        ngs, nw = self.n_gs, self.nw

        def cb(v, lam):
            f = 0.
            gf = np.zeros(self.n_v)
            g = np.zeros(self.n_g)
            lam_ = np.zeros(self.n_g) if lam is None else lam
            rows_J, cols_J, vals_J = [], [], []
            rows_H, cols_H, vals_H = [], [], []
            for kd, ks in groups.items():
                fun = self._fun(*kd)
                ks = np.asarray(ks)
                nb = len(ks)
                W = v[self.w_ind[ks]].T  # (nw, nb)
                lam_k = lam_.reshape(N, ngs)[ks].T
                Jv, gJ, gv, Jg, H = fun(W, times[ks][None, :], dts[ks][None, :], sps[ks].T,
                                        np.tile(tp[:, None], (1, nb)), np.tile(uold[:, None], (1, nb)), lam_k)
                Jv = np.broadcast_to(Jv, (nb,)) if np.ndim(Jv) else np.full(nb, Jv)
                gJ = np.broadcast_to(gJ.reshape(nw, -1), (nw, nb))
                gv = np.broadcast_to(gv.reshape(ngs, -1), (ngs, nb))
                Jg = np.broadcast_to(Jg.reshape(ngs, nw, -1), (ngs, nw, nb))
                H = np.broadcast_to(H.reshape(nw, nw, -1), (nw, nw, nb))
                f += Jv.sum()
                for i, k in enumerate(ks):
                    wi = self.w_ind[k]
                    np.add.at(gf, wi, gJ[:, i])
                    g[k * ngs:(k + 1) * ngs] = gv[:, i]
                    r = np.repeat(np.arange(k * ngs, (k + 1) * ngs), nw)
                    rows_J.append(r)
                    cols_J.append(np.tile(wi, ngs))
                    vals_J.append(Jg[:, :, i].reshape(-1))
                    rows_H.append(np.repeat(wi, nw))
                    cols_H.append(np.tile(wi, nw))
                    vals_H.append(H[:, :, i].reshape(-1))
            Jm = sp.csr_matrix((np.concatenate(vals_J), (np.concatenate(rows_J), np.concatenate(cols_J))),
                               shape=(self.n_g, self.n_v))
            Hm = sp.csr_matrix((np.concatenate(vals_H), (np.concatenate(rows_H), np.concatenate(cols_H))),
                               shape=(self.n_v, self.n_v))
            return f, gf, g, Jm, Hm
        return cb

    def bounds_and_guess(self, x0=None):
        s = self.spec
        lb = np.full(self.n_v, -np.inf)
        ub = np.full(self.n_v, np.inf)
        v0 = np.zeros(self.n_v)
        for k in range(s.N + 1):
            lb[self.x_ind[k]], ub[self.x_ind[k]], v0[self.x_ind[k]] = s.x_lb, s.x_ub, s.x_guess
        for k in range(s.N):
            lb[self.u_ind[k]], ub[self.u_ind[k]], v0[self.u_ind[k]] = s.u_lb, s.u_ub, s.u_guess
            if self.ncol:
                d = s.degree
                lb[self.c_ind[k]], ub[self.c_ind[k]] = np.tile(s.x_lb, d), np.tile(s.x_ub, d)
                v0[self.c_ind[k]] = np.tile(s.x_guess, d)
        if x0 is not None and s.fixed_x0:
            lb[self.x_ind[0]] = ub[self.x_ind[0]] = x0  # mpc.py:801-802
        return lb, ub, v0

    def solve(self, x0, v0=None, time=0., dts=None, sps=None, tp=None, uold=None, dt=None, ap=None, **ipm_opts):
        """One ``NMPC.optimize`` call (mpc.py:744-857) in scaled coordinates. Returns (IpmResult, x_pred, u_pred)."""
        s = self.spec
        if dts is None:
            dts = np.full(s.N, float(dt))
        if sps is None:
            sps = np.zeros((s.N, s.nsp))
        if tp is None:
            tp = np.zeros(s.ntp)
        if uold is None:
            uold = np.zeros(s.nuold)
        if s.nap:
            uold = np.concatenate([np.asarray(uold, dtype=float).reshape(-1), np.asarray(ap, dtype=float).reshape(-1)])
        lb, ub, vg = self.bounds_and_guess(x0)
        if v0 is None:
            v0 = vg
        res = solve_nlp(self.callback(time, np.asarray(dts, dtype=float), sps, tp, uold), v0, lb, ub,
                        psd_hessian=not self.exact_hessian, **ipm_opts)
        v = res.x
        x_pred = np.stack([v[i] for i in self.x_ind], axis=1)
        u_pred = np.stack([v[i] for i in self.u_ind], axis=1)
        return res, x_pred, u_pred
