"""Oracle (test infrastructure), SECOND restatement: flat NLPs of two concrete problems written straight from the
reference's transcription in plain NumPy - no import of ``hilo_mpc_b200``, no symbolic layer,
hand-coded right-hand sides, derivatives by complex-step differentiation.

Purpose: ``oracle.nmpc.FlatNlp`` consumes the ``OcpSpec`` that the product's ``NMPC.setup()`` / ``MHE.setup()`` build and
shares the product's symbolic differentiation, so an error in the product's front end (scaling, weights, reference
substitution, a wrong derivative) would be invisible to it.  The classes below share nothing with the product; the tests
(tests/test_independent_nlp.py) require that they and ``FlatNlp`` agree on ``f, g, grad f, jac g`` at random points and on
the minimiser.

* ``CstrNlp``  - BASELINE config 2: reference-tracking NMPC of the Klatt-Engell CSTR, Radau collocation
  (mpc.py:1455-1728: ``v = [x_0..x_N | u_0..u_{N-1} | X_k,1..deg per stage]``, per stage ``deg*nx`` collocation equations
  ``dt f(X_i) - sum_j C[j,i] X_j = 0`` then the continuity ``x_{k+1} - sum_j D[j] X_j = 0``, objective
  ``sum_k sum_i B[i] l(X_i, u_k) dt + E(x_f of the last interval)``; modeling.py:1180-1195), all in SCALED coordinates
  (base.py:1583-1591, modeling.py:310).
* ``ReactionMheNlp`` - the MHE test problem (2 concentrations + 1 rate constant) in the REFERENCE's own variable layout
  (mhe.py:596-743: ``v = [p | x_0..x_N | w_0..w_{N-1} | collocation]`` with ONE parameter variable), whereas the engine
  carries the parameter as an extra state with trivial dynamics: the minimisers must coincide.
"""
import numpy as np

from .collocation import polynomial_basis

# Klatt-Engell CSTR, SURVEY.md Appendix C
K10 = K20 = 1.287e12
K30 = 9.043e9
E1 = E2 = 9758.3
E3 = 8560.
DH_AB, DH_BC, DH_AD = 4.2, -11.0, -41.85
RHO, CP, CPK = 0.9342, 3.01, 2.0
A_R, V_R, M_K = 0.215, 10.01, 5.0
T_IN, K_W, CA0 = 130.0, 4032., 5.1
X_SCALE = np.array([1., 1., 100., 100.])
U_SCALE = np.array([10., 1000.])
X_LB, X_UB = np.array([.1, .1, 50., 50.]), np.array([2., 2., 140., 140.])
U_LB, U_UB = np.array([5., -8500.]), np.array([100., 0.])
Q_W = np.array([1., 10., 1., 1.])     # weights on the scaled states (mpc.py:207-263)
R_W = np.array([.1, .1])
U_SS = np.array([14.19, -6000.])


def cstr_ode(x, u):
    """Physical right-hand side; ``x (4, ...)``, ``u (2, ...)``; works for complex arguments."""
    cA, cB, TR, TK = x
    F, Qd = u
    k1 = K10 * np.exp(-E1 / (TR + 273.15))
    k2 = K20 * np.exp(-E2 / (TR + 273.15))
    k3 = K30 * np.exp(-E3 / (TR + 273.15))
    return np.array([
        F * (CA0 - cA) - k1 * cA - k3 * cA * cA,
        -F * cB + k1 * cA - k2 * cB,
        F * (T_IN - TR) + K_W * A_R / (RHO * CP * V_R) * (TK - TR)
        - (k1 * cA * DH_AB + k2 * cB * DH_BC + k3 * cA * cA * DH_AD) / (RHO * CP),
        (Qd + K_W * A_R * (TR - TK)) / (M_K * CPK)])


def cstr_steady_state(x_init=(1.3, .9, 128., 121.)):
    x = np.array(x_init, dtype=float)
    for _ in range(100):
        f = cstr_ode(x, U_SS)
        J = np.zeros((4, 4))
        for j in range(4):
            xp = x.astype(complex)
            xp[j] += 1e-30j
            J[:, j] = cstr_ode(xp, U_SS).imag / 1e-30
        dx = np.linalg.solve(J, -f)
        x = x + dx
        if np.abs(dx).max() < 1e-13:
            break
    return x


def _complex_step_jacobian(fun, v, m):
    n = v.size
    J = np.zeros((m, n))
    vc = v.astype(complex)
    for j in range(n):
        vc[j] += 1e-30j
        J[:, j] = np.atleast_1d(fun(vc)).imag / 1e-30
        vc[j] = v[j]
    return J


class CstrNlp:
    def __init__(self, N=10, dt=.005, degree=3, points='radau'):
        self.N, self.dt, self.deg = N, dt, degree
        self.nx, self.nu = 4, 2
        self.B, self.C, self.D, self.tau = polynomial_basis(degree, points)
        self.x_ref = cstr_steady_state() / X_SCALE
        self.u_ref = U_SS / U_SCALE
        nx, nu, nc = 4, 2, degree * 4
        self.n_v = (N + 1) * nx + N * nu + N * nc
        self.n_g = N * (nc + nx)
        self.ix = lambda k: slice(k * nx, (k + 1) * nx)
        o1 = (N + 1) * nx
        self.iu = lambda k: slice(o1 + k * nu, o1 + (k + 1) * nu)
        o2 = o1 + N * nu
        self.ic = lambda k, i: slice(o2 + k * nc + i * nx, o2 + k * nc + (i + 1) * nx)   # X_{k,i+1}

    def _scaled_ode(self, xs, us):
        return cstr_ode(xs * X_SCALE, us * U_SCALE) / X_SCALE

    def _lagrange(self, xs, us):
        dx, du = xs - self.x_ref, us - self.u_ref
        return np.sum(Q_W * dx * dx) + np.sum(R_W * du * du)

    def objective(self, v):
        J = 0.
        B, D, deg = self.B, self.D, self.deg
        for k in range(self.N):
            for i in range(1, deg + 1):
                J = J + B[i] * self._lagrange(v[self.ic(k, i - 1)], v[self.iu(k)]) * self.dt
        k = self.N - 1
        xf = D[0] * v[self.ix(k)]
        for i in range(1, deg + 1):
            xf = xf + D[i] * v[self.ic(k, i - 1)]
        dxf = xf - self.x_ref
        return J + np.sum(Q_W * dxf * dxf)      # Mayer term on x_f of the last interval (mpc.py:1681-1682)

    def constraints(self, v):
        C, D, deg, dt = self.C, self.D, self.deg, self.dt
        out = []
        for k in range(self.N):
            xk, uk = v[self.ix(k)], v[self.iu(k)]
            X = [xk] + [v[self.ic(k, i)] for i in range(deg)]
            for i in range(1, deg + 1):
                xp = sum(C[j, i] * X[j] for j in range(deg + 1))
                out.append(dt * self._scaled_ode(X[i], uk) - xp)
            xf = sum(D[j] * X[j] for j in range(deg + 1))
            out.append(v[self.ix(k + 1)] - xf)
        return np.concatenate(out)

    def gradient(self, v):
        return _complex_step_jacobian(self.objective, v, 1)[0]

    def jacobian(self, v):
        return _complex_step_jacobian(self.constraints, v, self.n_g)

    def bounds(self, x0_scaled):
        lb, ub = np.zeros(self.n_v), np.zeros(self.n_v)
        for k in range(self.N + 1):
            lb[self.ix(k)], ub[self.ix(k)] = X_LB / X_SCALE, X_UB / X_SCALE
        for k in range(self.N):
            lb[self.iu(k)], ub[self.iu(k)] = U_LB / U_SCALE, U_UB / U_SCALE
            for i in range(self.deg):
                lb[self.ic(k, i)], ub[self.ic(k, i)] = X_LB / X_SCALE, X_UB / X_SCALE
        lb[self.ix(0)] = ub[self.ix(0)] = x0_scaled            # mpc.py:801-802
        return lb, ub

    def guess(self):
        v = np.zeros(self.n_v)
        for k in range(self.N + 1):
            v[self.ix(k)] = self.x_ref
        for k in range(self.N):
            v[self.iu(k)] = self.u_ref
            for i in range(self.deg):
                v[self.ic(k, i)] = self.x_ref
        return v


class ReactionMheNlp:
    """ca' = f (1 - ca) - k ca,  cb' = k ca - f cb,  y = cb; horizon N, Radau-3; arrival weights on (x_0, k), stage weights
    on the measurement residual and the state noise; the reference's layout ``[k | x_0..x_N | w_0..w_{N-1} | colloc]``."""

    def __init__(self, N, dt, y_meas, u_meas, x_arr, p_arr, w_arr_x=(10., 10.), w_arr_p=1., w_meas=100.,
                 w_noise=(20., 20.), degree=3):
        self.N, self.dt, self.deg = N, dt, degree
        self.y, self.u = np.asarray(y_meas, dtype=float).reshape(N), np.asarray(u_meas, dtype=float).reshape(N)
        self.x_arr, self.p_arr = np.asarray(x_arr, dtype=float), float(p_arr)
        self.wx, self.wp, self.wy, self.ww = np.asarray(w_arr_x), w_arr_p, w_meas, np.asarray(w_noise)
        self.B, self.C, self.D, self.tau = polynomial_basis(degree, 'radau')
        nx, nc = 2, degree * 2
        self.n_v = 1 + (N + 1) * nx + N * nx + N * nc
        self.n_g = N * (nc + nx)
        self.ix = lambda k: slice(1 + k * nx, 1 + (k + 1) * nx)
        o1 = 1 + (N + 1) * nx
        self.iw = lambda k: slice(o1 + k * nx, o1 + (k + 1) * nx)
        o2 = o1 + N * nx
        self.ic = lambda k, i: slice(o2 + k * nc + i * nx, o2 + k * nc + (i + 1) * nx)

    @staticmethod
    def ode(x, f, k):
        return np.array([f * (1. - x[0]) - k * x[0], k * x[0] - f * x[1]])

    def objective(self, v):
        k = v[0]
        x0 = v[self.ix(0)]
        J = np.sum(self.wx * (x0 - self.x_arr) ** 2) + self.wp * (k - self.p_arr) ** 2      # ii == 0: arrival only
        for ii in range(1, self.N):                                                          # mhe.py:738-743
            x, w = v[self.ix(ii)], v[self.iw(ii)]
            J = J + self.wy * (x[1] - self.y[ii]) ** 2 + np.sum(self.ww * w * w)
        return J

    def constraints(self, v):
        C, D, deg, dt = self.C, self.D, self.deg, self.dt
        k = v[0]
        out = []
        for ii in range(self.N):
            xk = v[self.ix(ii)]
            X = [xk] + [v[self.ic(ii, i)] for i in range(deg)]
            for i in range(1, deg + 1):
                xp = sum(C[j, i] * X[j] for j in range(deg + 1))
                out.append(dt * self.ode(X[i], self.u[ii], k) - xp)
            xf = sum(D[j] * X[j] for j in range(deg + 1)) + v[self.iw(ii)]
            out.append(v[self.ix(ii + 1)] - xf)
        return np.concatenate(out)

    def gradient(self, v):
        return _complex_step_jacobian(self.objective, v, 1)[0]

    def jacobian(self, v):
        return _complex_step_jacobian(self.constraints, v, self.n_g)
