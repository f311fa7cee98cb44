"""Oracle (test infrastructure): discretisation constants.

Follows ``hilo_mpc/util/modeling.py:1092-1127`` (``RungeKutta._construct_polynomial_basis``) and the explicit
Butcher tableaus at ``hilo_mpc/util/modeling.py:1008-1085``.  ``ca.collocation_points`` (CasADi, not vendored) returns
the Gauss-Legendre / Radau-IIA abscissae on (0, 1]; they are recomputed here from the Legendre polynomials.
"""
import numpy as np
from numpy.polynomial import legendre as _leg


def collocation_points(degree: int, method: str = 'radau'):
    """Roots on [0, 1]: 'legendre' -> zeros of P_d(2t-1); 'radau' -> zeros of P_{d-1}(2t-1) - P_d(2t-1) (Radau IIA,
    right end point included). Degree 3 radau: (4-sqrt6)/10, (4+sqrt6)/10, 1."""
    d = int(degree)
    if method == 'legendre':
        c = np.zeros(d + 1)
        c[d] = 1.
        r = _leg.legroots(c)
    elif method == 'radau':
        c = np.zeros(d + 1)
        c[d - 1] = 1.
        c[d] = -1.
        r = _leg.legroots(c)
    else:
        raise ValueError(f"Unknown collocation points '{method}'")
    r = np.sort((np.real(r) + 1.) / 2.)
    if method == 'radau':
        r[-1] = 1.
    return [float(v) for v in r]


def polynomial_basis(degree: int, method: str = 'radau'):
    """Returns ``B, C, D, tau`` exactly as modeling.py:1103-1127 builds them with ``np.poly1d`` (``T = h * tau``)."""
    B = np.zeros(degree + 1)
    C = np.zeros((degree + 1, degree + 1))
    D = np.zeros(degree + 1)
    tau = [0.] + collocation_points(degree, method)
    for i in range(degree + 1):
        L = np.poly1d([1.])
        for j in range(degree + 1):
            if j != i:
                L *= np.poly1d([1., -tau[j]]) / (tau[i] - tau[j])
        D[i] = L(1.)
        Ldot = np.polyder(L)
        for j in range(degree + 1):
            C[i, j] = Ldot(tau[j])
        Lint = np.polyint(L)
        B[i] = Lint(1.)
    return B, C, D, np.array(tau)


# modeling.py:1008-1085 (only the tableaus reachable from the NMPC/MHE options: optimizer.py:1452-1456 maps
# 'rk4' -> order 4 'classic'; 'erk' order 1..4 -> forward_euler / midpoint / kutta / classic, modeling.py:1234-1245)
EXPLICIT_METHODS = {
    'forward_euler': {'A': np.array([[0.]]), 'b': np.array([1.]), 'c': np.array([0.])},
    'midpoint': {'A': np.array([[0., 0.], [.5, 0.]]), 'b': np.array([0., 1.]), 'c': np.array([0., .5])},
    'kutta': {'A': np.array([[0., 0., 0.], [.5, 0., 0.], [-1., 2., 0.]]), 'b': np.array([1 / 6, 2 / 3, 1 / 6]),
              'c': np.array([0., .5, 1.])},
    'classic': {'A': np.array([[0., 0., 0., 0.], [.5, 0., 0., 0.], [0., .5, 0., 0.], [0., 0., 1., 0.]]),
                'b': np.array([1 / 6, 1 / 3, 1 / 3, 1 / 6]), 'c': np.array([0., .5, .5, 1.])},
}


def tableau_for_order(order: int):
    return EXPLICIT_METHODS[{1: 'forward_euler', 2: 'midpoint', 3: 'kutta', 4: 'classic'}[int(order)]]
