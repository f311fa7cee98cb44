"""Discretisation constants used by the code generator (host side, computed once at ``setup()``).

Same definitions as the reference (``hilo_mpc/util/modeling.py:1092-1127`` Lagrange basis on ``tau = [0, c_1..c_d]``
with ``D_i = L_i(1)``, ``C_ij = L_i'(tau_j)``, ``B_i = int_0^1 L_i``; explicit tableaus ``modeling.py:1008-1085``) but
computed through ``numpy.polynomial`` in the monomial basis of the shifted Legendre / Radau nodes.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import Polynomial, legendre


def collocation_nodes(degree: int, points: str = 'radau') -> np.ndarray:
    d = int(degree)
    if d < 1:
        raise ValueError("Collocation degree must be at least 1")
    if points == 'legendre':
        nodes = legendre.leggauss(d)[0]
    elif points == 'radau':
        # Radau IIA: zeros of P_{d-1} - P_d on (-1, 1], i.e. the right end point is a node
        coef = np.zeros(d + 1)
        coef[d - 1], coef[d] = 1., -1.
        nodes = np.sort(np.real(legendre.legroots(coef)))
        nodes[-1] = 1.
    else:
        raise ValueError(f"Collocation points '{points}' not supported; choose 'radau' or 'legendre'")
    return (np.asarray(nodes) + 1.) / 2.


def collocation_coefficients(degree: int, points: str = 'radau'):
    """Returns ``B (d+1), C (d+1, d+1), D (d+1), tau (d+1)``."""
    tau = np.concatenate([[0.], collocation_nodes(degree, points)])
    n = degree + 1
    B = np.zeros(n)
    C = np.zeros((n, n))
    D = np.zeros(n)
    for i in range(n):
        L = Polynomial([1.])
        for j in range(n):
            if j != i:
                L = L * Polynomial([-tau[j], 1.]) / (tau[i] - tau[j])
        D[i] = L(1.)
        dL = L.deriv()
        C[i, :] = dL(tau)
        B[i] = L.integ()(1.)
    return B, C, D, tau


_TABLEAUS = {
    1: (np.array([[0.]]), np.array([1.]), np.array([0.])),
    2: (np.array([[0., 0.], [.5, 0.]]), np.array([0., 1.]), np.array([0., .5])),
    3: (np.array([[0., 0., 0.], [.5, 0., 0.], [-1., 2., 0.]]), np.array([1 / 6, 2 / 3, 1 / 6]),
        np.array([0., .5, 1.])),
    4: (np.array([[0., 0., 0., 0.], [.5, 0., 0., 0.], [0., .5, 0., 0.], [0., 0., 1., 0.]]),
        np.array([1 / 6, 1 / 3, 1 / 3, 1 / 6]), np.array([0., .5, .5, 1.])),
}


def tableau_for_order(order: int):
    """forward Euler / midpoint / Kutta-3 / classic RK4 — the defaults ``modeling.py:1234-1245`` picks per order."""
    if int(order) not in _TABLEAUS:
        raise NotImplementedError(f"Explicit Runge-Kutta discretization for order {order} is not yet implemented.")
    A, b, c = _TABLEAUS[int(order)]
    return {'A': A, 'b': b, 'c': c}
