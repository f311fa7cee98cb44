"""Canned models for the benchmark configurations (BASELINE.json ``configs``).

The reference ships a few canned models in ``hilo_mpc/library/models.py:30-296``; the ``nx=4, nu=2`` CSTR named by
BASELINE.json is *not* among them — it is the Klatt-Engell / van-de-Vusse reactor used
by the HILO-MPC paper's examples.  It is defined here through the public ``Model`` API.
"""
from __future__ import annotations

import numpy as np

from . import symbolic as sx
from .model import Model

# Klatt-Engell CSTR constants
CSTR = dict(k10=1.287e12, k20=1.287e12, k30=9.043e9, E1=9758.3, E2=9758.3, E3=8560., dH_AB=4.2, dH_BC=-11.0,
            dH_AD=-41.85, rho=0.9342, cp=3.01, cpK=2.0, A_R=0.215, V_R=10.01, m_K=5.0, T_in=130.0, k_w=4032.,
            cA0=5.1)


def cstr_rhs(x, u, c=CSTR, exp=None, side_reaction=None, k1_scale=1.):
    """Right-hand side for symbolic (``exp=sx.exp``) or NumPy (``exp=np.exp``) arguments; states
    ``[c_A, c_B, T_R, T_K]``, inputs ``[F, Q_dot]``.  ``side_reaction(c_A, T_R)``, if given, replaces the rate term
    ``k_3(T_R) c_A^2`` (BASELINE config 3: that term learned by a Gaussian process); ``k1_scale`` multiplies the
    pre-exponential factor ``k_10`` (BASELINE config 4: the parameter the MHE estimates)."""
    exp = exp or sx.exp
    cA, cB, TR, TK = x[0], x[1], x[2], x[3]
    F, Qd = u[0], u[1]
    k1 = k1_scale * c['k10'] * exp(-c['E1'] / (TR + 273.15))
    k2 = c['k20'] * exp(-c['E2'] / (TR + 273.15))
    k3 = c['k30'] * exp(-c['E3'] / (TR + 273.15))
    r3 = k3 * cA * cA if side_reaction is None else side_reaction(cA, TR)
    dcA = F * (c['cA0'] - cA) - k1 * cA - r3
    dcB = -F * cB + k1 * cA - k2 * cB
    dTR = F * (c['T_in'] - TR) + c['k_w'] * c['A_R'] / (c['rho'] * c['cp'] * c['V_R']) * (TK - TR) \
        - (k1 * cA * c['dH_AB'] + k2 * cB * c['dH_BC'] + r3 * c['dH_AD']) / (c['rho'] * c['cp'])
    dTK = (Qd + c['k_w'] * c['A_R'] * (TR - TK)) / (c['m_K'] * c['cpK'])
    return [dcA, dcB, dTR, dTK]


def cstr_side_reaction(cA, TR, c=CSTR, exp=np.exp):
    """The true rate term ``k_3(T_R) c_A^2`` (training target of the GP of config 3)."""
    return c['k30'] * exp(-c['E3'] / (TR + 273.15)) * cA * cA


def cstr_model(dt: float = 0.005, side_reaction=None, name: str = 'cstr_klatt_engell', with_parameter: bool = False) -> Model:
    """``with_parameter``: ``k1_scale`` becomes a model parameter (config 4: estimated by the MHE)."""
    model = Model(name=name, time_unit='h')
    x = model.set_dynamical_states(['c_A', 'c_B', 'T_R', 'T_K'])
    u = model.set_inputs(['F', 'Q_dot'])
    k1s = model.set_parameters(['k1_scale'])[0] if with_parameter else 1.
    model.set_measurements(['y_cB', 'y_TR'])
    model.set_measurement_equations([x[1], x[2]])
    model.set_dynamical_equations(cstr_rhs(x, u, side_reaction=side_reaction, k1_scale=k1s))
    model.setup(dt=dt)
    return model


def cstr_steady_state(u_ss=(14.19, -6000.), x_init=(1.3, 0.9, 128., 121.)):
    """Steady state for a constant input by Newton's method on the NumPy right-hand side (host-side problem setup)."""
    x = np.array(x_init, dtype=float)
    u = np.array(u_ss, dtype=float)
    for _ in range(100):
        f = np.array(cstr_rhs(x, u, exp=np.exp), dtype=float)
        J = np.zeros((4, 4))
        for j in range(4):
            xp = x.astype(complex)
            xp[j] += 1e-30j
            J[:, j] = np.array(cstr_rhs(xp, u, exp=np.exp)).imag / 1e-30
        step = np.linalg.solve(J, -f)
        x = x + step
        if np.abs(step).max() < 1e-13:
            break
    return x, u


CSTR_BOUNDS = dict(x_lb=[0.1, 0.1, 50., 50.], x_ub=[2., 2., 140., 140.], u_lb=[5., -8500.], u_ub=[100., 0.])
CSTR_SCALING = dict(x_scaling=[1., 1., 100., 100.], u_scaling=[10., 1000.])


def bicycle_rhs(x, u, lr=1.4, lf=1.8, sin=None, cos=None, tan=None, atan=None):
    """Kinematic bicycle of docs/docsource/examples/mpc_formula1.ipynb: states ``[px, py, v, phi]``, inputs
    ``[a, delta]``."""
    sin, cos, tan, atan = sin or sx.sin, cos or sx.cos, tan or sx.tan, atan or sx.atan
    beta = atan(lr / (lr + lf) * tan(u[1]))
    return [x[2] * cos(x[3] + beta), x[2] * sin(x[3] + beta), u[0], x[2] / lr * sin(beta)]
