"""Benchmark / test problem builders (BASELINE.json ``configs``), written against the public API only."""
from __future__ import annotations

import numpy as np

from .library import CSTR_BOUNDS, CSTR_SCALING, cstr_model, cstr_steady_state
from .nmpc import NMPC

# weights act on the scaled variables (mpc.py:207-263): c/1, T/100, F/10, Q/1000
CSTR_Q = [1., 10., 1., 1.]
CSTR_R = [.1, .1]


def cstr_nmpc(N: int = 50, dt: float = 0.005, integration_method: str = 'collocation', solver_options=None,
              setup: bool = True) -> NMPC:
    """Config 2: reference-tracking NMPC for the Klatt-Engell CSTR (nx=4, nu=2), default Radau-3 collocation."""
    model = cstr_model(dt)
    x_ss, u_ss = cstr_steady_state()
    nmpc = NMPC(model)
    nmpc.horizon = N
    nmpc.quad_stage_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
    nmpc.quad_stage_cost.add_inputs(names=['F', 'Q_dot'], ref=list(u_ss), weights=CSTR_R)
    nmpc.quad_terminal_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
    nmpc.set_box_constraints(**CSTR_BOUNDS)
    nmpc.set_scaling(**CSTR_SCALING)
    nmpc.set_initial_guess(x_guess=x_ss, u_guess=u_ss)
    nmpc.set_nlp_options(integration_method=integration_method)
    if solver_options:
        nmpc.set_solver_opts(solver_options)
    if setup:
        nmpc.setup()
    return nmpc


def cstr_initial_states(B: int, seed: int = 0) -> np.ndarray:
    """``x0 = x_ss + diag([.2,.2,5,5]) U(-1,1)`` with ``np.random.default_rng(seed)`` (cfg 2)."""
    x_ss, _ = cstr_steady_state()
    rng = np.random.default_rng(seed)
    return x_ss + np.array([.2, .2, 5., 5.]) * rng.uniform(-1., 1., size=(B, 4))
