"""Builds, for sm_100a, every generated module the tests / smoke / bench use, so the binaries travel with the tree."""
from __future__ import annotations


def build_all():
    from .estgen import build_est_module
    from .library import cstr_model
    from .model import Model
    paths = []
    m = Model()
    m.set_equations(equations="""
        dx_1/dt = -k_1*x_1(t) + u(k)
        dx_2/dt = k_1*x_1(t) - k_2*x_2(t)
        y(k) = x_2(t)
        """)
    m.discretize('erk', order=1, inplace=True)
    m.setup(dt=1.)
    paths.append(build_est_module(m))
    s = Model(discrete=True)
    s.set_dynamical_states('x')
    s.set_measurements('y')
    s.set_dynamical_equations('x/2 + 25*dt*x/(1 + x^2)')
    s.set_measurement_equations('x^2/20')
    s.setup(dt=1.)
    paths.append(build_est_module(s))
    paths.append(build_est_module(cstr_model(dt=0.005)))
    return paths
