"""Builds, for sm_100a, every generated module the tests / smoke / bench use, so the binaries travel with the tree."""
from __future__ import annotations

import numpy as np


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
    bio = Model()   # 3-state bioreactor of the UKF sigma-point golden (tests/test_KFs.py:695-712)
    bio.set_equations(equations="""
        dT/dt = alpha*(T_amb - T(t))
        dcB/dt = r - D(k)*cB(t)
        dcS/dt = -r/Y - D(k)*cS(t)

        y_0(k) = T(t)
        y_1(k) = cB(t)

        r = (mu_0 + mu_1*T(t))*cS(t)*cB(t)/(K + cS(t))
        """)
    bio.setup(dt=.1)
    paths.append(build_est_module(bio))
    # LMPC double integrator (BASELINE config 1 / tests/test_LMPC.py:8-33) and its plant model
    from .lmpc import LMPC
    di = Model(discrete=True)
    di.A = np.array([[1, di.dt], [0, 1]])
    di.B = np.array([[di.dt ** 2 / 2], [di.dt]])
    di.setup(dt=.5)
    paths.append(build_est_module(di))
    mpc = LMPC(di)
    mpc.Q, mpc.R, mpc.horizon = np.eye(2), 1, 10
    mpc.set_box_constraints(x_lb=[-5, -5], x_ub=[5, 5], u_lb=[-1], u_ub=[1])
    mpc.setup()
    paths.append(mpc.engine.module_path)
    # MHE test problem and BASELINE config 5 (MLP dynamics + path following)
    from .mhe import MHE
    rm = Model(name='reaction')
    x = rm.set_dynamical_states(['ca', 'cb'])
    u = rm.set_inputs(['f'])
    p = rm.set_parameters(['k'])
    rm.set_measurements(['y'])
    rm.set_measurement_equations([x[1]])
    rm.set_dynamical_equations([u[0] * (1. - x[0]) - p[0] * x[0], p[0] * x[0] - u[0] * x[1]])
    rm.setup(dt=.25)
    mhe = MHE(rm)
    mhe.horizon = 6
    mhe.quad_arrival_cost.add_states(weights=[10., 10.], guess=[.5, .2])
    mhe.quad_arrival_cost.add_parameters(weights=[1.], guess=[1.])
    mhe.quad_stage_cost.add_measurements(weights=[100.])
    mhe.quad_stage_cost.add_state_noise(weights=[20., 20.])
    mhe.set_box_constraints(x_lb=[0., 0.], x_ub=[1., 1.], p_lb=[.1], p_ub=[5.])
    mhe.setup()
    paths.append(mhe.engine.module_path)
    from .examples import cstr_mhe, robot_nn_nmpc
    paths.append(robot_nn_nmpc(N=5, hidden=8).engine.module_path)
    # BASELINE configs 4 and 5 at their stated problem sizes
    paths.append(cstr_mhe(N=20).engine.module_path)
    paths.append(robot_nn_nmpc(N=30, hidden=32).engine.module_path)
    return paths
