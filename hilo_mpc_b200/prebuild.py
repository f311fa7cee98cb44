"""Builds, for sm_100a, every generated module the tests / smoke / bench use, so the binaries travel with the tree."""
from __future__ import annotations

import numpy as np


def _est_first_order():
    from .estgen import build_est_module
    from .model import Model
    m = Model()
    m.set_equations(equations="""
        dx_1/dt = -k_1*x_1(t) + u(k)
        dx_2/dt = k_1*x_1(t) - k_2*x_2(t)
        y(k) = x_2(t)
        """)
    m.discretize('erk', order=1, inplace=True)
    m.setup(dt=1.)
    return build_est_module(m)


def _est_scalar():
    from .estgen import build_est_module
    from .model import Model
    s = Model(discrete=True)
    s.set_dynamical_states('x')
    s.set_measurements('y')
    s.set_dynamical_equations('x/2 + 25*dt*x/(1 + x^2)')
    s.set_measurement_equations('x^2/20')
    s.setup(dt=1.)
    return build_est_module(s)


def _est_cstr():
    from .estgen import build_est_module
    from .library import cstr_model
    return build_est_module(cstr_model(dt=0.005))


def _est_cstr_k1():
    from .estgen import build_est_module
    from .library import cstr_model
    return build_est_module(cstr_model(dt=0.005, name='cstr_k1', with_parameter=True))


def _est_bioreactor():
    from .estgen import build_est_module
    from .model import Model
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
    return build_est_module(bio)


def _double_integrator():
    from .model import Model
    di = Model(discrete=True)
    di.A = np.array([[1, di.dt], [0, 1]])
    di.B = np.array([[di.dt ** 2 / 2], [di.dt]])
    di.setup(dt=.5)
    return di


def _est_double_integrator():
    from .estgen import build_est_module
    return build_est_module(_double_integrator())


def _lmpc():
    # LMPC double integrator (BASELINE config 1 / tests/test_LMPC.py:8-33)
    from .lmpc import LMPC
    mpc = LMPC(_double_integrator())
    mpc.Q, mpc.R, mpc.horizon = np.eye(2), 1, 10
    mpc.set_box_constraints(x_lb=[-5, -5], x_ub=[5, 5], u_lb=[-1], u_ub=[1])
    mpc.setup(solver='riccati')   # the stage-structured path; the condensed dense-QP path needs no generated module
    return mpc.engine.module_path


def _mhe_reaction():
    from .mhe import MHE
    from .model import Model
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
    return mhe.engine.module_path


def _robot_small():
    from .examples import robot_nn_nmpc
    return robot_nn_nmpc(N=5, hidden=8).engine.module_path


def _mhe_cstr():   # BASELINE config 4 at its stated problem size
    from .examples import cstr_mhe
    return cstr_mhe(N=20).engine.module_path


def _robot_full():   # BASELINE config 5 at its stated problem size
    from .examples import robot_nn_nmpc
    return robot_nn_nmpc(N=30, hidden=32).engine.module_path


def _cstr(N, method):
    from .examples import cstr_nmpc
    return cstr_nmpc(N=N, integration_method=method).engine.module_path


def jobs():
    """Callables that each build one generated module and return its path (run concurrently by ``build_all``)."""
    from functools import partial
    out = [_est_first_order, _est_scalar, _est_cstr, _est_cstr_k1, _est_bioreactor, _est_double_integrator, _lmpc,
           _mhe_reaction, _robot_small, _mhe_cstr, _robot_full]
    for N, method in ((50, 'collocation'), (10, 'collocation'), (10, 'rk4')):
        out.append(partial(_cstr, N, method))
    return out


def _run(job):
    return job()


def build_all(workers: int = 6):
    """Every module the tests, the smoke test and the bench need.  One PROCESS per job: the symbolic layer interns its
    expression nodes in process-wide tables, which must not be shared between concurrently transcribed problems."""
    import multiprocessing as mp
    with mp.get_context('spawn').Pool(workers) as pool:
        return pool.map(_run, jobs(), chunksize=1)
