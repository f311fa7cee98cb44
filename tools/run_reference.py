"""Generate golden vectors with the REAL reference (HILO-MPC on CasADi + IPOPT / qpOASES) for the paths the reference's
own tests do not pin numerically: NMPC, LMPC ("parity unpinned").

This container has no CasADi (no network), so the script cannot run here; it is the hook a maintainer with a working
``pip install hilo-mpc casadi`` uses once:

    python tools/run_reference.py            # writes tests/golden/reference_nmpc_cstr.npz, reference_lmpc_di.npz

``tests/test_reference_goldens.py`` then compares the oracle (CPU) and the engine (GPU) with these files and is skipped as
long as they do not exist.  The problems are built with the same public calls as ``hilo_mpc_b200/examples.py`` (the API is
drop-in), on the same seeded initial states (``cstr_initial_states``), with IPOPT at ``tol = 1e-10``.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    try:
        import casadi as ca
        from hilo_mpc import LMPC, NMPC, Model
    except ImportError as exc:  # noqa: BLE001
        raise SystemExit(f"The reference is not importable here ({exc}); run this on a box with hilo-mpc + casadi.")
    from hilo_mpc_b200.examples import CSTR_Q, CSTR_R, cstr_initial_states
    from hilo_mpc_b200.library import CSTR_BOUNDS, CSTR_SCALING, cstr_rhs, cstr_steady_state
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)

    # ---- config 2 at a size IPOPT solves in seconds per instance: CSTR reference tracking, N = 10, Radau-3 collocation
    N, dt, B = 10, 0.005, 8
    model = Model(plot_backend=None)
    x = model.set_dynamical_states(['c_A', 'c_B', 'T_R', 'T_K'])
    u = model.set_inputs(['F', 'Q_dot'])
    model.set_dynamical_equations(cstr_rhs(x, u, exp=ca.exp))
    model.setup(dt=dt)
    x_ss, u_ss = cstr_steady_state()
    x0s = cstr_initial_states(B, seed=0)
    u0, xp, up, status = [], [], [], []
    for b in range(B):
        nmpc = NMPC(model)                               # fresh object per instance: cold IPOPT-default start
        nmpc.horizon = N
        nmpc.quad_stage_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
        nmpc.quad_stage_cost.add_inputs(names=['F', 'Q_dot'], ref=list(u_ss), weights=CSTR_R)
        nmpc.quad_terminal_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
        nmpc.set_box_constraints(**CSTR_BOUNDS)
        nmpc.set_scaling(**CSTR_SCALING)
        nmpc.set_initial_guess(x_guess=x_ss, u_guess=u_ss)
        nmpc.set_nlp_options(integration_method='collocation')
        nmpc.set_solver_opts({'ipopt.tol': 1e-10, 'ipopt.print_level': 0, 'print_time': False})
        nmpc.setup()
        u0.append(np.asarray(nmpc.optimize(x0s[b])).reshape(-1))
        xs_, us_, _ = nmpc.return_prediction()
        xp.append(np.asarray(xs_))
        up.append(np.asarray(us_))
        status.append(int(nmpc.solver_status_code))
    np.savez(os.path.join(out_dir, 'reference_nmpc_cstr.npz'), x0=x0s, u0=np.stack(u0), x_pred=np.stack(xp),
             u_pred=np.stack(up), status=np.array(status), N=N, dt=dt)

    # ---- config 1: LMPC double integrator (tests/test_LMPC.py:8-33), closed loop of 20 steps with qpOASES
    di = Model(plot_backend=None, discrete=True)
    di.A = np.array([[1, di.dt], [0, 1]])
    di.B = np.array([[di.dt ** 2 / 2], [di.dt]])
    di.setup(dt=.5)
    mpc = LMPC(di)
    mpc.Q, mpc.R, mpc.horizon = np.eye(2), 1, 10
    mpc.set_box_constraints(x_lb=[-5, -5], x_ub=[5, 5], u_lb=[-1], u_ub=[1])
    mpc.setup()
    xk = np.array([3.5, 1.5])
    xs_cl, us_cl = [xk.copy()], []
    A, Bm = np.array([[1, .5], [0, 1]]), np.array([.125, .5])
    for _ in range(20):
        uk = float(np.asarray(mpc.optimize(xk)).reshape(-1)[0])
        us_cl.append(uk)
        xk = A @ xk + Bm * uk
        xs_cl.append(xk.copy())
    np.savez(os.path.join(out_dir, 'reference_lmpc_di.npz'), x=np.stack(xs_cl), u=np.array(us_cl))
    print('golden files written to', out_dir)


if __name__ == '__main__':
    main()
