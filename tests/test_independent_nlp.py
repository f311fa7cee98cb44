"""The product-built problem description against a second, product-free restatement of the reference's NLP
(oracle/independent.py: hand-coded right-hand sides, complex-step derivatives, no symbolic layer, no OcpSpec).

Closes the blind spot of ``FlatNlp(nmpc.spec)``: a wrong scaling, weight, reference substitution or derivative in the
product's front end (nmpc.py / mhe.py / symbolic.py / costs.py) changes ``FlatNlp`` and the engine alike, but not
``oracle.independent``."""
import numpy as np
from scipy.optimize import minimize

from oracle.independent import CstrNlp, ReactionMheNlp, cstr_steady_state
from oracle.nmpc import FlatNlp


def _cstr_flat(N):
    from hilo_mpc_b200.examples import cstr_nmpc
    nmpc = cstr_nmpc(N=N, setup=False)
    nmpc._build_module = False      # problem description only, no CUDA module
    nmpc.setup()
    return nmpc, FlatNlp(nmpc.spec, exact_hessian=True)


def test_cstr_nlp_functions_agree_at_random_points():
    N = 10
    nmpc, flat = _cstr_flat(N)
    ind = CstrNlp(N=N, dt=nmpc.sampling_interval)
    assert flat.n_v == ind.n_v and flat.n_g == ind.n_g
    np.testing.assert_allclose(cstr_steady_state(), nmpc.spec.x_guess * np.asarray(nmpc._xs_full), rtol=1e-12)
    cb = flat.callback(0., np.full(N, nmpc.sampling_interval), np.zeros((N, 0)), np.zeros(0), np.zeros(0))
    rng = np.random.default_rng(0)
    lb, ub = ind.bounds(ind.x_ref)
    for trial in range(3):
        v = lb + (ub - lb) * rng.uniform(.1, .9, size=ind.n_v)
        f, gf, g, J, _ = cb(v, None)
        np.testing.assert_allclose(f, ind.objective(v), rtol=1e-12)
        np.testing.assert_allclose(g, ind.constraints(v), rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(gf, ind.gradient(v), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(J.toarray(), ind.jacobian(v), rtol=1e-10, atol=1e-12)


def test_cstr_minimiser_of_the_independent_nlp_equals_the_oracle_solution():
    """SLSQP on the product-free NLP (complex-step gradients) vs the oracle's interior-point solve of FlatNlp(spec)."""
    from hilo_mpc_b200.examples import cstr_initial_states
    N = 4   # SLSQP is dense: 76 variables, 64 equality constraints
    nmpc, flat = _cstr_flat(N)
    ind = CstrNlp(N=N, dt=nmpc.sampling_interval)
    x0 = cstr_initial_states(3, seed=4)[2] / np.asarray(nmpc._xs_full)
    res, xo, uo = flat.solve(x0, dt=nmpc.sampling_interval, finish_barrier=True, tol=1e-10)
    assert res.status == 1
    lb, ub = ind.bounds(x0)
    v0 = ind.guess()
    v0[ind.ix(0)] = x0
    sol = minimize(ind.objective, v0, jac=ind.gradient, bounds=list(zip(lb, ub)), method='SLSQP',
                   constraints=[{'type': 'eq', 'fun': ind.constraints, 'jac': ind.jacobian}],
                   options={'ftol': 1e-15, 'maxiter': 500})
    assert np.abs(ind.constraints(sol.x)).max() < 1e-9
    xs = np.stack([sol.x[ind.ix(k)] for k in range(N + 1)], axis=1)
    us = np.stack([sol.x[ind.iu(k)] for k in range(N)], axis=1)
    assert abs(sol.fun - res.f) <= 1e-9 * max(1., abs(res.f))
    np.testing.assert_allclose(xs, xo, atol=2e-7)
    np.testing.assert_allclose(us, uo, atol=5e-6)     # the Lagrange term is scaled by dt = 0.005: flat in u


def test_mhe_reference_layout_and_state_augmented_layout_have_the_same_minimiser():
    """The reference keeps ONE parameter variable (mhe.py:596-610); the engine's stage-structured form carries it as an
    extra state with trivial dynamics.  Same data -> same estimate."""
    from hilo_mpc_b200 import MHE
    from test_hostsim_lmpc_mhe import reaction_model
    N, dt = 6, .25
    rng = np.random.default_rng(1)
    y = .25 + .02 * rng.standard_normal(N)
    u = np.full(N, .4)
    mhe = MHE(reaction_model())
    mhe.horizon = N
    mhe.quad_arrival_cost.add_states(weights=[10., 10.], guess=[.5, .2])
    mhe.quad_arrival_cost.add_parameters(weights=[1.], guess=[1.])
    mhe.quad_stage_cost.add_measurements(weights=[100.])
    mhe.quad_stage_cost.add_state_noise(weights=[20., 20.])
    mhe.set_box_constraints(x_lb=[0., 0.], x_ub=[1., 1.], p_lb=[.1], p_ub=[5.])
    mhe._build_module = False
    mhe.setup()
    flat = FlatNlp(mhe.spec, exact_hessian=True)
    sp = np.stack([u, y], axis=1)
    res, xo, uo = flat.solve(None, dt=dt, sps=sp, ap=np.array([.5, .2, 1.]), finish_barrier=True, tol=1e-10)
    assert res.status == 1
    ind = ReactionMheNlp(N, dt, y, u, x_arr=[.5, .2], p_arr=1.)
    v0 = np.zeros(ind.n_v)
    v0[0] = 1.
    for k in range(N + 1):
        v0[ind.ix(k)] = [.5, .2]
    for k in range(N):
        for i in range(3):
            v0[ind.ic(k, i)] = [.5, .2]
    lb, ub = np.full(ind.n_v, -np.inf), np.full(ind.n_v, np.inf)
    lb[0], ub[0] = .1, 5.
    for k in range(N + 1):
        lb[ind.ix(k)], ub[ind.ix(k)] = 0., 1.
    for k in range(N):
        for i in range(3):
            lb[ind.ic(k, i)], ub[ind.ic(k, i)] = 0., 1.
    sol = minimize(ind.objective, v0, jac=ind.gradient, bounds=list(zip(lb, ub)), method='SLSQP',
                   constraints=[{'type': 'eq', 'fun': ind.constraints, 'jac': ind.jacobian}],
                   options={'ftol': 1e-15, 'maxiter': 500})
    assert np.abs(ind.constraints(sol.x)).max() < 1e-9
    assert abs(sol.fun - res.f) <= 1e-8 * max(1., abs(res.f))
    xs = np.stack([sol.x[ind.ix(k)] for k in range(N + 1)], axis=1)
    np.testing.assert_allclose(xs, xo[:2], atol=1e-6)
    np.testing.assert_allclose(sol.x[0], xo[2, -1], atol=1e-6)       # the parameter
    np.testing.assert_allclose(xo[2], xo[2, 0], atol=1e-9)           # constant along the horizon in the augmented form
