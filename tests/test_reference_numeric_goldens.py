"""Oracle and host logic against golden vectors produced by EXECUTING the reference's own source
(tools/make_reference_goldens.py; fixtures in tests/golden/reference_numeric.json): collocation constants
(modeling.py:1092-1127), Latin-hypercube normal sampling (pf.py:425-447), the resampling / roughening / moment block of
ParticleFilter.estimate (pf.py:405-421), the horizon of time-varying parameter values (mpc.py:293-333), the
equation-string parser (util/parsing.py), the export of trained networks (util/machine_learning.py:521-578) and one
interval of the collocation / explicit Runge-Kutta transcription (modeling.py:1129-1281), the quadratic cost classes
(modeling.py:38-818), the QP matrices of LMPC.setup (mpc.py:2179-2266), the horizon loops of NMPC._setup and MHE.setup and one UKF step
(kf.py:486-604)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_numeric.json')))


@pytest.mark.parametrize('method', ['radau', 'legendre'])
@pytest.mark.parametrize('deg', [1, 2, 3, 4, 5])
def test_collocation_constants_equal_the_reference(method, deg):
    from hilo_mpc_b200.transcription import collocation_coefficients
    from oracle.collocation import polynomial_basis
    g = GOLD['polynomial_basis']['values'][f'{method}{deg}']
    for fn in (polynomial_basis, collocation_coefficients):
        B, C, D, tau = fn(deg, method)
        np.testing.assert_allclose(B, g['B'], rtol=0, atol=1e-13)
        np.testing.assert_allclose(C, g['C'], rtol=0, atol=1e-11)
        np.testing.assert_allclose(D, g['D'], rtol=0, atol=1e-12)
        np.testing.assert_allclose(tau, g['T'], rtol=0, atol=1e-14)


@pytest.mark.parametrize('case', range(len(GOLD['lhsnorm']['cases'])))
def test_lhsnorm_equals_the_reference(case):
    from hilo_mpc_b200.estimators import lhsnorm as product
    from oracle.estimators import lhsnorm as oracle
    c = GOLD['lhsnorm']['cases'][case]
    for fn in (product, oracle):
        np.random.seed(c['seed'])
        x = fn(np.array(c['mu']), np.array(c['sigma']), c['n'])
        np.testing.assert_allclose(x, c['x'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('case', range(len(GOLD['pf_resample_block']['cases'])))
def test_pf_resampling_roughening_moments_equal_the_reference(case):
    """pf.py:405-421 with the legacy global RNG: ``np.random.choice`` consumes ``n_s`` uniforms, the roughening draws one
    Latin-hypercube sample.  The oracle's index search, roughening covariance and moments must reproduce the block."""
    from oracle.estimators import lhsnorm, pf_resample_indices, pf_roughening_covariance
    c = GOLD['pf_resample_block']['cases'][case]
    X, q, ns, nx = np.array(c['X_in']), np.array(c['q']), c['n_s'], c['n_x']
    np.random.seed(c['seed'])
    unif = np.random.random_sample(ns)
    idx = pf_resample_indices(q, unif)
    assert np.array_equal(idx, np.array(c['ind']))          # bit-exact
    Xr = X[:, idx]
    if c['roughening']:
        Xr = Xr + lhsnorm(np.zeros(nx), pf_roughening_covariance(Xr, c['K']), ns).T
    np.testing.assert_allclose(Xr, c['X_out'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(Xr.sum(axis=1) / ns, c['x_mean'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.cov(Xr), c['P'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('case', range(len(GOLD['tvp_horizon']['cases'])))
def test_time_varying_parameter_horizon_equals_the_reference(case):
    """``NMPC._get_tvp_parameters_values`` (mpc.py:293-333) executed iteration by iteration: inside the data the horizon
    is ``values[ci : ci + N]``; past its end the reference appends ``values[ci mod N]`` - reproduced as it is - and values
    shorter than the horizon are a TypeError.  ``optimize(tvp=...)`` takes the first N values as they are (mpc.py:342-355)."""
    import warnings
    import torch
    from hilo_mpc_b200 import Model, NMPC
    c = GOLD['tvp_horizon']['cases'][case]
    N, values = c['N'], c['values']
    m = Model(name=f'tvp_case{case}')
    x = m.set_dynamical_states(['a'])
    u = m.set_inputs(['u'])
    p = m.set_parameters(list(values))
    m.set_dynamical_equations([-x[0] + u[0] + sum(p[k] for k in range(len(values)))])
    m.setup(dt=.1)
    nmpc = NMPC(m)
    nmpc.horizon = N
    nmpc.quad_stage_cost.add_states(names=['a'], weights=[1.], ref=[0.])
    nmpc.set_time_varying_parameters(list(values), values=values)
    nmpc._build_module = False
    nmpc.setup()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for ci, hor in enumerate(c['horizons']):
            nmpc._n_iterations = ci
            sp = nmpc._stage_parameters(None, None, None, torch, torch.device('cpu'), 1)
            got = sp.reshape(N, -1).numpy()[:, :len(values)].T
            np.testing.assert_array_equal(got, np.asarray(hor))
    nmpc._n_iterations = 3
    fresh = {k: [float(100 + i) for i in range(N + 2)] for k in values}
    sp = nmpc._stage_parameters(None, fresh, None, torch, torch.device('cpu'), 1)
    np.testing.assert_array_equal(sp.reshape(N, -1).numpy()[:, 0], np.asarray(fresh[list(values)[0]][:N]))
    with pytest.raises(TypeError):
        nmpc._stage_parameters(None, {k: [1.] * (N - 1) for k in values}, None, torch, torch.device('cpu'), 1)


@pytest.mark.parametrize('case', range(len(GOLD['parser']['cases'])))
def test_equation_string_parser_equals_the_reference(case):
    """The reference's parser module (hilo_mpc/util/parsing.py) executed unmodified with a stub ``casadi``: which names
    become states / measurements / inputs / parameters and in which order, the constants, and the values of every
    right-hand side at random points must be what ``hilo_mpc_b200.parsing`` produces for the same strings."""
    from hilo_mpc_b200 import symbolic as sx
    from hilo_mpc_b200.parsing import parse_dynamic_equations
    c = GOLD['parser']['cases'][case]
    t, dt = sx.sym('t'), sx.sym('dt')
    model = parse_dynamic_equations(c['equations'].split('\n'), discrete=c['discrete'], t=t, dt=dt)
    for k in ('x', 'y', 'u', 'p'):
        assert [s.name for s in model[k]] == c['names'][k], k
    assert {k: float(v) for k, v in model['const'].items()} == c['const']
    syms = model['x'] + model['y'] + model['u'] + model['p'] + [t, dt]
    for kind in ('ode', 'meas'):
        F = sx.Function(kind, [('v', np.array(syms, dtype=object))], [('out', np.array(model[kind], dtype=object))])
        for pt in c['points']:
            vals = np.array([pt['env'][s.name] for s in syms])
            np.testing.assert_allclose(np.asarray(F(vals)).reshape(-1), pt['values'][kind], rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize('case', range(len(GOLD['net_to_graph']['cases'])))
def test_network_export_equals_the_reference(case):
    """``Activation`` + ``net_to_casadi_graph`` (util/machine_learning.py:44-118, :521-578) executed from source with a
    stub ``ca``: the labels of trained weights - one or two hidden layers, tanh / sigmoid / softplus / linear, with and
    without z-score scaling - must be what ``ANN.predict`` gives numerically and through its expression graph (call
    nodes and unrolled)."""
    from hilo_mpc_b200 import symbolic as sx
    from hilo_mpc_b200.ann import ANN, Layer
    c = GOLD['net_to_graph']['cases'][case]
    net = ANN([f'x{k}' for k in range(c['n_in'])], [f'y{k}' for k in range(c['n_out'])])
    net.add_layers(*[Layer.dense(h, a) for a, h in c['hidden']])
    net.setup()
    net.set_weights([np.array(w) for w in c['weights']], [np.array(b) for b in c['bias']],
                    input_scaling=c['input_scaling'], output_scaling=c['output_scaling'])
    net.call_node_min_hidden = 1
    z = sx.sym('z', c['n_in'])
    funs = []
    for mode in ('call', 'unroll'):
        net.symbolic_mode = mode
        funs.append(sx.Function('net', [('z', z)], [('y', net.predict(z))]))
    for pt in c['points']:
        x = np.array(pt['x'])
        np.testing.assert_allclose(net.predict(x[:, None])[:, 0], pt['out']['labels'], rtol=1e-12, atol=1e-13)
        for F in funs:
            np.testing.assert_allclose(np.asarray(F(x)).reshape(-1), pt['out']['labels'], rtol=1e-12, atol=1e-13)


def _toy_discretisation_problem(method, **opts):
    """The toy problem of ``tools/make_reference_goldens.py::_discretisation_goldens`` through the public API."""
    from hilo_mpc_b200 import Model, NMPC
    from hilo_mpc_b200 import symbolic as sx
    from oracle.nmpc import FlatNlp
    m = Model(name='toy_disc')
    x = m.set_dynamical_states(['x0', 'x1'])
    u = m.set_inputs(['u0'])
    p = m.set_parameters(['p0'])
    t = m.time
    m.set_dynamical_equations([x[1] + p[0] * sx.sin(t) * u[0], -x[0] * x[0] * x[1] + sx.cos(x[0]) * u[0] - .3 * t])
    m.setup(dt=.1)
    nmpc = NMPC(m)
    nmpc.horizon = 2
    nmpc.stage_cost.cost = x[0] * x[0] + .1 * u[0] * u[0] + x[1] * sx.sin(t)
    nmpc.set_nlp_options(integration_method=method, **opts)
    nmpc._build_module = False
    nmpc.setup()
    nlp = FlatNlp(nmpc.spec)
    e = nlp._stage_exprs(True, False, False)
    args = np.concatenate([e['w'], [e['t'], e['dt']], e['sp']])
    F = sx.Function('stage', [('v', args)], [('g', np.asarray(e['g'], dtype=object)), ('J', np.array([e['J']], dtype=object))])
    return F, nmpc.spec


@pytest.mark.parametrize('case', range(len(GOLD['discretisation']['collocation'])))
def test_collocation_interval_equals_the_reference(case):
    """``RungeKutta._collocation`` (modeling.py:1129-1210) executed from source on a toy problem with explicit time
    dependence and a Lagrange term: collocation equations ``dt f_i - sum_j C_ji X_j``, end state ``sum_i D_i X_i`` and
    quadrature ``sum_i B_i l_i dt`` - including the time ``t + T[i - 1]`` at which the reference evaluates point i
     - must be what the oracle's stage functions (and with them the NLP the engine is compared with)
    contain."""
    c = GOLD['discretisation']['collocation'][case]
    deg = c['degree']
    F, spec = _toy_discretisation_problem('collocation', degree=deg, collocation_points=c['points'])
    assert spec.nsp == 1
    for pt in c['evals']:
        env = pt['env']
        w = [env['x_0'], env['x_1'], env['u_0']] + [env[f'x_{k}_{i}'] for k in range(deg) for i in range(2)] + [0., 0.]
        g, J = F(np.array(w + [env['t_0'], env['dt_0'], env['p_0']]))
        g = np.asarray(g).reshape(-1)
        np.testing.assert_allclose(g[:2 * deg], pt['ce'], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(-g[2 * deg:], pt['xf'], rtol=1e-12, atol=1e-13)      # g_N = x_next - x_f at x_next = 0
        np.testing.assert_allclose(np.asarray(J).reshape(-1), pt['J'], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('case', range(len(GOLD['discretisation']['explicit'])))
def test_explicit_runge_kutta_step_equals_the_reference(case):
    """``RungeKutta._explicit`` (modeling.py:1213-1281) with the reference's Butcher tableaus for order 1-4."""
    c = GOLD['discretisation']['explicit'][case]
    F, spec = _toy_discretisation_problem('erk', order=c['order'])
    for pt in c['evals']:
        env = pt['env']
        g, J = F(np.array([env['x_0'], env['x_1'], env['u_0'], 0., 0., env['t_0'], env['dt_0'], env['p_0']]))
        np.testing.assert_allclose(-np.asarray(g).reshape(-1), pt['xf'], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(np.asarray(J).reshape(-1), pt['J'], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('case', range(len(GOLD['quadratic_cost']['cases'])))
def test_quadratic_stage_cost_equals_the_reference(case):
    """``QuadraticCost`` (modeling.py:89-531) executed from its class source the way ``NMPC._define_cost_terms`` drives it
    (mpc.py:207-242): scaled references (``ref / scale[ind]``), stabilisation terms, a full weight matrix and the
    input-change term with its ``u_old`` placeholder.  The Lagrange term the product hands to the kernels must be the same
    function of the scaled variables."""
    from hilo_mpc_b200 import Model, NMPC
    from hilo_mpc_b200 import symbolic as sx
    c = GOLD['quadratic_cost']['cases'][case]
    m = Model(name=f'cost_case{case}')
    x = m.set_dynamical_states(['c_A', 'c_B', 'T'])
    u = m.set_inputs(['F', 'Q'])
    m.set_dynamical_equations([-x[0] + u[0], x[0] - x[1], -x[2] + u[1]])
    m.setup(dt=.1)
    nmpc = NMPC(m)
    nmpc.horizon = 3
    q = nmpc.quad_stage_cost
    if c['label'] == 'reference_tracking_scaled':
        q.add_states(names=['c_A', 'T'], weights=[1., 10.], ref=[2.1, 110.])
        q.add_inputs(names=['F', 'Q'], weights=[.1, .2], ref=[14., -1100.])
    elif c['label'] == 'stabilisation_and_matrix_weight':
        q.add_states(names=['c_B'], weights=3.)
        q.add_inputs(names=['Q', 'F'], weights=np.array([[2., .5], [.5, 1.]]))
    else:
        q.add_states(names=['c_A', 'c_B', 'T'], weights=[1., 2., 3.], ref=[1., 1., 100.])
        q.add_inputs_change(names=['F'], weights=[5.])
        q.add_inputs_change(names=['Q'], weights=[7.])
    nmpc.set_scaling(x_scaling=c['x_scale'], u_scaling=c['u_scale'])
    nmpc._build_module = False
    nmpc.setup()
    s = nmpc.spec
    assert list(nmpc._du_ind) == c['ind_input_changes']
    lag = s.lag if s.lag_du is None else s.lag + s.lag_du
    args = np.concatenate([s.x, s.u, s.uold])
    F = sx.Function('lag', [('v', args)], [('l', np.array([lag], dtype=object))])
    for pt in c['points']:
        env = pt['env']
        vals = [env[n] for n in ('c_A', 'c_B', 'T', 'F', 'Q')] + [env[n] for n in c['u_old_names']]
        np.testing.assert_allclose(float(np.asarray(F(np.array(vals))).reshape(-1)[0]), pt['cost'], rtol=1e-12)


def test_mhe_quadratic_costs_equal_the_reference():
    """``MHEQuadraticCost`` (modeling.py:533-818) executed from its class source: arrival cost on states and parameters
    with guesses as placeholders, stage cost on measurements and state noise, all on PHYSICAL quantities (``_setup``
    substitutes ``x -> x * scale``, modeling.py:668-684).  ``MHE.setup`` must hand the same functions of the scaled
    variables to the kernels."""
    from hilo_mpc_b200 import MHE, Model
    from hilo_mpc_b200 import symbolic as sx
    g = GOLD['quadratic_cost']['mhe']
    m = Model(name='mhe_cost_case')
    x = m.set_dynamical_states(['a', 'b'])
    u = m.set_inputs(['f'])
    p = m.set_parameters(['k'])
    m.set_measurements(['y1', 'y2'])
    m.set_measurement_equations([x[0] * x[1] + p[0], x[1] * u[0]])
    m.set_dynamical_equations([-p[0] * x[0] + u[0], x[0] - x[1]])
    m.setup(dt=.1)
    mhe = MHE(m)
    mhe.horizon = 3
    mhe.quad_arrival_cost.add_states(weights=[10., 5.], guess=[.5, .2])
    mhe.quad_arrival_cost.add_parameters(weights=[2.], guess=[1.])
    mhe.quad_stage_cost.add_measurements(weights=[100., 10.])
    mhe.quad_stage_cost.add_state_noise(weights=[20., 30.])
    mhe.set_scaling(x_scaling=g['x_scale'], p_scaling=g['p_scale'], w_scaling=g['w_scale'])
    mhe._build_module = False
    mhe.setup()
    s = mhe.spec
    Fa = sx.Function('arr', [('v', np.concatenate([s.x, s.ap]))], [('c', np.array([s.arr], dtype=object))])
    Fs = sx.Function('lag', [('v', np.concatenate([s.x, s.u, s.sp]))], [('c', np.array([s.lag], dtype=object))])
    for pt in g['points']:
        e = pt['env']
        xa = [e['a'], e['b'], e['k']]                              # scaled states, then the scaled parameter
        arr = Fa(np.array(xa + [e[n] for n in g['arrival_placeholders']]))
        np.testing.assert_allclose(float(np.asarray(arr).reshape(-1)[0]), pt['arrival'], rtol=1e-12)
        w = [e[n] for n in g['w_names']]
        sp = [e['f']] + [e[n] for n in g['stage_placeholders']]    # [u_meas | y_meas]
        lag = Fs(np.array(xa + w + sp))
        np.testing.assert_allclose(float(np.asarray(lag).reshape(-1)[0]), pt['stage'], rtol=1e-12)


@pytest.mark.parametrize('case', range(len(GOLD['lmpc_matrices']['cases'])))
def test_lmpc_qp_matrices_equal_the_reference(case):
    """The matrix construction of ``LMPC.setup`` (mpc.py:2179-2266) executed from its source lines: ``H``, the equality
    block ``Aeq`` - whose input block really is ``kron(B, I_N)`` -, tiled bounds and index lists.  The
    oracle's flat QP must be these matrices entry by entry, and the product's condensed form (``bug_compat=True``) must
    be their elimination: ``Aeq [x_0; Phi x_0 + Gam u; u] = 0`` and the same objective."""
    from oracle.lmpc import lmpc_flat_qp
    c = GOLD['lmpc_matrices']['cases'][case]
    A, B, Q, R, P = (np.array(c[k]) for k in ('A', 'B', 'Q', 'R', 'P'))
    nx, nu, N = c['nx'], c['nu'], c['N']
    H, Aeq = lmpc_flat_qp(A, B, Q, R, P, N, bug_compat=True)
    np.testing.assert_array_equal(H, np.array(c['H']))
    np.testing.assert_array_equal(Aeq, np.array(c['Aeq']))
    assert not np.any(np.array(c['beq']))
    intended = lmpc_flat_qp(A, B, Q, R, P, N, bug_compat=False)[1]
    assert np.array_equal(intended, Aeq) == (nx == 1 and nu == 1)            # the two only coincide for scalars
    assert c['x_ind'] == [list(range(k * nx, (k + 1) * nx)) for k in range(N + 1)]
    assert c['u_ind'] == [list(range((N + 1) * nx + k * nu, (N + 1) * nx + (k + 1) * nu)) for k in range(N)]
    # product: condensed matrices of the same QP
    from hilo_mpc_b200 import LMPC, Model
    m = Model(discrete=True, name=f'lmpc_case{case}')
    m.A, m.B = A, B
    m.setup(dt=1.)
    mpc = LMPC(m, bug_compat=True)
    mpc.Q, mpc.R, mpc.P, mpc.horizon = Q, R, P, N
    mpc.set_box_constraints(x_lb=list(-np.arange(1, nx + 1.)), x_ub=list(2 * np.arange(1, nx + 1.)), u_lb=[-1.] * nu,
                            u_ub=[3.] * nu)
    mpc._build_module = False
    mpc.setup(solver='condensed')
    cd = mpc._condensed
    np.testing.assert_array_equal(np.concatenate([np.tile(mpc.spec.x_lb, N + 1), cd['ulb']]), np.array(c['v_lb']))
    np.testing.assert_array_equal(np.concatenate([np.tile(mpc.spec.x_ub, N + 1), cd['uub']]), np.array(c['v_ub']))
    rng = np.random.default_rng(case)
    for _ in range(3):
        x0, u = rng.normal(size=nx), rng.normal(size=N * nu)
        v = np.concatenate([x0, cd['Phi'] @ x0 + cd['Gam'] @ u, u])
        np.testing.assert_allclose(np.array(c['Aeq']) @ v, 0., atol=1e-11)
        full = .5 * v @ np.array(c['H']) @ v
        cond = .5 * u @ cd['H'] @ u + (cd['F'] @ x0) @ u
        v0 = np.concatenate([x0, cd['Phi'] @ x0, np.zeros(N * nu)])
        np.testing.assert_allclose(full - .5 * v0 @ np.array(c['H']) @ v0, cond, rtol=1e-10, atol=1e-11)


def test_nmpc_horizon_assembly_equals_the_reference():
    """The horizon loop of ``NMPC._setup`` (mpc.py, ``for ii in range(self._prediction_horizon)`` ... ``time += dt_ii``)
    executed from its source lines around the executed ``RungeKutta._collocation``: constraint vector (collocation and
    continuity equations interval by interval), objective (sum of the quadratures + Mayer term at ``time + dt`` of the last
    interval), running time with per-interval ``dt`` and the ``u_old`` rule (parameter in interval 0, the interval's own
    input afterwards - the input-change cost vanishes there).  The oracle's NLP callback, built from the
    product's ``OcpSpec`` of the same problem, must give the same ``f`` and ``g``."""
    from hilo_mpc_b200 import Model, NMPC
    from hilo_mpc_b200 import symbolic as sx
    from oracle.nmpc import FlatNlp
    G = GOLD['nmpc_assembly']
    N, deg = G['N'], G['degree']
    assert G['g_lb_ub_all_zero']
    m = Model(name='toy_assembly')
    x = m.set_dynamical_states(['x0', 'x1'])
    u = m.set_inputs(['u0'])
    p = m.set_parameters(['p0'])
    t = m.time
    m.set_dynamical_equations([x[1] + p[0] * sx.sin(t) * u[0], -x[0] * x[0] * x[1] + sx.cos(x[0]) * u[0] - .3 * t])
    m.setup(dt=.1)
    nmpc = NMPC(m)
    nmpc.horizon = N
    nmpc.stage_cost.cost = x[0] * x[0] + .1 * u[0] * u[0] + x[1] * sx.sin(t)
    nmpc.quad_stage_cost.add_inputs_change(names=['u0'], weights=[5.])
    nmpc.terminal_cost.cost = 2. * x[0] * x[0] + x[1] * x[1] + .1 * t * x[0]
    nmpc.set_nlp_options(integration_method='collocation', degree=deg, collocation_points='radau')
    nmpc._build_module = False
    nmpc.setup()
    nlp = FlatNlp(nmpc.spec)
    assert nlp.n_g == N * (deg * 2 + 2)
    for pt in G['points']:
        e = pt['env']
        v = np.zeros(nlp.n_v)
        for k in range(N + 1):
            v[nlp.x_ind[k]] = [e[f'X{k}_0'], e[f'X{k}_1']]
        for k in range(N):
            v[nlp.u_ind[k]] = [e[f'U{k}_0']]
            v[nlp.c_ind[k]] = [e[f'C{k}_{i}'] for i in range(deg * 2)]
        dts = np.array([e[f'PDT_{k}'] for k in range(N)])
        cb = nlp.callback(e['PTIME_0'], dts, np.full((N, 1), e['CP_0']), np.zeros(0), np.array([e['UOLD_0']]))
        f, _, g, _, _ = cb(v, None)
        np.testing.assert_allclose(g, pt['g'], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(f, pt['J'], rtol=1e-12)


def test_mhe_horizon_assembly_equals_the_reference():
    """The horizon loop of ``MHE.setup`` (mhe.py, ``for ii in range(self._horizon)`` ... ``time += dt_ii``) executed from
    its source lines on the reaction model of the MHE tests, around the executed ``RungeKutta._collocation`` and
    ``MHEQuadraticCost``: continuity with the additive state noise, arrival cost in interval 0 only, measurement / noise
    cost from interval 1 on, one parameter variable for the horizon.  ``oracle/independent.py::ReactionMheNlp`` - the
    product-free restatement that ``tests/test_independent_nlp.py`` ties to the form the engine solves - must give the same
    objective and constraints."""
    from oracle.independent import ReactionMheNlp
    G = GOLD['mhe_assembly']
    N, deg = G['N'], G['degree']
    for pt in G['points']:
        e = pt['env']
        nlp = ReactionMheNlp(N, G['dt'], [e[f'YM{k}'] for k in range(N)], [e[f'UM{k}'] for k in range(N)],
                             [e['XA_0'], e['XA_1']], e['PA_0'], degree=deg)
        v = np.zeros(nlp.n_v)
        v[0] = e['P_0']
        for k in range(N + 1):
            v[nlp.ix(k)] = [e[f'X{k}_0'], e[f'X{k}_1']]
        for k in range(N):
            v[nlp.iw(k)] = [e[f'W{k}_0'], e[f'W{k}_1']]
            for i in range(deg):
                v[nlp.ic(k, i)] = [e[f'C{k}_{2 * i}'], e[f'C{k}_{2 * i + 1}']]
        np.testing.assert_allclose(nlp.constraints(v), pt['g'], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(nlp.objective(v), pt['J'], rtol=1e-12)


@pytest.mark.parametrize('case', range(len(GOLD['ukf_step']['cases'])))
def test_unscented_kalman_filter_step_equals_the_reference(case):
    """``UnscentedKalmanFilter._setup_parameters / _setup_predict / _setup_update`` (kf.py:486-604) executed from source on
    numeric matrices with CasADi's shape semantics: weights, sigma points from the columns of the UPPER Cholesky factor,
    weighted mean / covariance, and the update that re-uses the propagated points - one predict + update step of a
    3-state bioreactor (explicit-Euler map, two measurements) for three (alpha, beta, kappa) settings.  The oracle's
    restatement and the kernel body (compiled for the host) must reproduce it."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'hostsim'))
    from est_runner import HostEst
    from hilo_mpc_b200 import Model
    from hilo_mpc_b200.estgen import build_est_module
    from oracle import estimators as est
    G = GOLD['ukf_step']
    c, par, dt = G['cases'][case], G['parameters'], G['dt']
    kw = dict(alpha=c['alpha'], beta=c['beta'], kappa=c['kappa'])

    def f(xc, uc, pc):
        T, cB, cS = xc
        r = (par['mu_0'] + par['mu_1'] * T) * cS * cB / (par['K'] + cS)
        return np.array([T + dt * par['alpha_'] * (par['T_amb'] - T), cB + dt * (r - uc[0] * cB),
                         cS + dt * (-r / par['Y'] - uc[0] * cS)])
    hfun = lambda xc, uc, pc: xc[:2]
    x, P, Q, R, u, y = (np.array(c[k]) for k in ('x', 'P', 'Q', 'R', 'u', 'y'))
    _, gamma, w = est.ukf_weights(3, **kw)
    np.testing.assert_allclose(w, c['weights'], rtol=1e-14)
    np.testing.assert_allclose(gamma, c['gamma'], rtol=1e-14)
    # alpha = 1e-3 makes the weights +-1e6: the weighted sums cancel six digits, and the reference's running sums, the
    # oracle's matrix products and the kernel's loops round differently - tolerances scale with |w_0|
    scale = max(1., abs(w[0, 0]))
    rt_mean, rt_cov = 1e-11 * scale, max(1e-7, 1e-9 * scale)
    xp, Pp, X = est.ukf_predict(f, x, P, u, None, Q, discrete=True, **kw)
    np.testing.assert_allclose(xp, c['x_pred'], rtol=rt_mean)
    np.testing.assert_allclose(Pp, c['P_pred'], rtol=rt_cov, atol=rt_cov * np.abs(c['P_pred']).max())
    np.testing.assert_allclose(X, c['X_prop'], rtol=1e-12)
    xu, Pu, yp = est.ukf_update(hfun, xp, Pp, X, y, u, None, R, **kw)
    np.testing.assert_allclose(xu, c['x_up'], rtol=max(1e-10, rt_mean))
    np.testing.assert_allclose(Pu, c['P_up'], rtol=rt_cov, atol=rt_cov * np.abs(c['P_up']).max())
    np.testing.assert_allclose(yp, c['y_pred'], rtol=rt_mean)
    # the kernel body on the host
    m = Model(discrete=True, name='bioreactor_euler')
    s = m.set_dynamical_states(['T', 'cB', 'cS'])
    d = m.set_inputs(['D'])
    m.set_measurements(['yT', 'yB'])
    r = (par['mu_0'] + par['mu_1'] * s[0]) * s[2] * s[1] / (par['K'] + s[2])
    m.set_dynamical_equations([s[0] + dt * par['alpha_'] * (par['T_amb'] - s[0]), s[1] + dt * (r - d[0] * s[1]),
                               s[2] + dt * (-r / par['Y'] - d[0] * s[2])])
    m.set_measurement_equations([s[0], s[1]])
    m.setup(dt=dt)
    h = HostEst(build_est_module(m, hostsim=True))
    out = h.call('ukf', x, P=P, y=y, u=u, Q=Q, R=R, dt=dt, **kw)
    assert out['status'][0] == 0
    np.testing.assert_allclose(out['x'][0], c['x_up'], rtol=max(1e-10, rt_mean))
    np.testing.assert_allclose(out['P'][0], c['P_up'], rtol=rt_cov, atol=rt_cov * np.abs(c['P_up']).max())
    np.testing.assert_allclose(out['X'][0].T, c['X_prop'], rtol=1e-12)
    np.testing.assert_allclose(out['y_pred'][0], c['y_pred'], rtol=rt_mean)
