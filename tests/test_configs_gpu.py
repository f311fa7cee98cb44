"""GPU tests of the BASELINE configurations 4 (MHE, CSTR, N=20) and 5 (path following with MLP dynamics, N=30) at their
stated problem sizes (reduced batch), and of the generalised multi-lane kernels they run on: evaluation kernel with 4 or
8 lanes per node incl. free x_0 / arrival cost / additive stage input / discrete-type objective, step kernel for the same
shapes, cp.async-ring Riccati sweep for any nu with IPOPT's inertia correction."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FAST = {'b200.eval_variant': 1, 'b200.step_variant': 2, 'b200.riccati_variant': 2}
PLAIN = {'b200.eval_variant': 0, 'b200.step_variant': 0, 'b200.riccati_variant': 0}


def _feed(mhe, u, y):
    est = (None, None)
    for k in range(u.shape[0]):
        mhe.add_measurements(y[k], u_meas=u[k])
        est = mhe.estimate()
    return est


def test_config4_mhe_fast_kernels_match_plain_kernels_and_oracle(cuda_device):
    """MHE of the CSTR with one estimated parameter, N=20 (nx=5 incl. the parameter state, nu=4 noise inputs, free x_0,
    arrival cost, discrete-type objective, Radau-3): the multi-lane kernels and the one-thread-per-node kernels solve
    the same NLP, and both equal the oracle's solve of the reference's flat NLP (mhe.py:596-743)."""
    from hilo_mpc_b200.examples import cstr_mhe, cstr_mhe_data
    from oracle.nmpc import FlatNlp
    N, B = 20, 48
    u, y, xt, ks = cstr_mhe_data(B, N)
    res = []
    for opts in (FAST, PLAIN):
        mhe = cstr_mhe(N=N, solver_options=dict(opts))
        d = None
        x_est, p_est = _feed(mhe, u, y)
        assert x_est.shape == (B, 4) and p_est.shape == (B, 1)
        assert (mhe.solver_status_code == 1).all()
        out = mhe._last
        res.append({k: out[k].clone() for k in ('x', 'u', 'iters')})
        if opts is FAST:
            d = mhe.engine.dims
            assert (d.nx, d.nu, d.N, d.fixed_x0) == (5, 4, 20, 0)
    assert (res[0]['iters'] - res[1]['iters']).abs().max().item() <= 1
    assert (res[0]['x'] - res[1]['x']).abs().max().item() < 1e-7
    assert (res[0]['u'] - res[1]['u']).abs().max().item() < 1e-6
    # oracle: the flat NLP of the reference, cold start, same data
    spec = mhe.spec
    nlp = FlatNlp(spec, exact_hessian=True)
    ap = np.concatenate([mhe.quad_arrival_cost.x_guess, mhe.quad_arrival_cost.p_guess])
    for b in (0, 17, 47):
        sp = np.concatenate([u[:, b], y[:, b]], axis=1)
        r, xo, uo = nlp.solve(None, dt=mhe.sampling_interval, sps=sp, ap=ap, finish_barrier=True, tol=1e-8)
        assert r.status == 1
        np.testing.assert_allclose(res[0]['x'][b].cpu().numpy(), xo.T, atol=2e-6)
        np.testing.assert_allclose(res[0]['u'][b].cpu().numpy(), uo.T, atol=2e-5)


def test_small_mhe_and_robot_fast_kernels_match_plain_kernels(cuda_device):
    """The 2-state + 1-parameter MHE (4 lanes per node) and the N=5 path-following problem with MLP dynamics (nx=5, nu=3:
    8 lanes per node, cp.async-ring sweep with a 3 x 3 Huu) through both kernel families."""
    import torch
    from hilo_mpc_b200 import MHE
    from hilo_mpc_b200.examples import robot_initial_states, robot_nn_nmpc
    from test_hostsim_lmpc_mhe import reaction_model
    rng = np.random.default_rng(3)
    B, N = 32, 6
    y = .25 + .02 * rng.standard_normal((N, B, 1))
    res = []
    for opts in (FAST, PLAIN):
        mhe = MHE(reaction_model())
        mhe.horizon = N
        mhe.quad_arrival_cost.add_states(weights=[10., 10.], guess=[.5, .2])
        mhe.quad_arrival_cost.add_parameters(weights=[1.], guess=[1.])
        mhe.quad_stage_cost.add_measurements(weights=[100.])
        mhe.quad_stage_cost.add_state_noise(weights=[20., 20.])
        mhe.set_box_constraints(x_lb=[0., 0.], x_ub=[1., 1.], p_lb=[.1], p_ub=[5.])
        mhe.setup(nlp_opts=dict(opts, **{'ipopt.tol': 1e-10}))
        _feed(mhe, np.full((N, 1), .4), y)
        assert (mhe.solver_status_code == 1).all()
        res.append(mhe._last['x'].clone())
    assert (res[0] - res[1]).abs().max().item() < 1e-8
    x0 = torch.tensor(robot_initial_states(64), device=cuda_device)
    res = []
    for opts in (FAST, PLAIN):
        nmpc = robot_nn_nmpc(N=5, hidden=8, setup=False)
        nmpc.set_solver_opts(dict(opts, **{'b200.regularization': 'mirror'}))
        nmpc.setup()
        nmpc.optimize(x0)
        st = nmpc.solver_status_code
        res.append((st.clone(), nmpc._last['x'].clone(), nmpc._last['u'].clone()))
    both = (res[0][0] == 1) & (res[1][0] == 1)
    assert both.float().mean().item() > .9
    # non-convex problem: rounding differences may send an instance to another local minimiser; the bulk must coincide
    same = ((res[0][1] - res[1][1]).abs().flatten(1).max(dim=1).values < 1e-6) & both
    assert same.float().mean().item() > .85


def test_config5_path_following_with_mlp_dynamics_full_size(cuda_device):
    """BASELINE config 5 at its stated problem size: 1 x 32 tanh MLP as the right-hand side, path variable, N=30, dt=1
    (mpc.py:1173-1204, machine_learning.py:521-578), cold start from the tiled guess.  The Lagrangian Hessian of this
    problem is indefinite far from the solution; with the eigenvalue-mirroring convexification of the stage Hessians
    every instance reaches a KKT point.  Parity: started from the engine's solution, the oracle's interior-point solve of
    the reference's flat NLP must stay at that point (basin-aware comparison of a non-convex problem)."""
    import torch
    from hilo_mpc_b200.examples import robot_initial_states, robot_nn_nmpc
    from oracle.nmpc import FlatNlp
    nmpc = robot_nn_nmpc(N=30, hidden=32, setup=False)
    nmpc.set_solver_opts({'b200.regularization': 'mirror', 'ipopt.max_iter': 1500})
    nmpc.setup()
    B = 96
    x0 = robot_initial_states(B)
    u0 = nmpc.optimize(torch.tensor(x0, device=cuda_device))
    assert u0.shape == (B, 2)
    st = nmpc.solver_status_code
    # status 2 = IPOPT's "solved to acceptable level" (15 consecutive iterates within acceptable_tol = 1e-6): the
    # convexified Newton steps converge linearly near the solution, a part of the batch stops there
    frac = ((st == 1) | (st == 2)).float().mean().item()
    assert frac >= .95, f'only {frac:.2%} of the instances reached a KKT point: {torch.bincount(st.clamp(min=0))}'
    assert (st == 1).float().mean().item() >= .5
    out = nmpc._last
    acc = (st == 2).cpu().numpy()
    if acc.any():
        sa = out['stats'].cpu().numpy()[acc]
        assert max(sa[:, 9].max(), sa[:, 10].max(), sa[:, 11].max()) <= 1e-6
    ok = (st == 1)
    theta = out['x'][ok][:, :, 4]
    assert (theta[:, 1:] - theta[:, :-1] >= .1 - 1e-6).all()      # u_theta >= 0.1
    stats = out['stats'].cpu().numpy()
    okn = ok.cpu().numpy()
    assert stats[okn, 10].max() < 1e-8 and stats[okn, 9].max() < 1e-8
    # tracking: the robot ends up on the path
    xs = out['x'][ok]
    px, py, th = xs[:, -1, 0], xs[:, -1, 1], xs[:, -1, 4]
    dist = ((px - (30. - 14. * torch.cos(th))) ** 2 + (py - (30. - 16. * torch.sin(th))) ** 2).sqrt()
    assert dist.median().item() < 1.
    # oracle, warm-started at the engine's point
    spec = nmpc.spec
    nlp = FlatNlp(spec, exact_hessian=True)
    idx = [int(i) for i in torch.nonzero(ok).flatten()[:2]]
    for b in idx:
        v0 = np.zeros(nlp.n_v)
        xb, ub, cb = (out[k][b].cpu().numpy() for k in ('x', 'u', 'xc'))
        for k in range(spec.N + 1):
            v0[nlp.x_ind[k]] = xb[k]
        for k in range(spec.N):
            v0[nlp.u_ind[k]] = ub[k]
            v0[nlp.c_ind[k]] = cb[k]
        x0b = np.concatenate([x0[b], [0.]])
        r, xo, uo = nlp.solve(x0b, v0=v0, dt=1., finish_barrier=True, tol=1e-8, mu_init=1e-6)
        assert r.status == 1
        np.testing.assert_allclose(xb, xo.T, atol=5e-5)
        np.testing.assert_allclose(ub, uo.T, atol=5e-5)


def test_nonconvex_instances_of_the_warp_sweep_are_redone_with_inertia_correction(cuda_device):
    """Pendulum swing-up (nx=2, nu=1: the warp-per-instance tensor-pipe sweep): far from the solution the exact Lagrangian
    Hessian is indefinite, the sweep flags the instance (ST_ICREQ) and the step kernel (step_variant 2) or the fix-up
    kernel (step_variant 0) redoes it with IPOPT's inertia correction.  Both routes must reach the same KKT points."""
    import torch
    from hilo_mpc_b200 import NMPC
    from hilo_mpc_b200 import symbolic as sx
    from hilo_mpc_b200.model import Model
    res = []
    for variant in (2, 0):
        model = Model(name='pendulum')
        x = model.set_dynamical_states(['th', 'om'])
        u = model.set_inputs(['tau'])
        model.set_dynamical_equations([x[1], -sx.sin(x[0]) - .1 * x[1] + u[0]])
        model.setup(dt=.1)
        nmpc = NMPC(model)
        nmpc.horizon = 20
        nmpc.quad_stage_cost.add_states(names=['th', 'om'], ref=[np.pi, 0.], weights=[10., 1.])
        nmpc.quad_stage_cost.add_inputs(names=['tau'], weights=[.1])
        nmpc.quad_terminal_cost.add_states(names=['th', 'om'], ref=[np.pi, 0.], weights=[100., 10.])
        nmpc.set_box_constraints(x_lb=[-10., -10.], x_ub=[10., 10.], u_lb=[-3.], u_ub=[3.])
        nmpc.set_initial_guess(x_guess=[0., 0.], u_guess=[0.])
        nmpc.set_solver_opts({'b200.step_variant': variant, 'ipopt.max_iter': 500})
        nmpc.setup()
        rng = np.random.default_rng(0)
        x0 = np.column_stack([rng.uniform(-.5, .5, 64), rng.uniform(-.5, .5, 64)])
        nmpc.optimize(torch.tensor(x0, device=cuda_device))
        assert nmpc.engine.dims.frag == 1
        st = nmpc.solver_status_code
        out = nmpc._last
        res.append((st.clone(), out['x'].clone(), out['stats'][:, 20].clone()))   # ST_DW: last inertia correction used
    ok = (res[0][0] == 1) & (res[1][0] == 1)
    assert ok.float().mean().item() > .9
    assert (res[0][2] > 0).any() and (res[1][2] > 0).any()     # the correction was needed: the non-convex path was taken
    same = ((res[0][1] - res[1][1]).abs().flatten(1).max(dim=1).values < 1e-6) & ok
    assert same.float().mean().item() > .8
