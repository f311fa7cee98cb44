"""GPU parity tests for the NMPC path: public API -> C ABI -> sm_100a kernels, compared with the CPU oracle on the
same seeded inputs (sizes the oracle finishes in seconds) and checked through size-independent properties at the full
BASELINE size.  Tolerances (scaled coordinates, fp64): |x - x_oracle| <= 2e-6, |u - u_oracle| <= 2e-5, both solvers
at KKT error <= 1e-8 and barrier parameter 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(nmpc):
    from oracle.nmpc import FlatNlp
    return FlatNlp(nmpc.spec, exact_hessian=True)


@pytest.mark.parametrize('method,N', [('collocation', 10), ('rk4', 10)])
def test_cstr_batch_matches_oracle(cuda_device, method, N):
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    nmpc = cstr_nmpc(N=N, integration_method=method)
    x0 = cstr_initial_states(64, seed=5)
    u0 = nmpc.optimize(torch.tensor(x0, device=cuda_device))
    assert u0.shape == (64, 2)
    status = nmpc.solver_status_code.cpu().numpy()
    assert (status == 1).all()
    x_pred, u_pred, _ = nmpc.return_prediction()
    xs, us = np.asarray(nmpc._xs_full), np.asarray(nmpc._us_full)
    nlp = _oracle(nmpc)
    for b in (0, 17, 63):
        res, xo, uo = nlp.solve(x0[b] / xs, dt=nmpc.sampling_interval, finish_barrier=True)
        assert res.status == 1
        np.testing.assert_allclose(x_pred[b].cpu().numpy() / xs, xo.T, atol=2e-6)
        np.testing.assert_allclose(u_pred[b].cpu().numpy() / us, uo.T, atol=2e-5)
        np.testing.assert_allclose(u0[b].cpu().numpy() / us, uo[:, 0], atol=5e-6)


def test_single_instance_api_and_closed_loop(cuda_device):
    """Drop-in call shape of the reference: x0 with nx entries -> (nu, 1) array; warm-started closed loop drives the
    plant (here: the prediction model) to the set-point."""
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    from hilo_mpc_b200.library import cstr_steady_state
    nmpc = cstr_nmpc(N=10)
    x_ss, u_ss = cstr_steady_state()
    x = cstr_initial_states(1, seed=9)[0]
    for _ in range(30):
        u = nmpc.optimize(x)
        assert u.shape == (2, 1)
        assert nmpc.solver_status_code == 1
        x_pred, u_pred, _ = nmpc.return_prediction()
        assert x_pred.shape == (4, 11) and u_pred.shape == (2, 10)
        x = x_pred[:, 1]
    assert np.abs((x - x_ss) / np.array([1., 1., 100., 100.])).max() < 2e-2
    assert nmpc.solution['u'].shape == (2, 30)


def test_full_size_properties(cuda_device):
    """BASELINE config 2 size (N=50, B=4096): every instance optimal, dynamics residual of the returned trajectory at
    rounding level, bounds respected, instance b solved alone equals instance b solved inside the batch."""
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    nmpc = cstr_nmpc(N=50)
    x0 = cstr_initial_states(4096, seed=0)
    u0 = nmpc.optimize(torch.tensor(x0, device=cuda_device))
    out = nmpc._last
    assert (out['status'] == 1).all()
    stats = out['stats'].cpu().numpy()
    assert stats[:, 10].max() < 1e-8  # primal infeasibility (collocation + continuity residuals)
    assert stats[:, 9].max() < 1e-8   # dual infeasibility
    spec = nmpc.spec
    x, u = out['x'].cpu().numpy(), out['u'].cpu().numpy()
    assert (x[:, 1:] >= spec.x_lb - 1e-12).all() and (x[:, 1:] <= spec.x_ub + 1e-12).all()
    assert (u >= spec.u_lb - 1e-12).all() and (u <= spec.u_ub + 1e-12).all()
    nmpc2 = cstr_nmpc(N=50)
    u_single = nmpc2.optimize(torch.tensor(x0[1234:1235], device=cuda_device))
    np.testing.assert_allclose(u_single.cpu().numpy(), u0[1234:1235].cpu().numpy(), rtol=0, atol=1e-9)


def test_full_size_independent_kkt_certificate(cuda_device):
    """BASELINE config 2 size, bench settings (warm-started closed loop, mu0 = 1e-8): for sampled instances of the
    4096-batch the returned point is checked with the ORACLE's functions, not the solver's own statistics - constraint
    residual of the flat NLP (mpc.py:1133-1787 transcription) <= 1e-8, bounds respected, and first-order optimality: a
    multiplier vector exists (least squares over the constraint multipliers, bound multipliers >= 0 on active bounds)
    that makes the Lagrangian gradient vanish to 1e-6.  Also: permuting the batch permutes the result bit for bit."""
    import torch
    from scipy.sparse.linalg import lsqr
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    B = 4096
    nmpc = cstr_nmpc(N=50, solver_options={'b200.warm_start': 1, 'b200.warm_start_mu': 1e-8})
    x0 = cstr_initial_states(B, seed=0)
    rng = np.random.default_rng(5)
    x = torch.tensor(x0, device=cuda_device)
    for k in range(3):   # third closed-loop step: warm-started solve like in bench.py
        nmpc.optimize(x)
        assert (nmpc.solver_status_code == 1).all()
        xp, _, _ = nmpc.return_prediction()
        x_solved = x.clone()
        x = xp[:, 1, :4] + torch.tensor(rng.uniform(-1, 1, size=(B, 4)) * np.array([.02, .02, .5, .5]), device=cuda_device)
    out = {k: nmpc._last[k].cpu().numpy() for k in ('x', 'u', 'xc')}
    xs = np.asarray(nmpc._xs_full)
    nlp = _oracle(nmpc)
    spec = nmpc.spec
    cb = nlp.callback(0., np.full(spec.N, nmpc.sampling_interval), np.zeros((spec.N, spec.nsp)), np.zeros(spec.ntp),
                      np.zeros(spec.nuold))
    for b in rng.choice(B, size=12, replace=False):
        lb, ub, _ = nlp.bounds_and_guess(x_solved[b].cpu().numpy() / xs)
        v = np.zeros(nlp.n_v)
        for k in range(spec.N + 1):
            v[nlp.x_ind[k]] = out['x'][b, k]
        for k in range(spec.N):
            v[nlp.u_ind[k]] = out['u'][b, k]
            v[nlp.c_ind[k]] = out['xc'][b, k]
        f, gf, g, J, _ = cb(v, None)
        assert np.abs(g).max() < 1e-8, f'instance {b}: constraint residual {np.abs(g).max():.2e}'
        assert (v >= lb - 1e-12).all() and (v <= ub + 1e-12).all()
        # stationarity on the variables that are strictly inside their bounds; on (nearly) active bounds the sign of the
        # remaining gradient component must be the one a non-negative bound multiplier can cancel
        # (mu = 1e-9 at the solution: a bound at distance s carries the multiplier mu/s, below 1e-6 for s > 1e-3)
        free = (v - lb > 1e-3) & (ub - v > 1e-3)
        lam = lsqr(J.T.tocsr()[free], -gf[free], atol=1e-14, btol=1e-14, iter_lim=20000)[0]
        r = gf + J.T @ lam
        assert np.abs(r[free]).max() < 1e-6, f'instance {b}: stationarity residual {np.abs(r[free]).max():.2e}'
        at_lb, at_ub = (v - lb <= 1e-3) & (ub > lb), (ub - v <= 1e-3) & (ub > lb)
        assert (r[at_lb] > -1e-6).all() and (r[at_ub] < 1e-6).all()
    # batch-permutation equivariance (no cross-instance coupling anywhere in the kernels)
    perm = torch.tensor(rng.permutation(B), device=cuda_device)
    a1 = cstr_nmpc(N=50)
    a2 = cstr_nmpc(N=50)
    xt = torch.tensor(x0, device=cuda_device)
    u1 = a1.optimize(xt)
    u2 = a2.optimize(xt[perm])
    assert torch.equal(u1[perm], u2)


def test_full_size_32_seeded_instances_vs_oracle(cuda_device):
    """BASELINE config 2 at its problem size (N=50, Radau-3, 904 variables per instance): 32 seeded instances of a
    4096-instance batch against the oracle's interior-point solve of the reference's flat NLP (scaled coordinates; the
    Lagrange term is scaled by dt = 0.005, hence the looser tolerance on the inputs)."""
    import torch
    from concurrent.futures import ProcessPoolExecutor
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    nmpc = cstr_nmpc(N=50)
    B = 4096
    x0 = cstr_initial_states(B, seed=11)
    nmpc.optimize(torch.tensor(x0, device=cuda_device))
    assert (nmpc.solver_status_code == 1).all()
    x_pred, u_pred, _ = nmpc.return_prediction()
    xs, us = np.asarray(nmpc._xs_full), np.asarray(nmpc._us_full)
    nlp = _oracle(nmpc)
    pick = np.random.default_rng(7).choice(B, size=32, replace=False)
    worst_x = worst_u = 0.
    for b in pick:
        res, xo, uo = nlp.solve(x0[b] / xs, dt=nmpc.sampling_interval, finish_barrier=True)
        assert res.status == 1
        worst_x = max(worst_x, np.abs(x_pred[b].cpu().numpy() / xs - xo.T).max())
        worst_u = max(worst_u, np.abs(u_pred[b].cpu().numpy() / us - uo.T).max())
    assert worst_x < 2e-6 and worst_u < 2e-5, (worst_x, worst_u)


def test_tma_riccati_matches_thread_per_instance_sweep(cuda_device):
    """The TMA-staged multi-lane Riccati kernels (1: first version, 2: producer warp + single-exchange stages) and the
    plain one-thread-per-instance sweep solve the same stage QPs: whole NMPC solves driven by any of them must agree to
    rounding level (N=50, ragged batch size)."""
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    x0 = torch.tensor(cstr_initial_states(1000 + 3, seed=21), device=cuda_device)
    res = []
    for variant in (0, 2, 6, 9):
        nmpc = cstr_nmpc(N=50, solver_options={'b200.riccati_variant': variant})
        nmpc.optimize(x0)
        out = nmpc._last
        assert (out['status'] == 1).all()
        res.append({k: out[k].clone() for k in ('x', 'u', 'lam', 'iters')})
    for other in range(1, len(res)):
        assert (res[0]['iters'] - res[other]['iters']).abs().max().item() <= 1
        assert (res[0]['x'] - res[other]['x']).abs().max().item() < 1e-8
        assert (res[0]['u'] - res[other]['u']).abs().max().item() < 1e-7
        assert (res[0]['lam'] - res[other]['lam']).abs().max().item() < 1e-6


def test_warm_start_mode_reaches_the_same_solution_in_fewer_iterations(cuda_device):
    """Closed loop with disturbances: the warm-start mode (multipliers kept, small initial barrier parameter) must return
    the same inputs as the IPOPT-default restart (mu = 0.1, z = 1) and need fewer interior-point iterations."""
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    x0 = torch.tensor(cstr_initial_states(512, seed=4), device=cuda_device)
    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(0)
    noise = (torch.rand(4, 512, 4, generator=gen, device=cuda_device, dtype=torch.float64) * 2 - 1) * \
        torch.tensor([.02, .02, .5, .5], dtype=torch.float64, device=cuda_device)
    runs = {}
    # warm = 0: IPOPT-default restart; 1e-4: library default of the warm-start mode; 1e-8: the setting bench.py uses
    for warm, mu0 in ((0, None), (1, 1e-4), (1, 1e-8)):
        opts = {'b200.warm_start': warm, 'ipopt.tol': 1e-9}
        if mu0 is not None:
            opts['b200.warm_start_mu'] = mu0
        nmpc = cstr_nmpc(N=50, solver_options=opts)
        x, us, its = x0.clone(), [], []
        for k in range(4):
            us.append(nmpc.optimize(x).clone())
            assert (nmpc.solver_status_code == 1).all()
            its.append(nmpc._last['iters'].double().mean().item())
            xp, _, _ = nmpc.return_prediction()
            x = xp[:, 1, :4] + noise[k]
        runs[(warm, mu0)] = (us, its)
    cold = runs[(0, None)]
    for key in ((1, 1e-4), (1, 1e-8)):
        for ua, ub in zip(cold[0], runs[key][0]):
            assert (ua - ub).abs().max().item() < 1e-4 * ua.abs().max().item()
        assert np.mean(runs[key][1][1:]) < np.mean(cold[1][1:])
    assert np.mean(runs[(1, 1e-8)][1][1:]) < np.mean(runs[(1, 1e-4)][1][1:])


def test_quad_eval_kernel_matches_thread_per_node_kernel(cuda_device):
    """The 4-lanes-per-node register-resident evaluation kernel and the plain one-thread-per-node kernel linearise the same
    NLP: complete solves driven by either must agree to rounding level (exact and Gauss-Newton Hessians, N=50)."""
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    x0 = torch.tensor(cstr_initial_states(777, seed=31), device=cuda_device)
    for hess in ('exact', 'gauss-newton'):
        res = []
        for variant in (0, 1):
            nmpc = cstr_nmpc(N=50, solver_options={'b200.eval_variant': variant, 'b200.hessian': hess})
            nmpc.optimize(x0)
            out = nmpc._last
            assert (out['status'] == 1).all()
            res.append({k: out[k].clone() for k in ('x', 'u', 'lam', 'iters')})
        assert (res[0]['iters'] - res[1]['iters']).abs().max().item() <= 1
        assert (res[0]['x'] - res[1]['x']).abs().max().item() < 1e-8
        assert (res[0]['u'] - res[1]['u']).abs().max().item() < 1e-7


def test_fused_step_kernels_match_separate_kernels(cuda_device):
    """Expand/Prep/Trial/Decide/Apply/Commit as one kernel, one CTA per instance.  Variant 1 (one thread per node) runs
    the same device functions in the same order as the separate launches: bit-identical solves.  Variant 2 (4 lanes per
    node, register-resident iterate) sums in a different order: same iteration counts up to one, same solution to
    rounding level.  The wide initial-state box makes the filter backtrack for part of the batch."""
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    x0 = torch.tensor(cstr_initial_states(1000, seed=5), device=cuda_device)
    res = []
    for variant in (0, 1, 2):
        nmpc = cstr_nmpc(N=50, solver_options={'b200.step_variant': variant})
        nmpc.optimize(x0)
        out = nmpc._last
        assert (out['status'] == 1).all()
        res.append({k: out[k].clone() for k in ('x', 'u', 'lam', 'iters')})
    assert torch.equal(res[0]['iters'], res[1]['iters'])
    assert torch.equal(res[0]['x'], res[1]['x'])
    assert torch.equal(res[0]['u'], res[1]['u'])
    assert (res[0]['iters'] - res[2]['iters']).abs().max().item() <= 1
    assert (res[0]['x'] - res[2]['x']).abs().max().item() < 1e-8
    assert (res[0]['u'] - res[2]['u']).abs().max().item() < 1e-7


def test_compaction_of_straggler_iterations_is_bitwise_neutral(cuda_device):
    """Late iterations run over the list of still-active instances (hb_ocp_call.no_compaction = 0): same kernels, same
    per-instance arithmetic, so iterates, iteration counts and statuses must be identical to the full-batch launches.
    Cold start with widely spread initial states and a poll after every iteration, so that the list mode is in use for
    most of the solve."""
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    x0 = cstr_initial_states(301, seed=33)
    x0[::3] += np.array([.3, -.2, 6., -4.])      # a third of the batch starts far away: long tail of iterations
    x0 = torch.tensor(x0, device=cuda_device)
    res = []
    for comp in (0, 1):
        nmpc = cstr_nmpc(N=10, solver_options={'b200.compaction': comp, 'b200.check_every': 1})
        nmpc.optimize(x0)
        out = nmpc._last
        assert (out['status'] == 1).all()
        res.append({k: out[k].clone() for k in ('x', 'u', 'lam', 'iters')})
    spread = (res[0]['iters'].max() - res[0]['iters'].min()).item()
    assert spread >= 2, 'test problem should have stragglers'
    for k in ('x', 'u', 'lam', 'iters'):
        assert torch.equal(res[0][k], res[1][k]), k


def test_sweep_entry_point_on_the_device(cuda_device):
    """hb_ocp_riccati on the GPU (B = 1000 random instances, N = 50): the warp-per-instance tensor-pipe sweep equals the
    one-thread-per-instance sweep and a plain NumPy recursion; indefinite R blocks take the regularising path."""
    import torch
    from hilo_mpc_b200.examples import cstr_nmpc
    from hilo_mpc_b200.stage_blocks import pack_stage_blocks, riccati_reference
    from test_hostsim_riccati import _random_stage_data
    eng = cstr_nmpc(N=50).engine
    eng._open()
    d = eng.dims
    assert d.frag == 1
    rng = np.random.default_rng(5)
    B = 1000
    for indefinite_every in (0, 64):
        A, Bm, b, Q, S, R, q, r, qb, rb, QN, qN, qbN = _random_stage_data(rng, B, d.N, d.nx, d.nu, indefinite_every)
        qa, qh = pack_stage_blocks(d, A, Bm, b, Q, S, R, q, r, qb, rb)
        qpN = np.concatenate([QN.reshape(B, -1), qN, qbN], axis=1)
        t = lambda a: torch.tensor(a, device=cuda_device)
        ref = eng.riccati(t(qa), t(qh), t(qpN), riccati_variant=0)
        out = eng.riccati(t(qa), t(qh), t(qpN), riccati_variant=2)
        assert (ref['status'] == 0).all() and (out['status'] == 0).all()
        # instances with an indefinite block: the regularised recursion of a non-convex QP is unstable over 50 stages
        # (solutions of magnitude 1e12), so the two summation orders only agree to a few digits there; all others to rounding
        hit = np.zeros(B * d.N, dtype=bool)
        if indefinite_every:
            hit[7::indefinite_every] = True
        hit = torch.tensor(hit.reshape(B, d.N).any(axis=1), device=cuda_device)
        for k in ('dx', 'du', 'lam'):
            err = (out[k] - ref[k]).abs().flatten(1).max(dim=1).values
            mag = ref[k].abs().flatten(1).max(dim=1).values.clamp(min=1.)
            assert (err[~hit] <= 1e-11 * mag[~hit]).all(), k
            if hit.any():
                assert torch.isfinite(out[k]).all() and (err[hit] <= 1e-2 * mag[hit]).all(), k
        assert ((out['edual'] - ref['edual']).abs()[~hit] <= 1e-12 * ref['edual'].abs()[~hit]).all()
        if not indefinite_every:
            dx, du = riccati_reference(A[7], Bm[7], b[7], Q[7], S[7], R[7], q[7], r[7], QN[7], qN[7])
            np.testing.assert_allclose(out['dx'][7].cpu().numpy(), dx, atol=1e-10)
            np.testing.assert_allclose(out['du'][7].cpu().numpy(), du, atol=1e-10)


def test_multi_start_and_free_initial_state(cuda_device):
    """``optimize(x0, runs=r)`` (mpc.py:728-740): r solves from randomly perturbed starting points (legacy NumPy RNG), the
    best converged one is kept - on the CSTR every start reaches the same minimiser.  ``fix_x0=False`` (mpc.py:803-807):
    x_0 becomes a decision variable inside the state bounds, compared with the oracle on the same flat NLP."""
    import copy
    import torch
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    from oracle.nmpc import FlatNlp
    x0 = cstr_initial_states(6, seed=11)
    x0t = torch.tensor(x0, device=cuda_device)
    base = cstr_nmpc(N=10)
    u_ref = base.optimize(x0t).clone()
    nmpc = cstr_nmpc(N=10)
    np.random.seed(5)
    u_ms = nmpc.optimize(x0t, runs=3)
    assert (nmpc.solver_status_code == 1).all()
    assert np.random.rand() != np.random.RandomState(5).rand()      # the legacy global RNG was consumed
    assert (u_ms - u_ref).abs().max().item() < 1e-5 * u_ref.abs().max().item()
    # free initial state
    free = cstr_nmpc(N=10)
    u_free = free.optimize(x0t, fix_x0=False)
    assert (free.solver_status_code == 1).all()
    out = free._last
    assert (out['stats'][:, 13] <= base._last['stats'][:, 13] + 1e-9).all()   # a relaxation of the pinned problem
    spec = copy.copy(free.spec)
    spec.fixed_x0 = False
    nlp = FlatNlp(spec, exact_hessian=True)
    xs, us = np.asarray(free._xs_full), np.asarray(free._us_full)
    for b in (0, 5):
        lb, ub, v0 = nlp.bounds_and_guess(None)
        v0[nlp.x_ind[0]] = x0[b] / xs
        res, xo, uo = nlp.solve(None, v0=v0, dt=free.sampling_interval, finish_barrier=True)
        assert res.status == 1
        np.testing.assert_allclose(out['x'][b].cpu().numpy(), xo.T, atol=2e-6)
        np.testing.assert_allclose(out['u'][b].cpu().numpy(), uo.T, atol=2e-5)
    assert u_free.shape == (6, 2)
