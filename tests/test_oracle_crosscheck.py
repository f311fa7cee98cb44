"""The oracle's interior-point solver (oracle/ipm.py, a restatement of IPOPT's published algorithm) against an INDEPENDENT
solver on the real nonlinear problem: SciPy's SLSQP on the same flat NLP (objective, gradient, constraints, Jacobian taken
from the oracle's transcription, nothing from its solver).  The reference holds no numeric assertion for NMPC (SURVEY.md
§8c, "parity unpinned"); this is the strongest pin available here next to the KKT certificate."""
import numpy as np
import scipy.optimize as sopt

from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
from oracle.nmpc import FlatNlp


def test_cstr_nmpc_oracle_matches_slsqp():
    N = 2
    nmpc = cstr_nmpc(N=N, setup=False)
    nmpc._build_module = False
    nmpc.setup()
    spec = nmpc.spec
    nlp = FlatNlp(spec, exact_hessian=True)
    x0 = cstr_initial_states(1, seed=4)[0] / np.asarray(nmpc._xs_full)
    res, xo, uo = nlp.solve(x0, dt=nmpc.sampling_interval, finish_barrier=True, tol=1e-10)
    assert res.status == 1 and res.kkt <= 1e-9
    lb, ub, v0 = nlp.bounds_and_guess(x0)
    cb = nlp.callback(0., np.full(N, nmpc.sampling_interval), np.zeros((N, spec.nsp)), np.zeros(spec.ntp),
                      np.zeros(spec.nuold))
    cons = {'type': 'eq', 'fun': lambda v: cb(v, None)[2], 'jac': lambda v: cb(v, None)[3].toarray()}
    r = sopt.minimize(lambda v: cb(v, None)[0], np.clip(v0, lb, ub), jac=lambda v: cb(v, None)[1],
                      bounds=list(zip(lb, ub)), constraints=[cons], method='SLSQP', options=dict(ftol=1e-15, maxiter=60))
    u_s = np.stack([r.x[i] for i in nlp.u_ind], axis=1)
    x_s = np.stack([r.x[i] for i in nlp.x_ind], axis=1)
    # the Lagrange term is scaled by dt = 0.005: the objective is flat in u, hence the looser input tolerance
    np.testing.assert_allclose(x_s, xo, atol=1e-7)
    np.testing.assert_allclose(u_s, uo, atol=2e-6)
    np.testing.assert_allclose(r.fun, res.x @ np.zeros_like(res.x) + cb(res.x, None)[0], rtol=1e-8)
    assert np.abs(cb(r.x, None)[2]).max() < 1e-9   # SLSQP's point is feasible


def test_mhe_oracle_matches_slsqp():
    """Moving-horizon estimation problem (free x_0, arrival cost, state noise, one estimated parameter; mhe.py:418-790)."""
    from test_hostsim_lmpc_mhe import _mhe
    from oracle.estimators import rk4
    N = 4
    mhe = _mhe(N)
    spec = mhe.spec
    rng = np.random.default_rng(0)
    x_true = np.array([.6, .25])
    f = lambda x, u, p: np.array([u[0] * (1 - x[0]) - p[0] * x[0], p[0] * x[0] - u[0] * x[1]])
    rows = []
    for _ in range(N):
        rows.append([.4, x_true[1] + rng.normal() * 5e-3])
        x_true = rk4(f, x_true, [.4], [1.7], .25, 20)
    sp = np.array(rows)
    ap = np.array([.5, .2, 1.])
    nlp = FlatNlp(spec, exact_hessian=True)
    res, xo, uo = nlp.solve(None, dt=.25, sps=sp, ap=ap, finish_barrier=True, tol=1e-10)
    assert res.status == 1 and res.kkt <= 1e-9
    lb, ub, v0 = nlp.bounds_and_guess(None)
    cb = nlp.callback(0., np.full(N, .25), sp, np.zeros(spec.ntp), np.concatenate([np.zeros(spec.nuold), ap]))
    cons = {'type': 'eq', 'fun': lambda v: cb(v, None)[2], 'jac': lambda v: cb(v, None)[3].toarray()}
    big = 1e6   # SLSQP wants finite bounds
    r = sopt.minimize(lambda v: cb(v, None)[0], np.clip(v0, lb, ub), jac=lambda v: cb(v, None)[1],
                      bounds=list(zip(np.maximum(lb, -big), np.minimum(ub, big))), constraints=[cons], method='SLSQP',
                      options=dict(ftol=1e-15, maxiter=80))
    x_s = np.stack([r.x[i] for i in nlp.x_ind], axis=1)
    u_s = np.stack([r.x[i] for i in nlp.u_ind], axis=1)
    np.testing.assert_allclose(x_s, xo, atol=1e-6)
    np.testing.assert_allclose(u_s, uo, atol=1e-6)
    assert np.abs(cb(r.x, None)[2]).max() < 1e-9
