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
