"""Kernel-logic parity on the CPU: the same kernel bodies that nvcc compiles for sm_100a are compiled for the host
(-DHB_HOSTSIM, test infrastructure) and compared with the oracle's flat-NLP solution of the identical transcription.
Tolerances: both solvers stop on the same central-path point (mu = 1e-9, KKT error <= 1e-8)."""
import numpy as np
import pytest

from hilo_mpc_b200.codegen import OcpGenerator
from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
from oracle.nmpc import FlatNlp
from runner import HostOcp


def _setup(N, method):
    nmpc = cstr_nmpc(N=N, integration_method=method, setup=False)
    nmpc._build_module = False
    nmpc.setup()
    return nmpc


@pytest.mark.parametrize('method', ['collocation', 'rk4'])
def test_cstr_host_kernels_match_oracle(method):
    N = 8
    nmpc = _setup(N, method)
    spec = nmpc.spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    xs = np.asarray(nmpc._xs_full)
    B = 4
    x0 = cstr_initial_states(B, seed=3) / xs
    x = np.tile(spec.x_guess, (B, N + 1, 1))
    u = np.tile(spec.u_guess, (B, N, 1))
    xc = np.tile(spec.x_guess, (B, N, spec.degree)) if spec.ncol else None
    out = host.solve(x0, x, u, xc, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub, dt=nmpc.sampling_interval)
    assert (out['status'] == 1).all()
    assert out['iters'].max() < 60
    nlp = FlatNlp(spec, exact_hessian=True)
    for b in range(2):
        res, xo, uo = nlp.solve(x0[b], dt=nmpc.sampling_interval, finish_barrier=True)
        assert res.status == 1 and res.kkt <= 1e-8
        np.testing.assert_allclose(out['x'][b], xo.T, atol=2e-6)
        np.testing.assert_allclose(out['u'][b], uo.T, atol=2e-5)
        np.testing.assert_allclose(out['u'][b, 0], uo[:, 0], atol=5e-6)


@pytest.mark.parametrize('dev', [[-.9, -.6, -30., -30.], [.6, .9, 11., 15.]])
def test_active_bounds_and_warm_start(dev):
    """Large initial deviations drive Q_dot to its upper bound (first case) / T_R to its upper state bound (second
    case); a second call warm-started from the solution (un-shifted, mpc.py:725-726) converges to the same point."""
    N = 8
    nmpc = _setup(N, 'collocation')
    spec = nmpc.spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    xs, us = np.asarray(nmpc._xs_full), np.asarray(nmpc._us_full)
    x_ss = spec.x_guess * xs
    x0 = ((x_ss + np.array(dev)) / xs)[None]
    x = np.tile(spec.x_guess, (1, N + 1, 1))
    u = np.tile(spec.u_guess, (1, N, 1))
    xc = np.tile(spec.x_guess, (1, N, spec.degree))
    # the Lagrange term is scaled by dt = 0.005, so the cost curvature in u is ~1e-3 and a KKT error of 1e-8 leaves
    # ~1e-5..1e-4 of slack in u; both solvers are therefore run to 1e-10 for a sharp comparison
    out = host.solve(x0, x, u, xc, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub, dt=nmpc.sampling_interval, tol=1e-10)
    assert out['status'][0] == 1
    nlp = FlatNlp(spec, exact_hessian=True)
    res, xo, uo = nlp.solve(x0[0], dt=nmpc.sampling_interval, finish_barrier=True, tol=1e-10)
    assert res.status == 1
    at_bound = ((np.abs(uo.T - spec.u_lb) < 1e-6) | (np.abs(uo.T - spec.u_ub) < 1e-6)).any() or \
        (np.abs(xo.T - spec.x_ub) < 5e-4).any()  # T_R rides 0.02 K below its bound (barrier-limited)
    assert at_bound, 'test problem should have active bounds'
    np.testing.assert_allclose(out['u'][0], uo.T, atol=2e-5)
    out2 = host.solve(x0, out['x'], out['u'], out['xc'], spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub,
                      dt=nmpc.sampling_interval, tol=1e-10)
    assert out2['status'][0] == 1
    np.testing.assert_allclose(out2['u'][0], out['u'][0], atol=1e-6)
