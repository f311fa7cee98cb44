"""Warp-per-instance tensor-pipe Riccati sweep (csrc/riccati6.cuh) on the CPU: the kernel body runs lane by lane on host
threads (tests/hostsim, warp_emu.cuh emulates mma.sync.m8n8k4.f64, shuffles and __syncwarp) and must reproduce the
one-thread-per-instance reference sweep (riccati_variant 0) inside a full interior-point solve - same iterates, same
iteration counts.  Covers the full-size fragment (nx=4, nu=2, CSTR) and the zero-padded one (nx=2, nu=1, LMPC)."""
import numpy as np
import pytest

from hilo_mpc_b200.codegen import OcpGenerator
from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
from runner import HostOcp
from test_hostsim_lmpc_mhe import _lmpc


def _compare(host, args, kw, variants=(2, 9)):
    ref = host.solve(*[a.copy() if isinstance(a, np.ndarray) else a for a in args], riccati_variant=0, **kw)
    assert (ref['status'] == 1).all()
    for v in variants:
        out = host.solve(*[a.copy() if isinstance(a, np.ndarray) else a for a in args], riccati_variant=v, **kw)
        assert (out['status'] == 1).all()
        assert (out['iters'] == ref['iters']).all()
        np.testing.assert_allclose(out['x'], ref['x'], atol=1e-12)
        np.testing.assert_allclose(out['u'], ref['u'], atol=1e-12)
        np.testing.assert_allclose(out['lam'], ref['lam'], atol=1e-11)


def test_cstr_sweep_matches_reference_sweep():
    N = 8
    nmpc = cstr_nmpc(N=N, integration_method='collocation', setup=False)
    nmpc._build_module = False
    nmpc.setup()
    spec = nmpc.spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    B = 5   # not a multiple of the 4 instances of a CTA
    x0 = cstr_initial_states(B, seed=3) / np.asarray(nmpc._xs_full)
    x = np.tile(spec.x_guess, (B, N + 1, 1))
    u = np.tile(spec.u_guess, (B, N, 1))
    xc = np.tile(spec.x_guess, (B, N, spec.degree))
    _compare(host, (x0, x, u, xc, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub), dict(dt=nmpc.sampling_interval))


def test_padded_fragment_nx2_nu1():
    mpc = _lmpc(10)
    spec = mpc.spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    x0 = np.array([[1., 1.], [3.5, 1.5], [-3., -1.]])
    B, N = 3, 10
    x = np.tile(spec.x_guess, (B, N + 1, 1))
    u = np.tile(spec.u_guess, (B, N, 1))
    _compare(host, (x0, x, u, None, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub), dict(dt=.5, tol=1e-10))
