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


def _random_stage_data(rng, B, N, nx, nu, indefinite_every=0):
    A = np.eye(nx) * .9 + .1 * rng.normal(size=(B, N, nx, nx))
    Bm = .3 * rng.normal(size=(B, N, nx, nu))
    b = .1 * rng.normal(size=(B, N, nx))
    G = rng.normal(size=(B, N, nx + nu, nx + nu))
    H = np.swapaxes(G, -1, -2) @ G + np.eye(nx + nu)
    Q, S, R = H[..., :nx, :nx].copy(), H[..., nx:, :nx].copy(), H[..., nx:, nx:].copy()
    if indefinite_every:
        R.reshape(B * N, nu, nu)[7::indefinite_every, 0, 0] = -5.
    q, r = rng.normal(size=(B, N, nx)), rng.normal(size=(B, N, nu))
    qb, rb = rng.normal(size=(B, N, nx)), rng.normal(size=(B, N, nu))
    QN = np.tile(2. * np.eye(nx) + .1, (B, 1, 1))
    qN, qbN = rng.normal(size=(B, nx)), rng.normal(size=(B, nx))
    return A, Bm, b, Q, S, R, q, r, qb, rb, QN, qN, qbN


@pytest.mark.parametrize('problem', ['cstr', 'lmpc'])
def test_sweep_entry_point_on_random_stage_data(problem):
    """hb_ocp_riccati (module entry hbm_ocp_riccati) on random positive-definite stage data: the warp-per-instance sweep
    (host emulation) equals the reference sweep AND a plain NumPy Riccati recursion; with some indefinite R blocks the
    regularising slow path of both sweeps must agree to rounding."""
    from hilo_mpc_b200.stage_blocks import pack_stage_blocks, riccati_reference
    if problem == 'cstr':
        nmpc = cstr_nmpc(N=8, integration_method='collocation', setup=False)
        nmpc._build_module = False
        nmpc.setup()
        spec = nmpc.spec
    else:
        spec = _lmpc(10).spec
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    d = host.dims
    assert d.frag == 1 and d.block_a == 32 and d.block_h == 48
    rng = np.random.default_rng(11)
    B = 6
    for indefinite_every in (0, 13):
        A, Bm, b, Q, S, R, q, r, qb, rb, QN, qN, qbN = _random_stage_data(rng, B, d.N, d.nx, d.nu, indefinite_every)
        qa, qh = pack_stage_blocks(d, A, Bm, b, Q, S, R, q, r, qb, rb)
        qpN = np.concatenate([QN.reshape(B, -1), qN, qbN], axis=1)
        ref = host.riccati(qa, qh, qpN, riccati_variant=0)
        out = host.riccati(qa, qh, qpN, riccati_variant=2)
        assert (ref['status'] == 0).all() and (out['status'] == 0).all()
        scale = max(1., np.abs(ref['du']).max())
        tol = 1e-12 if not indefinite_every else 1e-9
        np.testing.assert_allclose(out['dx'], ref['dx'], atol=tol * scale)
        np.testing.assert_allclose(out['du'], ref['du'], atol=tol * scale)
        np.testing.assert_allclose(out['lam'], ref['lam'], atol=1e-12 * max(1., np.abs(ref['lam']).max()))
        np.testing.assert_allclose(out['edual'], ref['edual'], rtol=1e-12)
        if not indefinite_every:
            for i in (0, B - 1):
                dx, du = riccati_reference(A[i], Bm[i], b[i], Q[i], S[i], R[i], q[i], r[i], QN[i], qN[i])
                np.testing.assert_allclose(out['dx'][i], dx, atol=1e-11)
                np.testing.assert_allclose(out['du'][i], du, atol=1e-11)
            # adjoint of the multiplier-form gradient: lam_N = qbar_N, lam_k = qbar_k + A_k^T lam_{k+1}
            lam = qbN[0].copy()
            for k in range(d.N - 1, -1, -1):
                lam = qb[0, k] + A[0, k].T @ lam
            np.testing.assert_allclose(out['lam'][0, 0], lam, atol=1e-11)
