"""Comparison with golden vectors of the REAL reference (CasADi + IPOPT / qpOASES), produced by ``tools/run_reference.py`` on
a box that has them.  Skipped while the files do not exist (this container has no CasADi)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet (needs the reference + CasADi: tools/run_reference.py)")
    return np.load(path)


def test_oracle_matches_reference_nmpc_golden():
    g = _load('reference_nmpc_cstr.npz')
    from hilo_mpc_b200.examples import cstr_nmpc
    from oracle.nmpc import FlatNlp
    nmpc = cstr_nmpc(N=int(g['N']), setup=False)
    nmpc._build_module = False
    nmpc.setup()
    xs, us = np.asarray(nmpc._xs_full), np.asarray(nmpc._us_full)
    nlp = FlatNlp(nmpc.spec, exact_hessian=True)
    for b in range(g['x0'].shape[0]):
        res, xo, uo = nlp.solve(g['x0'][b] / xs, dt=float(g['dt']), finish_barrier=True, tol=1e-10)
        assert res.status == 1 and int(g['status'][b]) == 1
        np.testing.assert_allclose(xo * xs[:, None], g['x_pred'][b], rtol=0, atol=1e-6 * np.abs(xs).max())
        np.testing.assert_allclose(uo[:, 0] * us, g['u0'][b], rtol=1e-6, atol=1e-6 * np.abs(us).max())


@pytest.mark.gpu
def test_engine_matches_reference_nmpc_golden(cuda_device):
    g = _load('reference_nmpc_cstr.npz')
    import torch
    from hilo_mpc_b200.examples import cstr_nmpc
    nmpc = cstr_nmpc(N=int(g['N']), solver_options={'ipopt.tol': 1e-10})
    u0 = nmpc.optimize(torch.tensor(g['x0'], device=cuda_device))
    assert (nmpc.solver_status_code == 1).all()
    us = np.asarray(nmpc._us_full)
    np.testing.assert_allclose(u0.cpu().numpy(), g['u0'], rtol=1e-6, atol=1e-6 * np.abs(us).max())


@pytest.mark.gpu
def test_engine_matches_reference_lmpc_golden(cuda_device):
    g = _load('reference_lmpc_di.npz')
    from test_hostsim_lmpc_mhe import double_integrator
    from hilo_mpc_b200 import LMPC
    mpc = LMPC(double_integrator())
    mpc.Q, mpc.R, mpc.horizon = np.eye(2), 1, 10
    mpc.set_box_constraints(x_lb=[-5, -5], x_ub=[5, 5], u_lb=[-1], u_ub=[1])
    mpc.setup()
    for k in range(len(g['u'])):
        uk = float(np.asarray(mpc.optimize(g['x'][k])).reshape(-1)[0])
        assert abs(uk - g['u'][k]) < 1e-6
