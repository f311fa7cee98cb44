"""Config 5 ingredients on the CPU: MLP graph (a9) inside the model + path-following augmentation (mpc.py:1173-1204)."""
import numpy as np
import pytest

from hilo_mpc_b200.codegen import OcpGenerator
from hilo_mpc_b200.examples import bicycle_ann, robot_initial_states, robot_nn_nmpc
from hilo_mpc_b200.library import bicycle_rhs
from oracle.nmpc import FlatNlp
from runner import HostOcp


def test_mlp_graph_matches_numeric_forward_pass_and_true_dynamics():
    from hilo_mpc_b200 import symbolic as sx
    net = bicycle_ann(32)
    z = sx.sym('z', 5)
    F = sx.Function('mlp', [('z', z)], [('y', net.predict(z))])
    rng = np.random.default_rng(1)
    v, phi = rng.uniform(0., 6., 50), rng.uniform(-6., 6., 50)   # heading outside [-pi, pi]: the features are periodic
    a, delta = rng.uniform(-1., 1., 50), rng.uniform(-.5, .5, 50)
    Z = np.vstack([v, np.sin(phi), np.cos(phi), a, delta])
    np.testing.assert_allclose(F(Z), net.predict(Z), rtol=1e-12, atol=1e-13)
    true = np.array([bicycle_rhs([0., 0., v[i], phi[i]], [a[i], delta[i]], sin=np.sin, cos=np.cos, tan=np.tan,
                                 atan=np.arctan) for i in range(50)]).T
    # the surrogate has to be a smooth, faithful stand-in for the trained network
    assert np.corrcoef(net.predict(Z).ravel(), true.ravel())[0, 1] > .98


@pytest.mark.parametrize('hidden', [8, 16])
def test_path_following_nn_nmpc_host_kernels_match_oracle(hidden):
    # hidden = 8: the network is unrolled into the expression graph; hidden = 16: it enters as call nodes, and the
    # generated Jacobian / Hessian functions share one walk over the weight table through the call cache
    N = 5
    nmpc = robot_nn_nmpc(N=N, setup=False, hidden=hidden)
    nmpc._build_module = False
    nmpc.setup()
    spec = nmpc.spec
    assert spec.nx == 5 and spec.nu == 3  # theta and u_theta were appended
    src = OcpGenerator(spec).source()
    assert ('f_jac_c' in src and 'f_hess_c' in src and 'NCALLCACHE = 84' in src) == (hidden == 16)
    host = HostOcp(OcpGenerator(spec).build(hostsim=True))
    B = 2
    x0 = np.column_stack([robot_initial_states(B), np.zeros(B)])
    x = np.tile(spec.x_guess, (B, N + 1, 1))
    u = np.tile(spec.u_guess, (B, N, 1))
    xc = np.tile(spec.x_guess, (B, N, spec.degree))
    out = host.solve(x0, x, u, xc, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub, dt=1., tol=1e-9, max_iter=300)
    assert (out['status'] == 1).all()
    nlp = FlatNlp(spec, exact_hessian=True)
    same_basin = 0
    for b in range(B):
        res, xo, uo = nlp.solve(x0[b], dt=1., finish_barrier=True, tol=1e-9, v0=None)
        assert res.status == 1
        # The path-following problem is non-convex: from a cold start two different interior-point implementations may
        # legitimately end in different local minimisers (SURVEY.md §7 "hard parts").  Same basin -> tight comparison.
        if np.abs(out['u'][b] - uo.T).max() < 1e-3:
            same_basin += 1
            np.testing.assert_allclose(out['x'][b], xo.T, atol=5e-6)
            np.testing.assert_allclose(out['u'][b], uo.T, atol=5e-6)
        # Started next to the oracle's minimiser, the engine must converge to exactly that KKT point.
        v = res.x
        xw = np.stack([v[i] for i in nlp.x_ind])[None] + 1e-3
        uw = np.stack([v[i] for i in nlp.u_ind])[None]
        cw = np.stack([v[i] for i in nlp.c_ind])[None] + 1e-3
        o2 = host.solve(x0[b:b + 1], xw, uw, cw, spec.x_lb, spec.x_ub, spec.u_lb, spec.u_ub, dt=1., tol=1e-9,
                        max_iter=300)
        assert o2['status'][0] == 1
        np.testing.assert_allclose(o2['x'][0], xo.T, atol=5e-6)
        np.testing.assert_allclose(o2['u'][0], uo.T, atol=5e-6)
    assert same_basin >= 1
