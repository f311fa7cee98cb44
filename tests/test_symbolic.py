"""The symbolic layer replaces CasADi SX on this path: derivatives are checked against complex-step differentiation
(independent of the AD rules) and the emitted C source is compiled and compared with the NumPy evaluator."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from hilo_mpc_b200 import symbolic as sx
from hilo_mpc_b200.library import cstr_rhs


def _complex_step_jac(fun, x, *rest):
    n = len(x)
    cols = []
    for j in range(n):
        xp = np.asarray(x, dtype=complex).copy()
        xp[j] += 1e-30j
        cols.append(np.asarray(fun(xp, *rest)).imag / 1e-30)
    return np.stack(cols, axis=1)


def test_jacobian_matches_complex_step_cstr():
    x, u = sx.sym('x', 4), sx.sym('u', 2)
    f = np.array(cstr_rhs(x, u), dtype=object)
    F = sx.Function('F', [('x', x), ('u', u)], [('f', f), ('fx', sx.jacobian(f, x)), ('fu', sx.jacobian(f, u))])
    rng = np.random.default_rng(0)
    for _ in range(5):
        xv = np.array([1.2, .9, 125., 120.]) + rng.normal(size=4) * np.array([.2, .2, 5., 5.])
        uv = np.array([14., -3000.]) + rng.normal(size=2) * np.array([3., 500.])
        fv, fx, fu = F(xv, uv)
        ref = np.array(cstr_rhs(xv, uv, exp=np.exp))
        np.testing.assert_allclose(fv, ref, rtol=1e-13)
        np.testing.assert_allclose(fx, _complex_step_jac(lambda a, b: F(a, b)[0], xv, uv), rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(fu, _complex_step_jac(lambda b, a: F(a, b)[0], uv, xv), rtol=1e-11, atol=1e-12)


def test_elementary_functions_and_hessian():
    x = sx.sym('x', 3)
    e = sx.sin(x[0]) * sx.exp(-x[1]) + sx.tanh(x[2]) ** 2 + sx.atan(x[0] * x[2]) + sx.sqrt(1 + x[1] ** 2) / (2 + sx.cos(x[2])) \
        + x[0] ** 3 + sx.log(3 + x[1]) + x[2] ** 1.5
    g = sx.gradient(e, x)
    H = sx.hessian(e, x)
    F = sx.Function('F', [('x', x)], [('e', [e]), ('g', g), ('H', H)])
    xv = np.array([.3, .7, 1.1])
    ev, gv, Hv = F(xv)
    gcs = _complex_step_jac(lambda a: F(a)[0], xv)[0]
    np.testing.assert_allclose(gv, gcs, rtol=1e-12)
    Hcs = _complex_step_jac(lambda a: F(a)[1], xv)
    np.testing.assert_allclose(Hv, Hcs, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(Hv, Hv.T, atol=1e-15)


def test_substitute_and_simplification():
    x = sx.sym('x', 2)
    e = x[0] * 0. + 1. * x[1] + (x[0] - x[0])
    assert e is x[1]
    s = sx.substitute(np.array([x[0] ** 2 + x[1]], dtype=object), x, [2., sx.sym('y')])
    assert sx.symbols_of(s)[0].name == 'y'
    assert sx.evaluate(s, sx.symbols_of(s), [3.])[0] == 7.


def test_batched_numpy_evaluation():
    x = sx.sym('x', 2)
    F = sx.Function('F', [('x', x)], [('o', [x[0] * x[1], x[0] + 1.])])
    X = np.arange(10.).reshape(2, 5)
    out = F(X)
    np.testing.assert_allclose(out, np.stack([X[0] * X[1], X[0] + 1.]))


def test_c_source_compiles_and_matches():
    x, u = sx.sym('x', 4), sx.sym('u', 2)
    f = np.array(cstr_rhs(x, u), dtype=object)
    F = sx.Function('cstr', [('x', x), ('u', u)], [('f', f), ('fx', sx.jacobian(f, x))])
    src = '#include <math.h>\n' + F.c_source(qualifiers='').replace('__restrict__', '')
    with tempfile.TemporaryDirectory() as tmp:
        c, so = os.path.join(tmp, 'f.c'), os.path.join(tmp, 'f.so')
        open(c, 'w').write(src)
        subprocess.run(['gcc', '-O1', '-shared', '-fPIC', c, '-o', so, '-lm'], check=True)
        lib = ctypes.CDLL(so)
        xv = np.array([1.2, .9, 125., 120.])
        uv = np.array([14., -3000.])
        fo, jo = np.zeros(4), np.zeros(16)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.cstr(xv.ctypes.data_as(dp), uv.ctypes.data_as(dp), fo.ctypes.data_as(dp), jo.ctypes.data_as(dp))
        fv, fx = F(xv, uv)
        np.testing.assert_allclose(fo, fv, rtol=1e-14)
        np.testing.assert_allclose(jo.reshape(4, 4), fx, rtol=1e-13)


def test_gp_call_node_derivatives_and_generated_c_code(tmp_path):
    """A GP mean with many training points enters expressions as ONE call node (symbolic.gp_mean_node): derivatives map
    mean -> gradient -> Hessian outputs of the same table loop.  Checks them against complex-step differentiation and the
    emitted C (hb::gp_se_eval, csrc/gp_device.cuh, compiled for the host) against the NumPy evaluation."""
    import ctypes as C
    import os
    import subprocess
    rng = np.random.default_rng(0)
    n, d = 300, 2
    Xs, c, il = rng.normal(size=(n, d)), rng.normal(size=n), np.array([1 / .7, 1 / 15.])
    gid = sx.register_gp_table(Xs, c, il)
    x = sx.sym('x', 3)
    m = sx.gp_mean_node(gid, [x[0] * 2. + x[2], x[1] + x[0]])
    e = m * x[0] + sx.sin(m) + x[2] ** 2
    g, H = sx.gradient(e, x), sx.hessian(e, x)
    F = sx.Function('f', [('x', x)], [('e', [e]), ('g', g), ('H', H)])
    pt = np.array([.3, -.2, .1])
    ev, gv, Hv = F(pt)
    gc = np.array([F(pt + 1e-30j * np.eye(3)[j])[0].imag / 1e-30 for j in range(3)]).reshape(-1)
    Hc = np.array([F(pt + 1e-30j * np.eye(3)[j])[1].imag / 1e-30 for j in range(3)])
    np.testing.assert_allclose(gv, gc, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(Hv, Hc, rtol=1e-10, atol=1e-11)
    # one table, found from the expression; a second mean node at another query gets its own call
    assert sx.gp_tables_used(np.array([e], dtype=object)) == [gid]
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'hilo_mpc_b200', 'csrc')
    src = '\n'.join(['#define HB_HOSTSIM', '#include "ocp_kernels.cuh"', sx.gp_tables_c_source([gid]),
                     'extern "C" {', F.c_source(qualifiers=''), '}'])
    assert src.count('gp_se_eval<2, 2>') == 1          # value, gradient and Hessian share ONE loop over the table
    cu = tmp_path / 'gp_fun.cpp'
    cu.write_text(src)
    so = tmp_path / 'gp_fun.so'
    subprocess.run(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-DHB_HOSTSIM', '-I', csrc, str(cu), '-o', str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    out_e, out_g, out_H = np.zeros(1), np.zeros(3), np.zeros(9)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.f(dp(pt), dp(out_e), dp(out_g), dp(out_H))
    np.testing.assert_allclose(out_e[0], ev, rtol=1e-13)
    np.testing.assert_allclose(out_g, gv, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out_H.reshape(3, 3), Hv, rtol=1e-11, atol=1e-12)


def test_mlp_call_node_equals_the_unrolled_graph_and_its_generated_c_code(tmp_path):
    """A one-hidden-layer network enters expressions as call nodes (value -> Jacobian -> per-output Hessian outputs of one
    loop over the hidden units, csrc/gp_device.cuh::mlp1_eval).  Value, Jacobian and the multiplier-contracted Hessian
    must equal those of the fully unrolled graph (machine_learning.py:521-578 restated), in NumPy and in the emitted C."""
    import ctypes as C
    import os
    import subprocess
    from hilo_mpc_b200.ann import ANN, Layer
    rng = np.random.default_rng(3)
    nz, nh, no = 5, 20, 4
    net = ANN([f'z{i}' for i in range(nz)], [f'y{i}' for i in range(no)])
    net.add_layers(Layer.dense(nh, 'tanh'))
    net.setup()
    net.set_weights([rng.normal(size=(nh, nz)), rng.normal(size=(no, nh))], [rng.normal(size=nh), rng.normal(size=no)],
                    input_scaling=(rng.normal(size=nz), rng.uniform(.5, 2., nz)),
                    output_scaling=(rng.normal(size=no), rng.uniform(.5, 2., no)))
    z, lam = sx.sym('z', nz), sx.sym('l', no)
    funs = {}
    for mode in ('call', 'unroll'):
        net.symbolic_mode = mode
        y = net.predict(sx.sin(z) * 2. + z)          # composed with an inner map: chain rule through the node
        funs[mode] = sx.Function('f', [('z', z), ('l', lam)],
                                 [('y', y), ('J', sx.jacobian(y, z)), ('H', sx.hessian(sx.dot(lam, y), z))])
    assert funs['call'].n_nodes() < funs['unroll'].n_nodes() / 5
    zv, lv = rng.normal(size=nz), rng.normal(size=no)
    ref = funs['unroll'](zv, lv)
    out = funs['call'](zv, lv)
    for a, b in zip(out, ref):
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(np.asarray(ref[0]).reshape(-1), net.predict((np.sin(zv) * 2. + zv)[:, None])[:, 0], rtol=1e-12)
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'hilo_mpc_b200', 'csrc')
    ids = sx.gp_tables_used(np.concatenate([o for o in funs['call'].outputs]))
    assert len(ids) == 1 and ids[0].startswith('mlp:')
    src = '\n'.join(['#define HB_HOSTSIM', '#include "ocp_kernels.cuh"', sx.gp_tables_c_source(ids), 'extern "C" {',
                     funs['call'].c_source(qualifiers=''), '}'])
    assert src.count('mlp1_eval<') == 1
    cu = tmp_path / 'mlp_fun.cpp'
    cu.write_text(src)
    so = tmp_path / 'mlp_fun.so'
    subprocess.run(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-DHB_HOSTSIM', '-I', csrc, str(cu), '-o', str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    o_y, o_J, o_H = np.zeros(no), np.zeros(no * nz), np.zeros(nz * nz)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.f(dp(zv), dp(lv), dp(o_y), dp(o_J), dp(o_H))
    np.testing.assert_allclose(o_y, np.asarray(ref[0]).reshape(-1), rtol=1e-12)
    np.testing.assert_allclose(o_J.reshape(no, nz), ref[1], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(o_H.reshape(nz, nz), ref[2], rtol=1e-10, atol=1e-11)


def test_call_cache_shares_one_table_walk_between_jacobian_and_hessian_functions(tmp_path):
    """``c_source(call_cache=...)``: the Jacobian function evaluates the call groups at order 2 and stores their outputs,
    the Hessian function reads them instead of calling again; both must equal the plain functions (GP and MLP nodes)."""
    import ctypes as C
    import os
    import subprocess
    from hilo_mpc_b200.ann import ANN, Layer
    rng = np.random.default_rng(4)
    nz, nh, no = 3, 18, 2
    net = ANN([f'z{i}' for i in range(nz)], [f'y{i}' for i in range(no)])
    net.add_layers(Layer.dense(nh, 'tanh'))
    net.setup()
    net.set_weights([rng.normal(size=(nh, nz)), rng.normal(size=(no, nh))], [rng.normal(size=nh), rng.normal(size=no)])
    gid = sx.register_gp_table(rng.normal(size=(40, 2)), rng.normal(size=40), np.array([.7, 1.3]))
    x, lam = sx.sym('x', nz), sx.sym('l', no + 1)
    y = list(net.predict(sx.as_vector([x[0], sx.sin(x[1]), x[2] * x[0]]))) + [sx.gp_mean_node(gid, [x[1], x[2]]) * x[0]]
    y = np.array(y, dtype=object)
    jac = sx.Function('jac', [('x', x)], [('y', y), ('J', sx.jacobian(y, x))])
    hes = sx.Function('hes', [('x', x), ('l', lam)], [('H', sx.hessian(sx.dot(lam, y), x))])
    cache = dict(mode='store', slots={}, size=0)
    src_jc = jac.c_source(qualifiers='', name='jac_c', call_cache=cache)
    assert cache['size'] == (no + no * nz + no * nz * (nz + 1) // 2) + (1 + 2 + 3) and len(cache['slots']) == 2
    src_hc = hes.c_source(qualifiers='', name='hes_c', call_cache=dict(mode='load', slots=cache['slots'], size=cache['size']))
    assert 'mlp1_eval' not in src_hc and 'gp_se_eval' not in src_hc      # nothing is walked a second time
    ids = sx.gp_tables_used(np.concatenate([o for o in jac.outputs] + [o for o in hes.outputs]))
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'hilo_mpc_b200', 'csrc')
    src = '\n'.join(['#define HB_HOSTSIM', '#include "ocp_kernels.cuh"', sx.gp_tables_c_source(ids), 'extern "C" {',
                     jac.c_source(qualifiers=''), hes.c_source(qualifiers=''), src_jc, src_hc, '}'])
    cu = tmp_path / 'cache_fun.cpp'
    cu.write_text(src)
    so = tmp_path / 'cache_fun.so'
    subprocess.run(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-DHB_HOSTSIM', '-I', csrc, str(cu), '-o', str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    xv, lv = rng.normal(size=nz), rng.normal(size=no + 1)
    ny = no + 1
    y0, J0, H0 = np.zeros(ny), np.zeros(ny * nz), np.zeros(nz * nz)
    y1, J1, H1, cc = np.zeros(ny), np.zeros(ny * nz), np.zeros(nz * nz), np.zeros(cache['size'])
    lib.jac(dp(xv), dp(y0), dp(J0))
    lib.hes(dp(xv), dp(lv), dp(H0))
    lib.jac_c(dp(xv), dp(y1), dp(J1), dp(cc))
    lib.hes_c(dp(xv), dp(lv), dp(H1), dp(cc))
    np.testing.assert_allclose(y1, y0, rtol=1e-14)
    np.testing.assert_allclose(J1, J0, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(H1, H0, rtol=1e-13, atol=1e-14)
    ref = hes(xv, lv)
    np.testing.assert_allclose(H1.reshape(nz, nz), np.asarray(ref), rtol=1e-10, atol=1e-11)
