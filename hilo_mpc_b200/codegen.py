"""CUDA code generation for transcribed problems.

The reference can emit C for its CasADi functions (``generate_c_code`` ``hilo_mpc/util/util.py:497-518`` ->
``ca.Importer``/``ca.external`` ``hilo_mpc/modules/base.py:328-358``) but the NLP itself is always interpreted by
the CasADi VM inside IPOPT.  Here the model/cost expressions are differentiated symbolically on the host and
emitted as ``__device__`` functions of a ``Prob`` struct; the generic batched solver templates in
``csrc/ocp_kernels.cuh`` are instantiated on it and compiled for sm_100a into one shared object per problem
(cached by content hash under ``hilo_mpc_b200/_gen/``; the built ``.so`` travels with the tree).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import symbolic as sx
from .ocp import OcpSpec
from .transcription import collocation_coefficients, tableau_for_order

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
GEN_DIR = os.path.join(_HERE, '_gen')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--shared',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']
QUAL = 'static HB_HD'


def nvcc_path() -> str:
    for cand in (os.environ.get('HB_NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found: the B200 engine needs the CUDA toolkit to build generated modules")


def _fmt(v: float) -> str:
    return sx._c_const(float(v))


def _table(name: str, values: Sequence[float]) -> str:
    vals = list(np.asarray(values, dtype=float).reshape(-1))
    if not vals:
        vals = [0.]
    body = ', '.join(_fmt(v) for v in vals)
    return f'  {QUAL} double {name}(int i) {{ constexpr double v[{len(vals)}] = {{{body}}}; return v[i]; }}'


def _indent(src: str, n: int = 2) -> str:
    pad = ' ' * n
    return '\n'.join(pad + ln for ln in src.split('\n'))


def cost_derivatives(total, ls_terms, rest, x, u, exact=False):
    """Gradient (exact) and Gauss-Newton Hessian blocks of a cost ``total = sum_i r_i^T W_i r_i + rest`` with respect
    to ``(x, u)``.  The least-squares part contributes ``2 J^T W J`` (exact whenever the residuals are affine, which
    is the reference's quadratic cost, modeling.py:283); ``rest`` contributes its exact Hessian."""
    xu = np.concatenate([sx.as_vector(x), sx.as_vector(u)])
    n = len(xu)
    g = sx.gradient(total, xu) if n else np.zeros(0, dtype=object)
    H = np.full((n, n), sx.Expr.const(0.), dtype=object)
    if exact:
        ls_terms, rest = [], total  # exact Hessian of the whole cost (what IPOPT gets from CasADi AD)
    for r, W in ls_terms:
        r = sx.as_vector(r)
        J = sx.jacobian(r, xu)
        W = np.asarray(W, dtype=float)
        H = H + 2. * (J.T @ (W @ J))
    if rest is not None and not (isinstance(rest, sx.Expr) and rest.is_const(0.)):
        H = H + sx.hessian(rest, xu)
    nx = len(sx.as_vector(x))
    return g[:nx], g[nx:], H[:nx, :nx], H[nx:, :nx], H[nx:, nx:]


class OcpGenerator:
    def __init__(self, spec: OcpSpec):
        self.spec = spec

    # ------------------------------------------------------------------ source
    def prob_struct(self) -> str:
        s = self.spec
        t, dt = [s.t], [s.dt]
        L: List[str] = ['struct Prob {']
        method = {'collocation': 0, 'erk': 1, 'discrete': 2}[s.method]
        deg = s.degree if s.method == 'collocation' else 1
        order = s.order if s.method == 'erk' else 1
        ints = dict(NX=s.nx, NU=s.nu, N=s.N, NC=s.ncol, NSP=s.nsp, NTP=s.ntp, NUOLD=s.nuold, NAP=s.nap,
                    METHOD=method, DEG=deg, ORDER=order, LAG_FROM=s.lag_from)
        for k, v in ints.items():
            L.append(f'  static constexpr int {k} = {int(v)};')
        bools = dict(OBJ_DISCRETE=(s.objective == 'discrete' or s.method == 'discrete'), FIXED_X0=s.fixed_x0,
                     HAS_GADD=s.gadd is not None, HAS_ARR=s.arr is not None,
                     EXACT_HESS=bool(s.exact_hessian) and s.method != 'erk')
        for k, v in bools.items():
            L.append(f'  static constexpr bool {k} = {"true" if v else "false"};')
        if s.method == 'collocation':
            B, C, D, tau = collocation_coefficients(s.degree, s.points)
        else:
            B, C, D, tau = np.zeros(2), np.zeros((2, 2)), np.zeros(2), np.zeros(2)
        L += [_table('CB', B), _table('CC', C), _table('CD', D), _table('CTAU', tau)]
        if s.method == 'erk':
            tab = tableau_for_order(s.order)
            L += [_table('RA', tab['A']), _table('RB', tab['b']), _table('RC', tab['c'])]
        else:
            L += [_table('RA', [0.]), _table('RB', [0.]), _table('RC', [0.])]
        L.append(_table('GADD', s.gadd if s.gadd is not None else [0.]))

        # ---- dynamics
        ins = [('x', s.x), ('u', s.u), ('sp', s.sp), ('t', t), ('dt', dt)]
        scal = ('t', 'dt')
        fx = sx.jacobian(s.ode, s.x)
        fu = sx.jacobian(s.ode, s.u)
        L.append(_indent(sx.Function('f_val', ins, [('f', s.ode)]).c_source(QUAL, scalar_inputs=scal)))
        L.append(_indent(sx.Function('f_jac', ins, [('f', s.ode), ('fx', fx), ('fu', fu)])
                         .c_source(QUAL, scalar_inputs=scal)))
        # Hessian of lam^T f with respect to (x, u): the constraint-curvature part of the Lagrangian Hessian
        lamf = sx.sym('lamf', s.nx)
        ex = bool(s.exact_hessian) and s.method != 'erk'
        if ex:
            Hf = sx.hessian(sx.dot(lamf, s.ode), np.concatenate([s.x, s.u]))
        else:
            Hf = np.full((s.nx + s.nu, s.nx + s.nu), sx.Expr.const(0.), dtype=object)
        f_hess = sx.Function('f_hess', ins + [('lamf', lamf)],
                             [('hxx', Hf[:s.nx, :s.nx]), ('hux', Hf[s.nx:, :s.nx]), ('huu', Hf[s.nx:, s.nx:])])
        L.append(_indent(f_hess.c_source(QUAL, scalar_inputs=scal)))
        # Call nodes (GP mean, MLP layer) are loops over a table; Jacobian and Hessian of a collocation point are evaluated
        # at the same point one after the other, so the Jacobian pass evaluates the loop once at order 2 and leaves its
        # outputs in a per-thread cache the Hessian pass reads (csrc/eval_quad.cuh, OcpEval): one walk instead of two
        cache = dict(mode='store', slots={}, size=0)
        if ex and s.method == 'collocation' and sx.gp_tables_used(list(s.ode)):
            src_jac = sx.Function('f_jac', ins, [('f', s.ode), ('fx', fx), ('fu', fu)]).c_source(
                QUAL, scalar_inputs=scal, name='f_jac_c', call_cache=cache)
            if cache['size'] > 0:
                L.append(_indent(src_jac))
                L.append(_indent(f_hess.c_source(QUAL, scalar_inputs=scal, name='f_hess_c',
                                                 call_cache=dict(mode='load', slots=cache['slots'], size=cache['size']))))
        L.append(f"  static constexpr int NCALLCACHE = {cache['size']};")
        # ---- Lagrange term
        lx, lu, lxx, lux, luu = cost_derivatives(s.lag, s.lag_ls, s.lag_rest, s.x, s.u, exact=ex)
        L.append(_indent(sx.Function('lag_raw', ins, [('l', [s.lag])]).c_source(QUAL, scalar_inputs=scal)))
        L.append(f'  {QUAL} double lag_val(const double* x, const double* u, const double* sp, double t, double dt) '
                 f'{{ double l; lag_raw(x, u, sp, t, dt, &l); return l; }}')
        L.append(_indent(sx.Function('lag_all', ins, [('l', [s.lag]), ('lx', lx), ('lu', lu), ('lxx', lxx),
                                                        ('lux', lux), ('luu', luu)])
                         .c_source(QUAL, scalar_inputs=scal)))
        # ---- input-change term (stage 0)
        ins_du = [('u', s.u), ('uo', s.uold)]
        du = s.lag_du if s.lag_du is not None else sx.Expr.const(0.)
        gdu = sx.gradient(du, s.u) if s.nu else np.zeros(0, dtype=object)
        Hdu = sx.hessian(du, s.u) if s.nu else np.zeros((0, 0), dtype=object)
        L.append(_indent(sx.Function('lagdu_raw', ins_du, [('l', [du])]).c_source(QUAL)))
        L.append(f'  {QUAL} double lagdu_val(const double* u, const double* uo) '
                 f'{{ double l; lagdu_raw(u, uo, &l); return l; }}')
        L.append(_indent(sx.Function('lagdu_all', ins_du, [('l', [du]), ('lu', gdu), ('luu', Hdu)]).c_source(QUAL)))
        # ---- Mayer term
        ins_m = [('x', s.x), ('tp', s.tp), ('t', t)]
        may = s.may if s.may is not None else sx.Expr.const(0.)
        Ex, _, Exx, _, _ = cost_derivatives(may, s.may_ls, s.may_rest, s.x, [], exact=ex)
        L.append(_indent(sx.Function('may_raw', ins_m, [('E', [may])]).c_source(QUAL, scalar_inputs=('t',))))
        L.append(f'  {QUAL} double may_val(const double* x, const double* tp, double t) '
                 f'{{ double E; may_raw(x, tp, t, &E); return E; }}')
        L.append(_indent(sx.Function('may_all', ins_m, [('E', [may]), ('Ex', Ex), ('Exx', Exx)])
                         .c_source(QUAL, scalar_inputs=('t',))))
        # ---- arrival cost (MHE)
        ap = s.ap if s.ap is not None else np.zeros(0, dtype=object)
        arr = s.arr if s.arr is not None else sx.Expr.const(0.)
        ins_a = [('x', s.x), ('ap', ap)]
        ax, _, axx, _, _ = cost_derivatives(arr, s.arr_ls, None if s.arr_ls else arr, s.x, [], exact=ex)
        L.append(_indent(sx.Function('arr_raw', ins_a, [('v', [arr])]).c_source(QUAL)))
        L.append(f'  {QUAL} double arr_val(const double* x, const double* ap) '
                 f'{{ double v; arr_raw(x, ap, &v); return v; }}')
        L.append(_indent(sx.Function('arr_all', ins_a, [('v', [arr]), ('g', ax), ('H', axx)]).c_source(QUAL)))
        L.append('};')
        return '\n'.join(L)

    def source(self) -> str:
        body = self.prob_struct()
        s = self.spec
        exprs = list(s.ode) + [s.lag, s.may if s.may is not None else sx.Expr.const(0.)]
        tables = sx.gp_tables_c_source(sx.gp_tables_used(np.array(exprs, dtype=object)))
        if tables:
            body = tables + '\n' + body
        # one namespace per problem: template instantiations (and their function-local statics, which the toolchain
        # emits as process-wide unique symbols) of different modules loaded into one process must not be merged
        ns = 'hbgen_' + hashlib.sha256(body.encode()).hexdigest()[:12]
        return '\n'.join([
            f'// Generated by hilo_mpc_b200/codegen.py for problem "{self.spec.name}" — do not edit.',
            '#include "ocp_module.cuh"',
            f'namespace {ns} {{',
            body,
            f'}}  // namespace {ns}',
            f'HB_DEFINE_OCP_MODULE({ns}::Prob)',
            ''])

    # ------------------------------------------------------------------ build
    def build(self, hostsim: bool = False, verbose: bool = False) -> str:
        """Write the module source under ``_gen/`` and compile it (nvcc for sm_100a; ``hostsim=True`` compiles the same
        kernel bodies for the host with g++ — used only by the CPU logic tests). Returns the path of the ``.so``."""
        src = self.source()
        return build_module(src, self.spec.name, hostsim=hostsim, verbose=verbose)


def _dep_hash() -> str:
    import glob
    h = hashlib.sha256()
    deps = sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.join(CSRC, '..', '..', 'include', 'hilo_b200.h')]
    for p in deps:
        if os.path.exists(p):
            with open(p, 'rb') as f:
                h.update(f.read())
    return h.hexdigest()


def build_module(src: str, name: str, hostsim: bool = False, verbose: bool = False,
                 extra_flags: Sequence[str] = ()) -> str:
    os.makedirs(GEN_DIR, exist_ok=True)
    # HB_NVCC_DEFINES='-DHB_STEPQ_MINB=3 ...': tuning experiments build separate, separately hashed modules
    tune = os.environ.get('HB_NVCC_DEFINES', '').split()
    digest = hashlib.sha256((src + _dep_hash() + ('host' if hostsim else 'sm100a') + ' '.join(tune)).encode()).hexdigest()[:16]
    safe = ''.join(c if c.isalnum() else '_' for c in name)
    base = os.path.join(GEN_DIR, f'{safe}_{digest}')
    so = base + ('_host.so' if hostsim else '.so')
    cu = base + '.cu'
    if os.path.exists(so) and os.path.getmtime(so) > 0:
        return so
    with open(cu, 'w') as f:
        f.write(src)
    tmp = so + f'.tmp{os.getpid()}'
    if hostsim:
        cmd = ['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-pthread', '-DHB_HOSTSIM', '-x', 'c++', '-I', CSRC, cu,
               '-o', tmp]
    else:
        cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + tune + ['-I', CSRC, cu, '-o', tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"Building generated module failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout, res.stderr, file=sys.stderr)
    os.replace(tmp, so)
    return so
