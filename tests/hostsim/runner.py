"""TEST INFRASTRUCTURE: run a generated OCP module that was compiled for the host (``-DHB_HOSTSIM``).

The product path always runs the sm_100a build of the same kernel bodies through ``libhilo_b200.so``; this runner
exists so that the *logic* of the kernels (linearisation, Riccati, line search, barrier updates) can be compared with
the oracle on a machine without a GPU.  Nothing under ``hilo_mpc_b200/`` imports it.
"""
import ctypes as C

import numpy as np


class OcpDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('nx', 'nu', 'N', 'nc', 'nsp', 'ntp', 'nuold', 'nap', 'stage_doubles',
                                       'elim_doubles', 'method', 'fixed_x0', 'n_state', 'block_a', 'block_h', 'frag')]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class OcpCall(C.Structure):
    _fields_ = [('B', C.c_int), ('n_alpha', C.c_int), ('max_iter', C.c_int), ('check_every', C.c_int),
                ('tol', C.c_double), ('acc_tol', C.c_double), ('mu_init', C.c_double), ('mu_min', C.c_double),
                ('time0', C.c_double), ('dt', C.c_double),
                ('x0', _DP), ('sp', _DP), ('sp_stride', C.c_longlong), ('tp', _DP), ('tp_stride', C.c_longlong),
                ('uold', _DP), ('uold_stride', C.c_longlong), ('ap', _DP), ('ap_stride', C.c_longlong),
                ('xlb', _DP), ('xub', _DP), ('ulb', _DP), ('uub', _DP),
                ('x', _DP), ('u', _DP), ('xc', _DP), ('lam', _DP), ('status', _IP), ('iters', _IP), ('stats', _DP),
                ('work', _DP), ('work_doubles', C.c_size_t), ('stream', C.c_void_p),
                ('n_iterations_launched', _IP), ('n_kernel_launches', C.POINTER(C.c_longlong)),
                ('kernel_ms', _DP), ('kernel_count', C.POINTER(C.c_longlong)), ('riccati_variant', C.c_int), ('warm_start', C.c_int), ('mu_init_warm', C.c_double), ('eval_variant', C.c_int), ('first_check', C.c_int),
                ('n_iterations_needed', C.POINTER(C.c_int)), ('step_variant', C.c_int), ('no_compaction', C.c_int),
                ('regularization', C.c_int), ('regularization_eps', C.c_double), ('regularization_mu', C.c_double), ('shift_warm_start', C.c_int),
                ('kernel_ms_by_iter', C.c_void_p), ('kernel_count_by_iter', C.c_void_p)]


def _dp(a):
    return a.ctypes.data_as(_DP)


class OcpRiccatiCall(C.Structure):
    _fields_ = [('B', C.c_int), ('riccati_variant', C.c_int), ('qa', C.c_void_p), ('qh', C.c_void_p),
                ('qpN', C.c_void_p), ('dx', C.c_void_p), ('du', C.c_void_p), ('lam', C.c_void_p),
                ('edual', C.c_void_p), ('status', C.c_void_p), ('work', C.c_void_p), ('work_doubles', C.c_size_t),
                ('stream', C.c_void_p)]


class HostOcp:
    def riccati(self, qa, qh, qpN, riccati_variant=0):
        """The sweep alone through the module's hbm_ocp_riccati entry (host build)."""
        d = self.dims
        qa, qh, qpN = (np.ascontiguousarray(v, dtype=np.float64) for v in (qa, qh, qpN))
        B = qa.shape[0]
        dx, du, lam = np.zeros((B, d.N + 1, d.nx)), np.zeros((B, d.N, d.nu)), np.zeros((B, d.N + 1, d.nx))
        edual, status = np.zeros(B), np.zeros(B, dtype=np.int32)
        self.lib.hbm_ocp_riccati_work_doubles.restype = C.c_size_t
        self.lib.hbm_ocp_riccati_work_doubles.argtypes = [C.c_int]
        work = np.zeros(self.lib.hbm_ocp_riccati_work_doubles(B))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        call = OcpRiccatiCall(B=B, riccati_variant=riccati_variant, qa=vp(qa), qh=vp(qh), qpN=vp(qpN), dx=vp(dx),
                              du=vp(du), lam=vp(lam), edual=vp(edual), status=vp(status), work=vp(work),
                              work_doubles=work.size, stream=None)
        self.lib.hbm_ocp_riccati.argtypes = [C.POINTER(OcpRiccatiCall)]
        rc = self.lib.hbm_ocp_riccati(C.byref(call))
        if rc != 0:
            raise RuntimeError(f"hbm_ocp_riccati returned {rc}")
        return dict(dx=dx, du=du, lam=lam, edual=edual, status=status)

    def __init__(self, so_path):
        self.lib = C.CDLL(so_path)
        self.lib.hbm_ocp_dims.argtypes = [C.POINTER(OcpDims)]
        self.lib.hbm_ocp_work_doubles.restype = C.c_size_t
        self.lib.hbm_ocp_work_doubles.argtypes = [C.c_int, C.c_int]
        self.lib.hbm_ocp_solve.argtypes = [C.POINTER(OcpCall)]
        self.lib.hbm_ocp_solve.restype = C.c_int
        self.dims = OcpDims()
        self.lib.hbm_ocp_dims(C.byref(self.dims))

    def solve(self, x0, x, u, xc, xlb, xub, ulb, uub, dt, time0=0., sp=None, tp=None, uold=None, ap=None, tol=1e-8,
              max_iter=200, n_alpha=8, mu_init=.1, riccati_variant=0, regularization=0, regularization_eps=1e-4, regularization_mu=0., warm_start=0, mu_init_warm=1e-4,
              shift_warm_start=0, work=None):
        d = self.dims
        B = x.shape[0]
        f8 = lambda a, shape: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shape))
        x0 = f8(x0, (B, d.nx))
        x = f8(x, (B, d.N + 1, d.nx)).copy()
        u = f8(u, (B, d.N, d.nu)).copy()
        xc = f8(xc if d.nc else np.zeros((B, d.N, 0)), (B, d.N, d.nc)).copy()
        if xc.size == 0:
            xc = np.zeros(1)

        def opt(a, per, n):
            if a is None:
                return np.zeros(max(1, per)), 0
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            if a.size == per:
                return a.reshape(-1), 0
            return a.reshape(B, -1), per
        spv, sps = opt(sp, d.N * d.nsp, B)
        tpv, tps = opt(tp, d.ntp, B)
        uov, uos = opt(uold, d.nuold, B)
        apv, aps = opt(ap, d.nap, B)
        lam = np.zeros((B, d.N + 1, d.nx))
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B + 1, dtype=np.int32)
        stats = np.zeros((B, d.n_state))
        nwork = self.lib.hbm_ocp_work_doubles(B, n_alpha)
        if work is None:
            work = np.zeros(nwork)   # pass the `work` of a previous solve back in for a warm start with multipliers
        launched = C.c_int(0)
        b = [f8(v, (-1,)) for v in (xlb, xub, ulb, uub)]
        call = OcpCall(B=B, n_alpha=n_alpha, max_iter=max_iter, check_every=1, tol=tol, acc_tol=1e-6, mu_init=mu_init,
                       mu_min=tol / 10., time0=time0, dt=dt, x0=_dp(x0), sp=_dp(spv), sp_stride=sps, tp=_dp(tpv),
                       tp_stride=tps, uold=_dp(uov), uold_stride=uos, ap=_dp(apv), ap_stride=aps,
                       xlb=_dp(b[0]), xub=_dp(b[1]), ulb=_dp(b[2]), uub=_dp(b[3]), x=_dp(x), u=_dp(u), xc=_dp(xc),
                       lam=_dp(lam), status=status.ctypes.data_as(_IP), iters=iters.ctypes.data_as(_IP),
                       stats=_dp(stats), work=_dp(work), work_doubles=nwork, stream=None,
                       n_iterations_launched=C.pointer(launched), riccati_variant=riccati_variant,
                       regularization=regularization, regularization_eps=regularization_eps,
                       regularization_mu=regularization_mu, warm_start=warm_start, mu_init_warm=mu_init_warm,
                       shift_warm_start=shift_warm_start)
        rc = self.lib.hbm_ocp_solve(C.byref(call))
        if rc != 0:
            raise RuntimeError(f"hbm_ocp_solve returned {rc}")
        return dict(x=x, u=u, xc=xc.reshape(B, d.N, d.nc) if d.nc else None, lam=lam, status=status,
                    iters=iters[:B], stats=stats, launched=launched.value, work=work, n_alpha=n_alpha)
