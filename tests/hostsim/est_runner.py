"""TEST INFRASTRUCTURE: run a generated estimator module compiled for the host (-DHB_HOSTSIM); see runner.py."""
import ctypes as C

import numpy as np

from hilo_mpc_b200.estgen import EstCall, EstDims


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostEst:
    def __init__(self, so_path):
        self.lib = C.CDLL(so_path)
        self.dims = EstDims()
        self.lib.hbm_est_dims(C.byref(self.dims))
        self.lib.hbm_est_work_doubles.restype = C.c_size_t
        self.lib.hbm_est_work_doubles.argtypes = [C.c_int, C.c_int]

    def call(self, which, x, P=None, X=None, y=None, u=None, p=None, Q=None, R=None, w=None, v=None, unif=None, dt=1.,
             substeps=1, do_predict=True, do_update=True, alpha=1e-3, beta=2., kappa=0., t0=0.):
        d = self.dims
        f8 = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        x = f8(x).reshape(-1, d.nx).copy()
        B = x.shape[0]
        P = None if P is None else np.broadcast_to(f8(P).reshape(-1, d.nx, d.nx), (B, d.nx, d.nx)).copy()
        y = None if y is None else f8(y).reshape(B, d.nye)
        u = np.zeros(max(d.nu, 1)) if u is None else f8(u).reshape(-1)
        p = np.zeros(max(d.np, 1)) if p is None else f8(p).reshape(-1)
        us = d.nu if u.size == B * d.nu and B > 1 else 0
        ps = d.np if p.size == B * d.np and B > 1 else 0
        Q, R = f8(Q), f8(R)
        ns = 0
        if which == 'pf':
            X = f8(X).reshape(B, -1, d.nx).copy()
            ns = X.shape[1]
        elif which == 'ukf':
            X = np.zeros((B, 2 * d.nx + 1, d.nx)) if X is None else f8(X).reshape(B, 2 * d.nx + 1, d.nx).copy()
        ypred = np.zeros((B, d.nye))
        status = np.zeros(B, dtype=np.int32)
        q = np.zeros((B, max(ns, 1)))
        idx = np.zeros((B, max(ns, 1)), dtype=np.int64)
        work = np.zeros(self.lib.hbm_est_work_doubles(B, ns))
        w, v, unif = f8(w), f8(v), f8(unif)
        call = EstCall(B=B, substeps=substeps, n_particles=ns, do_predict=int(do_predict), do_update=int(do_update),
                       dt=dt, t0=t0, alpha=alpha, beta=beta, kappa=kappa, x=_p(x), P=_p(P), X=_p(X), y=_p(y), u=_p(u),
                       u_stride=us, p=_p(p), p_stride=ps, Q=_p(Q), R=_p(R), y_pred=_p(ypred), w=_p(w), v=_p(v),
                       unif=_p(unif), q=_p(q), idx=_p(idx), work=_p(work), work_doubles=work.size, status=_p(status),
                       stream=None)
        fn = getattr(self.lib, f'hbm_{which}_step')
        fn.argtypes = [C.POINTER(EstCall)]
        rc = fn(C.byref(call))
        if rc != 0:
            raise RuntimeError(f'hbm_{which}_step returned {rc}')
        return dict(x=x, P=P, X=X, y_pred=ypred, status=status, q=q, idx=idx, work=work)
