"""Device-side solve engine for one transcribed problem: owns the generated module handle and the batch's tensors.

This is what ``NMPC._optimize`` / ``MHE.estimate`` call instead of ``self._solver(x0=..., lbx=..., p=...)``
(``hilo_mpc/modules/controller/mpc.py:712-742``): one call solves the NLP of every instance in the batch.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from .codegen import OcpGenerator
from .ocp import OcpSpec

INF = 1e20  # "no bound" marker understood by the kernels (|bound| >= 1e19)


def _finite(a):
    a = np.asarray(a, dtype=np.float64).copy()
    a[np.isposinf(a)] = INF
    a[np.isneginf(a)] = -INF
    return a


class OcpEngine:
    def __init__(self, spec: OcpSpec, device=None, build: bool = True):
        self.spec = spec
        self.module_path = OcpGenerator(spec).build() if build else None
        self._handle = None
        self._device = device
        self._batch = 0
        self._t: Dict[str, object] = {}
        self.dims = None
        self.last_launched = 0

    # ------------------------------------------------------------------ lazy device setup
    def _open(self):
        if self._handle is not None:
            return
        torch = _lib.require_cuda()
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.hb_ocp_open(self.module_path.encode(), C.byref(h)), 'hb_ocp_open')
        self._handle = h
        d = _lib.OcpDims()
        _lib.check(lib.hb_ocp_get_dims(h, C.byref(d)), 'hb_ocp_get_dims')
        self.dims = d
        s = self.spec
        assert (d.nx, d.nu, d.N, d.nc) == (s.nx, s.nu, s.N, s.ncol), "generated module does not match the problem"
        if self._device is None:
            self._device = torch.device('cuda', torch.cuda.current_device())
        dev = self._device
        f8 = dict(dtype=torch.float64, device=dev)
        self._t['xlb'] = torch.tensor(_finite(s.x_lb), **f8)
        self._t['xub'] = torch.tensor(_finite(s.x_ub), **f8)
        self._t['ulb'] = torch.tensor(_finite(s.u_lb), **f8)
        self._t['uub'] = torch.tensor(_finite(s.u_ub), **f8)
        self._t['dummy'] = torch.zeros(8, **f8)

    def _alloc(self, B: int, n_alpha: int):
        torch = _lib.require_cuda()
        if self._batch == B and self._t.get('n_alpha') == n_alpha:
            return
        d, dev = self.dims, self._device
        f8 = dict(dtype=torch.float64, device=dev)
        s = self.spec
        t = self._t
        t['x'] = torch.tensor(np.tile(s.x_guess, (B, d.N + 1, 1)), **f8)
        t['u'] = torch.tensor(np.tile(s.u_guess, (B, d.N, 1)), **f8)
        t['xc'] = torch.tensor(np.tile(s.x_guess, (B, d.N, s.degree if d.nc else 0)).reshape(B, d.N, d.nc), **f8) \
            if d.nc else torch.zeros(1, **f8)
        t['lam'] = torch.zeros(B, d.N + 1, d.nx, **f8)
        t['status'] = torch.zeros(B, dtype=torch.int32, device=dev)
        t['iters'] = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        t['stats'] = torch.zeros(B, d.n_state, **f8)
        nwork = _lib.load().hb_ocp_work_doubles(self._handle, B, n_alpha)
        t['work'] = torch.empty(nwork, **f8)
        t['n_alpha'] = n_alpha
        self._batch = B
        self._warm = False

    def reset_guess(self, v0: Optional[dict] = None):
        """Forget the warm start (next solve starts from the tiled guess, mpc.py:1465-1506) or install ``v0``."""
        self._batch = 0 if v0 is None else self._batch
        self._solved_once = False
        if v0 is not None:
            for k in ('x', 'u', 'xc'):
                if k in v0 and v0[k] is not None:
                    self._t[k].copy_(v0[k])

    # ------------------------------------------------------------------ solve
    def solve(self, x0, dt: float, time0: float = 0., sp=None, tp=None, uold=None, ap=None, tol: float = 1e-8,
              acc_tol: float = 1e-6, max_iter: int = 3000, n_alpha: int = 8, mu_init: float = .1,
              mu_min: Optional[float] = None, check_every: int = 4, stream=None, kernel_ms=None, kernel_count=None,
              riccati_variant: int = 2, warm_start: int = 0, mu_init_warm: float = 1e-4,
              eval_variant: int = 1, step_variant: int = 2, adaptive_poll: bool = True, compaction: bool = True,
              regularization: int = 0, regularization_eps: float = 1e-4, regularization_mu: float = 0., shift_warm_start: int = 0,
              kernel_ms_by_iter=None,
              kernel_count_by_iter=None):
        """``x0``: CUDA tensor ``[B, nx]`` (scaled). Optional ``sp [B|1, N, nsp]``, ``tp``, ``uold``, ``ap`` tensors.
        Returns dict of device tensors: x [B,N+1,nx], u [B,N,nu], xc, lam, status, iters, stats."""
        torch = _lib.require_cuda()
        self._open()
        lib = _lib.load()
        d = self.dims
        x0 = x0.to(device=self._device, dtype=torch.float64).contiguous()
        B = x0.shape[0]
        self._alloc(B, n_alpha)
        t = self._t

        def opt(a, per):
            if a is None or per == 0:
                return t['dummy'], 0
            a = a.to(device=self._device, dtype=torch.float64).contiguous()
            if a.numel() == per:
                return a, 0
            if a.numel() != per * B:
                raise ValueError(f"Parameter tensor has {a.numel()} elements, expected {per} or {per * B}")
            return a, per
        spv, sps = opt(sp, d.N * d.nsp)
        tpv, tps = opt(tp, d.ntp)
        uov, uos = opt(uold, d.nuold)
        apv, aps = opt(ap, d.nap)
        launched = C.c_int(0)
        needed = C.c_int(0)
        warm = int(bool(warm_start) and getattr(self, '_solved_once', False) and self._solved_batch == B)
        n_launch = C.c_longlong(0)
        st = torch.cuda.current_stream(self._device) if stream is None else stream
        call = _lib.OcpCall(
            B=B, n_alpha=n_alpha, max_iter=max_iter, check_every=check_every, tol=tol, acc_tol=acc_tol,
            mu_init=mu_init, mu_min=(tol / 10. if mu_min is None else mu_min), time0=time0, dt=dt,
            x0=x0.data_ptr(), sp=spv.data_ptr(), sp_stride=sps, tp=tpv.data_ptr(), tp_stride=tps,
            uold=uov.data_ptr(), uold_stride=uos, ap=apv.data_ptr(), ap_stride=aps,
            xlb=t['xlb'].data_ptr(), xub=t['xub'].data_ptr(), ulb=t['ulb'].data_ptr(), uub=t['uub'].data_ptr(),
            x=t['x'].data_ptr(), u=t['u'].data_ptr(), xc=t['xc'].data_ptr(), lam=t['lam'].data_ptr(),
            status=t['status'].data_ptr(), iters=t['iters'].data_ptr(), stats=t['stats'].data_ptr(),
            work=t['work'].data_ptr(), work_doubles=t['work'].numel(), stream=st.cuda_stream,
            n_iterations_launched=C.pointer(launched), n_kernel_launches=C.pointer(n_launch),
            kernel_ms=None if kernel_ms is None else kernel_ms.ctypes.data_as(_lib._DP),
            kernel_count=None if kernel_count is None else kernel_count.ctypes.data_as(C.POINTER(C.c_longlong)),
            riccati_variant=riccati_variant,
            warm_start=warm, first_check=(getattr(self, 'last_needed', 0) if (warm and adaptive_poll) else 0),
            n_iterations_needed=C.pointer(needed),
            mu_init_warm=mu_init_warm, eval_variant=eval_variant, step_variant=step_variant,
            no_compaction=int(not compaction), regularization=int(regularization),
            regularization_eps=float(regularization_eps), regularization_mu=float(regularization_mu),
            shift_warm_start=int(shift_warm_start),
            kernel_ms_by_iter=None if kernel_ms_by_iter is None else kernel_ms_by_iter.ctypes.data,
            kernel_count_by_iter=None if kernel_count_by_iter is None else kernel_count_by_iter.ctypes.data)
        _lib.check(lib.hb_ocp_solve(self._handle, C.byref(call)), 'hb_ocp_solve')
        self.last_launched = launched.value
        self.last_needed = needed.value
        self._solved_once, self._solved_batch = True, B
        self.last_kernel_launches = n_launch.value
        self.total_kernel_launches = getattr(self, 'total_kernel_launches', 0) + n_launch.value
        self._keep = (x0, spv, tpv, uov, apv)  # keep inputs alive until the stream has consumed them
        return dict(x=t['x'], u=t['u'], xc=t['xc'] if d.nc else None, lam=t['lam'], status=t['status'],
                    iters=t['iters'][:B], stats=t['stats'])

    def riccati(self, qa, qh, qpN, riccati_variant: int = 2):
        """The Riccati sweep alone (``hb_ocp_riccati``): ``qa [B, N, block_a]``, ``qh [B, N, block_h]`` (see
        ``stage_blocks.pack_stage_blocks``), ``qpN [B, nx*nx + 2 nx]`` CUDA tensors -> dict of ``dx``, ``du``, ``lam``,
        ``edual``, ``status`` (device tensors)."""
        torch = _lib.require_cuda()
        self._open()
        lib = _lib.load()
        d = self.dims
        f8 = dict(dtype=torch.float64, device=self._device)
        qa, qh, qpN = (v.to(**f8).contiguous() for v in (qa, qh, qpN))
        B = qa.shape[0]
        out = dict(dx=torch.zeros(B, d.N + 1, d.nx, **f8), du=torch.zeros(B, d.N, d.nu, **f8),
                   lam=torch.zeros(B, d.N + 1, d.nx, **f8), edual=torch.zeros(B, **f8),
                   status=torch.zeros(B, dtype=torch.int32, device=self._device))
        work = torch.empty(lib.hb_ocp_riccati_work_doubles(self._handle, B), **f8)
        call = _lib.OcpRiccatiCall(B=B, riccati_variant=riccati_variant, qa=qa.data_ptr(), qh=qh.data_ptr(),
                                   qpN=qpN.data_ptr(), dx=out['dx'].data_ptr(), du=out['du'].data_ptr(),
                                   lam=out['lam'].data_ptr(), edual=out['edual'].data_ptr(),
                                   status=out['status'].data_ptr(), work=work.data_ptr(), work_doubles=work.numel(),
                                   stream=torch.cuda.current_stream(self._device).cuda_stream)
        _lib.check(lib.hb_ocp_riccati(self._handle, C.byref(call)), 'hb_ocp_riccati')
        self._keep_ric = (qa, qh, qpN, work)
        return out

    def close(self):
        if self._handle is not None:
            _lib.load().hb_ocp_close(self._handle)
            self._handle = None

    def __del__(self):  # noqa: D105
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
