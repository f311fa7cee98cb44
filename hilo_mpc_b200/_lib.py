"""ctypes binding of ``libhilo_b200.so`` (the C ABI in ``include/hilo_b200.h``).

There is no CPU fallback: if the library cannot be loaded, or no CUDA device is present when a solve is requested,
the call raises.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from .build import LIB_PATH, build_core

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class OcpDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('nx', 'nu', 'N', 'nc', 'nsp', 'ntp', 'nuold', 'nap', 'stage_doubles',
                                       'elim_doubles', 'method', 'fixed_x0', 'n_state', 'block_a', 'block_h', 'frag')]


class OcpRiccatiCall(C.Structure):   # mirrors struct hb_ocp_riccati_call, include/hilo_b200.h
    _fields_ = [('B', C.c_int), ('riccati_variant', C.c_int), ('qa', C.c_void_p), ('qh', C.c_void_p),
                ('qpN', C.c_void_p), ('dx', C.c_void_p), ('du', C.c_void_p), ('lam', C.c_void_p),
                ('edual', C.c_void_p), ('status', C.c_void_p), ('work', C.c_void_p), ('work_doubles', C.c_size_t),
                ('stream', C.c_void_p)]


class OcpCall(C.Structure):
    _fields_ = [('B', C.c_int), ('n_alpha', C.c_int), ('max_iter', C.c_int), ('check_every', C.c_int),
                ('tol', C.c_double), ('acc_tol', C.c_double), ('mu_init', C.c_double), ('mu_min', C.c_double),
                ('time0', C.c_double), ('dt', C.c_double),
                ('x0', C.c_void_p), ('sp', C.c_void_p), ('sp_stride', C.c_longlong),
                ('tp', C.c_void_p), ('tp_stride', C.c_longlong),
                ('uold', C.c_void_p), ('uold_stride', C.c_longlong), ('ap', C.c_void_p), ('ap_stride', C.c_longlong),
                ('xlb', C.c_void_p), ('xub', C.c_void_p), ('ulb', C.c_void_p), ('uub', C.c_void_p),
                ('x', C.c_void_p), ('u', C.c_void_p), ('xc', C.c_void_p), ('lam', C.c_void_p),
                ('status', C.c_void_p), ('iters', C.c_void_p), ('stats', C.c_void_p),
                ('work', C.c_void_p), ('work_doubles', C.c_size_t), ('stream', C.c_void_p),
                ('n_iterations_launched', _IP), ('n_kernel_launches', C.POINTER(C.c_longlong)),
                ('kernel_ms', _DP), ('kernel_count', C.POINTER(C.c_longlong)), ('riccati_variant', C.c_int), ('warm_start', C.c_int), ('mu_init_warm', C.c_double), ('eval_variant', C.c_int), ('first_check', C.c_int),
                ('n_iterations_needed', C.POINTER(C.c_int)), ('step_variant', C.c_int), ('no_compaction', C.c_int),
                ('regularization', C.c_int), ('regularization_eps', C.c_double), ('regularization_mu', C.c_double), ('shift_warm_start', C.c_int),
                ('kernel_ms_by_iter', C.c_void_p), ('kernel_count_by_iter', C.c_void_p)]


_lib: Optional[C.CDLL] = None


class EngineUnavailable(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load (building on first use if the toolkit is present) the core library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    try:
        build_core()   # no-op unless the library is missing or older than its sources (nvcc needed only then)
    except Exception as exc:  # noqa: BLE001
        if not os.path.exists(path):
            raise EngineUnavailable(f"libhilo_b200.so is missing and could not be built: {exc}") from exc
        raise EngineUnavailable(f"libhilo_b200.so is older than its sources and could not be rebuilt: {exc}") from exc
    try:
        lib = C.CDLL(path)
    except OSError as exc:
        raise EngineUnavailable(f"Could not load {path}: {exc}. The B200 engine has no CPU fallback.") from exc
    lib.hb_version.restype = C.c_int
    lib.hb_last_error.restype = C.c_char_p
    lib.hb_device_info.argtypes = [C.c_int, _IP, C.POINTER(C.c_size_t), _IP, _IP]
    lib.hb_ocp_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.hb_ocp_get_dims.argtypes = [C.c_void_p, C.POINTER(OcpDims)]
    lib.hb_ocp_work_doubles.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.hb_ocp_work_doubles.restype = C.c_size_t
    lib.hb_ocp_solve.argtypes = [C.c_void_p, C.POINTER(OcpCall)]
    lib.hb_ocp_riccati.argtypes = [C.c_void_p, C.POINTER(OcpRiccatiCall)]
    lib.hb_ocp_riccati_work_doubles.argtypes = [C.c_void_p, C.c_int]
    lib.hb_ocp_riccati_work_doubles.restype = C.c_size_t
    lib.hb_ocp_close.argtypes = [C.c_void_p]
    _register_optional(lib)
    _lib = lib
    return lib


ABI_VERSION = 8   # HB_ABI_VERSION of include/hilo_b200.h this binding was written against


def _register_optional(lib: C.CDLL) -> None:
    """argtypes of the model-independent entry points, exactly as declared in include/hilo_b200.h."""
    vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    fit = [i, i, vp, vp, vp, d, d, d, vp, vp, vp, vp, vp, vp, vp]
    pred = [i, i, i, vp, vp, vp, d, d, d, vp, vp, vp, vp, vp, vp]
    sigs = {
        'hb_pf_resample': ([i, i, vp, vp, vp, vp, vp], i),
        'hb_lmpc_work_doubles': ([i, i, i], sz),
        'hb_lmpc_solve': ([i, i, i, i] + [vp] * 13 + [d, i, vp, sz, vp], i),
        'hb_gp_fit_work_doubles': ([i], sz),
        'hb_gp_predict_work_doubles': ([i, i], sz),
        'hb_gp_fit': (fit, i),
        'hb_gp_predict': (pred, i),
        'hb_gp_fit_kernel': ([i] + fit, i),
        'hb_gp_predict_kernel': ([i] + pred, i),
        'hb_gp_lml_grad_work_doubles': ([i], sz),
        'hb_gp_lml_grad': ([i, i, i, vp, vp, d, d, vp, vp, vp, vp, vp], i),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    v = lib.hb_version()
    if v != ABI_VERSION:
        raise EngineUnavailable(f"libhilo_b200.so reports ABI version {v}, this package expects {ABI_VERSION}: "
                                f"rebuild it (python -m hilo_mpc_b200.build --force)")


def last_error() -> str:
    return load().hb_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}: {last_error()}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise EngineUnavailable("No CUDA device is visible. hilo_mpc_b200 runs its solvers on a B200 GPU and has "
                                "no CPU fallback.")
    return torch
