"""The C-ABI library builds for sm_100a on a CPU-only box, loads, and exports every symbol include/hilo_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

from hilo_mpc_b200.build import build_core

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'hilo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(hb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build_core())
    names = declared_functions()
    assert 'hb_ocp_solve' in names and len(names) >= 8
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'symbols declared in include/hilo_b200.h but not exported: {missing}'
    lib.hb_version.restype = ctypes.c_int
    assert lib.hb_version() >= 1


def test_misuse_is_reported_not_thrown():
    lib = ctypes.CDLL(build_core())
    lib.hb_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    rc = lib.hb_ocp_open(b'/nonexistent/module.so', ctypes.byref(h))
    assert rc < 0 and b'hb_ocp_open' in lib.hb_last_error()
    assert lib.hb_ocp_solve(None, None) < 0


def test_product_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from hilo_mpc_b200 import _lib
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
    nmpc = cstr_nmpc(N=10)
    with pytest.raises(_lib.EngineUnavailable):
        nmpc.optimize(cstr_initial_states(1)[0])
