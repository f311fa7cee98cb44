"""Build the core C-ABI library ``libhilo_b200.so`` for sm_100a (in-tree, so the binary travels with the snapshot)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

from .codegen import CSRC, nvcc_path

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libhilo_b200.so')
CORE_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--shared',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def core_sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale(target: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = glob.glob(os.path.join(CSRC, '*.cu')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(_HERE, '..', 'include', 'hilo_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_core(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB_PATH):
        return LIB_PATH
    tmp = LIB_PATH + f'.tmp{os.getpid()}'
    cmd = [nvcc_path()] + CORE_FLAGS + ['-Xptxas', '-v'] * bool(verbose) + ['-I', CSRC] + core_sources() + \
          ['-o', tmp, '-ldl']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"Building libhilo_b200.so failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout, res.stderr, file=sys.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build_core(force='--force' in sys.argv, verbose='-v' in sys.argv))
