"""Build the CUDA shared library in-tree (mystic_b200/_lib/libmystic_b200.so) for sm_100a.

nvcc cross-compiles without a GPU.  The library is the product: there is no CPU fallback,
`mystic_b200._native` raises if it is missing.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, '_lib')
LIBPATH = os.path.join(LIBDIR, 'libmystic_b200.so')

SOURCES = ['de_abi.cu']
HEADERS = ['philox.cuh', 'de_device.cuh', 'de_kernels.cuh', 'de_tma.cuh', 'de_cheb_mma.cuh', 'de_small.cuh', os.path.join('..', '..', 'include', 'mystic_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    # host code computes reference-order constants (chebyshev abscissae): no FMA contraction
    '-Xcompiler', '-fPIC,-ffp-contract=off',
    '-shared',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build mystic_b200 CUDA library')
    return exe


def _stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile if the library is missing or older than its sources; return its path."""
    if not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    tmp = LIBPATH + '.tmp.%d' % os.getpid()
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', tmp]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed:\n' + proc.stdout)
    os.replace(tmp, LIBPATH)
    if verbose:
        print(proc.stdout, file=sys.stderr)
    return LIBPATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
