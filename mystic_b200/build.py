"""Build the CUDA shared library in-tree (mystic_b200/_lib/libmystic_b200.so) for sm_100a.

nvcc cross-compiles without a GPU.  The kernel instantiations are spread over several
translation units that compile in parallel.  The library is the product: there is no CPU
fallback, `mystic_b200._native` raises if it is missing.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, '_lib')
OBJDIR = os.path.join(LIBDIR, 'obj')
LIBPATH = os.path.join(LIBDIR, 'libmystic_b200.so')

SOURCES = ['de_abi.cu', 'inst_ldg_exp.cu', 'inst_ldg_bin.cu', 'inst_tma_512.cu', 'inst_tma_1024.cu', 'inst_tma_768.cu', 'inst_tma_exp.cu', 'inst_tma_exp_odd.cu', 'inst_tma_pool.cu', 'inst_tma_pool4.cu', 'inst_tma_pool1.cu', 'inst_small.cu', 'inst_subwarp.cu',
           'inst_rows_rosen.cu', 'inst_rows_corana.cu', 'inst_rows_griewangk.cu', 'inst_rows_bulk_rosen.cu', 'inst_rows_bulk_corana.cu', 'inst_rows_bulk_griewangk.cu', 'inst_datafit.cu', 'inst_separable.cu', 'probe.cu']
HEADERS = ['philox.cuh', 'de_device.cuh', 'de_kernels.cuh', 'de_misc.cuh', 'de_tma.cuh', 'de_tma_pool.cuh', 'de_tma_exp.cuh', 'de_cheb_mma.cuh', 'de_small.cuh', 'de_subwarp.cuh', 'de_rows.cuh', 'de_rows_bulk.cuh',
           os.path.join('..', '..', 'include', 'mystic_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    # host code computes reference-order constants (chebyshev abscissae): no FMA contraction
    '-Xcompiler', '-fPIC,-ffp-contract=off',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build mystic_b200 CUDA library')
    return exe


def _deps():
    return [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]


def _stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(args):
    src, obj, verbose, extra = args
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    return src, proc.returncode, proc.stdout


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile if the library is missing or older than its sources; return its path.
    `extra_flags` / `out` build a tuning variant (e.g. -DMB2_ROWS_CPT=1) next to the product library;
    MB2_LIBRARY=<path> makes mystic_b200._native load it."""
    if out is None and not force and not _stale():
        return LIBPATH
    os.makedirs(OBJDIR, exist_ok=True)
    tag = '%d.%s' % (os.getpid(), os.path.basename(out) if out else 'lib')
    jobs = [(os.path.join(CSRC, s), os.path.join(OBJDIR, s[:-3] + '.%s.o' % tag), verbose, tuple(extra_flags)) for s in SOURCES]
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        results = list(pool.map(_compile, jobs))
    objs = [j[1] for j in jobs]
    try:
        for src, rc, log in results:
            if verbose:
                print(log, file=sys.stderr)
            if rc != 0:
                raise RuntimeError('nvcc failed on %s:\n%s' % (src, log))
        target = out or LIBPATH
        tmp = target + '.tmp.%d' % os.getpid()
        link = subprocess.run([_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + objs + ['-o', tmp],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
        if link.returncode != 0:
            raise RuntimeError('link failed:\n' + link.stdout)
        os.replace(tmp, target)
    finally:
        for o in objs:
            if os.path.exists(o):
                os.remove(o)
    return out or LIBPATH


if __name__ == '__main__':
    # python -m mystic_b200.build [--force] [-v] [--out PATH -DNAME=VALUE ...]
    argv = sys.argv[1:]
    out = argv[argv.index('--out') + 1] if '--out' in argv else None
    print(build(force='--force' in argv, verbose='-v' in argv, extra_flags=[a for a in argv if a.startswith('-D')], out=out))
