"""The C-ABI library builds, loads and exports every symbol include/mystic_b200.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, 'include', 'mystic_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mb2_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from mystic_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'missing export %s' % n
    from mystic_b200 import _native
    assert sorted(_native.EXPORTS) == names
    assert _native.lib.mb2_abi_version() == 2


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under mystic_b200/ may reference it"""
    pkg = os.path.join(REPO, 'mystic_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'de_oracle' not in text.replace('oracle/de_oracle.py', ''), f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from mystic_b200 import DifferentialEvolutionSolver2, models
    s = DifferentialEvolutionSolver2(3, 8)
    s.SetRandomInitialPoints([-1.] * 3, [1.] * 3)
    with pytest.raises(RuntimeError):
        s.Step(models.rosen)


def test_missing_library_fails_loudly():
    """no CPU fallback: without the CUDA library the package cannot drive a solver (the import of the native
    layer raises and says how to build it)"""
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MB2_LIBRARY=os.path.join(repo, 'mystic_b200', '_lib', 'no_such_library.so'))
    code = ('import mystic_b200._native as n\n')
    proc = subprocess.run([sys.executable, '-c', code], cwd=repo, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          universal_newlines=True, timeout=120)
    assert proc.returncode != 0
    assert 'is missing' in proc.stdout and 'no CPU fallback' in proc.stdout, proc.stdout[-2000:]


def test_python_constants_match_the_header():
    """the cost ids of mystic_b200.models and the kernel-family names of the engine are the enums of
    include/mystic_b200.h"""
    import os
    import re
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'mystic_b200.h')).read()
    enums = dict((k, int(v)) for k, v in re.findall(r'\b(MB2_[A-Z0-9_]+)\s*=\s*(\d+)', hdr))
    for name, value in vars(models).items():
        if name.startswith('COST_') and isinstance(value, int):
            assert enums['MB2_' + name] == value, name
    costs = [c for c in vars(models).values() if isinstance(c, models.DeviceCost)]
    assert len(set(c.__name__ for c in costs)) >= 26
    fam = dict((k[len('MB2_KERNEL_'):].lower(), v) for k, v in enums.items() if k.startswith('MB2_KERNEL_'))
    names = DeviceEngine.KERNEL_FAMILIES
    assert len(names) == len(fam)
    alias = {'warp_per_row': 'warp_per_row', 'tma_bin': 'tma_bin', 'tma_exp': 'tma_exp', 'rows': 'rows', 'rows_per_warp': 'rows_per_warp',
             'thread_per_row': 'thread_per_row', 'dmma': 'dmma', 'islands': 'islands', 'rows_bulk': 'rows_bulk', 'none': 'none', 'tma_pool': 'tma_pool'}
    for i, n in enumerate(names):
        assert fam[alias[n]] == i, n


def _build_c_demo(out):
    import os
    import subprocess
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    cmd = ['gcc', '-std=c99', '-O2', '-Wall', '-Werror', os.path.join(repo, 'examples', 'c_abi_demo.c'), '-I' + os.path.join(repo, 'include'),
           '-I' + os.path.join(cuda, 'include'), '-L' + os.path.join(repo, 'mystic_b200', '_lib'), '-lmystic_b200',
           '-L' + os.path.join(cuda, 'lib64'), '-lcudart', '-Wl,-rpath,' + os.path.join(repo, 'mystic_b200', '_lib'), '-o', out]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=300)


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/mystic_b200.h compiles as C99 and examples/c_abi_demo.c links against the library with nothing but
    libcudart: the boundary carries no C++ or torch types"""
    import shutil
    if shutil.which('gcc') is None:
        import pytest
        pytest.skip('no gcc')
    from mystic_b200 import build as _build
    _build.build()
    proc = _build_c_demo(str(tmp_path / 'c_abi_demo'))
    assert proc.returncode == 0, proc.stdout[-3000:]
