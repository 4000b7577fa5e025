"""The scripts under examples/ run as they are (one GPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('script,expect', [('rosenbrock3.py', 'generations 39, evaluations 1600'),
                                           ('corana_penalty.py', 'bestEnergy'),
                                           ('chebyshev8_dmma.py', 'fitted'),
                                           ('sharded_griewangk.py', '1 GPUs'),
                                           ('lagrange_multipliers.py', 'constraints met')])
def test_example_runs(script, expect):
    proc = subprocess.run([sys.executable, os.path.join(REPO, 'examples', script)], cwd=REPO, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, universal_newlines=True, timeout=280,
                          env=dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', '')))
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert expect in proc.stdout, proc.stdout[-3000:]


def test_c_abi_demo_runs(tmp_path):
    """examples/c_abi_demo.c: the solver's generation loop driven from plain C through include/mystic_b200.h"""
    import shutil
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    from tests.test_abi_symbols import _build_c_demo
    exe = str(tmp_path / 'c_abi_demo')
    proc = _build_c_demo(exe)
    assert proc.returncode == 0, proc.stdout[-3000:]
    run = subprocess.run([exe], cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=120)
    assert run.returncode == 0, run.stdout[-3000:]
    assert 'generation 200' in run.stdout and 'expected error: unknown cost id 99' in run.stdout, run.stdout
