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
                                           ('sharded_griewangk.py', '1 GPUs')])
def test_example_runs(script, expect):
    proc = subprocess.run([sys.executable, os.path.join(REPO, 'examples', script)], cwd=REPO, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, universal_newlines=True, timeout=280,
                          env=dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', '')))
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert expect in proc.stdout, proc.stdout[-3000:]
