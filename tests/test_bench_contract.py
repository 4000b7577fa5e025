"""bench.py's reference arm runs on the host cores (no GPU needed): one JSON line with the contract's keys;
our arm refuses to run without a CUDA device (the engine has no CPU fallback)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py')] + list(args), cwd=REPO, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, universal_newlines=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    proc = _run('--impl', 'reference', '--steps', '2', '--warmup', '1', '--ref-np', '512')
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'].startswith('DE candidate-evals/sec')
    for k in ('value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['value'] > 0 and d['higher_is_better'] is True and d['dtype'] == 'f64' and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    proc = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                           '--warmup', '1', '--ref-np', '256'], cwd=REPO, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          universal_newlines=True, timeout=300)
    assert proc.returncode == 0 and proc.stdout.strip() == ''


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a GPU is present')
    proc = _run('--steps', '1', '--warmup', '1')
    assert proc.returncode != 0
    assert 'no CPU fallback' in (proc.stdout + proc.stderr)
