"""The multi-GPU best exchange (stores into peer memory fused into the argmin epilogue,
mb2_exchange_* in include/mystic_b200.h) on a ONE-GPU box: two processes share cuda:0, map each
other's mailboxes through CUDA IPC and run tests/dist_gpu_check.py, which compares every rank's
shard with the oracle's emulation of all shards (populations and best solution bit-exact)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_exchange_between_two_processes_on_one_gpu():
    env = dict(os.environ, MB2_DIST_ONE_GPU='1', MB2_EXCHANGE='peer')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29571', os.path.join(REPO, 'tests', 'dist_gpu_check.py')]
    proc = subprocess.run(cmd, cwd=REPO, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True,
                          timeout=240)
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert 'multi-gpu parity ok (world 2)' in proc.stdout, proc.stdout[-3000:]
    assert proc.stdout.count('batched Solve == stepwise Solve: True') == 2, proc.stdout[-3000:]


@pytest.mark.parametrize('mode', ['step', 'batch'])
def test_a_silent_peer_fails_the_step_and_the_batched_solve(mode):
    """the epilogue's wait times out, the stale mailbox is not merged, and both mb2_read_result (per-generation
    path) and mb2_run_result (device-resident loop) report it"""
    env = dict(os.environ, MB2_EXCHANGE='peer', MB2_EXCHANGE_TIMEOUT_MS='300')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29573' if mode == 'step' else '29575', os.path.join(REPO, 'tests', 'dist_gpu_timeout.py'), mode]
    proc = subprocess.run(cmd, cwd=REPO, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True,
                          timeout=240)
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert 'timeout detected' in proc.stdout and 'NO ERROR RAISED' not in proc.stdout, proc.stdout[-3000:]
