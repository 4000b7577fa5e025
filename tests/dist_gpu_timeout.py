#!/usr/bin/env python
"""A peer that stops stepping must not go unnoticed (the batched path never read the exchange's error
flag).  Two ranks on cuda:0 (gloo group, MB2_EXCHANGE_TIMEOUT_MS shortened): both run generation 0, then rank 1
stops while rank 0 first steps once (mb2_read_result must raise) -- or, with argv[1] == 'batch', calls Solve()
(mb2_run_result must raise).  Prints 'timeout detected' on rank 0."""
import os
import sys
import time

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from mystic_b200 import DifferentialEvolutionSolver2, models, termination  # noqa: E402
from mystic_b200._native import NativeError  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    batch = len(sys.argv) > 1 and sys.argv[1] == 'batch'
    torch.cuda.set_device(0)
    dist.init_process_group('gloo')
    D, NP = 12, 200
    s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=77, distributed=True, strategy='Best1Exp')
    s.SetRandomInitialPoints([-3.] * D, [3.] * D)
    s.SetStrictRanges([-3.] * D, [3.] * D, tight=True)
    s.SetEvaluationLimits(generations=50)
    s.SetTermination(termination.VTR(-1.0))
    s.Step(models.rosen)          # generation 0 with a working exchange
    assert s._engine.peer_exchange
    dist.barrier()
    if rank == 0:
        try:
            if batch:
                s.Solve()
            else:
                s.Step()
            print('NO ERROR RAISED', flush=True)
        except NativeError as e:
            assert 'did not deliver' in str(e), str(e)
            print('timeout detected after %d generations: %s' % (s.generations, e), flush=True)
    else:
        time.sleep(2.0)           # never joins the exchange
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)                   # the abandoned mailbox state is not worth a clean teardown


if __name__ == '__main__':
    main()
