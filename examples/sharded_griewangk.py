#!/usr/bin/env python
"""BASELINE config 5 at a reduced size, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/sharded_griewangk.py
Each rank owns a contiguous block of candidates; donors are shard-local; the generation best is exchanged every
step (stores into peer memory fused into the argmin epilogue, or an NCCL all-gather with MB2_EXCHANGE=nccl)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run in place

import torch
import torch.distributed as dist

from mystic_b200 import DifferentialEvolutionSolver2
from mystic_b200.models import griewangk
from mystic_b200.strategy import Best1Bin
from mystic_b200.termination import VTR

world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
D, NP = 256, (1 << 17) * world
solver = DifferentialEvolutionSolver2(D, NP, seed=5, distributed=world > 1)
solver.SetRandomInitialPoints([-400.] * D, [400.] * D)
solver.SetStrictRanges([-400.] * D, [400.] * D, tight=True)
solver.SetEvaluationLimits(generations=200)
solver.Solve(griewangk, VTR(1e-3), strategy=Best1Bin, CrossProbability=0.9, ScalingFactor=0.8)
if int(os.environ.get('RANK', '0')) == 0:
    print(solver.Terminated(info=True))
    print('%d GPUs, NP = %d: generations %d, bestEnergy %.6g' % (world, NP, solver.generations, solver.bestEnergy))
if world > 1:
    dist.destroy_process_group()
