#!/usr/bin/env python
"""torchrun: the two sharded cases of bench.shard_parity with the comparison spelled out per rank"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mystic_b200 import DifferentialEvolutionSolver2, models, termination
from oracle import de_oracle as orc
from tests.dist_gpu_check import emulate

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
for strategy, cost, D, NP, lo, hi in (('Best1Bin', 'rosen', 256, 64 * world + 259, -5., 5.), ('Rand1Exp', 'corana', 64, 96 * world + 5, -1000., 1000.),
                                      ('Best1Bin', 'rosen', 256, 64 * 8 + 3, -5., 5.)):
    seed, gens = 0xD15C0 + D, 4
    s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=seed, distributed=True, strategy=strategy)
    s.SetRandomInitialPoints([lo] * D, [hi] * D)
    s.SetStrictRanges([lo] * D, [hi] * D, tight=True)
    s.SetTermination(termination.VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 6)
    s.Step(getattr(models, cost))
    hists = [s.bestEnergy]
    for _ in range(gens):
        s.Step()
    engs, hist = emulate(world, NP, D, seed, strategy, getattr(models, cost).cost_id, (orc.BOUNDS_CLAMP, np.full(D, lo), np.full(D, hi)), gens)
    mine = engs[rank]
    pop_ok = np.array_equal(s.population, mine.population())
    bad = np.flatnonzero((s.population != mine.population()).any(axis=1))
    print('rank %d %s/%s D=%d NP=%d: population %s (bad rows %s), best %s, history %s | dev %s | orc %s | family %s' % (
        rank, strategy, cost, D, NP, pop_ok, bad[:8].tolist(), np.array_equal(np.array(s.bestSolution), mine.state.bestSolution),
        np.allclose(np.array(s.energy_history), np.array(hist), rtol=1e-12, atol=0), ['%.17g' % e for e in s.energy_history],
        ['%.17g' % e for e in hist], s._engine.kernel_family()), flush=True)
    del s
dist.destroy_process_group()
