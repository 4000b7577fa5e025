#!/usr/bin/env python
"""BASELINE config 4 at a reduced size: Corana 64-D, Rand1Exp, SetStrictRanges(tight=True) and the penalty
quadratic_inequality(sum(x) - 100 <= 0) (mystic/examples2/g01_alt.py style), in-kernel Philox draws."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run in place

from mystic_b200 import DifferentialEvolutionSolver2
from mystic_b200.models import corana
from mystic_b200.penalty import linear, quadratic_inequality
from mystic_b200.strategy import Rand1Exp
from mystic_b200.termination import ChangeOverGeneration

D, NP = 64, 1 << 16


@quadratic_inequality(linear([1.] * D, 100.), k=100, h=5)     # condition(x) = G.x - h as data for the kernel
def penalty(x):
    return 0.0


solver = DifferentialEvolutionSolver2(D, NP, seed=4)          # rng='philox'
solver.SetRandomInitialPoints([-1000.] * D, [1000.] * D)
solver.SetStrictRanges([-1000.] * D, [1000.] * D, tight=True)
solver.SetPenalty(penalty)
solver.SetEvaluationLimits(generations=300)
solver.Solve(corana, ChangeOverGeneration(1e-8, 50), strategy=Rand1Exp, CrossProbability=0.9, ScalingFactor=0.8)
print(solver.Terminated(info=True))
print('generations %d, evaluations %d, bestEnergy %.6g, sum(x) = %.6g' % (
    solver.generations, solver.evaluations, solver.bestEnergy, sum(solver.bestSolution)))
