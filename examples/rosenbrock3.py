#!/usr/bin/env python
"""BASELINE config 1 (mystic/examples/test_rosenbrock3b.py): DifferentialEvolutionSolver2, Rosenbrock
3-D, NP = 40, Best1Exp, VTR(1e-4) -- with rng='reference' the B200 solver makes the reference's own draws, so
`random.seed(123)` gives the reference's run: 39 generations, 1600 evaluations, bestEnergy 7.710567500336223e-05."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run in place

import random

from mystic_b200 import DifferentialEvolutionSolver2
from mystic_b200.models import rosen
from mystic_b200.monitors import VerboseMonitor
from mystic_b200.strategy import Best1Exp
from mystic_b200.termination import VTR

random.seed(123)
solver = DifferentialEvolutionSolver2(3, 40, rng='reference')
solver.SetRandomInitialPoints([0.] * 3, [2.] * 3)
solver.SetEvaluationLimits(generations=99999)
solver.SetGenerationMonitor(VerboseMonitor(10))
solver.Solve(rosen, VTR(1e-4), strategy=Best1Exp, CrossProbability=0.5, ScalingFactor=0.6)
print('generations %d, evaluations %d, bestEnergy %.17g' % (solver.generations, solver.evaluations, solver.bestEnergy))
print('bestSolution', list(solver.bestSolution))
assert (solver.generations, solver.evaluations) == (39, 1600)
