#!/usr/bin/env python
"""mystic/examples/test_ffit.py on the FP64 tensor cores: fit the coefficients of the 8th-order Chebyshev
polynomial (chebyshev8cost, 9 coefficients), Best1Exp CR = 1.0 F = 0.9, initial points in +-100.
tensor_cores=True runs the coefficient x abscissa contraction as DMMA (not bit-faithful to the Horner order of
the reference; the default, tensor_cores=False, is)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run in place

from mystic_b200 import DifferentialEvolutionSolver2
from mystic_b200.models import chebyshev8cost, CHEBYSHEV_TARGETS
from mystic_b200.strategy import Best1Exp
from mystic_b200.termination import VTR

ND, NP = 9, 1 << 14
solver = DifferentialEvolutionSolver2(ND, NP, seed=123, tensor_cores=True)
solver.SetRandomInitialPoints([-100.] * ND, [100.] * ND)
solver.SetEvaluationLimits(generations=2000)
solver.Solve(chebyshev8cost, VTR(0.01), strategy=Best1Exp, CrossProbability=1.0, ScalingFactor=0.9)
print(solver.Terminated(info=True))
print('generations %d, bestEnergy %.6g' % (solver.generations, solver.bestEnergy))
print('fitted  ', ['%.3f' % c for c in solver.bestSolution])
print('target  ', ['%.3f' % c for c in CHEBYSHEV_TARGETS[8]])
