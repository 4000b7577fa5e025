#!/usr/bin/env python
"""Augmented-Lagrangian outer loop with mystic's lagrange penalties (mystic/penalty.py:459-601) on the device:
minimise Rosenbrock 8-D subject to sum(x) = 6 and x0 + x1 <= 1.5.  Every outer iteration runs one DE solve with the
current multipliers, stores the condition values at the solution (`penalty.store`) and steps the penalty
(`penalty.iter`): the decorators' own bookkeeping, a few scalars on the host; the kernel sees k*h**n and the
accumulated multiplier of each term."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run in place

from mystic_b200 import DifferentialEvolutionSolver2
from mystic_b200.models import rosen
from mystic_b200.penalty import lagrange_equality, lagrange_inequality, linear
from mystic_b200.strategy import Best1Exp
from mystic_b200.termination import ChangeOverGeneration

D, NP = 8, 4096


@lagrange_inequality(linear([1., 1.] + [0.] * (D - 2), 1.5), k=20, h=3)
@lagrange_equality(linear([1.] * D, 6.), k=20, h=3)
def penalty(x):
    return 0.0


x = None
for outer in range(4):
    solver = DifferentialEvolutionSolver2(D, NP, seed=11 + outer)
    solver.SetRandomInitialPoints([-2.] * D, [2.] * D)
    solver.SetStrictRanges([-2.] * D, [2.] * D, tight=True)
    solver.SetPenalty(penalty)
    solver.SetEvaluationLimits(generations=400)
    solver.Solve(rosen, ChangeOverGeneration(1e-10, 40), strategy=Best1Exp)
    x = solver.bestSolution
    print('outer %d: generations %3d, energy %.6f, sum(x) - 6 = %+.2e, x0 + x1 - 1.5 = %+.2e'
          % (outer, solver.generations, solver.bestEnergy, sum(x) - 6., x[0] + x[1] - 1.5))
    penalty.store(x)      # the condition values at this solution feed the multipliers of the next iteration
    penalty.iter()
print('multipliers after %d iterations: stored conditions %s' % (penalty.iteration(), [t['y'] for t in penalty.terms]))
assert abs(sum(x) - 6.) < 2e-2 and x[0] + x[1] - 1.5 < 2e-2
print('constraints met')
