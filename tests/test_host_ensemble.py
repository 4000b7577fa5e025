"""Host logic of mystic_b200.ensemble (BuckshotSolver / LatticeSolver over nested DE2 runs that advance
together on the device).  The generation step comes from the test-only OracleEngine, whose islands
are stand-alone oracle shards; tests/test_gpu_islands.py proves the device's islands equal such shards."""
import random

import numpy as np
import pytest

from mystic_b200 import DifferentialEvolutionSolver2, models, termination
from mystic_b200.ensemble import BuckshotSolver, LatticeSolver
from tests._oracle_engine import OracleEngine


def test_buckshot_members_stop_independently_and_the_best_is_reported():
    random.seed(3)
    s = BuckshotSolver(3, 6, engine=OracleEngine, seed=42)
    s.SetNestedSolver(DifferentialEvolutionSolver2(3, 20, strategy='Best1Exp'))
    s.SetStrictRanges([-2.] * 3, [2.] * 3)
    s.SetEvaluationLimits(generations=60)
    s.Solve(models.rosen, termination.VTR(1e-3))
    iters, evals, energies = s._all_iters, s._all_evals, s._all_bestEnergy
    assert len(iters) == 6 and len(set(iters)) > 1                     # members stop at different generations
    msgs = s.Terminated(info=True, all=True)
    for g, e, fe, msg in zip(iters, evals, energies, msgs):
        assert (fe <= 1e-3 and msg.startswith('VTR')) or (g == 60 and msg.startswith('EvaluationLimits'))
        assert e <= 20 * (g + 1)
    assert s.bestEnergy == min(energies) and s._total_evals == sum(evals) and s._total_iters == sum(iters)
    k = energies.index(s.bestEnergy)
    assert list(s.bestSolution) == list(s._all_bestSolution[k]) and s.generations == iters[k]
    assert models is not None and abs(float(np.sum(100 * (s.bestSolution[1:] - s.bestSolution[:-1] ** 2) ** 2 +
                                                   (1 - s.bestSolution[:-1]) ** 2)) - s.bestEnergy) < 1e-12
    assert s.energy_history[-1] == s.bestEnergy and len(s.energy_history) == s.generations + 1


def test_a_member_is_exactly_a_stand_alone_shard_run():
    """member k of the ensemble = a DE2 run on rows [k*NP, (k+1)*NP) with the same draws"""
    random.seed(11)
    S, NP, D = 4, 16, 5
    ens = LatticeSolver(D, (2, 2, 1, 1, 1), engine=OracleEngine, seed=7)
    ens.SetNestedSolver(DifferentialEvolutionSolver2(D, NP, strategy='Rand1Exp', CrossProbability=0.8, ScalingFactor=0.6))
    ens.SetStrictRanges([-4.] * D, [4.] * D, tight=True)
    ens.SetEvaluationLimits(generations=9)
    ens.Solve(models.rastrigin, termination.VTR(-1.0))
    assert ens._npts == S and ens._all_iters == [9] * S
    # the lattice: centres of the 2 x 2 x 1 x 1 x 1 bins
    assert sorted(tuple(p) for p in ens._init_solution) == sorted([(-2., -2., 0., 0., 0.), (-2., 2., 0., 0., 0.),
                                                                  (2., -2., 0., 0., 0.), (2., 2., 0., 0., 0.)])
    pop0 = ens._engine._isl  # the oracle shards
    for k in range(S):
        e = OracleEngine(NP, D, S * NP, k * NP)
        e.set_cost(models.rastrigin.cost_id)
        e.set_bounds(2, np.full(D, -4.), np.full(D, 4.))
        e.set_strategy(2, 0.6, 0.8)            # Rand1Exp
        e.set_rng_philox(7)
        e.upload_population(ens._engine._pop[k * NP:(k + 1) * NP])
        e.eval_initial()
        for _ in range(9):
            e.step()
        assert np.array_equal(e.population(), pop0[k].population())
        assert e.read_result()[0] == ens._all_bestEnergy[k]


def test_nested_class_takes_the_ensemble_settings_and_errors_are_raised():
    random.seed(5)
    s = BuckshotSolver(2, 3, engine=OracleEngine, seed=1)
    s.SetNestedSolver(DifferentialEvolutionSolver2, NP=12)
    s.SetStrictRanges([0., 0.], [5., 5.])
    s.SetEvaluationLimits(generations=5)
    s.Solve(models.zimmermann)
    assert all(g <= 5 for g in s._all_iters) and s.Terminated() is True
    with pytest.raises(TypeError):
        s.SetNestedSolver(object())
    with pytest.raises(ValueError):
        s.SetNestedSolver(DifferentialEvolutionSolver2(3, 10))
    with pytest.raises(TypeError):
        BuckshotSolver(2, 2, engine=OracleEngine).Solve(lambda x: sum(x))     # python costs cannot run on the device
    with pytest.raises(ValueError):
        s.SetInitialPoints([[1., 1.]])


def test_zero_start_coordinates_are_spread_like_the_reference():
    """abstract_solver.py:499-502: x0[j] == 0 -> members drawn from [-radius, +radius] in that coordinate (a lattice
    over symmetric bounds with an odd number of bins has such centres; x0*(1 +- radius) would pin the coordinate)"""
    random.seed(5)
    ens = LatticeSolver(3, (2, 1, 1), engine=OracleEngine, seed=3)
    ens.SetNestedSolver(DifferentialEvolutionSolver2(3, 20, strategy='Best1Exp'))
    ens.SetStrictRanges([-2.] * 3, [2.] * 3)
    ens.SetEvaluationLimits(generations=200)
    ens.Solve(models.rosen, termination.VTR(1e-6))
    pop0 = ens._engine._pop
    for s in range(2):
        rows = pop0[s * 20:(s + 1) * 20]
        assert list(rows[0]) == list(ens._init_solution[s])               # member 0 is the start point itself
        assert np.abs(rows[1:, 1:]).max() <= 0.05 and np.abs(rows[1:, 1:]).min() > 0.0
        assert len(set(rows[1:, 1])) == 19                                 # not pinned to one value
    assert ens.bestEnergy < 1e-6                                           # every coordinate was explored


@pytest.mark.parametrize('make', [lambda: LatticeSolver(4, engine=OracleEngine, seed=1),
                                  lambda: BuckshotSolver(4, 1, engine=OracleEngine, seed=1)])
def test_an_ensemble_of_one_nested_solver_runs(make):
    """LatticeSolver(dim >= 4) with the default nbins = 8 has one bin per dimension, i.e. ONE nested solver"""
    random.seed(9)
    ens = make()
    assert ens._npts == 1
    ens.SetNestedSolver(DifferentialEvolutionSolver2(4, 24, strategy='Best1Bin'))
    ens.SetStrictRanges([-3.] * 4, [3.] * 4)
    ens.SetEvaluationLimits(generations=25)
    ens.Solve(models.rosen, termination.VTR(1e-9))
    assert len(ens._all_iters) == 1 and 1 <= ens.generations <= 25
    assert ens.bestEnergy == ens._all_bestEnergy[0] and len(ens.bestSolution) == 4
    assert ens.Terminated(info=True)


@pytest.mark.parametrize('term', [('VTR', 1e-3), ('ChangeOverGeneration', 1e-4, 8), ('NormalizedChangeOverGeneration', 1e-3, 6),
                                  ('EvaluationLimits', None, 900)])
def test_batched_termination_equals_the_per_generation_loop(term):
    """conditions the device can test run 64 generations per host round trip, every nested solver stopping on its own:
    same stop generations, messages, histories and best members as testing on the host after every generation"""
    runs = []
    for batch in (64, 7, 0):
        random.seed(21)
        ens = BuckshotSolver(3, 7, engine=OracleEngine, seed=5)
        ens._batch_generations = batch
        ens.SetNestedSolver(DifferentialEvolutionSolver2(3, 20, strategy='Best1Exp'))
        ens.SetStrictRanges([-2.] * 3, [2.] * 3)
        ens.SetEvaluationLimits(generations=150)
        cond = getattr(termination, term[0])(*term[1:])
        ens.Solve(models.rosen, cond)
        runs.append((ens._all_iters, ens._all_evals, ens._all_bestEnergy, [list(x) for x in ens._all_bestSolution],
                     ens.Terminated(info=True, all=True), [m.energy_history for m in ens._members]))
    assert runs[0] == runs[2] and runs[1] == runs[2]
    assert len(set(runs[0][0])) > 1 or term[0] == 'EvaluationLimits'      # the nested solvers stop at different generations
