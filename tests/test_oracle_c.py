"""The C restatement (oracle/de_oracle.c) against the recordings of the unmodified reference
(replay mode, bit for bit) and against the numpy oracle in Philox mode."""
import numpy as np
import pytest

from oracle import c_port
from oracle import de_oracle as orc
from tests import _golden as G


@pytest.mark.parametrize('name', G.names())
def test_c_oracle_reproduces_reference(name):
    meta, z = G.load(name)
    lo, hi, mode = G.bounds_of(meta)
    o = c_port.COracle(z['pop0'], meta['strategy'], meta['CR_eff'], meta['F_eff'], meta['cost'],
                       G.extra_of(meta, z), lo, hi, mode, meta.get('penalty') or ())
    for g in range(meta['generations'] + 1):
        if g == 0:
            o.step0()
        else:
            o.step(replay=dict(donors=z['donors'][g - 1], nstart=z['nstart'][g - 1], uni=z['uni'][g - 1]))
        tag = '%s gen %d' % (name, g)
        G.assert_bits_equal(o.trial, z['trial'][g], tag + ' trial')
        G.assert_bits_equal(o.population, z['pop'][g], tag + ' pop')
        G.assert_bits_equal(o.best_x, z['bestX'][g], tag + ' bestX')
        if meta['cost'] in G.COMPENSATED_SUM_COSTS or meta['cost'] in G.NUMPY_EXP_COSTS:
            np.testing.assert_allclose(o.trial_e, z['trialE'][g], rtol=1e-14)
            np.testing.assert_allclose(o.popEnergy, z['popE'][g], rtol=1e-14)
        else:
            G.assert_bits_equal(o.trial_e, z['trialE'][g], tag + ' trialE')
            G.assert_bits_equal(o.popEnergy, z['popE'][g], tag + ' popE')
            G.assert_bits_equal(o.bestEnergy, z['bestE'][g], tag + ' bestE')
        assert o.evaluations == int(z['evals'][g])


@pytest.mark.parametrize('strategy,cost,D,NP,mode', [
    ('Best1Bin', 'rosen', 37, 50, orc.BOUNDS_CLAMP), ('Rand1Exp', 'corana', 64, 40, orc.BOUNDS_CLAMP),
    ('RandToBest1Bin', 'griewangk', 20, 33, orc.BOUNDS_REJECT), ('Rand2Exp', 'rosen', 130, 30, orc.BOUNDS_NONE),
    ('Best2Bin', 'chebyshev8cost', 9, 24, orc.BOUNDS_NONE)])
def test_c_oracle_philox_equals_numpy_oracle(strategy, cost, D, NP, mode):
    seed = 0xC0FFEE + D
    lo, hi = np.full(D, -4.), np.full(D, 4.)
    pop0 = orc.philox_init_population(seed, NP, lo * 1.2, hi * 1.2)
    G.assert_bits_equal(c_port.init_population(seed, NP, lo * 1.2, hi * 1.2), pop0, 'init')
    st = orc.DEState(pop0, lo if mode else None, hi if mode else None, mode)
    f = orc.make_cost(cost)
    o = c_port.COracle(pop0, strategy, 0.9, 0.8, cost, (), lo, hi, mode)
    orc.step0(st, f)
    o.step0()
    for g in range(1, 7):
        orc.step(st, strategy, 0.9, 0.8, f, philox=dict(seed=seed, offset=0))
        o.step(seed=seed)
        G.assert_bits_equal(o.trial, st.trialSolution, 'gen %d trial' % g)
        G.assert_bits_equal(o.population, st.population, 'gen %d pop' % g)
        G.assert_bits_equal(o.trial_e, st.trialEnergy, 'gen %d trialE' % g)
        G.assert_bits_equal(o.best_x, st.bestSolution, 'gen %d bestX' % g)
        assert o.bestEnergy == st.bestEnergy
