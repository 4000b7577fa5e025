"""Trajectory parity AT THE SIZE of BASELINE.json's configurations (round-1 verdict: the full-size cases were
only covered by property tests, a wrong donor index at row 40 000 would have passed).

The device (Philox mode, through the C ABI) and the C/OpenMP oracle port run generation 0 and two
generations from the same seed; after every generation the WHOLE population and the best member must be
bit-identical, the energies within 1e-12, and the counters equal -- with one documented exception, selections of
trials whose energy ties with the parent's to within that 1e-12 (see the loop below; at config 4 one row in two
million: a trial that swaps -1000 and +1000 on two columns has the parent's cost exactly, and only the rounding of
the penalty's running sum tells them apart).  The launch geometry is the one the
bench uses (148 persistent CTAs, every stage of the rings reused many times)."""
import numpy as np
import pytest

from oracle import c_port
from oracle import de_oracle as orc
from tests import _golden as G
from tests.test_gpu_parity import energy_close

pytestmark = pytest.mark.gpu


def _bits_equal(a, b, what):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, what
    same = a.view(np.uint64) == b.view(np.uint64)
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError('%s: %d of %d elements differ, first at %s: %r vs %r'
                             % (what, len(bad), a.size, tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])]))


def _run(NP, D, strategy, cost, extra, lo, hi, bounds, CR, F, pen, family, gens=2, tensor_cores=False, seed=0x5EED, check=None):
    from mystic_b200 import models, penalty
    from mystic_b200.engine import DeviceEngine
    if not c_port.available():
        pytest.skip('the C oracle port is not built')
    c_port.load().orc_set_num_threads(len(__import__('os').sched_getaffinity(0)))
    lo_v, hi_v = np.full(D, lo), np.full(D, hi)
    eng = DeviceEngine(NP, D)
    dev = getattr(models, cost)
    eng.set_cost(dev.device_id(tensor_cores, D), dev.params(tuple(extra)))
    eng.set_bounds(bounds, lo_v, hi_v)
    terms = ()
    if pen:
        def p0(x):
            return 0.0
        pf = penalty.quadratic_inequality(penalty.linear([pen['G']] * D, pen['h']), k=pen['k'], h=pen['hpow'])(p0)
        eng.set_penalty(*pf.device_terms(D))
        terms = [dict(ptype='quadratic_inequality', G=[pen['G']] * D, h=pen['h'], k=pen['k'], hpow=pen['hpow'])]
    eng.set_strategy(orc.STRATEGIES[strategy][0], F, CR)
    eng.set_rng_philox(seed)
    eng.init_population(seed, lo_v, hi_v)
    o = c_port.COracle(c_port.init_population(seed, NP, lo_v, hi_v), strategy, CR, F, cost, tuple(extra), lo_v, hi_v, bounds, terms,
                       trials=False)
    _bits_equal(eng.population(), o.population, 'initial population')
    accepted = 0
    ties = []
    for g in range(gens + 1):
        if g == 0:
            eng.eval_initial()
            o.step0()
        else:
            eng.step()
            o.step(seed=seed, keep_trials=False)
        best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
        pop = eng.population()
        en = eng.energies()
        # Near-ties.  Energies agree to 1e-12, not bit for bit (the device sums in its canonical tree, the reference
        # left to right), so where a trial's energy equals its parent's to within that tolerance the strict `<` of
        # the selection (differential_evolution.py:604) may legitimately fall the other way.  Such rows are verified
        # one by one -- the device kept exactly the parent or took a trial with the oracle's trial energy -- counted,
        # and the oracle's copy is aligned with the device so that later generations compare like with like.
        differ = np.flatnonzero((pop.view(np.uint64) != o.population.view(np.uint64)).any(axis=1))
        assert len(differ) <= 8, '%s generation %d: %d rows differ' % (cost, g, len(differ))
        for c in differ:
            assert g > 0
            parent, e_par = o.pop[o.cur ^ 1][c], o.en[o.cur ^ 1][c]
            te = o.trial_e[c]
            assert abs(te - e_par) <= 1e-12 * max(abs(e_par), 1.0), \
                '%s generation %d row %d differs and is no near-tie: parent %r trial %r' % (cost, g, c, e_par, te)
            if np.array_equal(pop[c], parent):       # the device rejected what the oracle accepted
                assert abs(en[c] - e_par) <= 1e-12 * max(abs(e_par), 1.0)
                n_acc += 1
            else:                                     # the device accepted what the oracle rejected
                assert np.array_equal(o.population[c], parent) and abs(en[c] - te) <= 1e-12 * max(abs(te), 1.0)
                n_acc -= 1
            o.population[c] = pop[c]
            o.popEnergy[c] = en[c]
            ties.append((g, int(c)))
        _bits_equal(pop, o.population, '%s generation %d population' % (cost, g))
        energy_close(en, o.popEnergy, pop if cost == 'griewangk' else None, cost, '%s generation %d energies' % (cost, g))
        if not differ.size:
            _bits_equal(best_x, o.best_x, 'generation %d best member' % g)
        # the port reports the member that became best in THIS generation (-1: none did), the device the member
        # that last became best
        assert (o.best_idx.value < 0 or differ.size or best_idx == o.best_idx.value) and n_inf == o.n_inf.value \
            and n_acc == o.n_acc.value, (g, best_idx, o.best_idx.value, n_inf, o.n_inf.value, n_acc, o.n_acc.value)
        assert abs(best_e - o.bestEnergy) <= 1e-12 * max(1.0, abs(o.bestEnergy))
        if g:
            accepted += n_acc
        del pop
    assert eng.kernel_family() in (family if isinstance(family, tuple) else (family,)), eng.kernel_family()
    if check is not None:
        check(eng)
    grid, block, smem = eng.launch_info()
    print('%s %s D=%d NP=%d: %d generations bit-identical to the C oracle port%s; %d trials accepted; grid %d x %d'
          % (strategy, cost, D, NP, gens, (' but for %d near-tie selections %s' % (len(ties), ties)) if ties else '', accepted,
             grid, block))
    eng.close()
    return accepted


def test_config2_rosenbrock_1024d_np65536_best1bin():
    """BASELINE config 2 on K2p (in place: 7 consumer groups x 2 warps + the producer warp) or K2a (double buffer, when
    more than 0.30 of the trials were accepted): every one of the 65 536 rows, every generation"""
    _run(65536, 1024, 'Best1Bin', 'rosen', (), -5., 5., orc.BOUNDS_CLAMP, 0.9, 0.8, None, ('tma_pool', 'tma_bin'))


def test_config2_shape_rand1exp():
    """the same rows under exponential crossover (K2e: parent row + donor windows by bulk copy)"""
    _run(65536, 1024, 'Rand1Exp', 'rosen', (), -5., 5., orc.BOUNDS_CLAMP, 0.9, 0.8, None, 'tma_exp')


def test_config4_corana_64d_np1m_rand1exp_penalty():
    """BASELINE config 4 on K2f: 1 048 576 rows, clamp + quadratic_inequality(sum(x) - 100)"""
    acc = _run(1048576, 64, 'Rand1Exp', 'corana', (), -1000., 1000., orc.BOUNDS_CLAMP, 0.9, 0.8,
               dict(G=1.0, h=100., k=100, hpow=5), 'rows_bulk')
    assert acc > 1000   # selections really happen at this size


def test_config5_griewangk_256d_np2m_best1bin():
    """BASELINE config 5, one GPU's shard (2 097 152 rows) on K2a with 32 one-warp groups"""
    _run(2097152, 256, 'Best1Bin', 'griewangk', (), -400., 400., orc.BOUNDS_CLAMP, 0.9, 0.8, None, 'tma_bin')


def test_config3_chebyshev8_dmma_np65536():
    """BASELINE config 3's kernel (FP64 DMMA contraction, M = 4096) against the port's Horner evaluation on 65 536
    rows: trial vectors / selections are compared through the populations (bit-exact), energies to 1e-12; the
    DMMA association differs from Horner's in the last ulps, so the number of abscissae within 1e-9 of the
    |p(x)| = 1 threshold (where a term may flip by up to 4) is reported -- 0 expected on random rows"""
    from mystic_b200 import models
    NP, D, M, seed = 65536, 9, 4096, 0xC3
    lo_v, hi_v = np.full(D, -100.), np.full(D, 100.)
    pop0 = c_port.init_population(seed, NP, lo_v, hi_v)
    xs = orc.chebyshev_points(M)
    sub = pop0[::64]
    px = np.stack([orc._polyeval(sub, x) for x in xs], axis=1)
    near = int((np.abs(np.abs(px) - 1.0) < 1e-9).sum())
    print('chebyshev8cost M=%d: %d of %d sampled points within 1e-9 of the threshold' % (M, near, px.size))
    _run(NP, D, 'Best1Exp', 'chebyshev8cost', (M,), -100., 100., orc.BOUNDS_NONE, 1.0, 0.9, None, 'dmma', tensor_cores=True,
         seed=seed)
