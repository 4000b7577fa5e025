"""How a generation's rows leave the staged kernels (include/mystic_b200.h, mb2_set_commit_mode).

The reference keeps one population and a trialSolution list and copies the accepted trials over their parents
when the whole generation has been evaluated (mystic/differential_evolution.py:603-609).  In place, the kernels of
rows of 16 dimensions and more (K2a / K2e / K2f / K2d / K2b) write only the accepted trials to the other buffer and de_commit_accepted_kernel moves
them over their parents; with the double buffer every selected row is written and the buffers are swapped.
Both must give exactly the oracle's populations, energies, best members and counters, generation after
generation, step by step and inside a device-side batch (mb2_run), whatever the accepted fraction; `auto`
follows the accepted fraction of the last generation the host read."""
import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G
from tests.test_gpu_exp_tma import _engine
from tests.test_gpu_parity import energy_close

pytestmark = pytest.mark.gpu

CASES = [
    # (dim, NP, strategy, cost, bounds(lo,hi,mode), penalty terms, CR, F, gens, kernel family)
    (1024, 96, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6, 'tma_bin'),          # BASELINE config 2's rows
    (256, 200, 'Best1Bin', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 8, 'tma_bin'),  # config 5's rows
    (512, 70, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_REJECT), None, 0.2, 0.5, 8, 'tma_bin'),          # many accepted trials
    (1024, 150, 'Rand1Exp', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6, 'tma_exp'),
    (300, 61, 'RandToBest1Exp', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), None, 1.0, 0.8, 5, 'tma_exp'),   # whole rows
    (130, 77, 'Best1Exp', 'rosen', None, None, 0.5, 0.4, 10, 'tma_exp'),                                   # short windows, high acceptance
    (64, 1001, 'Rand1Exp', 'corana', (-1000., 1000., orc.BOUNDS_CLAMP),
     [dict(ptype='quadratic_inequality', G=[1.] * 64, h=100., k=100, hpow=5)], 0.9, 0.8, 8, 'rows_bulk'),  # BASELINE config 4
    (128, 203, 'Rand2Exp', 'griewangk', (-400., 400., orc.BOUNDS_REJECT), None, 0.97, 0.8, 6, 'rows_bulk'),
    (32, 333, 'Best1Exp', 'rosen', (-2., 2., orc.BOUNDS_CLAMP), None, 0.5, 0.5, 12, 'rows_bulk'),          # last warp partly idle
    (64, 301, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 8, 'rows'),                # K2d: binomial crossover on short rows
    (128, 90, 'Best1Bin', 'griewangk', (-400., 400., orc.BOUNDS_REJECT), None, 0.3, 0.5, 8, 'rows'),
    (301, 64, 'Rand1Exp', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6, 'warp_per_row'),        # K2b (MB2_TMA_ODD=0 below)
    (1001, 40, 'Best1Bin', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 5, 'warp_per_row'),
    (301, 65, 'Rand1Exp', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6, 'tma_exp'),             # odd row length, staged
    (1001, 41, 'Best1Bin', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 5, 'tma_bin'),
]
MODES = ['in_place', 'double_buffer', 'auto']


def _start(case, mode, seed):
    dim, NP, strat, cost, bounds, pen, CR, F, gens, family = case
    eng, lo, hi, bmode = _engine(dim, NP, strat, cost, bounds, pen, CR, F)
    eng.set_commit_mode(mode)
    ilo = np.full(dim, -3.0) if bounds is None else lo * 1.25   # start partly outside: gen-0 clip
    ihi = np.full(dim, 3.0) if bounds is None else hi * 1.25
    eng.set_rng_philox(seed)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi), lo, hi, bmode)
    return eng, st


def _check(eng, st, cost, tag):
    best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
    pop = eng.population()
    G.assert_bits_equal(pop, st.population, tag + ' population')
    energy_close(eng.energies(), st.popEnergy, pop, cost, tag + ' energies')
    G.assert_bits_equal(best_x, st.bestSolution, tag + ' bestX')
    energy_close([best_e], [st.bestEnergy], best_x, cost, tag + ' bestE')
    assert n_inf == int(np.isinf(st.trialEnergy).sum()), tag
    assert n_acc == int(st.accepted.sum()), tag
    return n_acc


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-%s-D%d' % (c[2], c[3], c[0]))
def test_step_by_step_equals_the_oracle_in_every_commit_mode(case, mode, monkeypatch):
    monkeypatch.delenv('MB2_COMMIT_MODE', raising=False)
    dim, NP, strat, cost, bounds, pen, CR, F, gens, family = case
    monkeypatch.setenv('MB2_TMA_ODD', '0' if family == 'warp_per_row' else '1')
    seed = 0xC0117000 + dim * 7 + NP
    eng, st = _start(case, mode, seed)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost, pen or ())
    _check(eng, st, cost, 'gen 0')
    frac = None   # accepted fraction of the last generation the host has read
    seen = set()
    for g in range(1, gens + 1):
        eng.step()
        in_place = eng.last_commit_in_place()
        # (in place, Best1Bin rows of 768..1536 dimensions run the pooled kernel K2p: tests/test_gpu_pool.py)
        assert eng.kernel_family() == ('tma_pool' if family == 'tma_bin' and in_place and 768 <= dim <= 1536 and dim % 2 == 0 else family)
        if mode == 'auto':
            assert in_place == (frac is None or frac < (0.30 if dim > 128 else 0.12)), (g, frac)
        else:
            assert in_place == (mode == 'in_place')
        seen.add(in_place)
        orc.step(st, strat, CR, F, ocost, pen or (), philox=dict(seed=seed, offset=0))
        frac = _check(eng, st, cost, '%s gen %d' % (mode, g)) / NP
    eng.close()


@pytest.mark.parametrize('mode', ['in_place', 'double_buffer'])
@pytest.mark.parametrize('case', [CASES[1], CASES[3], CASES[6]], ids=lambda c: '%s-%s-D%d' % (c[2], c[3], c[0]))
def test_device_side_batches_equal_the_oracle_in_every_commit_mode(case, mode, monkeypatch):
    """mb2_run: generations enqueued ahead of the termination test; the launches after the stop are no-ops and the
    population the host reads afterwards is the one of the generation that stopped"""
    from mystic_b200 import termination
    monkeypatch.delenv('MB2_COMMIT_MODE', raising=False)
    dim, NP, strat, cost, bounds, pen, CR, F, gens, family = case
    seed = 0xBA7C0000 + dim * 3 + NP
    eng, st = _start(case, mode, seed)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost, pen or ())
    best_e = eng.read_result()[0]
    history = [best_e]
    term = termination.device_descriptor(termination.EvaluationLimits(generations=7))
    done_total = 0
    for batch in (3, 16):   # the first batch ends with its log full, the second on the condition (after 4 more)
        recs, stopped = eng.run(batch, term, history, done_total, NP * (1 + done_total), None, None)
        assert eng.last_commit_in_place() == (mode == 'in_place')
        for r in recs:
            orc.step(st, strat, CR, F, ocost, pen or (), philox=dict(seed=seed, offset=0))
            history.append(r[0])
            energy_close([r[0]], [st.bestEnergy], r[4:], cost, 'batch record bestE')
            G.assert_bits_equal(r[4:], st.bestSolution, 'batch record bestX')
            assert int(r[3]) == int(st.accepted.sum())
        done_total += len(recs)
        _check(eng, st, cost, '%s after %d generations' % (mode, done_total))
    assert done_total == 7 and stopped
    # and the engine goes on from there
    eng.step()
    orc.step(st, strat, CR, F, ocost, pen or (), philox=dict(seed=seed, offset=0))
    _check(eng, st, cost, '%s generation after the batches' % mode)
    eng.close()


def test_environment_overrides_the_context_setting(monkeypatch):
    case = CASES[0]
    eng, st = _start(case, 'double_buffer', 5)
    eng.eval_initial()
    monkeypatch.setenv('MB2_COMMIT_MODE', '2')
    eng.step()
    assert eng.last_commit_in_place()
    monkeypatch.setenv('MB2_COMMIT_MODE', '1')
    eng.step()
    assert not eng.last_commit_in_place()
    monkeypatch.delenv('MB2_COMMIT_MODE')
    eng.set_commit_mode('in_place')
    eng.step()
    assert eng.last_commit_in_place()
    with pytest.raises(Exception):
        eng.set_commit_mode(7)
    eng.close()


def test_other_kernel_families_keep_the_double_buffer(monkeypatch):
    """the rows-per-warp / thread-per-row / DMMA kernels (rows of a few dimensions) write every selected row"""
    monkeypatch.delenv('MB2_COMMIT_MODE', raising=False)
    for dim, NP, strat in ((7, 33, 'Best1Bin'), (10, 50, 'Rand1Exp')):
        eng, lo, hi, bmode = _engine(dim, NP, strat, 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8)
        eng.set_commit_mode('in_place')
        eng.set_rng_philox(3)
        eng.init_population(3, lo, hi)
        eng.eval_initial()
        eng.step()
        assert eng.kernel_family() in ('rows_per_warp', 'thread_per_row')
        assert not eng.last_commit_in_place()
        eng.close()


def test_auto_follows_the_accepted_fraction_without_host_reads(monkeypatch):
    """mb2_step loops that never read a result: the epilogue kernel leaves every generation's accepted count in pinned
    host memory, and the next launch commits in place iff that fraction is below the threshold (0.12 for these short
    rows) -- checked against the oracle's accepted counts, generation by generation"""
    import torch
    monkeypatch.delenv('MB2_COMMIT_MODE', raising=False)
    case = (32, 333, 'Best1Exp', 'rosen', (-2., 2., orc.BOUNDS_CLAMP), None, 0.5, 0.5, 40, 'rows_bulk')
    dim, NP, strat, cost, bounds, pen, CR, F, gens, family = case
    seed = 0xA0700000
    eng, st = _start(case, 'auto', seed)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost, ())
    torch.cuda.synchronize()
    prev = None
    seen = set()
    for g in range(1, gens + 1):
        eng.step()
        torch.cuda.synchronize()   # (not a read of the result: the hint is whatever the epilogue wrote)
        expect = prev is None or prev / NP < 0.12
        assert eng.last_commit_in_place() == expect, (g, prev)
        seen.add(expect)
        orc.step(st, strat, CR, F, ocost, (), philox=dict(seed=seed, offset=0))
        prev = int(st.accepted.sum())
    assert seen == {True, False}, 'the case should cross the threshold (accepted counts: last %d of %d)' % (prev, NP)
    _check(eng, st, cost, 'after %d generations without reads' % gens)
    eng.close()
