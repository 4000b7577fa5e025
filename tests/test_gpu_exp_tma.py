"""K2e (mystic_b200/csrc/de_tma_exp.cuh): exponential crossover over TMA-staged long rows.

Nine of the reference's ten strategies run the exponential loop (mystic/strategy.py:48-56); for
rows longer than 128 dimensions they take this kernel.  Same bars as tests/test_gpu_parity.py:
trial vectors, populations and best members bit-exact against the oracle, energies <= 1e-12 —
over every ring geometry, with donor windows that fit the staged segments and ones that do not,
windows that wrap, cover the whole row or are empty, rows the context has and has not checked
against the strict range, and recorded (replay) as well as Philox draws.  The warp-per-row kernel
(de_kernels.cuh) is the second witness: same device functions, so bit-identical results."""
import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G
from tests.test_gpu_parity import _device_cost, energy_close

pytestmark = pytest.mark.gpu

GEOMETRIES = {
    'default': {},
    'segments_never_fit': {'MB2_XTMA_CAP': '2'},                      # plain-load donors, undo from global
    'deep_ring_one_warp': {'MB2_XTMA_GROUPS': '3', 'MB2_XTMA_GWARPS': '1', 'MB2_XTMA_STAGES': '4'},
    'eight_warps': {'MB2_XTMA_GROUPS': '2', 'MB2_XTMA_GWARPS': '8', 'MB2_XTMA_STAGES': '2', 'MB2_XTMA_CAP': '16'},
    'two_warps': {'MB2_XTMA_GROUPS': '5', 'MB2_XTMA_GWARPS': '2', 'MB2_XTMA_STAGES': '1'},
}


@pytest.fixture(params=sorted(GEOMETRIES))
def geometry(request, monkeypatch):
    for k in ('MB2_XTMA_CAP', 'MB2_XTMA_GROUPS', 'MB2_XTMA_GWARPS', 'MB2_XTMA_STAGES', 'MB2_TMA_EXP'):
        monkeypatch.delenv(k, raising=False)
    for k, v in GEOMETRIES[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


CASES = [
    # (dim, NP, strategy, cost, bounds(lo,hi,mode), penalty terms, CR, F, gens)
    (1024, 150, 'Rand1Exp', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 4),
    (130, 77, 'Best1Exp', 'rosen', None, None, 0.99, 0.8, 4),
    (300, 61, 'RandToBest1Exp', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), None, 1.0, 0.8, 3),   # whole row
    (256, 90, 'Best2Exp', 'griewangk', (-400., 400., orc.BOUNDS_REJECT), None, 0.97, 0.8, 4),
    (2048, 40, 'Rand2Exp', 'rosen', (-3., 3., orc.BOUNDS_REJECT), None, 0.9, 0.7, 3),
    (512, 64, 'Rand1Bin', 'corana', (-1000., 1000., orc.BOUNDS_CLAMP),
     [dict(ptype='quadratic_inequality', G=[1.] * 512, h=100., k=100, hpow=5)], 0.9, 0.8, 4),
    (200, 50, 'Best2Bin', 'ackley', (-30., 30., orc.BOUNDS_CLAMP), None, 0.95, 0.8, 3),
    (134, 45, 'RandToBest1Bin', 'sphere', None, None, 0.0, 0.8, 3),                                 # empty windows
    (6002, 12, 'Rand2Bin', 'griewangk', None, None, 0.999, 0.8, 2),
]


def _engine(dim, NP, strat, cost, bounds, pen, CR, F):
    from mystic_b200 import penalty
    from mystic_b200.engine import DeviceEngine
    lo = hi = None
    mode = orc.BOUNDS_NONE
    if bounds:
        lo, hi, mode = np.full(dim, bounds[0]), np.full(dim, bounds[1]), bounds[2]
    eng = DeviceEngine(NP, dim, keep_trials=True)
    c, dev_extra = _device_cost(cost, (), dim)
    eng.set_cost(c.cost_id, c.params(dev_extra))
    eng.set_bounds(mode, lo, hi)
    if pen:
        def p(x):
            return 0.0
        for t in pen:
            p = getattr(penalty, t['ptype'])(penalty.linear(t['G'], t['h']), k=t['k'], h=t['hpow'])(p)
        eng.set_penalty(*p.device_terms(dim))
    eng.set_strategy(orc.STRATEGIES[strat][0], F, CR)
    return eng, lo, hi, mode


def _compare(eng, st, cost, tag):
    best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
    trial, trial_e = eng.trials()
    G.assert_bits_equal(trial, st.trialSolution, tag + ' trial')
    energy_close(trial_e, st.trialEnergy, trial, cost, tag + ' trialE')
    G.assert_bits_equal(eng.population(), st.population, tag + ' population')
    energy_close(eng.energies(), st.popEnergy, eng.population(), cost, tag + ' popE')
    G.assert_bits_equal(best_x, st.bestSolution, tag + ' bestX')
    energy_close([best_e], [st.bestEnergy], best_x, cost, tag + ' bestE')
    assert n_inf == int(np.isinf(st.trialEnergy).sum()), tag
    assert n_acc == int(st.accepted.sum()), tag


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-%s-D%d' % (c[2], c[3], c[0]))
def test_philox_trajectory_matches_oracle(case, geometry):
    dim, NP, strat, cost, bounds, pen, CR, F, gens = case
    seed = 0xE2E00000 + dim * 17 + NP
    eng, lo, hi, mode = _engine(dim, NP, strat, cost, bounds, pen, CR, F)
    ilo = np.full(dim, -3.0) if bounds is None else lo * 1.25   # start partly outside: gen-0 clip
    ihi = np.full(dim, 3.0) if bounds is None else hi * 1.25
    eng.set_rng_philox(seed)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi), lo, hi, mode)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost, pen or ())
    for g in range(1, gens + 1):
        eng.step()
        assert eng.kernel_family() == 'tma_exp'
        orc.step(st, strat, CR, F, ocost, pen or (), philox=dict(seed=seed, offset=0))
        _compare(eng, st, cost, '%s gen %d' % (geometry, g))
    eng.close()


def _synthetic_draws(rng, NP, D, k, CR, lengths):
    """what the reference's random.sample / randrange / random would have produced (strategy.py:21-27,
    48-56): k distinct donors != candidate, a start index, and L uniforms below CR followed by one
    that is not (NaN = never drawn)"""
    donors = np.zeros((NP, 5), dtype=np.int64)
    for c in range(NP):
        donors[c, :k] = rng.choice(np.delete(np.arange(NP), c), size=k, replace=False)
    nstart = rng.integers(0, D, size=NP)
    uni = np.full((NP, D + 1), np.nan)
    for c in range(NP):
        L = int(lengths[c])
        uni[c, :L] = rng.uniform(0.0, CR, size=L) * 0.999
        if L < D:
            uni[c, L] = CR + (1.0 - CR) * rng.uniform(0.0, 1.0)
    return donors, nstart, uni


@pytest.mark.parametrize('strat,cost,dim,bounds', [
    ('Best1Exp', 'rosen', 1024, (-5., 5., orc.BOUNDS_CLAMP)),
    ('Rand2Exp', 'griewangk', 258, (-400., 400., orc.BOUNDS_REJECT)),
    ('RandToBest1Exp', 'rastrigin', 140, None),
])
def test_replayed_draws_match_oracle(strat, cost, dim, bounds, geometry):
    """validation mode: recorded draws in; window lengths cover 0, 1, the staged capacity +-1, D-1
    and D, start indices include 0, odd ones and D-1 (wrapping windows)"""
    NP, CR, F = 60, 0.9, 0.8
    k = orc.STRATEGIES[strat][1]
    rng = np.random.default_rng(dim * 7 + k)
    eng, lo, hi, mode = _engine(dim, NP, strat, cost, bounds, None, CR, F)
    pop0 = rng.uniform(-3., 3., size=(NP, dim)) if bounds is None else rng.uniform(lo * 1.1, hi * 1.1, size=(NP, dim))
    eng.upload_population(pop0)
    st = orc.DEState(pop0, lo, hi, mode)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost)
    special = [0, 1, 2, 3, 15, 16, 17, 63, 64, 65, 66, 67, 68, 69, 70, dim - 2, dim - 1, dim]
    for g in range(1, 4):
        lengths = np.minimum(rng.geometric(1.0 - CR, size=NP) - 1, dim)
        lengths[:len(special)] = special
        donors, nstart, uni = _synthetic_draws(rng, NP, dim, k, CR, lengths)
        nstart[:6] = [0, 1, dim - 1, dim - 2, dim - 3, 7]
        nstart[len(special) - 3:len(special)] = [dim - 1, 5, 1]
        eng.set_rng_replay(donors, nstart, uni)
        eng.step()
        assert eng.kernel_family() == 'tma_exp'
        orc.step(st, strat, CR, F, ocost, (), replay=dict(donors=donors, nstart=nstart, uni=uni))
        _compare(eng, st, cost, '%s gen %d' % (geometry, g))
    eng.close()


@pytest.mark.parametrize('mode', [orc.BOUNDS_CLAMP, orc.BOUNDS_REJECT])
def test_rows_not_known_to_be_in_range(mode, geometry):
    """SetStrictRanges after generation 0 (abstract_solver.py:334-405): parents may lie outside the new
    range, so the reference's clamp / test touches every element of a trial, not only the window"""
    dim, NP, strat, cost, CR, F = 192, 70, 'Rand1Exp', 'rosen', 0.9, 0.8
    seed = 0xABCD
    eng, _, _, _ = _engine(dim, NP, strat, cost, None, None, CR, F)
    eng.set_rng_philox(seed)
    ilo, ihi = np.full(dim, -3.0), np.full(dim, 3.0)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi))
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost)
    lo, hi = np.full(dim, -2.0), np.full(dim, 2.5)
    eng.set_bounds(mode, lo, hi)
    st.lo, st.hi, st.bounds_mode = lo, hi, mode
    for g in range(1, 4):
        eng.step()
        assert eng.kernel_family() == 'tma_exp'
        orc.step(st, strat, CR, F, ocost, (), philox=dict(seed=seed, offset=0))
        _compare(eng, st, cost, '%s gen %d' % (geometry, g))
    eng.close()


@pytest.mark.parametrize('strat,cost,dim,NP', [('Rand1Exp', 'rosen', 1024, 6000), ('Best1Exp', 'griewangk', 256, 20000),
                                               ('Rand2Exp', 'corana', 160, 9000)])
def test_equals_warp_per_row_kernel_bit_for_bit(strat, cost, dim, NP, monkeypatch):
    """many rows per group (ring wrap-around, both mbarrier parities, every CTA of a full grid): the TMA
    ring and the plain-load kernel run the same device functions, so populations AND energies agree
    bit for bit"""
    out = {}
    for tma in ('1', '0'):
        monkeypatch.setenv('MB2_TMA_EXP', tma)
        eng, lo, hi, mode = _engine(dim, NP, strat, cost, (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8)
        eng.set_rng_philox(99)
        eng.init_population(99, lo, hi)
        eng.eval_initial()
        for g in range(5):
            eng.step()
        assert eng.kernel_family() == ('tma_exp' if tma == '1' else 'warp_per_row')
        out[tma] = (eng.population().copy(), eng.energies().copy(), eng.read_result())
        eng.close()
    G.assert_bits_equal(out['1'][0], out['0'][0], 'population')
    G.assert_bits_equal(out['1'][1], out['0'][1], 'energies')
    a, b = out['1'][2], out['0'][2]
    assert a[:5] == b[:5]
    G.assert_bits_equal(a[5], b[5], 'best member')


@pytest.mark.parametrize('dim', [130, 384, 1000, 2050, 4096, 8190, 12000, 14000])
@pytest.mark.parametrize('strat,CR', [('Best1Exp', 0.5), ('Rand1Exp', 1.0), ('Best2Exp', 0.9), ('Rand2Exp', 0.999), ('Best1Bin', 0.9)])
def test_every_supported_row_length_launches_and_matches(dim, strat, CR, monkeypatch):
    """the ring geometries are derived from the row length, the donor count and CR (de_abi.cu: tma_geometry,
    exp_tma_geometry): every combination up to the longest row a CTA can hold must fit the 227 KB of shared
    memory, launch, and agree with the oracle"""
    for k in ('MB2_XTMA_CAP', 'MB2_XTMA_GROUPS', 'MB2_XTMA_GWARPS', 'MB2_XTMA_STAGES', 'MB2_TMA_EXP'):
        monkeypatch.delenv(k, raising=False)
    NP, cost, F, seed = 24, 'sphere' if dim % 3 else 'rosen', 0.7, 4242 + dim
    eng, lo, hi, mode = _engine(dim, NP, strat, cost, (-4., 4., orc.BOUNDS_CLAMP), None, CR, F)
    eng.set_rng_philox(seed)
    eng.init_population(seed, lo, hi)
    st = orc.DEState(orc.philox_init_population(seed, NP, lo, hi), lo, hi, mode)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost)
    for g in range(1, 3):
        eng.step()
        orc.step(st, strat, CR, F, ocost, (), philox=dict(seed=seed, offset=0))
        _compare(eng, st, cost, 'D=%d %s gen %d (%s)' % (dim, strat, g, eng.kernel_family()))
    assert eng.kernel_family() in ('tma_exp', 'tma_bin', 'tma_pool', 'warp_per_row')
    eng.close()
