"""Rows of ODD length on the staged long-row kernels (de_tma.cuh K2a, de_tma_exp.cuh K2e).

8*D bytes is then not a multiple of 16 and every other row starts 8 bytes off a 16-byte boundary, which
`cp.async.bulk` does not take: the kernels move the 16-byte aligned COVER of each row / donor window (one element
more on one side) and address the buffers with a per-row phase; the selected row leaves by plain stores.  Round 1
left these rows on the warp-per-row kernel (2.2-2.4x slower at 1023 dimensions).  Covered here: both kernels and
every base strategy; populations with an odd and an even number of rows (the last row of an odd count has a cover
that would end past the buffer: short copy + plain load); donors of both phases; windows that are empty, wrap,
start at element 0 / D-1, cover the whole row; segments that do not fit the staging buffers; bounds by clamp,
rejection and none; both commit modes; generation 0; replayed and Philox draws.  Bars as everywhere: trial
vectors, populations and best members bit-exact against the oracle, energies <= 1e-12."""
import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G
from tests.test_gpu_exp_tma import _compare, _engine, _synthetic_draws

pytestmark = pytest.mark.gpu

ENV = ('MB2_XTMA_CAP', 'MB2_XTMA_GROUPS', 'MB2_XTMA_GWARPS', 'MB2_XTMA_STAGES', 'MB2_TMA_EXP', 'MB2_TMA', 'MB2_TMA_ODD',
       'MB2_TMA_GROUPS', 'MB2_TMA_GWARPS', 'MB2_TMA_STAGES', 'MB2_COMMIT_MODE')
VARIANTS = {
    'default': {},
    'in_place': {'MB2_COMMIT_MODE': '2'},
    'double_buffer': {'MB2_COMMIT_MODE': '1'},
    'segments_never_fit': {'MB2_XTMA_CAP': '2', 'MB2_COMMIT_MODE': '1'},
    'deep_rings': {'MB2_XTMA_GROUPS': '3', 'MB2_XTMA_GWARPS': '1', 'MB2_XTMA_STAGES': '4',
                   'MB2_TMA_GROUPS': '2', 'MB2_TMA_GWARPS': '2', 'MB2_TMA_STAGES': '2'},
}


@pytest.fixture(params=sorted(VARIANTS))
def variant(request, monkeypatch):
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    for k, v in VARIANTS[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


CASES = [
    # (dim, NP, strategy, cost, bounds(lo,hi,mode), CR, F, gens, family)
    (1023, 97, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), 0.9, 0.8, 4, 'tma_bin'),
    (257, 120, 'Best1Bin', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), 0.9, 0.8, 5, 'tma_bin'),
    (513, 64, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_REJECT), 0.3, 0.5, 5, 'tma_bin'),
    (2049, 21, 'Best1Bin', 'sphere', None, 0.9, 0.8, 3, 'tma_bin'),
    (129, 50, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), 0.0, 0.8, 3, 'tma_bin'),            # only the start index crosses
    (1023, 151, 'Rand1Exp', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), 0.9, 0.8, 4, 'tma_exp'),
    (301, 64, 'RandToBest1Exp', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), 1.0, 0.8, 3, 'tma_exp'),  # whole rows
    (131, 77, 'Best1Exp', 'rosen', None, 0.99, 0.8, 4, 'tma_exp'),                                      # long windows, wraps
    (255, 90, 'Best2Exp', 'griewangk', (-400., 400., orc.BOUNDS_REJECT), 0.97, 0.8, 4, 'tma_exp'),
    (2051, 41, 'Rand2Exp', 'rosen', (-3., 3., orc.BOUNDS_REJECT), 0.9, 0.7, 3, 'tma_exp'),
    (135, 45, 'RandToBest1Bin', 'sphere', None, 0.0, 0.8, 3, 'tma_exp'),                                # empty windows
    (6001, 12, 'Rand2Bin', 'griewangk', None, 0.999, 0.8, 2, 'tma_exp'),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-%s-D%d-NP%d' % (c[2], c[3], c[0], c[1]))
def test_philox_trajectory_matches_oracle(case, variant):
    dim, NP, strat, cost, bounds, CR, F, gens, family = case
    seed = 0x0DD00000 + dim * 17 + NP
    eng, lo, hi, mode = _engine(dim, NP, strat, cost, bounds, None, CR, F)
    ilo = np.full(dim, -3.0) if bounds is None else lo * 1.25   # start partly outside: gen-0 clip
    ihi = np.full(dim, 3.0) if bounds is None else hi * 1.25
    eng.set_rng_philox(seed)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi), lo, hi, mode)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    assert eng.kernel_family() == 'tma_exp'   # generation 0 streams through the staged kernel too
    orc.step0(st, ocost, ())
    _compare(eng, st, cost, '%s gen 0' % variant)
    for g in range(1, gens + 1):
        eng.step()
        assert eng.kernel_family() == family
        orc.step(st, strat, CR, F, ocost, (), philox=dict(seed=seed, offset=0))
        _compare(eng, st, cost, '%s gen %d' % (variant, g))
    eng.close()


@pytest.mark.parametrize('strat,cost,dim,NP,bounds', [
    ('Best1Exp', 'rosen', 1023, 61, (-5., 5., orc.BOUNDS_CLAMP)),
    ('Rand2Exp', 'griewangk', 259, 60, (-400., 400., orc.BOUNDS_REJECT)),
    ('RandToBest1Exp', 'rastrigin', 141, 33, None),
    ('Best1Bin', 'rosen', 333, 47, (-5., 5., orc.BOUNDS_CLAMP)),
])
def test_replayed_draws_match_oracle(strat, cost, dim, NP, bounds, variant):
    """validation mode: window lengths 0, 1, around the staged capacity, D-1 and D; start indices 0, 1, D-1, ...;
    the LAST row (whose cover is the short one when NP is odd) is made parent and donor of wrapping windows"""
    CR, F = 0.9, 0.8
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
    binomial = strat == 'Best1Bin'
    for g in range(1, 4):
        lengths = np.minimum(rng.geometric(1.0 - CR, size=NP) - 1, dim)
        lengths[:len(special)] = special
        donors, nstart, uni = _synthetic_draws(rng, NP, dim, k, CR, lengths)
        if binomial:   # one uniform per dimension (strategy.py:77-85)
            uni = rng.uniform(0.0, 1.0, size=(NP, dim + 1))
        nstart[:6] = [0, 1, dim - 1, dim - 2, dim - 3, 7]
        nstart[len(special) - 3:len(special)] = [dim - 1, 5, 1]
        # the last row as a donor of windows that end at the row's end, wrap, and cover everything
        for c, (ns, L) in zip((20, 21, 22, 23), ((dim - 4, 4), (dim - 3, 9), (0, dim), (dim - 1, 1))):
            others = [r for r in range(NP - 1) if r != c][:k - 1]
            donors[c, :k] = [NP - 1] + others
            nstart[c] = ns
            if not binomial:
                uni[c, :] = np.nan
                uni[c, :L] = CR * 0.5
                if L < dim:
                    uni[c, L] = 0.95
        eng.set_rng_replay(donors, nstart, uni)
        eng.step()
        assert eng.kernel_family() == ('tma_bin' if binomial else 'tma_exp')
        orc.step(st, strat, CR, F, ocost, (), replay=dict(donors=donors, nstart=nstart, uni=uni))
        _compare(eng, st, cost, '%s gen %d' % (variant, g))
    eng.close()


@pytest.mark.parametrize('strat,cost,dim,NP', [('Rand1Exp', 'rosen', 1023, 6001), ('Best1Bin', 'rosen', 1023, 5000),
                                               ('Best1Exp', 'griewangk', 255, 20001), ('Best1Bin', 'griewangk', 257, 9000)])
def test_equals_warp_per_row_kernel_bit_for_bit(strat, cost, dim, NP, monkeypatch):
    """many rows per group (ring wrap-around, both mbarrier parities, every CTA of a full grid) against the plain-load
    kernel the odd rows ran on before (MB2_TMA_ODD=0): the same device functions, so populations AND energies agree
    bit for bit"""
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    out = {}
    for tma in ('1', '0'):
        monkeypatch.setenv('MB2_TMA_ODD', tma)
        eng, lo, hi, mode = _engine(dim, NP, strat, cost, (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8)
        eng.set_rng_philox(99)
        eng.init_population(99, lo, hi)
        eng.eval_initial()
        for g in range(5):
            eng.step()
        assert eng.kernel_family() == (('tma_bin' if strat == 'Best1Bin' else 'tma_exp') if tma == '1' else 'warp_per_row')
        out[tma] = (eng.population().copy(), eng.energies().copy(), eng.read_result())
        eng.close()
    G.assert_bits_equal(out['1'][0], out['0'][0], 'population')
    G.assert_bits_equal(out['1'][1], out['0'][1], 'energies')
    a, b = out['1'][2], out['0'][2]
    assert a[:5] == b[:5]
    G.assert_bits_equal(a[5], b[5], 'best member')


@pytest.mark.parametrize('dim', [129, 383, 1001, 2051, 4097, 8191, 12001])
@pytest.mark.parametrize('strat,CR', [('Best1Exp', 0.5), ('Rand1Exp', 1.0), ('Rand2Exp', 0.999), ('Best1Bin', 0.9)])
def test_every_supported_odd_row_length_launches_and_matches(dim, strat, CR, monkeypatch):
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    NP, cost, F, seed = 25, 'sphere' if dim % 3 else 'rosen', 0.7, 4242 + dim
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
