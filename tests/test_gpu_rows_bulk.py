"""K2f (mystic_b200/csrc/de_rows_bulk.cuh): the short-row generation step with cp.async.bulk staging.

Rows of 16..128 dimensions under the exponential-crossover strategies (BASELINE config 4) take this
kernel once generation 0 has clipped them into the strict range.  Bars as in tests/test_gpu_parity.py:
trial vectors, populations and best members bit-exact against the oracle, energies <= 1e-12 — over
ring geometries, donor pools too small for the windows (plain-load fallback, restore from global
memory), windows that wrap / cover the whole row / are empty, populations that leave idle groups in
the last warp, replayed and Philox draws.  K2d (de_rows.cuh) is the second witness: same draws, same
mutation and cost code, so bit-identical populations and energies."""
import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G
from tests.test_gpu_exp_tma import _compare, _engine, _synthetic_draws

pytestmark = pytest.mark.gpu

GEOMETRIES = {
    'default': {},
    'pool_of_one_slot': {'MB2_RB_POOL': '1'},            # almost nothing is staged
    'pool_of_nine_slots': {'MB2_RB_POOL': '9'},          # some rows of a stage staged, some not
    'three_warps_four_stages': {'MB2_RB_WARPS': '3', 'MB2_RB_STAGES': '4'},
    'twelve_warps_two_stages': {'MB2_RB_WARPS': '12', 'MB2_RB_STAGES': '2'},
}
ENV = ('MB2_RB_POOL', 'MB2_RB_WARPS', 'MB2_RB_STAGES', 'MB2_ROWS_BULK', 'MB2_SMALL_KERNEL')


@pytest.fixture(params=sorted(GEOMETRIES))
def geometry(request, monkeypatch):
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    for k, v in GEOMETRIES[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


CASES = [
    # (dim, NP, strategy, cost, bounds(lo,hi,mode), penalty terms, CR, F, gens)
    (64, 1001, 'Rand1Exp', 'corana', (-1000., 1000., orc.BOUNDS_CLAMP),
     [dict(ptype='quadratic_inequality', G=[1.] * 64, h=100., k=100, hpow=5)], 0.9, 0.8, 5),        # BASELINE config 4
    (128, 203, 'Rand2Exp', 'griewangk', (-400., 400., orc.BOUNDS_REJECT), None, 0.97, 0.8, 4),
    (100, 77, 'Best2Exp', 'rosen', None, None, 0.98, 0.7, 4),
    (64, 337, 'RandToBest1Exp', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP), None, 1.0, 0.8, 3),   # whole rows
    (48, 133, 'Best1Exp', 'corana', (-1000., 1000., orc.BOUNDS_REJECT),
     [dict(ptype='quadratic_inequality', G=[1.] * 48, h=100., k=100, hpow=5),
      dict(ptype='linear_equality', G=[0.5] * 48, h=3., k=10, hpow=2)], 0.9, 0.8, 4),
    (16, 59, 'Rand1Bin', 'rosen', (-2., 2., orc.BOUNDS_CLAMP), None, 0.0, 0.8, 3),                  # empty windows
    (20, 90, 'Best2Bin', 'rosen', (-2., 2., orc.BOUNDS_CLAMP), None, 0.6, 0.8, 4),
    (72, 64, 'Rand2Bin', 'rosen', None, None, 0.9, 0.8, 4),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-%s-D%d' % (c[2], c[3], c[0]))
def test_philox_trajectory_matches_oracle(case, geometry):
    dim, NP, strat, cost, bounds, pen, CR, F, gens = case
    seed = 0xB01C0000 + dim * 17 + NP
    eng, lo, hi, mode = _engine(dim, NP, strat, cost, bounds, pen, CR, F)
    ilo = np.full(dim, -3.0) if bounds is None else lo * 1.25   # start partly outside: gen-0 clip
    ihi = np.full(dim, 3.0) if bounds is None else hi * 1.25
    eng.set_rng_philox(seed)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi), lo, hi, mode)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    assert eng.kernel_family() == 'rows'
    orc.step0(st, ocost, pen or ())
    for g in range(1, gens + 1):
        eng.step()
        assert eng.kernel_family() == 'rows_bulk'
        orc.step(st, strat, CR, F, ocost, pen or (), philox=dict(seed=seed, offset=0))
        _compare(eng, st, cost, '%s gen %d' % (geometry, g))
    eng.close()


@pytest.mark.parametrize('strat,cost,dim,bounds', [
    ('Rand1Exp', 'corana', 64, (-1000., 1000., orc.BOUNDS_CLAMP)),
    ('Best1Exp', 'rosen', 128, (-5., 5., orc.BOUNDS_REJECT)),
    ('Rand2Exp', 'griewangk', 36, None),
])
def test_replayed_draws_match_oracle(strat, cost, dim, bounds, geometry):
    """validation mode: recorded draws in; window lengths 0, 1, around the chunk and pool sizes, D-1 and D;
    start indices 0, inside a chunk, D-1 (wrapping windows)"""
    NP, CR, F = 75, 0.9, 0.8
    k = orc.STRATEGIES[strat][1]
    rng = np.random.default_rng(dim * 11 + k)
    eng, lo, hi, mode = _engine(dim, NP, strat, cost, bounds, None, CR, F)
    pop0 = rng.uniform(-3., 3., size=(NP, dim)) if bounds is None else rng.uniform(lo * 1.1, hi * 1.1, size=(NP, dim))
    eng.upload_population(pop0)
    st = orc.DEState(pop0, lo, hi, mode)
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost)
    special = [0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 15, 16, 17, dim - 5, dim - 4, dim - 2, dim - 1, dim]
    for g in range(1, 4):
        lengths = np.minimum(rng.geometric(1.0 - CR, size=NP) - 1, dim)
        lengths[:len(special)] = special
        donors, nstart, uni = _synthetic_draws(rng, NP, dim, k, CR, lengths)
        nstart[:8] = [0, 1, dim - 1, dim - 2, dim - 3, 7, 3, 4]
        nstart[len(special) - 3:len(special)] = [dim - 1, 5, 1]
        eng.set_rng_replay(donors, nstart, uni)
        eng.step()
        assert eng.kernel_family() == 'rows_bulk'
        orc.step(st, strat, CR, F, ocost, (), replay=dict(donors=donors, nstart=nstart, uni=uni))
        _compare(eng, st, cost, '%s gen %d' % (geometry, g))
    eng.close()


def test_rows_not_known_to_be_in_range_keep_the_tiled_kernel(monkeypatch):
    """SetStrictRanges after generation 0: the reference clamps every element of a trial, which K2f does
    not do, so the step stays on K2d"""
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    dim, NP = 64, 50
    eng, _, _, _ = _engine(dim, NP, 'Rand1Exp', 'rosen', None, None, 0.9, 0.8)
    eng.set_rng_philox(5)
    eng.init_population(5, np.full(dim, -3.0), np.full(dim, 3.0))
    eng.eval_initial()
    eng.step()
    assert eng.kernel_family() == 'rows_bulk'
    eng.set_bounds(orc.BOUNDS_CLAMP, np.full(dim, -2.0), np.full(dim, 2.5))
    eng.step()
    assert eng.kernel_family() == 'rows'
    eng.close()


@pytest.mark.parametrize('mode', [orc.BOUNDS_CLAMP, orc.BOUNDS_REJECT])
def test_rows_found_in_range_take_the_bulk_kernel(mode, monkeypatch):
    """a strict range set after generation 0 (or rows installed by the caller: restart files) that every row
    already satisfies: the engine tests the population once and then treats it like a clipped one"""
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    dim, NP, strat, cost, CR, F, seed = 64, 333, 'Rand1Exp', 'rosen', 0.9, 0.8, 77
    eng, _, _, _ = _engine(dim, NP, strat, cost, None, None, CR, F)
    eng.set_rng_philox(seed)
    ilo, ihi = np.full(dim, -3.0), np.full(dim, 3.0)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi))
    ocost = orc.make_cost(cost, ())
    eng.eval_initial()
    orc.step0(st, ocost)
    lo, hi = np.full(dim, -3.0), np.full(dim, 3.5)      # the initial box: nothing is outside
    eng.set_bounds(mode, lo, hi)
    st.lo, st.hi, st.bounds_mode = lo, hi, mode
    for g in range(1, 4):
        eng.step()
        assert eng.kernel_family() == 'rows_bulk'
        orc.step(st, strat, CR, F, ocost, (), philox=dict(seed=seed, offset=0))
        _compare(eng, st, cost, 'gen %d' % g)
    eng.close()


@pytest.mark.parametrize('strat,cost,dim,NP,pen', [
    ('Rand1Exp', 'corana', 64, 300001, [dict(ptype='quadratic_inequality', G=[1.] * 64, h=100., k=100, hpow=5)]),
    ('Best1Exp', 'griewangk', 128, 100003, None),
    ('Rand2Exp', 'rosen', 32, 250000, None),
])
def test_equals_tiled_rows_kernel_bit_for_bit(strat, cost, dim, NP, pen, monkeypatch):
    """a full grid, many chunks per warp (ring wrap-around, both mbarrier parities), a last chunk with
    idle groups: K2f and K2d share draws, mutation and cost code, so populations AND energies agree"""
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    out = {}
    for bulk in ('1', '0'):
        monkeypatch.setenv('MB2_ROWS_BULK', bulk)
        eng, lo, hi, mode = _engine(dim, NP, strat, cost, (-5., 5., orc.BOUNDS_CLAMP), pen, 0.9, 0.8)
        eng.set_rng_philox(1234)
        eng.init_population(1234, lo, hi)
        eng.eval_initial()
        for g in range(4):
            eng.step()
        assert eng.kernel_family() == ('rows_bulk' if bulk == '1' else 'rows')
        out[bulk] = (eng.population().copy(), eng.energies().copy(), eng.read_result())
        eng.close()
    G.assert_bits_equal(out['1'][0], out['0'][0], 'population')
    G.assert_bits_equal(out['1'][1], out['0'][1], 'energies')
    a, b = out['1'][2], out['0'][2]
    assert a[:5] == b[:5]
    G.assert_bits_equal(a[5], b[5], 'best member')
