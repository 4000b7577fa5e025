"""K2p (mystic_b200/csrc/de_tma_pool.cuh): the pooled producer / consumer kernel that runs Best1Bin generations of
768..1536-dimensional rows when they are committed in place (BASELINE config 2's shape).

It replaces K2a's per-group stages by two rings of row buffers (P parent + second-donor pairs, Q first-donor
buffers) that a producer warp fills in row order; the trial, cost and selection code is K2a's.  What can go wrong is
therefore the plumbing: ring slots and mbarrier parities once a CTA has processed more rows than the rings hold,
batches of 16 rows in the producer (a last partial batch, rows past the end), consumer groups that outnumber the
first-donor ring (Q >= G is enforced by the host), recorded draws.  Cases here give every CTA of a FULL grid several
ring revolutions and compare, as everywhere, with the oracle: populations and best members bit for bit, energies to
1e-12 (mystic/differential_evolution.py:585-623 is the generation they restate).  K2a (MB2_TMA_POOL=0) is the second
witness: the same device functions, so bit-identical energies too."""
import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G
from tests.test_gpu_exp_tma import _engine
from tests.test_gpu_fullsize_parity import _run
from tests.test_gpu_parity import energy_close

pytestmark = pytest.mark.gpu

POOL_ENV = ('MB2_TMA_POOL', 'MB2_POOL_P', 'MB2_POOL_Q', 'MB2_POOL_G', 'MB2_POOL_GW', 'MB2_POOL_NPROD', 'MB2_POOL_SHORT', 'MB2_COMMIT_MODE')

GEOMETRIES = {
    'default': {},
    'small_rings': {'MB2_POOL_P': '2', 'MB2_POOL_Q': '3', 'MB2_POOL_G': '8'},       # G is clipped to Q; constant back-pressure
    'eight_groups': {'MB2_POOL_P': '8', 'MB2_POOL_Q': '10', 'MB2_POOL_G': '8'},
    'four_warps': {'MB2_POOL_GW': '4'},
    'one_producer': {'MB2_POOL_NPROD': '1'},
    'one_warp_groups': {'MB2_POOL_GW': '1', 'MB2_POOL_G': '11'},
    'one_group': {'MB2_POOL_G': '1', 'MB2_POOL_P': '3', 'MB2_POOL_Q': '2'},
}


@pytest.fixture(params=sorted(GEOMETRIES))
def geometry(request, monkeypatch):
    for k in POOL_ENV:
        monkeypatch.delenv(k, raising=False)
    for k, v in GEOMETRIES[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


def _is_pool(eng):
    return eng.kernel_family() == 'tma_pool'


@pytest.mark.parametrize('D,NP,cost,lo,hi,bounds', [
    (1024, 6000, 'rosen', -5., 5., orc.BOUNDS_CLAMP),          # 41 rows per CTA: three producer batches, the last one partial
    (768, 4737, 'griewangk', -400., 400., orc.BOUNDS_REJECT),  # the shortest rows of the kernel, an odd row count
    (1536, 3000, 'rosen', -2., 2., orc.BOUNDS_CLAMP),          # 12 KB rows: 17 buffers, Q = 8 < default groups + 1
    (1280, 2500, 'corana', -1000., 1000., orc.BOUNDS_CLAMP),   # rows that are not whole trips of 512 elements
])
def test_full_grid_trajectory_equals_the_c_oracle(D, NP, cost, lo, hi, bounds, geometry, monkeypatch):
    """every CTA of a full grid revolves its rings several times; three generations against the C oracle port"""
    monkeypatch.setenv('MB2_COMMIT_MODE', '2')   # in place whatever the accepted fraction (`auto` leaves it above 0.30)
    seen = []
    _run(NP, D, 'Best1Bin', cost, (), lo, hi, bounds, 0.9, 0.8, None, 'tma_pool', gens=3, check=lambda eng: seen.append(_is_pool(eng)))
    assert seen == [True]


def test_pool_and_k2a_give_identical_bits(monkeypatch):
    """the same generations with the pool switched off (K2a, in place) and on: populations AND energies bit for bit"""
    D, NP, gens = 1024, 5000, 4
    out = {}
    for pool in ('0', '1'):
        for k in POOL_ENV:
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv('MB2_TMA_POOL', pool)
        eng, lo, hi, mode = _engine(D, NP, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8)
        eng.set_commit_mode('in_place')
        eng.set_rng_philox(77)
        eng.init_population(77, lo * 1.25, hi * 1.25)
        eng.eval_initial()
        for g in range(gens):
            eng.step()
            assert eng.kernel_family() == ('tma_pool' if pool == '1' else 'tma_bin')
        out[pool] = (eng.population(), eng.energies(), eng.read_result())
        eng.close()
    G.assert_bits_equal(out['0'][0], out['1'][0], 'population')
    G.assert_bits_equal(out['0'][1], out['1'][1], 'energies')
    assert out['0'][2][:5] == out['1'][2][:5]


def test_double_buffer_keeps_k2a(monkeypatch):
    """the pool needs rejected rows to stay where they are: a double-buffered generation runs K2a"""
    for k in POOL_ENV:
        monkeypatch.delenv(k, raising=False)
    eng, lo, hi, mode = _engine(1024, 64, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8)
    eng.set_commit_mode('double_buffer')
    eng.set_rng_philox(3)
    eng.init_population(3, lo, hi)
    eng.eval_initial()
    eng.step()
    assert eng.kernel_family() == 'tma_bin'
    eng.close()


def test_replayed_draws_through_the_pool(geometry):
    """validation mode (recorded donors, start indices and one uniform per dimension, strategy.py:77-85): the producer
    reads the donors lane-parallel, the consumers the uniforms; 20 rows per CTA so that the rings revolve"""
    D, NP, CR, F = 1024, 2960, 0.9, 0.8
    rng = np.random.default_rng(2024)
    eng, lo, hi, mode = _engine(D, NP, 'Best1Bin', 'rosen', (-5., 5., orc.BOUNDS_CLAMP), None, CR, F)
    eng.set_commit_mode('in_place')
    pop0 = rng.uniform(lo * 1.1, hi * 1.1, size=(NP, D))
    eng.upload_population(pop0)
    st = orc.DEState(pop0, lo, hi, mode)
    ocost = orc.make_cost('rosen', ())
    eng.eval_initial()
    orc.step0(st, ocost)
    for g in range(1, 3):
        donors = np.zeros((NP, 5), dtype=np.int64)
        for c in range(NP):
            donors[c, :2] = rng.choice(np.delete(np.arange(NP), c), size=2, replace=False) if c < 64 else \
                (np.array([c + 1, c + 2]) + rng.integers(0, NP - 3)) % NP
        same = (donors[:, 0] == np.arange(NP)) | (donors[:, 1] == np.arange(NP)) | (donors[:, 0] == donors[:, 1])
        donors[same, 0] = (np.arange(NP)[same] + 1) % NP
        donors[same, 1] = (np.arange(NP)[same] + 2) % NP
        nstart = rng.integers(0, D, size=NP)
        nstart[:4] = [0, 1, D - 1, D - 2]
        uni = rng.uniform(0.0, 1.0, size=(NP, D + 1))
        eng.set_rng_replay(donors, nstart, uni)
        eng.step()
        assert _is_pool(eng)
        orc.step(st, 'Best1Bin', CR, F, ocost, (), replay=dict(donors=donors, nstart=nstart, uni=uni))
        best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
        pop = eng.population()
        G.assert_bits_equal(pop, st.population, 'generation %d population' % g)
        energy_close(eng.energies(), st.popEnergy, pop, 'rosen', 'generation %d energies' % g)
        G.assert_bits_equal(best_x, st.bestSolution, 'generation %d best member' % g)
        assert n_acc == int(st.accepted.sum()) and n_inf == int(np.isinf(st.trialEnergy).sum())
    eng.close()
