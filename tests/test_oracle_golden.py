"""The CPU oracle (oracle/de_oracle.py) against recordings of the unmodified reference.

Free-running: the oracle starts from the recorded initial population, consumes the
reference's recorded draws generation by generation and must reproduce, BIT FOR BIT,
every trial vector, trial energy, population, energy vector, best energy / solution and
the evaluation counter of mystic.DifferentialEvolutionSolver2."""
import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G


@pytest.mark.parametrize('name', G.names())
def test_oracle_reproduces_reference(name):
    meta, z = G.load(name)
    lo, hi, mode = G.bounds_of(meta)
    cost = orc.make_cost(meta['cost'], G.extra_of(meta, z))
    pen = meta.get('penalty') or ()
    st = orc.DEState(z['pop0'], lo, hi, mode)
    orc.step0(st, cost, pen)
    ngen = meta['generations']
    for g in range(ngen + 1):
        if g > 0:
            replay = dict(donors=z['donors'][g - 1], nstart=z['nstart'][g - 1], uni=z['uni'][g - 1])
            orc.step(st, meta['strategy'], meta['CR_eff'], meta['F_eff'], cost, pen, replay=replay)
        G.assert_bits_equal(st.trialSolution, z['trial'][g], '%s gen %d trial' % (name, g))
        G.assert_bits_equal(st.population, z['pop'][g], '%s gen %d pop' % (name, g))
        if meta['cost'] in G.COMPENSATED_SUM_COSTS:
            # builtin sum() of the recording interpreter (3.12) is compensated for runs of
            # exact floats; see oracle/de_oracle.py:cost_griewangk
            for a, b in ((st.trialEnergy, z['trialE'][g]), (st.popEnergy, z['popE'][g]),
                         (st.bestEnergy, z['bestE'][g])):
                np.testing.assert_allclose(a, b, rtol=1e-14, atol=0)
        else:
            G.assert_bits_equal(st.trialEnergy, z['trialE'][g], '%s gen %d trialE' % (name, g))
            G.assert_bits_equal(st.popEnergy, z['popE'][g], '%s gen %d popE' % (name, g))
            G.assert_bits_equal(st.bestEnergy, z['bestE'][g], '%s gen %d bestE' % (name, g))
        G.assert_bits_equal(st.bestSolution, z['bestX'][g], '%s gen %d bestX' % (name, g))
        assert st.evaluations == int(z['evals'][g]), (name, g)
    assert st.evaluations == meta['evaluations']


def test_model_known_answers():
    """SURVEY.md 8(c) known-answer vectors, produced by the reference in the build container."""
    f = orc.make_cost
    assert f('rosen')([[0.8, 1.2, 0.7]])[0] == 86.19999999999997
    assert f('rosen')([[1., 1., 1.]])[0] == 0.0
    assert f('rosen')([np.linspace(-2, 2, 7)])[0] == 3608.567901234568
    assert f('corana')([[0.3, -0.7, 1.234, 5.0]])[0] == 859.6112499999999
    assert f('corana')([[0.3, -0.7, 1.234, 5.0] + [9.0] * 60])[0] == 859.6112499999999
    assert f('corana')([[0.04, -0.04, 0.01, 0.0]])[0] == 0.0
    assert f('griewangk')([[1., 2., 3.]])[0] == 1.0170279701835734
    assert f('griewangk')([[0.] * 10])[0] == 0.0
    assert f('griewangk')([np.linspace(-400, 400, 10)])[0] == 164.05614975873598
    assert f('zimmermann')([[7., 2.]])[0] == 0.0
    assert f('zimmermann')([[1., 1.]])[0] == 7.0
    assert f('zimmermann')([[-1., 6.]])[0] == 1600.0
    assert f('chebyshev8cost')([orc.CHEBYSHEV_TARGETS[8]])[0] == 0.0
    assert f('chebyshev8cost')([[1.] * 9])[0] == 7823.744334789609
    assert f('chebyshev8cost', (4096,))([[1.] * 9])[0] == 22615.651073208195
    assert f('chebyshev8cost', (4096,))([[100, -50, 3, 7, -20, 1, 0.5, -2, 1]])[0] == 2014464.4595239898
    # the other mystic.models functions (values of the unmodified reference in the build container; the first
    # three, the minima its docstrings quote: pohlheim.py:276, 311, 350)
    for name, x, want in [
            ('branins', [3.141592653589793, 2.275], 0.39788735772973816),
            ('easom', [3.141592653589793, 3.141592653589793], -1.0),
            ('goldstein', [0.0, -1.0], 3.0),
            ('peaks', [0.22827892, -1.62553496], -6.551133332835841),
            ('venkat91', [4.0, 4.0], -19.668329370585823),
            ('fosc3d', [-0.215018, 0.240356], -4.501069742528923),
            ('nmin51', [-0.02440313, 0.21061247], -3.306868647458795),
            ('shekel', [-31.5, -31.9], 0.01562551162942235),
            ('ellipsoid', [1.0, 2.0, 3.0, -4.0], 50.0),
            ('paviani', [9.35] * 10, -45.77846739623988),
            ('sphere', [1.0, 2.0, 3.0], 14.0),
            ('rastrigin', [0.5, -1.5], 42.5),
            ('schwefel', [420.9687465, 420.9687465], -837.9657745448675),
            ('ackley', [0.1, -0.2, 0.3], 2.2544445866053593),
            ('step', [0.5, -1.5, 6.0], 54.4),
            ('michal', [2.20289811, 1.57078059], -1.8013033991765086),
            ('powers', [0.5, -0.5, 2.0], 16.375)]:
        got = f(name)([x])[0]
        assert got == want or abs(got - want) <= 1e-14 * abs(want), (name, got, want)
    pen = [dict(ptype='quadratic_inequality', G=[1., 1., 1.], h=1.0, k=100, hpow=5)]
    assert orc.penalty_value([[1., 1., 1.]], pen)[0] == 800.0
    assert orc.penalty_value([[0., 0., 0.]], pen)[0] == 0.0


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = orc.philox4x32(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(g) for g in got) == want


def test_philox_donors_are_valid_and_uniform():
    NP, D = 11, 7
    cand = np.arange(NP)
    counts = np.zeros((NP, NP))
    for gen in range(1, 400):
        donors, nstart = orc.philox_donors(0x5EED, gen, cand, NP, D, 5)
        assert ((donors >= 0) & (donors < NP)).all()
        assert (donors != cand[:, None]).all()
        srt = np.sort(donors, axis=1)
        assert (srt[:, 1:] != srt[:, :-1]).all()
        assert ((nstart >= 0) & (nstart < D)).all()
        for c in range(NP):
            counts[c, donors[c, 0]] += 1
    off = counts[~np.eye(NP, dtype=bool)]
    assert abs(off.mean() - 399 / 10.) < 1e-9
    assert off.min() > 15 and off.max() < 70


# ---- draw protocol 2 of the Philox mode (mystic_b200/csrc/philox.cuh): known answers and distributions ----
def test_exponential_length_is_a_count_of_leading_zero_bits_at_cr_one_half():
    """CR = 0.5: thr = 2^31, T[l] = 2^(32-l), so L = #{l : w < 2^(32-l)} = number of leading zero bits of w"""
    tab = orc.exp_length_table(0.5, 1000)
    assert [int(t) for t in tab] == [1 << (32 - l) for l in range(1, 33)]
    w = np.array([0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 0x40000000, 0x3FFFFFFF, 1, 0], dtype=np.uint32)
    assert orc.exp_length(w, 0.5, 1000).tolist() == [0, 0, 1, 1, 2, 31, 32]
    assert orc.exp_length(w, 0.5, 5).tolist() == [0, 0, 1, 1, 2, 5, 5]          # capped at D
    assert orc.exp_length(w, 1.0, 7).tolist() == [7] * 7 and orc.exp_length(w, 0.0, 7).tolist() == [0] * 7


def test_exponential_length_follows_the_reference_loop_distribution():
    """strategy.py:48-56 draws random() < CR until the first failure: P(L >= l) = CR^l, capped at D"""
    rng = np.random.default_rng(3)
    w = rng.integers(0, 1 << 32, size=400000, dtype=np.uint64).astype(np.uint32)
    for CR, D in ((0.9, 64), (0.9, 1024), (0.5, 9), (0.97, 30)):
        L = orc.exp_length(w, CR, D)
        assert L.min() >= 0 and L.max() <= D
        for l in (1, 2, 5, 10, 20, 40):
            if l <= D:
                assert abs((L >= l).mean() - CR ** l) < 4e-3, (CR, D, l)
        assert abs(L.mean() - CR * (1 - CR ** D) / (1 - CR)) < 0.05


def test_binomial_fields_are_bernoulli_cr_and_cover_every_dimension_once():
    f = orc.philox_xover_fields(0x1234, 7, np.arange(2000), 100)
    assert f.shape == (2000, 104) and f.max() < 65536
    for CR in (0.9, 0.25):
        hit = f[:, :100] < orc.crossover_threshold16(CR)
        assert abs(hit.mean() - CR) < 4e-3
    # dimension 8q+e reads bits 16*(e&1).. of word e>>1 of block q
    x = orc.philox4x32(np.uint32(3), np.uint32(5), np.uint32(7), np.uint32(orc.STREAM_XOVER), np.uint32(0x1234), np.uint32(0))
    words = [int(np.asarray(v).ravel()[0]) for v in x]
    want = [(words[e >> 1] >> (16 * (e & 1))) & 0xFFFF for e in range(8)]
    assert f[5, 24:32].tolist() == want
