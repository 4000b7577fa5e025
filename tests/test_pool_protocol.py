"""Host-side checks of K2p (mystic_b200/csrc/de_tma_pool.cuh), the pooled producer / consumer kernel that runs
BASELINE config 2's generations (Best1Bin, rows committed in place: mystic/differential_evolution.py:585-609).  No GPU.

1. The geometry the host picks (`mb2_probe_pool_geometry`, with every MB2_POOL_* override the sweeps use) satisfies what
   the kernel relies on: 2P + Q row buffers and bestSolution fit a CTA's shared memory, Q >= G, the builds' thread bounds.
2. A discrete model of the kernel's mbarrier protocol — parity tests exactly as the kernel makes them, consumers and
   producers interleaved by a random scheduler — never deadlocks, never hands a consumer a row that has not been issued,
   and never lets a copy land in a buffer that is still in use.  The model is what showed, before any GPU time was spent,
   that a consumer's parity test passes one phase early when the groups outnumber the first-donor ring (Q < G) unless the
   producer tags the slot with the row number: both the host-side clamp and the tag are in the product, and the model
   runs with the tag check switched off to show the failure it guards against.
"""
import ctypes
import random

import pytest

POOL_ENV = ('MB2_TMA_POOL', 'MB2_POOL_P', 'MB2_POOL_Q', 'MB2_POOL_G', 'MB2_POOL_GW', 'MB2_POOL_NPROD', 'MB2_POOL_SHORT')
SMEM_PER_CTA = 227 * 1024          # opt-in maximum of an sm_100 CTA
STATIC_SMEM_BOUND = 8 * 1024       # barriers, tags, scratch (cuobjdump: 5.9-6.9 KB per instantiation)


def _geometry(dim):
    from mystic_b200 import _native
    out = (ctypes.c_int32 * 6)()
    rc = _native.lib.mb2_probe_pool_geometry(dim, out)
    return None if rc != 0 else dict(zip(('P', 'Q', 'G', 'GW', 'NPROD', 'smem'), list(out)))


@pytest.fixture
def clean_env(monkeypatch):
    for k in POOL_ENV:
        monkeypatch.delenv(k, raising=False)
    return monkeypatch


def _check(dim, g):
    row = (dim * 8 + 127) // 128 * 128
    assert g['P'] >= 2 and g['Q'] >= 2 and 1 <= g['G'] <= g['Q'], (dim, g)      # Q >= G: see the module docstring
    assert g['smem'] == row * (1 + 2 * g['P'] + g['Q']), (dim, g)
    assert g['smem'] + STATIC_SMEM_BOUND <= SMEM_PER_CTA, (dim, g)
    assert g['GW'] in (1, 2, 4) and g['NPROD'] in (1, 2)
    warps = g['G'] * g['GW'] + g['NPROD']
    assert warps <= (20 if g['GW'] == 1 else 17), (dim, g)                         # __launch_bounds__(640 / 544)
    assert g['GW'] == 1 or g['G'] <= 15                                             # named barriers 1..15
    trips = -(-(dim // 2) // (4 * 32 * g['GW']))
    assert trips <= 8, (dim, g)                                                     # 8 crossover bits per trip in a 64-bit mask


def test_default_geometry_of_every_row_length(clean_env):
    seen = 0
    for dim in range(2, 2200, 2):
        g = _geometry(dim)
        if dim < 768 or dim > 1536:
            assert g is None, dim          # K2a keeps these rows (short rows: MB2_POOL_SHORT, measured slower)
            continue
        assert g is not None, dim
        _check(dim, g)
        seen += 1
    assert seen == 385
    assert _geometry(1023) is None and _geometry(1025) is None      # rows of odd length stay on K2a's aligned covers
    assert _geometry(1024) == dict(P=7, Q=12, G=7, GW=2, NPROD=2, smem=8192 * 27)   # BASELINE config 2 (profiles/README.md)


@pytest.mark.parametrize('env', [
    {'MB2_POOL_P': '2', 'MB2_POOL_Q': '3', 'MB2_POOL_G': '8'}, {'MB2_POOL_P': '8', 'MB2_POOL_Q': '10', 'MB2_POOL_G': '8'},
    {'MB2_POOL_GW': '4'}, {'MB2_POOL_GW': '4', 'MB2_POOL_G': '9'}, {'MB2_POOL_NPROD': '1'}, {'MB2_POOL_GW': '1', 'MB2_POOL_G': '30'},
    {'MB2_POOL_G': '1', 'MB2_POOL_P': '3', 'MB2_POOL_Q': '2'}, {'MB2_POOL_P': '40', 'MB2_POOL_Q': '40'}, {'MB2_POOL_SHORT': '1'},
])
def test_overrides_cannot_break_the_invariants(clean_env, env):
    for k, v in env.items():
        clean_env.setenv(k, v)
    for dim in (256, 300, 512, 766, 768, 1000, 1024, 1280, 1408, 1536):
        g = _geometry(dim)
        if g is None:
            assert dim < 768 and 'MB2_POOL_SHORT' not in env
            continue
        _check(dim, g)


def test_switch_off(clean_env):
    clean_env.setenv('MB2_TMA_POOL', '0')
    assert _geometry(1024) is None


# ---------------------------------------------------------------------------------------------------------------------
# the protocol model


class Barrier(object):
    """an mbarrier reduced to what the protocol uses: a count of completed phases; try_wait.parity(p) succeeds when the
    phase of parity p has completed, i.e. when the CURRENT phase has the other parity"""

    def __init__(self):
        self.completed = 0

    def passed(self, parity):
        return (self.completed & 1) != parity


def simulate(G, P, Q, nrows, nprod=2, tag_check=True, seed=0, max_steps=2000000, slow_landing=0.0):
    """-> 'ok', or a description of the first violation.  Rows 0..nrows-1 of one CTA; row k uses pair k % P and
    first-donor buffer k % Q; consumer group g takes rows g, g + G, ...; producers issue in row order (one producer
    does both rings; two split them: the first the pairs, the second the first-donor buffers and the tag).
    `slow_landing`: probability per scheduler round that a row whose copies have all been requested has NOT landed yet
    (0: the bytes land with the last request) -- copies complete out of order on the real memory system."""
    rng = random.Random(seed)
    full = [Barrier() for _ in range(Q)]
    full_arrivals = [0] * Q                     # arrivals of the current phase of full[q] (a phase needs nprod of them)
    pair_free = [Barrier() for _ in range(P)]
    d0_free = [Barrier() for _ in range(Q)]
    tag = [None] * Q
    pair_owner = [None] * P                     # row whose data the buffer holds / is receiving (None: free)
    d0_owner = [None] * Q
    prod = [dict(k=0) for _ in range(nprod)]
    landing = []                                # slots whose copies are all requested but have not landed
    cons = [dict(k=g, state='wait') for g in range(G)]
    done = 0
    for _ in range(max_steps):
        if done == nrows:
            return 'ok'
        progress = False
        for q in list(landing):
            if rng.random() >= slow_landing:
                landing.remove(q)
                full[q].completed += 1
                progress = True
        actors = [('p', i) for i in range(nprod)] + [('c', g) for g in range(G)]
        rng.shuffle(actors)
        for kind, i in actors:
            if rng.random() < 0.35:
                continue
            if kind == 'p':
                k = prod[i]['k']
                if k >= nrows:
                    continue
                p, q = k % P, k % Q
                do_pairs, do_d0 = i == 0, i == nprod - 1
                if do_pairs and k >= P and not pair_free[p].passed(((k // P) - 1) & 1):
                    continue
                if k >= Q and not d0_free[q].passed(((k // Q) - 1) & 1):
                    continue
                if do_pairs:
                    if pair_owner[p] is not None:
                        return 'row %d issued into pair %d still owned by row %s' % (k, p, pair_owner[p])
                    pair_owner[p] = k
                if do_d0:
                    if d0_owner[q] is not None:
                        return 'row %d issued into buffer %d still owned by row %s' % (k, q, d0_owner[q])
                    d0_owner[q] = k
                    tag[q] = k
                full_arrivals[q] += 1
                if full_arrivals[q] == nprod:   # every copy requested: the phase completes when the bytes have landed
                    full_arrivals[q] = 0
                    landing.append(q)
                prod[i]['k'] = k + 1
                progress = True
            else:
                c = cons[i]
                k = c['k']
                if k >= nrows:
                    continue
                p, q = k % P, k % Q
                if c['state'] == 'wait':
                    if not full[q].passed((k // Q) & 1):
                        continue
                    if tag_check and tag[q] != k:
                        continue                # another row's tag: wait again
                    if d0_owner[q] != k or pair_owner[p] != k:
                        return 'group %d took row %d from buffers owned by rows %s / %s' % (i, k, pair_owner[p], d0_owner[q])
                    c['state'] = 'build'
                elif c['state'] == 'build':
                    pair_owner[p] = None
                    pair_free[p].completed += 1
                    c['state'] = 'cost'
                else:
                    d0_owner[q] = None
                    d0_free[q].completed += 1
                    c['k'] = k + G
                    c['state'] = 'wait'
                    done += 1
                progress = True
        if not progress and done < nrows and not landing:
            # nobody could move in this round; with random skipping that is only a deadlock if it persists
            stuck = True
            for kind, i in actors:
                if kind == 'p':
                    k = prod[i]['k']
                    if k < nrows:
                        p, q = k % P, k % Q
                        ok = (not (i == 0 and k >= P) or pair_free[p].passed(((k // P) - 1) & 1)) and \
                             (k < Q or d0_free[q].passed(((k // Q) - 1) & 1))
                        stuck = stuck and not ok
                else:
                    c = cons[i]
                    if c['k'] < nrows:
                        q = c['k'] % Q
                        ok = c['state'] != 'wait' or (full[q].passed((c['k'] // Q) & 1) and (not tag_check or tag[q] == c['k']))
                        stuck = stuck and not ok
            if stuck:
                return 'deadlock with %d of %d rows done' % (done, nrows)
    return 'no result after %d steps' % max_steps


@pytest.mark.parametrize('G,P,Q', [(7, 7, 12), (8, 8, 10), (8, 7, 12), (8, 6, 14), (7, 8, 10), (6, 7, 12), (4, 9, 8), (3, 2, 3),
                                   (1, 3, 2), (2, 2, 2), (8, 4, 8), (17, 29, 50)])
@pytest.mark.parametrize('nprod', [1, 2])
def test_protocol_with_product_geometries(G, P, Q, nprod):
    """Q >= G, as the host enforces: every schedule completes, with and without the tag"""
    for nrows in (1, G, 16, 17, 443):
        for seed in range(3):
            assert simulate(G, P, Q, nrows, nprod, tag_check=True, seed=seed) == 'ok'
            assert simulate(G, P, Q, nrows, nprod, tag_check=False, seed=seed) == 'ok'


@pytest.mark.parametrize('nprod', [1, 2])
def test_copies_that_land_late_need_the_tag(nprod):
    """Q > G (config 2: 7 groups, 12 buffers): the slot's previous row has been ISSUED when a group looks at its next row's
    barrier, but its copies may still be in flight; with the tag every schedule completes, without it a late row is
    taken for the next one"""
    bad = 0
    for seed in range(30):
        assert simulate(7, 7, 12, 120, nprod, tag_check=True, seed=seed, slow_landing=0.97) == 'ok'
        bad += simulate(7, 7, 12, 120, nprod, tag_check=False, seed=seed, slow_landing=0.97) != 'ok'
    assert bad > 0


@pytest.mark.parametrize('nprod', [1, 2])
def test_groups_outnumbering_the_ring_need_the_tag(nprod):
    """Q < G: a group looks at its next row's barrier before the slot's previous row has even been issued, the parity
    test passes one phase early, and only the row tag keeps the group from taking another row's data"""
    bad = 0
    for seed in range(40):
        assert simulate(5, 3, 2, 60, nprod, tag_check=True, seed=seed) == 'ok'
        bad += simulate(5, 3, 2, 60, nprod, tag_check=False, seed=seed) != 'ok'
    assert bad > 0
