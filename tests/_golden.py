"""Helpers shared by the oracle and GPU parity tests: load a recorded reference run."""
import json
import os

import numpy as np

from oracle import de_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz'))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(str(z['meta']))
    return meta, z


FIT_COSTS = ('polyfit', 'lorentzian', 'decay')
# costs the reference sums with the builtin sum(): CPython 3.12 compensates runs of exact floats
# (oracle/de_oracle.py:cost_griewangk), so their ENERGIES are compared to 1e-14, everything else bit for bit
COMPENSATED_SUM_COSTS = ('griewangk', 'rastrigin', 'schwefel', 'ackley', 'michal', 'powers', 'ellipsoid')

# costs the reference evaluates with numpy's vectorised exp (models/br8.py:32): within 1 ulp of libm's, so the C
# port's ENERGIES are compared to 1e-14 (the numpy oracle reproduces them bit for bit)
NUMPY_EXP_COSTS = ('decay',)


def extra_of(meta, z):
    """ExtraArgs of the recorded run for orc.make_cost / c_port.COracle: data-fit costs carry their
    evaluation points and observations in the recording"""
    if meta['cost'] in FIT_COSTS:
        return (z['fit_pts'], z['fit_obs'], z['fit_sigma'] if 'fit_sigma' in z.files else 1.0)
    return tuple(meta.get('ExtraArgs') or ())


def device_cost_of(meta, z):
    """the mystic_b200.models object of the recorded run"""
    from mystic_b200 import models
    if meta['cost'] == 'polyfit':
        return models.poly.CostFactory2(z['fit_pts'], z['fit_obs'], meta['dim'])
    if meta['cost'] == 'lorentzian':
        return models.lorentzian.CostFactory2(z['fit_pts'], z['fit_obs'], meta['dim'])
    if meta['cost'] == 'decay':
        return models.decay.CostFactory2(z['fit_pts'], z['fit_obs'], meta['dim'])
    return getattr(models, meta['cost'])


def bounds_of(meta):
    b = meta.get('bounds')
    if b is None:
        return None, None, orc.BOUNDS_NONE
    lo, hi, tight = b
    return np.asarray(lo, float), np.asarray(hi, float), (orc.BOUNDS_CLAMP if tight else orc.BOUNDS_REJECT)


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bits_equal(a, b, what=''):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError('%s: %d mismatches, first at %s: %r vs %r' % (what, len(bad), i, a[i], b[i]))


def restore_mt_state(z):
    """put stdlib `random` into the state the reference had right before Solve()"""
    import random
    random.setstate((3, tuple(int(v) for v in z['mt_state']), None))


def device_penalty_of(meta):
    """the recorded run's penalty stack built with mystic_b200.penalty's decorators, penalty iterations included:
    the same store(x, i) / iter(n) calls oracle/host_callables.build_reference_penalty made on the reference's
    decorators (outermost term first; a store passes through to the wrapped terms, which then store their own points)"""
    from mystic_b200 import penalty

    def pen(x):
        return 0.0
    for t in meta['penalty']:
        pen = getattr(penalty, t['ptype'])(penalty.linear(t['G'], t['h']), k=t['k'], h=t['hpow'])(pen)
    n = max([t.get('n', 0) for t in meta['penalty']] or [0])
    if n:
        for ti in reversed(range(len(meta['penalty']))):
            sub = penalty.DevicePenalty(pen.terms[:ti + 1])
            for i, x in enumerate(meta['penalty'][ti].get('store_points', ())):
                sub.store(x, i)
        pen.iter(n)
        for t, d in zip(meta['penalty'], pen.terms):
            if t['ptype'].startswith('lagrange'):
                assert d['y'] == t['stored'], (d['y'], t['stored'])   # the host's condition values equal the reference's
    return pen
