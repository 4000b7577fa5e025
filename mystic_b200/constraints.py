"""Device-expressible constraints: the bounds clamp.

The reference turns SetStrictRanges(tight=True) into `mystic.constraints.boundsconstrain`
(mystic/constraints.py:1582-1622), whose generated code is x[i]=max(lo,x[i]); x[i]=min(hi,x[i])
(mystic/symbolic.py:1139-1144), and applies it to every trial through `and_`
(mystic/constraints.py:526-601).  Arbitrary Python/sympy constraint functions cannot run inside
the kernel and are rejected with TypeError.
"""
import numpy as np


class BoundsClamp(object):
    """x' = min(max(x, lo), hi), as data for the kernel"""

    def __init__(self, lo, hi):
        self.lo = np.asarray(lo, dtype=np.float64)
        self.hi = np.asarray(hi, dtype=np.float64)
        if self.lo.shape != self.hi.shape or np.any(self.lo > self.hi):
            raise ValueError('each min[i] must be <= the corresponding max[i]')

    def __call__(self, x):
        raise NotImplementedError('BoundsClamp is applied inside the device kernel')


def boundsconstrain(min, max, **kwds):
    """mystic/constraints.py:1582-1622 for the clamping case (symbolic=True or clip=True)"""
    symbolic = kwds['symbolic'] if 'symbolic' in kwds else True
    clip = kwds['clip'] if 'clip' in kwds else True
    if symbolic and not clip:
        raise NotImplementedError('symbolic must clip to the nearest bound')
    if not clip:
        raise NotImplementedError('clip=False (impose_bounds re-mapping) is not available on the device')
    return BoundsClamp(min, max)


def resolve(constraints):
    if isinstance(constraints, BoundsClamp):
        return constraints
    raise TypeError("constraints %r cannot run on the device: the B200 engine supports the bounds clamp "
                    "(mystic_b200.constraints.boundsconstrain / SetStrictRanges(tight=True)) only" % (constraints,))
