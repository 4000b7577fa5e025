"""Device-expressible penalties with the decorator API of mystic.penalty.

The reference's penalty decorators wrap an arbitrary Python `condition(x)`
(mystic/penalty.py:41-157, 341-457).  The device needs data, so the condition here is a linear
form `linear(G, h)`: G.x - h.  Usage is the reference's:

    from mystic_b200.penalty import quadratic_inequality, linear
    @quadratic_inequality(linear([1.]*64, 100.), k=100, h=5)
    def penalty(x): return 0.0
    solver.SetPenalty(penalty)

`SetPenalty` also accepts a function decorated with the UNMODIFIED mystic.penalty decorators as
long as every condition in the stack is a `LinearCondition` (k, h and the iteration count are
read from the decorator's closure).
"""

PTYPES = {'quadratic_equality': 0, 'linear_equality': 1, 'quadratic_inequality': 2, 'linear_inequality': 3}


class LinearCondition(object):
    """condition(x) = G.x - h, as data.  Not evaluated on the host."""

    def __init__(self, G, h=0.0):
        self.G = [float(g) for g in G]
        self.h = float(h)

    def __call__(self, x, *args, **kwds):
        raise NotImplementedError('LinearCondition is evaluated inside the device kernel '
                                  '(for a host-callable twin see oracle/host_callables.py)')

    def __repr__(self):
        return 'linear(G[%d], h=%r)' % (len(self.G), self.h)


def linear(G, h=0.0):
    return LinearCondition(G, h)


class DevicePenalty(object):
    """A stack of penalty terms; term 0 is the innermost decorator."""

    def __init__(self, terms):
        self.terms = list(terms)  # dicts: ptype, condition, k, h, n

    # interface of the reference's decorated function (penalty.py:80-98)
    @property
    def ptype(self):
        return self.terms[-1]['ptype']

    @property
    def func(self):
        return self.terms[-1]['condition']

    def iter(self, i=None):
        for t in self.terms:
            t['n'] = t['n'] + 1 if i is None else i

    def iteration(self):
        return self.terms[-1]['n']

    def clear(self):
        for t in self.terms:
            t['n'] = 0

    def __call__(self, x, *args, **kwds):
        raise NotImplementedError('device penalties are evaluated inside the fused kernel')

    def device_terms(self, dim):
        """-> (ptype ids, G rows, h, kscale) for mb2_set_penalty"""
        ptypes, G, h, ks = [], [], [], []
        for t in self.terms:
            cond = t['condition']
            if len(cond.G) != dim:
                raise ValueError('penalty condition has %d coefficients, solver has nDim=%d' % (len(cond.G), dim))
            ptypes.append(PTYPES[t['ptype']])
            G.append(list(cond.G))
            h.append(cond.h)
            ks.append(float(t['k'] * pow(t['h'], t['n'])))  # _k = k * pow(h, n), penalty.py:87
        return ptypes, G, h, ks


def _decorator(ptype):
    def factory(condition=None, args=None, kwds=None, k=100, h=5):
        if not isinstance(condition, LinearCondition):
            raise TypeError('mystic_b200.penalty.%s needs a mystic_b200.penalty.linear(G, h) condition; '
                            'arbitrary Python conditions cannot run on the device' % ptype)
        if args or kwds:
            raise TypeError('linear conditions take no args/kwds')

        def dec(f):
            inner = f.terms if isinstance(f, DevicePenalty) else []
            return DevicePenalty(inner + [dict(ptype=ptype, condition=condition, k=k, h=h, n=0)])
        return dec
    factory.__name__ = ptype
    return factory


quadratic_equality = _decorator('quadratic_equality')
linear_equality = _decorator('linear_equality')
quadratic_inequality = _decorator('quadratic_inequality')
linear_inequality = _decorator('linear_inequality')


def resolve(penalty):
    """DevicePenalty | reference-decorated function with LinearCondition conditions -> DevicePenalty."""
    if isinstance(penalty, DevicePenalty):
        return penalty
    terms = []
    f = penalty
    while hasattr(f, 'ptype') and hasattr(f, 'func'):
        if f.ptype not in PTYPES or not isinstance(f.func, LinearCondition):
            raise TypeError('penalty %r (%s) is not device-expressible' % (f, getattr(f, 'ptype', '?')))
        cells = dict(zip(f.__code__.co_freevars, (c.cell_contents for c in f.__closure__)))
        terms.append(dict(ptype=f.ptype, condition=f.func, k=cells['k'], h=cells['h'], n=f.iteration()))
        f = cells['f']
    if not terms:
        raise TypeError("penalty %r is not a device penalty: build it with mystic_b200.penalty decorators "
                        "(or mystic.penalty decorators over mystic_b200.penalty.linear conditions)" % (penalty,))
    terms.reverse()  # innermost first
    return DevicePenalty(terms)
