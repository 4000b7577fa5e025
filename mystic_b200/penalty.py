"""Device-expressible penalties with the decorator API of mystic.penalty.

The reference's penalty decorators wrap an arbitrary Python `condition(x)`
(mystic/penalty.py:41-601: quadratic / linear / uniform / lagrange equality and inequality, barrier inequality).  The device needs data, so the condition here is a linear
form `linear(G, h)`: G.x - h.  Usage is the reference's:

    from mystic_b200.penalty import quadratic_inequality, linear
    @quadratic_inequality(linear([1.]*64, 100.), k=100, h=5)
    def penalty(x): return 0.0
    solver.SetPenalty(penalty)

`SetPenalty` also accepts a function decorated with the UNMODIFIED mystic.penalty decorators as
long as every condition in the stack is a `LinearCondition` (k, h and the iteration count are
read from the decorator's closure).
"""

PTYPES = {'quadratic_equality': 0, 'linear_equality': 1, 'quadratic_inequality': 2, 'linear_inequality': 3,
          'uniform_equality': 4, 'uniform_inequality': 5, 'barrier_inequality': 6, 'lagrange_inequality': 7,
          'lagrange_equality': 8}
# the decorators' default multipliers (penalty.py:41,100,159,218,276,341,400,459,532)
DEFAULT_K = {'quadratic_equality': 100, 'linear_equality': 100, 'quadratic_inequality': 100, 'linear_inequality': 100,
             'uniform_equality': float('inf'), 'uniform_inequality': float('inf'), 'barrier_inequality': 100,
             'lagrange_inequality': 20, 'lagrange_equality': 20}
LAGRANGE = ('lagrange_inequality', 'lagrange_equality')


def condition_value(cond, x):
    """G.x - h on the host, left to right without fused multiply-add (the order of oracle/host_callables.py);
    used by store(): one dot product per penalty iteration, not a cost evaluation"""
    acc = 0.0
    for g, v in zip(cond.G, x):
        acc = acc + g * float(v)
    return acc - cond.h


def lagrange_state(ptype, k, h, n, stored):
    """-> (_k, multiplier) after the decorator's loop over the first n stored condition values
    (penalty.py:511-515 beta of lagrange_inequality, :579-582 lam of lagrange_equality)"""
    mult, _k = 0., k
    for i in range(n):
        y = stored[i] if i < len(stored) else 0.0
        if ptype == 'lagrange_inequality':
            mult += 2. * _k * max(-mult / (2. * _k), y)
        else:
            mult += 2. * _k * y
        _k *= h
    return _k, mult


class LinearCondition(object):
    """condition(x) = G.x - h, as data: the generation step evaluates it inside the device kernel for every
    candidate.  Calling it evaluates ONE point on the host -- what the decorators' bookkeeping needs
    (`store(x)` of the lagrange kinds, `error(x)`: penalty.py:489-500, 471-477), never the solver."""

    def __init__(self, G, h=0.0):
        self.G = [float(g) for g in G]
        self.h = float(h)

    def __call__(self, x, *args, **kwds):
        return condition_value(self, x)

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

    def store(self, x, i=None):
        """record the condition value at x for penalty iteration i (the lagrange kinds only: penalty.py:489-500,
        557-568; the other decorators pass the call through to the function they wrap)"""
        for t in self.terms:
            if t['ptype'] in LAGRANGE:
                y = condition_value(t['condition'], x)
                ys = t.setdefault('y', [])
                j = t['n'] if i is None else i
                if j >= len(ys):
                    ys.extend([0.] * (j - len(ys)) + [y])
                else:
                    ys[j] = y

    def stored(self, i=None):
        """of the outermost term, like the reference's decorated function"""
        ys = self.terms[-1].get('y', [])
        if i is None:
            return ys[:]
        try:
            return ys[i]
        except IndexError:
            return 0.0

    def clear(self):
        for t in self.terms:
            t['n'] = 0
            t['y'] = []

    def __call__(self, x, *args, **kwds):
        raise NotImplementedError('device penalties are evaluated inside the fused kernel')

    def device_terms(self, dim):
        """-> (ptype ids, G rows, h, kscale, multipliers) for mb2_set_penalty_ex"""
        ptypes, G, h, ks, mult = [], [], [], [], []
        for t in self.terms:
            cond = t['condition']
            if len(cond.G) != dim:
                raise ValueError('penalty condition has %d coefficients, solver has nDim=%d' % (len(cond.G), dim))
            ptypes.append(PTYPES[t['ptype']])
            G.append(list(cond.G))
            h.append(cond.h)
            if t['ptype'] in LAGRANGE:
                _k, m = lagrange_state(t['ptype'], t['k'], t['h'], t['n'], t.get('y', []))
                ks.append(float(_k))
                mult.append(float(m))
            else:
                ks.append(float(t['k']) * pow(t['h'], t['n']))  # _k = k * pow(h, n), penalty.py:87
                mult.append(0.0)
        return ptypes, G, h, ks, mult


def _decorator(ptype):
    def factory(condition=None, args=None, kwds=None, k=DEFAULT_K[ptype], h=5):
        if not isinstance(condition, LinearCondition):
            raise TypeError('mystic_b200.penalty.%s needs a mystic_b200.penalty.linear(G, h) condition; '
                            'arbitrary Python conditions cannot run on the device' % ptype)
        if args or kwds:
            raise TypeError('linear conditions take no args/kwds')

        def dec(f):
            inner = f.terms if isinstance(f, DevicePenalty) else []
            return DevicePenalty(inner + [dict(ptype=ptype, condition=condition, k=k, h=h, n=0, y=[])])
        return dec
    factory.__name__ = ptype
    return factory


quadratic_equality = _decorator('quadratic_equality')
linear_equality = _decorator('linear_equality')
quadratic_inequality = _decorator('quadratic_inequality')
linear_inequality = _decorator('linear_inequality')
uniform_equality = _decorator('uniform_equality')
uniform_inequality = _decorator('uniform_inequality')
barrier_inequality = _decorator('barrier_inequality')
lagrange_inequality = _decorator('lagrange_inequality')
lagrange_equality = _decorator('lagrange_equality')


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
        terms.append(dict(ptype=f.ptype, condition=f.func, k=cells['k'], h=cells['h'], n=f.iteration(),
                          y=list(f.stored()) if f.ptype in LAGRANGE else []))
        f = cells['f']
    if not terms:
        raise TypeError("penalty %r is not a device penalty: build it with mystic_b200.penalty decorators "
                        "(or mystic.penalty decorators over mystic_b200.penalty.linear conditions)" % (penalty,))
    terms.reverse()  # innermost first
    return DevicePenalty(terms)
