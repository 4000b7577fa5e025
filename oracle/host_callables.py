"""CPU callables that let the UNMODIFIED reference solver run the same constrained
problems the device engine runs.  TEST INFRASTRUCTURE ONLY (see oracle/README.md):
imported by oracle/record_reference.py, tests/ and bench.py's reference arm, never
by the product package `mystic_b200`.

The product describes a penalty condition as data (`mystic_b200.penalty.linear`:
one row G[D] and an offset h, meaning  G.x - h).  The reference needs a Python
callable `condition(x)` (mystic/penalty.py:341-398 calls `condition(x,*args,**kwds)`).
`HostLinearCondition` is that callable, with the arithmetic order pinned
(left-to-right multiply-add, no fused multiply-add, then `- h`) so that the numpy
restatement in oracle/de_oracle.py can reproduce it bit for bit.
"""


class HostLinearCondition(object):
    """condition(x) = ((G[0]*x[0] + G[1]*x[1]) + ...) - h   (left to right)."""

    def __init__(self, G, h):
        self.G = [float(g) for g in G]
        self.h = float(h)

    def __call__(self, x, *args, **kwds):
        acc = 0.0
        G = self.G
        for j in range(len(G)):
            acc = acc + G[j] * float(x[j])
        return acc - self.h


def build_reference_penalty(terms):
    """Stack mystic.penalty decorators for a list of term dicts.

    terms: [{'ptype': 'quadratic_inequality', 'G': [...], 'h': .., 'k': 100, 'hpow': 5}, ...]
    terms[0] is applied FIRST (innermost); the reference evaluates the outermost
    term first and adds the inner result to it (mystic/penalty.py:88,147,388,447).
    Requires the reference on sys.path (record_reference.py sets it up).
    """
    import mystic.penalty as mp

    def base(x):
        return 0.0

    f = base
    for t in terms:
        dec = getattr(mp, t['ptype'])(HostLinearCondition(t['G'], t['h']),
                                      k=t.get('k', 100), h=t.get('hpow', 5))
        f = dec(f)
    # penalty iterations done before the run (the lagrange decorators accumulate their multiplier from the
    # condition values `store(x, i)` recorded at each: penalty.py:489-500, 511-515).  `store_points` of a term:
    # one point per iteration; the resulting values go back into the term as 'stored' so that the oracle and
    # the device can be handed the same numbers.
    n = max([t.get('n', 0) for t in terms] or [0])
    if n:
        g = f
        for t in reversed(terms):      # outermost first
            for i, x in enumerate(t.get('store_points', ())):
                g.store(x, i)          # (passes through to the wrapped functions as well; each keeps its own list)
            g = dict(zip(g.__code__.co_freevars, (c.cell_contents for c in g.__closure__)))['f']
        f.iter(n)
        g = f
        for t in reversed(terms):
            if t['ptype'].startswith('lagrange'):
                t['stored'] = list(g.stored())
            t['n'] = n
            g = dict(zip(g.__code__.co_freevars, (c.cell_contents for c in g.__closure__)))['f']
    return f
