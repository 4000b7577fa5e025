"""Strategy tokens with the names of mystic.strategy (mystic/strategy.py:34-311).

The reference passes strategy FUNCTIONS to Solve() and stores them by `__name__`
(mystic/differential_evolution.py:682,692).  The engine runs mutation + crossover inside the
fused CUDA kernel, so these are tokens: the solver maps `__name__` to a device enum, which also
means the unmodified `mystic.strategy.Best1Bin` etc. can be passed instead.  Calling a token
raises -- there is no host implementation.
"""

# name -> (device id, number of donors, uses bestSolution)
STRATEGY_IDS = {
    'Best1Exp': (0, 2, True), 'Best1Bin': (1, 2, True), 'Rand1Exp': (2, 3, False),
    'RandToBest1Exp': (3, 2, True), 'Best2Exp': (4, 4, True), 'Rand2Exp': (5, 5, False),
    'Rand1Bin': (6, 3, False), 'RandToBest1Bin': (7, 2, True), 'Best2Bin': (8, 4, True),
    'Rand2Bin': (9, 5, False),
}


def _token(name, doc):
    def strategy(inst, candidate):
        raise NotImplementedError('%s runs inside the fused device kernel; it is not callable on the host' % name)
    strategy.__name__ = name
    strategy.__qualname__ = name
    strategy.__doc__ = doc
    return strategy


Best1Exp = _token('Best1Exp', 'trial = best + scale*(candidate1 - candidate2); mutates until random stop')
Best1Bin = _token('Best1Bin', 'trial = best + scale*(candidate1 - candidate2); mutates at random')
Rand1Exp = _token('Rand1Exp', 'trial = candidate1 + scale*(candidate2 - candidate3); mutates until random stop')
RandToBest1Exp = _token('RandToBest1Exp', 'trial += scale*(best - trial) + scale*(candidate1 - candidate2)')
Best2Exp = _token('Best2Exp', 'trial = best + scale*(candidate1 + candidate2 - candidate3 - candidate4)')
Rand2Exp = _token('Rand2Exp', 'trial = candidate1 + scale*(candidate2 + candidate3 - candidate4 - candidate5)')
Rand1Bin = _token('Rand1Bin', 'as Rand1Exp (the reference runs the exponential loop, strategy.py:203-227)')
RandToBest1Bin = _token('RandToBest1Bin', 'as RandToBest1Exp (strategy.py:229-255)')
Best2Bin = _token('Best2Bin', 'as Best2Exp (strategy.py:257-283)')
Rand2Bin = _token('Rand2Bin', 'as Rand2Exp (strategy.py:285-311)')


def strategy_id(strategy):
    """name or function -> (id, k, uses_best); unknown names fall back to Best1Bin like the
    reference's getattr(strategy, self.strategy, strategy.Best1Bin) (differential_evolution.py:682)."""
    name = strategy if isinstance(strategy, str) else getattr(strategy, '__name__', 'Best1Bin')
    return STRATEGY_IDS.get(name, STRATEGY_IDS['Best1Bin']), (name if name in STRATEGY_IDS else 'Best1Bin')
