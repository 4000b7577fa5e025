"""Device cost descriptors with the names of mystic.models.

    rosen       mystic/models/dejong.py:89-101
    corana      mystic/models/storn.py:57-96
    griewangk   mystic/models/storn.py:120-139
    zimmermann  mystic/models/storn.py:161-191
    sphere, step   mystic/models/dejong.py:54-66, 166-192
    rastrigin, schwefel, ackley, michal, powers   mystic/models/pohlheim.py:60-71, 123-132, 186-204, 226-240, 153-163
    branins, easom, goldstein, ellipsoid   mystic/models/pohlheim.py:261-282, 303-313, 336-355, 92-102
    peaks, venkat91, fosc3d, nmin51, shekel, paviani   mystic/models/nag.py:48-59, venkataraman.py:38-52,
                wolfram.py:44-64, 87-103, dejong.py:252-278, schittkowski.py:46-71
    chebyshev{2,4,6,8,16}cost   mystic/models/poly.py:124-149, mystic/models/_model_helper.py:10-33
    poly.CostFactory / poly.CostFactory2, lorentzian.CostFactory / lorentzian.CostFactory2, decay.CostFactory / decay.CostFactory2
                data-fit residual costs, mystic/models/abstract_model.py:146-161 ->
                mystic/forward_model.py:204-260 with mystic/models/poly.py:36-54,
                mystic/models/lorentzian.py:32-37 and mystic/models/br8.py:27-56

A descriptor names a cost function compiled into the fused kernel.  `resolve()` also recognises
the unmodified reference callables (mystic.models.rosen is `Rosenbrock().function`, etc.) so
user scripts written against mystic keep working.  Arbitrary Python callables are rejected
with TypeError: the engine has no CPU path.
"""

COST_ROSEN, COST_CORANA, COST_GRIEWANGK, COST_ZIMMERMANN, COST_CHEBYSHEV, COST_CHEBYSHEV_MMA = range(6)
COST_POLYFIT, COST_POLYFIT_MMA, COST_LORENTZIAN = 6, 7, 8
COST_SPHERE, COST_RASTRIGIN, COST_SCHWEFEL, COST_ACKLEY, COST_STEP, COST_MICHAL, COST_POWERS = 9, 10, 11, 12, 13, 14, 15
COST_DECAY = 26
(COST_BRANINS, COST_EASOM, COST_GOLDSTEIN, COST_PEAKS, COST_VENKAT91, COST_FOSC3D, COST_NMIN51, COST_SHEKEL, COST_ELLIPSOID,
 COST_PAVIANI) = range(16, 26)

CHEBYSHEV_TARGETS = {  # mystic/models/_model_helper.py:10-14
    2: [2., 0., -1.],
    4: [8., 0., -8., 0., 1.],
    6: [32., 0., -48., 0., 18., 0., -1.],
    8: [128., 0., -256., 0., 160., 0., -32., 0., 1.],
    16: [32768., 0., -131072., 0., 212992., 0., -180224., 0., 84480., 0., -21504., 0., 2688., 0., -128., 0., 1.],
}


class DeviceCost(object):
    """A cost function that exists as a device function of the fused kernel."""

    def __init__(self, name, cost_id, order=None, doc=''):
        self.__name__ = name
        self.cost_id = cost_id
        self.order = order
        self.__doc__ = doc

    def device_id(self, tensor_cores=False, dim=None):
        """cost id handed to mb2_set_cost.  The Chebyshev costs have a second implementation that
        runs the coefficient x abscissa contraction on the FP64 tensor cores (DMMA); it is not
        bit-faithful to the reference's Horner evaluation, so it is opt-in (tensor_cores=True)."""
        if tensor_cores and self.cost_id == COST_CHEBYSHEV and (dim is None or dim <= 17):
            return COST_CHEBYSHEV_MMA
        return self.cost_id

    def params(self, extra_args=()):
        """numeric parameter vector handed to mb2_set_cost"""
        if self.cost_id in (COST_CHEBYSHEV, COST_CHEBYSHEV_MMA):
            M = int(extra_args[0]) if extra_args else 61  # chebyshevNcost(trial, M=61), poly.py:124-149
            return [float(M), float(self.order)] + [float(c) for c in CHEBYSHEV_TARGETS[self.order]]
        if extra_args:
            raise TypeError('%s() takes no ExtraArgs' % self.__name__)
        return []

    def __call__(self, x, *extra_args, **kwds):
        """Evaluate on the device (one row, or a [n, D] batch) through mb2_eval_cost."""
        from .engine import evaluate_rows
        import numpy as np
        arr = np.asarray(x, dtype=np.float64)
        out = evaluate_rows(self, np.atleast_2d(arr), extra_args, tensor_cores=kwds.get('tensor_cores', False))
        return float(out[0]) if arr.ndim == 1 else out

    def __repr__(self):
        return '<mystic_b200 device cost %s>' % self.__name__


class DataFitCost(DeviceCost):
    """sum(((F(params)(x_i) - obs_i)/sigma)**2) over the evaluation points: what
    mystic.forward_model.CostFactory.getCostFunction builds for ONE model with the default L2
    metric (forward_model.py:204-260).  F is the polynomial (numpy.poly1d) or the lorentzian model."""

    def __init__(self, name, cost_id, pts, observations, nparams, sigma=1.0):
        import numpy as np
        DeviceCost.__init__(self, name, cost_id, doc=self.__doc__)
        self.pts = np.ascontiguousarray(pts, dtype=np.float64).ravel()
        self.observations = np.ascontiguousarray(observations, dtype=np.float64).ravel()
        if self.pts.shape != self.observations.shape or self.pts.size < 1:
            raise ValueError('evaluation points and observations must be two equally long, non-empty vectors')
        self.nparams = int(nparams)
        # getCostFunction(sigma=None) does not divide; an array divides point by point (forward_model.py:252-255)
        if sigma is not None and np.ndim(sigma) > 0:
            self.sigma = np.ascontiguousarray(sigma, dtype=np.float64).ravel()
            if self.sigma.shape != self.pts.shape:
                raise ValueError('sigma must be a scalar or have one entry per evaluation point')
            if not np.all(self.sigma != 0.0):
                raise ZeroDivisionError('sigma')
        else:
            self.sigma = 1.0 if sigma is None else float(sigma)
            if self.sigma == 0.0:
                raise ZeroDivisionError('sigma')

    def device_id(self, tensor_cores=False, dim=None):
        """the polynomial fit has a second implementation on the FP64 tensor cores (the coefficient x
        point contraction as DMMA); like the Chebyshev one it is not bit-faithful to the Horner order
        of numpy.polyval, so it is opt-in (tensor_cores=True)."""
        if tensor_cores and self.cost_id == COST_POLYFIT and (dim is None or dim <= 17) and isinstance(self.sigma, float):
            return COST_POLYFIT_MMA
        return self.cost_id

    def params(self, extra_args=()):
        if extra_args:
            raise TypeError('%s() takes no ExtraArgs' % self.__name__)
        if not isinstance(self.sigma, float):
            return [float(self.pts.size), 1.0] + self.pts.tolist() + self.observations.tolist() + self.sigma.tolist()
        return [float(self.pts.size), self.sigma] + self.pts.tolist() + self.observations.tolist()


class _FitModel(object):
    """mystic.models.poly / mystic.models.lorentzian as far as the DE path needs them
    (abstract_model.py:103-161): evaluate, ForwardFactory, CostFactory, CostFactory2."""

    def __init__(self, name, cost_id, nparams=None):
        self.__name__ = name
        self._cost_id = cost_id
        self._nparams = nparams

    def evaluate(self, coeffs, x):
        """the forward model on the host, in the reference's operation order (used to make observations)"""
        import numpy as np
        x = np.asarray(x, dtype=np.float64)
        if self._cost_id == COST_POLYFIT:
            val = 0 * x                      # mystic/math/poly.py:37-39
            for c in coeffs:
                val = c + val * x
            return val
        if self._cost_id == COST_DECAY:
            a1, a2, a3, a4, a5 = [float(c) for c in coeffs]      # models/br8.py:27-32
            return a1 + a2*np.exp(-x/a4) + a3*np.exp(-x/a5)
        a1, a2, a3, A0, E0, G0, n = [float(c) for c in coeffs]   # models/lorentzian.py:32-37
        return (a1 + a2*x + a3*x*x + A0 * ( G0/(2*np.pi) )/( (x-E0)*(x-E0)+(G0/2)*(G0/2) ))/n

    def ForwardFactory(self, coeffs):
        coeffs = list(coeffs)
        return lambda evalpts: self.evaluate(coeffs, evalpts)

    def _check(self, nparams):
        if self._nparams is not None and int(nparams) != self._nparams:
            raise ValueError('%s has %d parameters' % (self.__name__, self._nparams))

    def _sigma(self, datapts):
        """models/br8.py:47,55: the decay model weighs every residual by sqrt(datapts); the others by 1"""
        import numpy as np
        return np.sqrt(np.asarray(datapts, dtype=np.float64)) if self._cost_id == COST_DECAY else 1.0

    def CostFactory(self, target, pts):
        """cost function from a list of target coefficients and evaluation points (abstract_model.py:146-153)"""
        self._check(len(target))
        datapts = self.evaluate(target, pts)
        return DataFitCost(self.__name__ + '.cost', self._cost_id, pts, datapts, len(target), self._sigma(datapts))

    def CostFactory2(self, pts, datapts, nparams):
        """cost function from data points and evaluation points (abstract_model.py:155-161)"""
        self._check(nparams)
        return DataFitCost(self.__name__ + '.cost', self._cost_id, pts, datapts, nparams, self._sigma(datapts))


poly = _FitModel('poly', COST_POLYFIT)
lorentzian = _FitModel('lorentz', COST_LORENTZIAN, 7)
decay = _FitModel('decay', COST_DECAY, 5)          # mystic.models.decay (br8.py: Bevington's dual exponential decay)

rosen = DeviceCost('rosen', COST_ROSEN, doc="Rosenbrock's saddle, sum(100*(x[1:]-x[:-1]**2)**2 + (1-x[:-1])**2)")
corana = DeviceCost('corana', COST_CORANA, doc="Corana's parabola (first four coordinates)")
griewangk = DeviceCost('griewangk', COST_GRIEWANGK, doc="Griewangk's function")
zimmermann = DeviceCost('zimmermann', COST_ZIMMERMANN, doc="Zimmermann's function (2-D)")
sphere = DeviceCost('sphere', COST_SPHERE, doc='De Jong sphere, sum(x_i^2)')
rastrigin = DeviceCost('rastrigin', COST_RASTRIGIN, doc="Rastrigin's function, 10 N + sum(x_i^2 - 10 cos(2 pi x_i))")
schwefel = DeviceCost('schwefel', COST_SCHWEFEL, doc="Schwefel's function, sum(-x_i sin(sqrt(|x_i|)))")
ackley = DeviceCost('ackley', COST_ACKLEY, doc="Ackley's path function")
step = DeviceCost('step', COST_STEP, doc="De Jong's step function")
michal = DeviceCost('michal', COST_MICHAL, doc="Michalewicz's function")
powers = DeviceCost('powers', COST_POWERS, doc="Pohlheim's sum of different powers")
branins = DeviceCost('branins', COST_BRANINS, doc="Branins's rcos function (2-D)")
easom = DeviceCost('easom', COST_EASOM, doc="Easom's function (2-D)")
goldstein = DeviceCost('goldstein', COST_GOLDSTEIN, doc="Goldstein-Price's function (2-D)")
peaks = DeviceCost('peaks', COST_PEAKS, doc="NAG's peaks function (2-D)")
venkat91 = DeviceCost('venkat91', COST_VENKAT91, doc="Venkataraman's sinc function (2-D)")
fosc3d = DeviceCost('fosc3d', COST_FOSC3D, doc="Wolfram's fOsc3D function (2-D)")
nmin51 = DeviceCost('nmin51', COST_NMIN51, doc="Wolfram's NMinimize test function 51 (2-D)")
shekel = DeviceCost('shekel', COST_SHEKEL, doc="Shekel's foxholes (2-D)")
ellipsoid = DeviceCost('ellipsoid', COST_ELLIPSOID, doc="Pohlheim's rotated hyper-ellipsoid, sum of squared prefix sums")
paviani = DeviceCost('paviani', COST_PAVIANI, doc="Paviani's function (x_i in (2, 10))")
chebyshev2cost = DeviceCost('chebyshev2cost', COST_CHEBYSHEV, 2, 'order-2 Chebyshev fitting cost')
chebyshev4cost = DeviceCost('chebyshev4cost', COST_CHEBYSHEV, 4, 'order-4 Chebyshev fitting cost')
chebyshev6cost = DeviceCost('chebyshev6cost', COST_CHEBYSHEV, 6, 'order-6 Chebyshev fitting cost')
chebyshev8cost = DeviceCost('chebyshev8cost', COST_CHEBYSHEV, 8, 'order-8 Chebyshev fitting cost')
chebyshev16cost = DeviceCost('chebyshev16cost', COST_CHEBYSHEV, 16, 'order-16 Chebyshev fitting cost')

_BY_NAME = dict((c.__name__, c) for c in (rosen, corana, griewangk, zimmermann, sphere, rastrigin, schwefel, ackley, step, michal, powers, branins, easom,
                                          goldstein, peaks, venkat91, fosc3d, nmin51, shekel, ellipsoid, paviani, chebyshev2cost, chebyshev4cost,
                                          chebyshev6cost, chebyshev8cost, chebyshev16cost))
_BY_REFERENCE_CLASS = {'Rosenbrock': rosen, 'Corana': corana, 'Griewangk': griewangk, 'Zimmermann': zimmermann,
                       'Sphere': sphere, 'Rastrigin': rastrigin, 'Schwefel': schwefel, 'Ackley': ackley,
                       'Step': step, 'Michalewicz': michal, 'DifferentPowers': powers, 'Branins': branins, 'Easom': easom,
                       'GoldsteinPrice': goldstein, 'Peaks': peaks, 'Sinc': venkat91, 'fOsc3D': fosc3d, 'NMinimize51': nmin51,
                       'Shekel': shekel, 'HyperEllipsoid': ellipsoid, 'Paviani': paviani}


def _resolve_reference_costfactory(cost):
    """The closure mystic.forward_model.CostFactory.getCostFunction returns (forward_model.py:241-258),
    e.g. from mystic.models.poly.CostFactory(target, pts): recognised when it holds exactly ONE model
    whose ForwardFactory belongs to mystic.models.poly.Polynomial or mystic.models.lorentzian.Lorentzian,
    no input checker / output filter, and the default L2 metric."""
    import numpy as np
    code = getattr(cost, '__code__', None)
    if code is None or getattr(cost, '__module__', '') != 'mystic.forward_model' or not cost.__closure__:
        return None
    cells = dict(zip(code.co_freevars, (c.cell_contents for c in cost.__closure__)))
    if not all(k in cells for k in ('self', 'evalpts', 'observations', 'sigma', 'metric')):
        return None
    cf = cells['self']
    fac, nin = getattr(cf, '_forwardFactories', None), getattr(cf, '_inputs', None)
    if not fac or len(fac) != 1:
        raise TypeError('the B200 engine evaluates CostFactory costs with exactly one forward model')
    owner = getattr(fac[0], '__self__', None)
    kind = {'Polynomial': COST_POLYFIT, 'Lorentzian': COST_LORENTZIAN, 'BevingtonDecay': COST_DECAY}.get(type(owner).__name__)
    if kind is None or not type(owner).__module__.startswith('mystic.models'):
        raise TypeError('only the poly, lorentzian and decay forward models of mystic.models exist on the device')
    from_null = lambda f: type(f).__name__ in ('NullChecker', 'function') and f(np.ones(1), np.ones(1)) is None  # noqa: E731
    ident = lambda f: np.array_equal(f(np.array([1.5, -2.0])), np.array([1.5, -2.0]))  # noqa: E731
    if not from_null(cf._inputCheckers[0]) or not ident(cf._outputFilters[0]):
        raise TypeError('input checkers / output filters of a CostFactory cannot run on the device')
    probe = np.array([1.0, -2.0, 0.5])
    if float(cells['metric'](probe)) != 5.25:
        raise TypeError('only the default L2 metric, sum(x*x), exists on the device')
    name = {COST_POLYFIT: 'poly.cost', COST_LORENTZIAN: 'lorentz.cost', COST_DECAY: 'decay.cost'}[kind]
    return DataFitCost(name, kind, cells['evalpts'], cells['observations'], nin[0], cells['sigma'])


def resolve(cost):
    """DeviceCost | reference callable | name -> DeviceCost, else TypeError."""
    if isinstance(cost, DeviceCost):
        return cost
    if isinstance(cost, str) and cost in _BY_NAME:
        return _BY_NAME[cost]
    mod = getattr(cost, '__module__', '') or ''
    owner = getattr(cost, '__self__', None)
    if owner is not None and getattr(cost, '__name__', '') == 'function':
        # mystic.models.rosen = Rosenbrock().function (dejong.py:288), storn.py:201-203
        kls = type(owner).__name__
        if type(owner).__module__.startswith('mystic.models') and kls in _BY_REFERENCE_CLASS:
            if kls == 'Rosenbrock' and getattr(owner, 'axis', None) is not None:
                raise TypeError('Rosenbrock(axis=...) is not supported on the device')
            return _BY_REFERENCE_CLASS[kls]
    fit = _resolve_reference_costfactory(cost)
    if fit is not None:
        return fit
    name = getattr(cost, '__name__', None)
    if mod.startswith('mystic.models') and name in _BY_NAME:
        return _BY_NAME[name]  # mystic.models.poly.chebyshev8cost etc.
    raise TypeError(
        "cost %r is not a device cost function: the B200 engine evaluates the cost inside the fused "
        "CUDA kernel and has no CPU path. Use mystic_b200.models.{rosen,corana,griewangk,zimmermann,"
        "sphere,rastrigin,schwefel,ackley,step,michal,powers,branins,easom,goldstein,peaks,venkat91,fosc3d,nmin51,shekel,ellipsoid,paviani,chebyshevNcost,poly.CostFactory,lorentzian.CostFactory} or the "
        "corresponding mystic.models objects." % (cost,))
