"""CPU restatement (numpy) of mystic's DifferentialEvolutionSolver2 generation step.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py.  The product package
`mystic_b200` never imports anything from oracle/.

PARITY PINNED: every function below is checked bit-for-bit against recordings of the
unmodified reference (tests/golden/*.npz, produced by oracle/record_reference.py from
/root/reference) in tests/test_oracle_golden.py, plus the known-answer vectors of
SURVEY.md 8(c).

All citations are file:line under /root/reference/.  Layout: population [NP, D] float64
row-major (the reference's list-of-lists, abstract_solver.py:118-120), energies [NP].
"""
import math

import numpy as np

# ----------------------------------------------------------------------------------------
# strategies: name -> (device id, number of donors k, base kind, crossover kind)
# mystic/strategy.py: Best1Exp 34-59, Best1Bin 61-88, Rand1Exp 90-115, RandToBest1Exp 119-145,
# Best2Exp 147-173, Rand2Exp 175-201, Rand1Bin 203-227, RandToBest1Bin 229-255,
# Best2Bin 257-283, Rand2Bin 285-311.  Only Best1Bin is binomial; the other "Bin" names
# run the exponential loop (code, not docstring, is the semantics).
# ----------------------------------------------------------------------------------------
BASE_BEST1, BASE_RAND1, BASE_RANDTOBEST1, BASE_BEST2, BASE_RAND2 = range(5)
XOVER_EXP, XOVER_BIN = 0, 1

STRATEGIES = {
    'Best1Exp':       (0, 2, BASE_BEST1, XOVER_EXP),
    'Best1Bin':       (1, 2, BASE_BEST1, XOVER_BIN),
    'Rand1Exp':       (2, 3, BASE_RAND1, XOVER_EXP),
    'RandToBest1Exp': (3, 2, BASE_RANDTOBEST1, XOVER_EXP),
    'Best2Exp':       (4, 4, BASE_BEST2, XOVER_EXP),
    'Rand2Exp':       (5, 5, BASE_RAND2, XOVER_EXP),
    'Rand1Bin':       (6, 3, BASE_RAND1, XOVER_EXP),
    'RandToBest1Bin': (7, 2, BASE_RANDTOBEST1, XOVER_EXP),
    'Best2Bin':       (8, 4, BASE_BEST2, XOVER_EXP),
    'Rand2Bin':       (9, 5, BASE_RAND2, XOVER_EXP),
}

BOUNDS_NONE, BOUNDS_REJECT, BOUNDS_CLAMP = 0, 1, 2

PTYPES = {'quadratic_equality': 0, 'linear_equality': 1,
          'quadratic_inequality': 2, 'linear_inequality': 3,
          'uniform_equality': 4, 'uniform_inequality': 5, 'barrier_inequality': 6,
          'lagrange_inequality': 7, 'lagrange_equality': 8}


# ----------------------------------------------------------------------------------------
# cost models (a4)
# ----------------------------------------------------------------------------------------
def cost_rosen(x):
    """mystic/models/dejong.py:98-101: numpy elementwise terms, then numpy.sum (pairwise)."""
    x = np.asarray(x, dtype=np.float64)
    x = np.atleast_2d(x)
    if x.shape[1] < 2:
        return np.zeros(x.shape[0])
    return np.sum(100.0 * (x[:, 1:] - x[:, :-1] ** 2.0) ** 2.0 + (1 - x[:, :-1]) ** 2.0, axis=-1)


_pow = np.frompyfunc(math.pow, 2, 1)
_cos = np.frompyfunc(math.cos, 1, 1)
_sin = np.frompyfunc(math.sin, 1, 1)
_exp = np.frompyfunc(math.exp, 1, 1)
_sqrt = np.frompyfunc(math.sqrt, 1, 1)


def cost_corana(x):
    """mystic/models/storn.py:77-96.  len>=4 uses x[0..3] only; shorter inputs are scattered
    through _d=[0,3,1,2] (x[_d.index(i)] = _x[i]); strictly sequential accumulation; math.pow."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, D = x.shape
    d = [1., 1000., 10., 100.]
    _d = [0, 3, 1, 2]
    if D < 4:
        _x = np.zeros((n, 4))
        _x[:, :D] = x
        xx = np.zeros((n, 4))
        for i in range(4):
            xx[:, _d.index(i)] = _x[:, i]
    else:
        xx = x[:, :4]
    r = np.zeros(n)
    for j in range(4):
        xj = xx[:, j]
        zj = np.floor(np.abs(xj / 0.2) + 0.49999) * np.sign(xj) * 0.2
        near = np.abs(xj - zj) < 0.05
        a = 0.15 * _pow(zj - 0.05 * np.sign(zj), 2).astype(np.float64) * d[j]
        b = d[j] * xj * xj
        r = r + np.where(near, a, b)
    return r


def cost_griewangk(x):
    """mystic/models/storn.py:135-139: sum([c*c for c in coeffs])/4000., running product of
    math.cos(c_i/sqrt(i+1.0)), term1 - term2 + 1.

    NOTE on `sum`: the restatement accumulates plainly left to right, which is what the
    reference computes under Python <= 3.11, and under 3.12 whenever the row holds numpy
    scalars (rows that came out of ndarray population rows: strict ranges, Best* mutants).
    CPython 3.12's builtin sum() switches to Neumaier-compensated accumulation for runs of
    exact Python floats (Python/bltinmodule.c), so under the recording interpreter some
    reference energies differ from this restatement in the last bit (<= 2 ulp observed).
    tests/test_oracle_golden.py therefore checks griewangk ENERGIES to 1e-14 relative and
    everything else (trial vectors, selections, populations, best solutions) bit for bit."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, D = x.shape
    s = np.zeros(n)
    for i in range(D):
        s = s + x[:, i] * x[:, i]
    term1 = s / 4000.
    term2 = np.ones(n)
    for i in range(D):
        term2 = term2 * _cos(x[:, i] / math.sqrt(i + 1.0)).astype(np.float64)
    return term1 - term2 + 1


def cost_zimmermann(x):
    """mystic/models/storn.py:182-191 (D == 2)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    x0, x1 = x[:, 0], x[:, 1]
    f8 = 9 - x0 - x1
    c0 = np.where(x0 < 0, -100 * x0, 0.0)
    c1 = np.where(x1 < 0, -100 * x1, 0.0)
    xx = (x0 - 3.) * (x0 - 3) + (x1 - 2.) * (x1 - 2)
    c2 = np.where(xx > 16, 100 * (xx - 16), 0.0)
    c3 = np.where(x0 * x1 > 14, 100 * (x0 * x1 - 14.), 0.0)
    # python max(f8,c0,c1,c2,c3): first maximal element; values only, so np.maximum is equal
    return np.maximum(np.maximum(np.maximum(np.maximum(f8, c0), c1), c2), c3)


CHEBYSHEV_TARGETS = {  # mystic/models/_model_helper.py:10-14
    2: [2., 0., -1.],
    4: [8., 0., -8., 0., 1.],
    6: [32., 0., -48., 0., 18., 0., -1.],
    8: [128., 0., -256., 0., 160., 0., -32., 0., 1.],
    16: [32768., 0., -131072., 0., 212992., 0., -180224., 0., 84480., 0., -21504., 0., 2688., 0., -128., 0., 1],
}


def _polyeval(coeffs, x):
    """mystic/math/poly.py:37-39: val = 0*x; for c in coeffs: val = c + val*x."""
    val = 0 * x
    for j in range(coeffs.shape[1]):
        val = coeffs[:, j] + val * x
    return val


def chebyshev_points(M):
    """x_i as the reference accumulates them: x=-1.0; x += 2.0/(M-1) (_model_helper.py:19-25)."""
    xs = np.empty(M)
    x = -1.0
    dx = 2.0 / (M - 1)
    for i in range(M):
        xs[i] = x
        x += dx
    return xs


def cost_chebyshev(x, order=8, M=61):
    """mystic/models/_model_helper.py:16-33 with target = chebyshev{order}coeffs."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n = x.shape[0]
    target = np.asarray(CHEBYSHEV_TARGETS[order], dtype=np.float64)[None, :]
    result = np.zeros(n)
    for xi in chebyshev_points(int(M)):
        px = _polyeval(x, xi)
        out = (px < -1) | (px > 1)
        result = result + np.where(out, (1 - px) * (1 - px), 0.0)
    for xe in (1.2, -1.2):
        px = _polyeval(x, xe) - _polyeval(target, xe)
        result = result + np.where(px < 0, px * px, 0.0)
    return result


def _f(a):
    return np.asarray(a, dtype=np.float64)


def cost_sphere(x):
    """mystic/models/dejong.py:63-66: f = 0.; for c in coeffs: f += c*c"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    f = np.zeros(x.shape[0])
    for i in range(x.shape[1]):
        f = f + x[:, i] * x[:, i]
    return f


def cost_rastrigin(x):
    """mystic/models/pohlheim.py:132: 10.*len(coeffs) + sum(c*c - 10.*cos(2*pi*c) for c in coeffs)
    (plain left-to-right sum, see cost_griewangk's note on CPython 3.12's compensated sum())"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    s = np.zeros(x.shape[0])
    for i in range(x.shape[1]):
        c = x[:, i]
        s = s + (c * c - 10. * _f(_cos(2 * math.pi * c)))
    return 10. * x.shape[1] + s


def cost_schwefel(x):
    """mystic/models/pohlheim.py:71: sum(-c * sin(sqrt(abs(c))) for c in coeffs)"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    s = np.zeros(x.shape[0])
    for i in range(x.shape[1]):
        c = x[:, i]
        s = s + (-c) * _f(_sin(_f(_sqrt(np.abs(c)))))
    return s


def cost_ackley(x):
    """mystic/models/pohlheim.py:199-204"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n = x.shape[1]
    a, b, d = 20., 0.2, 2 * math.pi
    s1 = np.zeros(x.shape[0])
    s2 = np.zeros(x.shape[0])
    for i in range(n):
        c = x[:, i]
        s1 = s1 + c * c
        s2 = s2 + _f(_cos(d * c))
    f0 = -a * _f(_exp(-b * _f(_sqrt(s1 / n))))
    f1 = -_f(_exp(s2 / n)) + a + math.exp(1)
    return f0 + f1


def cost_step(x):
    """mystic/models/dejong.py:184-192"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    f = np.full(x.shape[0], 30.)
    for i in range(x.shape[1]):
        c = x[:, i]
        with np.errstate(invalid='ignore'):
            t = np.where(np.abs(c) <= 5.12, np.floor(c), np.where(c > 5.12, 30 * (c - 5.12), 30 * (5.12 - c)))
        f = f + t
    return f


def cost_michal(x):
    """mystic/models/pohlheim.py:237-240 (plain left-to-right sum)"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    s = np.zeros(x.shape[0])
    _pow = np.frompyfunc(lambda a, b: a ** b, 2, 1)
    for i in range(x.shape[1]):
        c = x[:, i]
        w = _f(_sin(_f((i + 1) * _f(_pow(c, 2)) / math.pi)))
        s = s + _f(_sin(c)) * _f(_pow(w, 20))
    return -s


def cost_powers(x):
    """mystic/models/pohlheim.py:162-163"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    s = np.zeros(x.shape[0])
    _pow = np.frompyfunc(lambda a, b: a ** b, 2, 1)
    for i in range(x.shape[1]):
        s = s + _f(_pow(np.abs(x[:, i]).astype(object), i + 2))
    return s



def cost_decay(x, pts, obs, sigma=1.0):
    """the same residual cost with the dual exponential decay of mystic/models/br8.py:27-32,
    coeffs = (a1,a2,a3,a4,a5); sigma is sqrt(observations) in the reference's own factories (br8.py:47,55)"""
    from numpy import exp
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    t = np.asarray(pts, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64)
    out = np.empty(x.shape[0])
    for r in range(x.shape[0]):
        a1, a2, a3, a4, a5 = [float(v) for v in x[r]]
        with np.errstate(all='ignore'):
            try:
                y = a1 + a2*exp(-t/a4) + a3*exp(-t/a5)
            except ZeroDivisionError:
                y = np.full_like(t, np.nan)
            d = (y - obs) / sigma
            out[r] = np.sum(d * d)
    return out


# ---- closed-form functions of a few variables: the reference's own expressions on Python floats, row by row
def _rowwise(f):
    def cost(x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return np.array([f([float(v) for v in row]) for row in x], dtype=np.float64)
    return cost


def _branins(coeffs):
    """mystic/models/pohlheim.py:278-282"""
    pi, cos = math.pi, math.cos
    x1, x2 = coeffs
    a = 1.; b = 5.1/(4*pi**2); c = 5./pi; d = 6.; e = 10.; f = 1./(8*pi)
    f0 = a * (x2 - b * x1**2 + c * x1 - d)**2
    f1 = e * (1 - f) * cos(x1) + e
    return f0 + f1


def _easom(coeffs):
    """mystic/models/pohlheim.py:312-313"""
    pi, cos, exp = math.pi, math.cos, math.exp
    x0, x1 = coeffs
    return -cos(x0) * cos(x1) * exp(-((x0-pi)**2+(x1-pi)**2))


def _goldstein(coeffs):
    """mystic/models/pohlheim.py:352-355"""
    x0, x1 = coeffs
    f0 = 19 - 14*x0 + 3*x0**2 - 14*x1 + 6*x0*x1 + 3*x1**2
    f1 = 18 - 32*x0 + 12*x0**2 + 48*x1 - 36*x0*x1 + 27*x1**2
    return (1 + (x0 + x1 + 1)**2 * f0) * (30 + (2*x0 - 3*x1)**2 * f1)


def _peaks(coeffs):
    """mystic/models/nag.py:55-59"""
    exp = math.exp
    x, y = coeffs
    result = 3.*(1. - x)**2*exp(-x**2 - (y + 1.)**2) - \
            10.*(x*(1./5.) - x**3 - y**5)*exp(-x**2 - y**2) - \
            1./3.*exp(-(x + 1.)**2 - y**2)
    return result


def _venkat91(coeffs):
    """mystic/models/venkataraman.py:50-52"""
    x, y = coeffs
    R = ((x-4.)**2 + (y-4.)**2 + 0.1)**.5
    return -20. * math.sin(R)/R


def _fosc3d(coeffs):
    """mystic/models/wolfram.py:60-64"""
    exp, sin = math.exp, math.sin
    x, y = coeffs
    func = -4. * exp(-x*x - y*y) + sin(6. * x) * sin(5. * y)
    penalty = 0
    if y < 0: penalty = 100.*y*y
    return func + penalty


def _nmin51(coeffs):
    """mystic/models/wolfram.py:101-103"""
    exp, sin = math.exp, math.sin
    x, y = coeffs
    return exp(sin(50.*x)) + sin(60.*exp(y)) + sin(70.*sin(x)) + \
           sin(sin(80.*y)) - sin(10.*(x + y)) + 1./4.*(x**2 + y**2)


def _shekel(coeffs):
    """mystic/models/dejong.py:266-278"""
    A = [-32., -16., 0., 16., 32.]
    a1 = A * 5
    a2 = [c for c in A for _ in range(5)]
    x, y = coeffs
    r = 0.0
    for i in range(25):
        z = 1.0*i + pow(x-a1[i], 6) + pow(y-a2[i], 6)
        if z: r += 1.0/z
        else: r += math.inf
    return 1.0/(0.002 + r)


def _ellipsoid(coeffs):
    """mystic/models/pohlheim.py:101-102 (the builtin sum(): compensated since CPython 3.12)"""
    ncoeffs = range(len(coeffs))
    return sum(sum(coeffs[0:i+1])**2 for i in ncoeffs)


def _paviani(coeffs):
    """mystic/models/schittkowski.py:64-71"""
    from functools import reduce
    s = 0.
    for x in coeffs:
        if x >= 10. or x <= 2.:
            s = math.inf
            break
        s += math.log(x-2.)**2 + math.log(10.-x)**2
    p = reduce(lambda x, y: x*y, coeffs)**(.2)
    return s - p


FIXED_COSTS = {'branins': _branins, 'easom': _easom, 'goldstein': _goldstein, 'peaks': _peaks, 'venkat91': _venkat91,
               'fosc3d': _fosc3d, 'nmin51': _nmin51, 'shekel': _shekel, 'ellipsoid': _ellipsoid, 'paviani': _paviani}
# magnitude of the terms the functions combine: the scale energies are compared on (their transcendental
# terms differ by <= 2 ulp between libm and CUDA, and they cancel)
FIXED_COST_SCALE = {'branins': 10., 'easom': 1e-3, 'goldstein': 1., 'peaks': 10., 'venkat91': 20., 'fosc3d': 4., 'nmin51': 6.,
                    'shekel': 1e-3, 'ellipsoid': 0., 'paviani': 10.}


def cost_polyfit(x, pts, obs, sigma=1.0):
    """mystic/forward_model.py:241-258 with the polynomial model of mystic/models/poly.py:36-54:
    F(params) = numpy.poly1d(params) whose evaluation is numpy.polyval's
    `y = zeros_like(x); for pv in p: y = y*x + pv`; x = (y - observations)/sigma; metric
    numpy.sum(x*x) (abstract_model.py: sigma = 1.0, metric L2)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    pts = np.asarray(pts, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64)
    out = np.empty(x.shape[0])
    for r in range(x.shape[0]):
        y = np.zeros_like(pts)
        for pv in x[r]:
            y = y * pts + pv
        d = (y - obs) / sigma
        out[r] = np.sum(d * d)
    return out


def cost_lorentzian(x, pts, obs, sigma=1.0):
    """the same residual cost with the lorentzian peak of mystic/models/lorentzian.py:32-37,
    coeffs = (a1,a2,a3,A0,E0,G0,n), evaluated in the reference's own expression order."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    pts = np.asarray(pts, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64)
    pi = np.pi
    out = np.empty(x.shape[0])
    for r in range(x.shape[0]):
        a1, a2, a3, A0, E0, G0, n = [float(v) for v in x[r]]
        X = pts
        with np.errstate(all='ignore'):
            try:
                y = (a1 + a2*X + a3*X*X + A0 * ( G0/(2*pi) )/( (X-E0)*(X-E0)+(G0/2)*(G0/2) ))/n
            except ZeroDivisionError:   # n == 0.0 as a python float; cannot happen inside a strict range n > 0
                y = np.full_like(X, np.nan)
            d = (y - obs) / sigma
            out[r] = np.sum(d * d)
    return out


def make_cost(name, extra_args=()):
    """name -> f(X[NP,D]) -> [NP]; names as exported by mystic.models / mystic.models.poly.
    Data-fit costs: name 'polyfit' / 'lorentzian' with extra_args = (pts, observations[, sigma])."""
    if name in ('polyfit', 'lorentzian', 'decay'):
        pts, obs = extra_args[0], extra_args[1]
        sigma = extra_args[2] if len(extra_args) > 2 else 1.0
        sigma = np.asarray(sigma, dtype=np.float64) if np.ndim(sigma) > 0 else float(sigma)
        f = {'polyfit': cost_polyfit, 'lorentzian': cost_lorentzian, 'decay': cost_decay}[name]
        return lambda X: f(X, pts, obs, sigma)
    if name == 'rosen':
        return cost_rosen
    if name == 'corana':
        return cost_corana
    if name == 'griewangk':
        return cost_griewangk
    if name == 'zimmermann':
        return cost_zimmermann
    if name in ('sphere', 'rastrigin', 'schwefel', 'ackley', 'step', 'michal', 'powers'):
        return {'sphere': cost_sphere, 'rastrigin': cost_rastrigin, 'schwefel': cost_schwefel, 'ackley': cost_ackley,
                'step': cost_step, 'michal': cost_michal, 'powers': cost_powers}[name]
    if name in FIXED_COSTS:
        return _rowwise(FIXED_COSTS[name])
    if name.startswith('chebyshev') and name.endswith('cost'):
        order = int(name[len('chebyshev'):-len('cost')])
        M = int(extra_args[0]) if extra_args else 61
        return lambda X: cost_chebyshev(X, order, M)
    raise KeyError(name)


# ----------------------------------------------------------------------------------------
# penalty (a5): mystic/penalty.py:41-601 stacked decorators; tools.py:380-389
# ----------------------------------------------------------------------------------------
def lagrange_state(ptype, k, h, n, stored):
    """(_k, multiplier) after the decorator's loop over the first n stored condition values:
    penalty.py:511-515 (beta of lagrange_inequality), :579-582 (lam of lagrange_equality)"""
    mult, _k = 0., k
    for i in range(n):
        y = stored[i] if i < len(stored) else 0.0     # stored(i) answers 0.0 past the end (:503-506)
        if ptype == 'lagrange_inequality':
            mult += 2. * _k * max(-mult / (2. * _k), y)
        else:
            mult += 2. * _k * y
        _k *= h
    return _k, mult


def penalty_value(x, terms):
    """terms[0] is the innermost decorator.  Outermost is evaluated first and the inner
    result is added to it: t_last + (... + (t_0 + 0.0)).  condition = sequential dot - h
    (oracle/host_callables.py)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, D = x.shape
    total = np.zeros(n)  # the undecorated penalty returns 0.0
    for t in terms:
        G = np.asarray(t['G'], dtype=np.float64)
        acc = np.zeros(n)
        for j in range(D):
            acc = acc + G[j] * x[:, j]
        pf = acc - float(t['h'])
        ptype = t['ptype']
        if 'kscale' in t:                      # as handed to the device: multiplier and accumulated lagrange multiplier
            _k, mult = t['kscale'], t.get('mult', 0.0)
        elif ptype in ('lagrange_inequality', 'lagrange_equality'):
            _k, mult = lagrange_state(ptype, t.get('k', 20), t.get('hpow', 5), t.get('n', 0), t.get('stored', ()))
        else:
            _k, mult = t.get('k', 100) * pow(t.get('hpow', 5), t.get('n', 0)), 0.0
        if ptype == 'quadratic_equality':      # penalty.py:88
            v = float(_k) * _pow(pf, 2).astype(np.float64)
        elif ptype == 'linear_equality':       # penalty.py:147
            v = float(_k) * np.abs(pf)
        elif ptype == 'quadratic_inequality':  # penalty.py:388
            v = float(2 * _k) * _pow(np.maximum(0., pf), 2).astype(np.float64)
        elif ptype == 'linear_inequality':     # penalty.py:447
            v = float(2 * _k) * np.abs(np.maximum(0., pf))
        elif ptype == 'uniform_equality':      # penalty.py:208: _k if pf else 0.0
            v = np.where(pf != 0.0, float(_k), 0.0)
        elif ptype == 'uniform_inequality':    # penalty.py:266: _k if pf > 0 else 0.0
            v = np.where(pf > 0.0, float(_k), 0.0)
        elif ptype == 'barrier_inequality':    # penalty.py:322-326: inf if pf > 0 else -.5/_k*log(-pf) (numpy.log)
            with np.errstate(divide='ignore', invalid='ignore'):
                v = np.where(pf > 0.0, np.inf, -.5 / _k * np.log(-np.where(pf > 0.0, -1.0, pf)))
        elif ptype == 'lagrange_inequality':   # penalty.py:516-517: mpf = max(-beta/(2.*_k), pf); Python max keeps the first unless the second is larger
            lo_ = -mult / (2. * _k)
            mpf = np.where(pf > lo_, pf, lo_)
            v = float(_k) * _pow(mpf, 2).astype(np.float64) + mult * mpf
        elif ptype == 'lagrange_equality':     # penalty.py:583
            v = float(_k) * _pow(pf, 2).astype(np.float64) + mult * pf
        else:
            raise KeyError(ptype)
        total = v + total
    return total


# ----------------------------------------------------------------------------------------
# Philox4x32-10 counter RNG (Salmon et al., SC'11; Random123 v1.14 philox.h) and the draw
# protocol shared with the device kernels (mystic_b200/csrc/philox.cuh).
# ----------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)

STREAM_DONORS, STREAM_XOVER, STREAM_INIT = 0, 1, 2


def philox4x32(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10; all args uint32 arrays (broadcastable). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def _mulhi64(a, b):
    """floor(a*b / 2^64) for uint64 arrays a and python int / uint64 b (b < 2^32 here)."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.uint64(b)
    a_lo = a & _MASK32
    a_hi = a >> np.uint64(32)
    # b < 2^32 so b_hi = 0:  a*b = a_hi*b*2^32 + a_lo*b
    lo = a_lo * b
    mid = a_hi * b + (lo >> np.uint64(32))
    return mid >> np.uint64(32)


def philox_donors(seed, generation, cand, NP, D, k, offset=0, with_xword=False):
    """Donor indices [n,k] (distinct, != cand, uniform over ordered k-tuples) and start index.

    stream STREAM_DONORS, counter (block, cand, generation, stream); block b supplies the
    64-bit words r_{2b} = x0<<32|x1 and r_{2b+1} = x2<<32|x3.  Donor i uses r_i:
    t = mulhi64(r_i, NP-1-i) is a rank among the not-yet-excluded indices, converted to an index by
    walking the sorted exclusion list {cand, earlier donors}.  nstart = mulhi64(r_k, D).
    Replaces random.sample(...)/random.randrange(D) (strategy.py:27,42) in Philox mode.

    `cand` are GLOBAL candidate indices (they key the counter); donors are drawn from the shard
    [offset, offset+NP) and returned as LOCAL row indices, excluding the candidate's own local row."""
    cand = np.asarray(cand, dtype=np.int64)
    n = cand.shape[0]
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    words = []
    for b in range((k + 1 + 1) // 2):
        x0, x1, x2, x3 = philox4x32(np.uint32(b), cand.astype(np.uint32), np.uint32(generation),
                                    np.uint32(STREAM_DONORS), k0, k1)
        words.append((x0.astype(np.uint64) << np.uint64(32)) | x1.astype(np.uint64))
        words.append((x2.astype(np.uint64) << np.uint64(32)) | x3.astype(np.uint64))
    donors = np.zeros((n, k), dtype=np.int64)
    excl = np.full((n, k + 1), np.iinfo(np.int64).max, dtype=np.int64)  # sorted ascending, padded
    excl[:, 0] = cand - int(offset)
    for i in range(k):
        t = _mulhi64(words[i], NP - 1 - i).astype(np.int64)
        idx = t.copy()
        for e in range(i + 1):
            idx = idx + (idx >= excl[:, e]).astype(np.int64)
        donors[:, i] = idx
        excl[:, i + 1] = idx
        excl.sort(axis=1)
    nstart = _mulhi64(words[k], D).astype(np.int64)
    if with_xword:   # low half of the start-index word: the geometric draw of the exponential crossover length
        return donors, nstart, (words[k] & _MASK32).astype(np.uint32)
    return donors, nstart


RNG_PROTOCOL = 2   # mystic_b200/csrc/philox.cuh kRngProtocol


def crossover_threshold(CR):
    """floor(CR * 2^32) in [0, 2^32]: the success probability of the exponential crossover's length table"""
    cr = min(max(float(CR), 0.0), 1.0)
    return int(math.floor(cr * 4294967296.0))


def crossover_threshold16(CR):
    """Bernoulli(CR) as `field < thr16` on a 16-bit field of a Philox word (binomial crossover), thr16 in [0, 2^16]"""
    cr = min(max(float(CR), 0.0), 1.0)
    return int(math.floor(cr * 65536.0))


def exp_length_table(CR, D):
    """T[1..nmax] of the geometric draw (philox.cuh, protocol 2): T[0] = 2^32, T[l] = (T[l-1] * thr) >> 32 in integer
    arithmetic, thr = floor(CR * 2^32); nmax = min(D, last l with T[l] > 0).  None when CR >= 1 (length D)."""
    thr = crossover_threshold(CR)
    if thr >= 1 << 32:
        return None
    tab, t = [], 1 << 32
    while len(tab) < D:
        t = (t * thr) >> 32
        if t == 0:
            break
        tab.append(t)
    return np.asarray(tab, dtype=np.uint64)


def exp_length(xword, CR, D):
    """exponential crossover length (strategy.py:48-56: number of leading draws below CR, at most D) as ONE geometric
    draw: L = #{l >= 1 : w < T[l]}, w = low 32 bits of the start-index word"""
    tab = exp_length_table(CR, D)
    w = np.asarray(xword, dtype=np.uint64)
    if tab is None:
        return np.full(w.shape, D, dtype=np.int64)
    if tab.size == 0:
        return np.zeros(w.shape, dtype=np.int64)
    # T decreases: count the entries above w
    return (tab.size - np.searchsorted(tab[::-1], w, side='right')).astype(np.int64)


def philox_xover_fields(seed, generation, cand, D):
    """[n, 8*ceil(D/8)] crossover fields of stream 1 (binomial crossover): dimension 8q+e reads bits
    16*(e&1)..16*(e&1)+15 of word e>>1 of block q"""
    cand = np.asarray(cand, dtype=np.int64)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    nblk = (D + 7) // 8
    b = np.arange(nblk, dtype=np.uint32)[None, :]
    x = philox4x32(b, cand.astype(np.uint32)[:, None], np.uint32(generation), np.uint32(STREAM_XOVER), k0, k1)
    words = np.stack(x, axis=-1)                                   # [n, nblk, 4]
    fields = np.stack([words & np.uint32(0xFFFF), words >> np.uint32(16)], axis=-1)   # [n, nblk, 4, 2]
    return fields.reshape(cand.shape[0], nblk * 8)


def philox_init_population(seed, NP, lo, hi, offset=0):
    """Device-side SetRandomInitialPoints (abstract_solver.py:532-534, distributionally):
    pop[c][j] = lo[j] + (hi[j]-lo[j]) * u,  u = (r64 >> 11) * 2^-53, r64 from stream 2,
    block j//2 (word pair j%2), counter (block, c, 0, STREAM_INIT)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    D = lo.shape[0]
    cand = (np.arange(NP, dtype=np.int64) + offset)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    nblk = (D + 1) // 2
    b = np.arange(nblk, dtype=np.uint32)[None, :]
    x0, x1, x2, x3 = philox4x32(b, cand.astype(np.uint32)[:, None], np.uint32(0), np.uint32(STREAM_INIT), k0, k1)
    r_even = (x0.astype(np.uint64) << np.uint64(32)) | x1.astype(np.uint64)
    r_odd = (x2.astype(np.uint64) << np.uint64(32)) | x3.astype(np.uint64)
    r = np.stack([r_even, r_odd], axis=-1).reshape(NP, nblk * 2)[:, :D]
    u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo[None, :] + (hi - lo)[None, :] * u


# ----------------------------------------------------------------------------------------
# the generation step (a1, a2, a3, a6, a7)
# ----------------------------------------------------------------------------------------
def crossover_mask_replay(xover, D, nstart, uni, CR):
    """[n, D] bool mask from the reference's own uniforms (NaN = not drawn).
    Bin (strategy.py:77-85): i==n or u_i < CR.  Exp (strategy.py:48-56): L = number of leading
    draws with u < CR, capped at D; dims n, n+1, ... (mod D)."""
    n = nstart.shape[0]
    idx = np.arange(D)[None, :]
    if xover == XOVER_BIN:
        with np.errstate(invalid='ignore'):
            return (idx == nstart[:, None]) | (uni[:, :D] < CR)
    with np.errstate(invalid='ignore'):
        ok = uni[:, :D] < CR                      # NaN -> False
    L = np.where(ok.all(axis=1), D, np.argmin(ok, axis=1))
    rel = (idx - nstart[:, None]) % D
    return rel < L[:, None]


def crossover_mask_philox(xover, D, nstart, CR, fields=None, xword=None):
    """Bin: i == n or field_i < thr16 (fields from philox_xover_fields).  Exp: the window of exp_length(xword) dims"""
    idx = np.arange(D)[None, :]
    if xover == XOVER_BIN:
        ok = fields[:, :D].astype(np.uint32) < np.uint32(crossover_threshold16(CR))
        return (idx == nstart[:, None]) | ok
    L = exp_length(xword, CR, D)
    rel = (idx - nstart[:, None]) % D
    return rel < L[:, None]


def mutate(base, pop, best, donors, F):
    """Mutant vector for every dimension, with the reference's evaluation order (no FMA)."""
    p = [pop[donors[:, i]] for i in range(donors.shape[1])]
    b = best[None, :]
    if base == BASE_BEST1:       # strategy.py:52-54, 83-85
        return b + F * (p[0] - p[1])
    if base == BASE_RAND1:       # strategy.py:108-110, 220-222
        return p[0] + F * (p[1] - p[2])
    if base == BASE_RANDTOBEST1:  # strategy.py:137-140, 247-250: t += F*(best-t) + F*(p1-p2)
        return pop + (F * (b - pop) + F * (p[0] - p[1]))
    if base == BASE_BEST2:       # strategy.py:164-168, 274-278
        return b + F * (p[0] + p[1] - p[2] - p[3])
    if base == BASE_RAND2:       # strategy.py:192-196, 302-306
        return p[0] + F * (p[1] + p[2] - p[3] - p[4])
    raise KeyError(base)


class DEState(object):
    """Solver state the step reads/writes (names follow abstract_solver.py:112-122)."""

    def __init__(self, population, lo=None, hi=None, bounds_mode=BOUNDS_NONE):
        self.population = np.array(population, dtype=np.float64)
        NP, D = self.population.shape
        self.popEnergy = np.full(NP, np.inf)
        self.bestEnergy = np.inf
        self.bestSolution = self.population[0].copy()
        self.bestIndex = -1          # index of the member that last became best (-1: none yet)
        self.evaluations = 0
        self.generation = -1         # last generation evaluated (0 = initial population)
        self.bounds_mode = bounds_mode
        self.lo = None if lo is None else np.asarray(lo, dtype=np.float64)
        self.hi = None if hi is None else np.asarray(hi, dtype=np.float64)
        self.trialSolution = None
        self.trialEnergy = None
        self.accepted = None


def evaluate(trial, cost, state, penalty_terms=()):
    """wrap_penalty(wrap_bounds(wrap_function(raw))) -- differential_evolution.py:515-521,
    tools.py:386-389, 420-426: OOB -> inf without calling the cost, then + penalty."""
    if state.bounds_mode != BOUNDS_NONE:
        oob = ((trial < state.lo[None, :]) | (trial > state.hi[None, :])).any(axis=1)
    else:
        oob = np.zeros(trial.shape[0], dtype=bool)
    raw = np.full(trial.shape[0], np.inf)
    if (~oob).any():
        raw[~oob] = cost(trial[~oob])
    if penalty_terms:
        with np.errstate(invalid='ignore'):
            return raw + penalty_value(trial, penalty_terms)
    return raw + 0.0


def select(state, trial, trialE):
    """differential_evolution.py:596-613."""
    state.evaluations += int(trialE.shape[0] - np.isinf(trialE).sum())
    accept = trialE < state.popEnergy
    state.popEnergy = np.where(accept, trialE, state.popEnergy)
    state.population = np.where(accept[:, None], trial, state.population)
    cand = np.where(accept & (trialE < state.bestEnergy), trialE, np.inf)
    i = int(np.argmin(cand))              # first occurrence = lowest index (strict '<' scan)
    if cand[i] < state.bestEnergy:
        state.bestEnergy = float(cand[i])
        state.bestSolution = trial[i].copy()
        state.bestIndex = i
    state.trialSolution = trial
    state.trialEnergy = trialE
    state.accepted = accept


def apply_bounds(trial, state):
    """Mode B (tight=True): x[i]=max(lo,x[i]); x[i]=min(hi,x[i]) (symbolic.py:1139-1144)."""
    if state.bounds_mode == BOUNDS_CLAMP:
        return np.minimum(np.maximum(trial, state.lo[None, :]), state.hi[None, :])
    return trial


def step0(state, cost, penalty_terms=()):
    """Generation 0: differential_evolution.py:516-520 (clip population into the strict range,
    at=True), :559-567, :575-582 with strategy=None, :589, :603-613."""
    if state.bounds_mode != BOUNDS_NONE:
        state.population = state.population.clip(state.lo, state.hi)
    state.bestSolution = state.population[0].copy()
    state.bestEnergy = np.inf
    trial = apply_bounds(state.population.copy(), state)
    trialE = evaluate(trial, cost, state, penalty_terms)
    select(state, trial, trialE)
    state.generation = 0


def step(state, strategy, CR, F, cost, penalty_terms=(), replay=None, philox=None):
    """One generation g >= 1.  Exactly one of:
       replay = dict(donors[NP,>=k], nstart[NP], uni[NP,D+1])  -- the reference's recorded draws
       philox = dict(seed=int, offset=int)                      -- counter RNG (generation = g)"""
    _, k, base, xover = STRATEGIES[strategy]
    pop = state.population
    NP, D = pop.shape
    g = state.generation + 1
    if replay is not None:
        donors = np.asarray(replay['donors'])[:, :k].astype(np.int64)
        nstart = np.asarray(replay['nstart']).astype(np.int64)
        mask = crossover_mask_replay(xover, D, nstart, np.asarray(replay['uni']), CR)
    else:
        cand = np.arange(NP, dtype=np.int64) + int(philox.get('offset', 0))
        donors, nstart, xword = philox_donors(philox['seed'], g, cand, NP, D, k, offset=int(philox.get('offset', 0)),
                                              with_xword=True)
        fields = philox_xover_fields(philox['seed'], g, cand, D) if xover == XOVER_BIN else None
        mask = crossover_mask_philox(xover, D, nstart, CR, fields=fields, xword=xword)
    mutant = mutate(base, pop, state.bestSolution, donors, F)
    trial = np.where(mask, mutant, pop)
    trial = apply_bounds(trial, state)
    trialE = evaluate(trial, cost, state, penalty_terms)
    select(state, trial, trialE)
    state.generation = g
    return donors, nstart, mask
