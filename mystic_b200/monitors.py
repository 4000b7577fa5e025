"""Monitors for solvers without a mystic installation (interface of mystic/monitors.py:75-371).

Unmodified `mystic.monitors` objects are accepted by the B200 solver: it only calls
`monitor(x, y, id)`, `len(monitor)`, `.x/.y/._y/._x/.info()/.prepend()`.  These stand-ins keep
the same attributes so `mystic.termination` conditions and user code read them the same way.
"""
import sys


class Null(object):
    """does nothing, reliably (monitors.py:75-102)"""
    _inst = None
    x = _x = ()
    y = _y = ()
    _id = ()
    _info = ()
    k = None
    _npts = None
    label = None

    def __new__(cls, *args, **kwds):
        if cls._inst is None:
            cls._inst = object.__new__(cls)
        return cls._inst

    def __call__(self, *args, **kwds):
        return self

    def __len__(self):
        return 0

    def __bool__(self):
        return False

    def __repr__(self):
        return 'Null()'

    def info(self, message):
        return

    def prepend(self, monitor):
        return

    def extend(self, monitor):
        return


def _aslist(x):
    if hasattr(x, 'tolist'):
        return x.tolist()
    if isinstance(x, (list, tuple)):
        return [_aslist(i) for i in x]
    return x


class Monitor(object):
    """logs (x, y, id) triples (monitors.py:106-371)"""

    def __init__(self, **kwds):
        self._x = []
        self._y = []
        self._id = []
        self._info = []
        self.k = kwds.get('k', None)
        self._npts = kwds.get('npts', None)
        self.label = kwds.get('label', 'ChiSquare')

    def __len__(self):
        return len(self._x)

    def info(self, message):
        self._info.append('%s' % message)

    def __call__(self, x, y, id=None, **kwds):
        self._x.append(_aslist(x))
        y = _aslist(y)
        if self.k is not None:
            y = [self.k * v for v in y] if isinstance(y, list) else self.k * y
        self._y.append(y)
        self._id.append(id)

    def __getitem__(self, i):
        return self.x[i], self.y[i]

    def _compatible(self, monitor):
        if isinstance(monitor, Null) or monitor is Null:
            return Monitor()
        if not all(hasattr(monitor, a) for a in ('_x', '_y', '_id')):
            raise TypeError("'%s' is not a monitor instance" % monitor)
        return monitor

    def extend(self, monitor):
        monitor = self._compatible(monitor)
        self._x.extend(monitor._x)
        self._y.extend(monitor._y)
        self._id.extend(monitor._id)
        self._info.extend(getattr(monitor, '_info', []))

    def prepend(self, monitor):
        monitor = self._compatible(monitor)
        self._x[:0] = list(monitor._x)
        self._y[:0] = list(monitor._y)
        self._id[:0] = list(monitor._id)
        self._info[:0] = list(getattr(monitor, '_info', []))

    @property
    def x(self):
        return self._x

    @property
    def y(self):
        if self.k in (None, 1):
            return self._y
        return [[v / self.k for v in y] if isinstance(y, list) else y / self.k for y in self._y]

    @property
    def id(self):
        return self._id

    def min(self):
        ys = self.y
        i = min(range(len(ys)), key=lambda j: ys[j])
        return self[i]


class VerboseMonitor(Monitor):
    """prints y every `interval` calls, x every `xinterval` (monitors.py:373-428)"""

    def __init__(self, interval=10, xinterval=float('inf'), all=True, **kwds):
        super(VerboseMonitor, self).__init__(**kwds)
        self._yinterval = interval
        self._xinterval = xinterval
        self._all = all

    def __call__(self, x, y, id=None, **kwds):
        super(VerboseMonitor, self).__call__(x, y, id, **kwds)
        step = len(self) - 1
        who = '' if id is None else ' id(%s)' % id
        if self._yinterval is not float('inf') and int(step % self._yinterval) == 0:
            print('Generation %d has %s: %s%s' % (step, self.label, self._y[-1], who))
            sys.stdout.flush()
        if self._xinterval != float('inf') and int(step % self._xinterval) == 0:
            print('Generation %d has fit parameters:\n %s%s' % (step, self._x[-1], who))
            sys.stdout.flush()
