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


def _is_seq(v):
    return isinstance(v, (list, tuple)) or type(v).__name__ == 'ndarray'


def _unscaled(self, k_applied):
    """the last y as it is reported: the stored value has the monitor's k applied (monitors.py:355-357)"""
    y = self._y[-1]
    if k_applied or self.k is None:
        return y
    return [v / self.k for v in y] if isinstance(y, list) else y / self.k


def _interval(v):
    return float('inf') if (not v or v != v) else v


def _screen_report(self, x, y, id, best, k, yint, xint):
    """the lines VerboseMonitor / VerboseLoggingMonitor print (monitors.py:391-420, 534-563)"""
    step = len(self) - 1
    tag = '' if id is None else '[id: %s] ' % (id,)
    if yint != float('inf') and int(step % yint) == 0:
        v = _unscaled(self, k)
        who = ''
        if _is_seq(y) and not self._all:
            who, v = ' best', v[best]
        print('%sGeneration %s has%s %s: %s' % (tag, step, who, self.label, v))
    if xint != float('inf') and int(step % xint) == 0:
        if not _is_seq(x):
            who, text = '', ' %s' % (self._x[-1],)
        elif self._all:
            who, text = '', '\n %s' % (self._x[-1],)
        else:
            who, text = ' best', '\n %s' % (self._x[-1][best],)
        print('%sGeneration %s has%s fit parameters:%s' % (tag, step, who, text))
    sys.stdout.flush()


class VerboseMonitor(Monitor):
    """prints y every `interval` calls, x every `xinterval` (monitors.py:373-428)"""

    def __init__(self, interval=10, xinterval=float('inf'), all=True, **kwds):
        super(VerboseMonitor, self).__init__(**kwds)
        self._yinterval = _interval(interval)
        self._xinterval = _interval(xinterval)
        self._all = all

    def info(self, message):
        super(VerboseMonitor, self).info(message)
        print('%s' % message)

    def __call__(self, x, y, id=None, best=0, k=False, **kwds):
        super(VerboseMonitor, self).__call__(x, y, id)
        _screen_report(self, x, y, id, best, k, self._yinterval, self._xinterval)


class LoggingMonitor(Monitor):
    """writes `  (step[, id])     y   x` lines to a file every `interval` calls, with the reference's
    header and layout (monitors.py:431-514), so log readers written for mystic keep working"""

    def __init__(self, interval=1, filename='log.txt', new=False, all=True, info=None, **kwds):
        import datetime
        import os
        super(LoggingMonitor, self).__init__(**kwds)
        self._filename = filename
        self._yinterval = self._xinterval = _interval(interval)
        head = new or not os.path.exists(filename) or 'label' in kwds
        with open(filename, 'w' if new else 'a') as f:
            if info or head:
                f.write('# %s\n' % datetime.datetime.now().ctime())
                if info:
                    f.write('# %s\n' % (info,))
                if head:
                    f.write('# ___#___  __%s__  __params__\n' % self.label)
        self._all = all

    def info(self, message):
        super(LoggingMonitor, self).info(message)
        with open(self._filename, 'a') as f:
            f.write('# %s\n' % (message,))

    def __call__(self, x, y, id=None, best=0, k=False, **kwds):
        super(LoggingMonitor, self).__call__(x, y, id)
        if self._yinterval == float('inf') or int((len(self) - 1) % self._yinterval) != 0:
            return
        v = _unscaled(self, k)
        if _is_seq(y) and not self._all:
            v = v[best]
        xs = self._x[-1]
        if _is_seq(x) and not self._all:
            xs = xs[best]
        text = '%s' % (xs,) if _is_seq(xs) else '[%s]' % (xs,)
        step = (len(self) - 1,) if id is None else (len(self) - 1, id)
        with open(self._filename, 'a') as f:
            f.write('  %s     %s   %s\n' % (step, v, text))

    def __reduce__(self):
        state = dict(_x=self._x, _y=self._y, _id=self._id, _info=self._info, k=self.k, label=self.label)
        return (self.__class__, (self._yinterval, self._filename, False, self._all, None), state)

    def __setstate__(self, state):
        self.__dict__.update(state)


class VerboseLoggingMonitor(LoggingMonitor):
    """LoggingMonitor that also prints like VerboseMonitor (monitors.py:516-585)"""

    def __init__(self, interval=1, yinterval=10, xinterval=float('inf'), filename='log.txt', new=False, all=True,
                 info=None, **kwds):
        super(VerboseLoggingMonitor, self).__init__(interval, filename, new, all, info, **kwds)
        self._vyinterval = _interval(yinterval)
        self._vxinterval = _interval(xinterval)

    def info(self, message):
        super(VerboseLoggingMonitor, self).info(message)
        print('%s' % message)

    def __call__(self, x, y, id=None, best=0, k=False, **kwds):
        super(VerboseLoggingMonitor, self).__call__(x, y, id, best, k)
        _screen_report(self, x, y, id, best, k, self._vyinterval, self._vxinterval)

    def __reduce__(self):
        state = dict(_x=self._x, _y=self._y, _id=self._id, _info=self._info, k=self.k, label=self.label)
        return (self.__class__, (self._yinterval, self._vyinterval, self._vxinterval, self._filename, False, self._all, None),
                state)
