import sys, time, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
from mystic_b200 import engine as E
from mystic_b200._native import lib, check
NP, D = 65536, 1024
torch.zeros(1, device='cuda')
def T(label, f, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
    print('%-28s %8.3f ms' % (label, (time.perf_counter() - t) * 1e3)); return r
for rep in range(3):
    print('--- rep', rep)
    f64 = dict(dtype=torch.float64, device='cuda')
    a = T('2 x torch.empty pop', lambda: [torch.empty((NP, D), **f64) for _ in range(2)])
    b = T('2 x torch.full energy', lambda: [torch.full((NP,), float('inf'), **f64) for _ in range(2)])
    ctx = ctypes.c_void_p()
    T('mb2_create', lambda: check(lib.mb2_create(0, NP, D, NP, 0, ctypes.byref(ctx))))
    T('mb2_destroy', lambda: lib.mb2_destroy(ctx))
    eng = T('DeviceEngine()', E.DeviceEngine, NP, D, NP, 0)
    T('close', eng.close)
    del a, b, eng
