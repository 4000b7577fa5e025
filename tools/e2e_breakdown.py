import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from mystic_b200 import DifferentialEvolutionSolver2, models, termination
from mystic_b200 import engine as E
NP, D = 65536, 1024
hp = torch.empty((NP, D), dtype=torch.float64).pin_memory().numpy()
hp[:] = np.random.default_rng(0).uniform(-5, 5, size=(NP, D))
def T(label, f, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
    print('%-28s %8.3f ms' % (label, (time.perf_counter() - t) * 1e3)); return r
for rep in range(2):
    print('--- rep', rep)
    s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=1, strategy='Best1Bin')
    s.SetStrictRanges([-5.] * D, [5.] * D, tight=True)
    s.SetTermination(termination.VTR(-1.0)); s.SetEvaluationLimits(generations=10**9, evaluations=10**18)
    T('SetPopulation', s.SetPopulation, hp, local=True)
    eng = T('DeviceEngine()', E.DeviceEngine, NP, D, NP, 0)
    T('upload_population', eng.upload_population, hp)
    T('set_bounds', eng.set_bounds, 2, np.full(D, -5.), np.full(D, 5.))
    T('set_cost', eng.set_cost, 0, [])
    T('eval_initial', eng.eval_initial)
    T('read_result', eng.read_result)
    eng.set_strategy(1, 0.8, 0.9); eng.set_rng_philox(1)
    T('20 x step', lambda: [eng.step() for _ in range(20)])
    T('20 x (step+read_result)', lambda: [(eng.step(), eng.read_result()) for _ in range(20)])
    eng.close(); del eng
    T('solver.Step (gen 0, all)', s.Step, models.rosen)
    T('20 x solver.Step', lambda: [s.Step() for _ in range(20)])
    del s
