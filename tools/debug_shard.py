import sys
sys.path.insert(0, '.')
import numpy as np
from mystic_b200 import models
from mystic_b200.engine import DeviceEngine
from oracle import de_oracle as orc
from tests._oracle_engine import OracleEngine
for (D, strat, cost, npl, npg, off, lohi) in [(256, 'Best1Bin', 'rosen', 66, 131, 0, 5.), (256, 'Best1Bin', 'rosen', 65, 131, 66, 5.),
                                          (64, 'Rand1Exp', 'corana', 99, 197, 0, 1000.), (64, 'Rand1Exp', 'corana', 98, 197, 99, 1000.),
                                          (256, 'Best1Bin', 'rosen', 700, 1400, 700, 5.)]:
    seed = 0xD15C0 + D
    lo, hi = np.full(D, -lohi), np.full(D, lohi)
    engs = []
    for E in (DeviceEngine, OracleEngine):
        e = E(npl, D, npg, off)
        e.set_cost(getattr(models, cost).cost_id)
        e.set_bounds(orc.BOUNDS_CLAMP, lo, hi)
        e.set_rng_philox(seed)
        e.init_population(seed, lo, hi)
        e.eval_initial()
        engs.append(e)
    for g in range(4):
        for e in engs:
            e.set_strategy(orc.STRATEGIES[strat][0], 0.8, 0.9)
            e.step()
        a, b = engs[0].population(), engs[1].population()
        bad = np.flatnonzero((a != b).any(axis=1))
        print(D, strat, npl, off, 'gen', g + 1, 'differing rows', bad[:10].tolist(), len(bad), engs[0].kernel_family(), engs[0].launch_info())
