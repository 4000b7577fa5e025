"""Small runs of the bulk-copy kernels (K2f, K2e, K2a), meant for `compute-sanitizer --tool memcheck|racecheck
python tools/sanitize_case.py` on a machine where the sanitizer may run (it is closed on the round-1 GPU pool, so
bad accesses are hunted with small cases against the oracle instead: tests/test_gpu_exp_tma.py,
tests/test_gpu_rows_bulk.py)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mystic_b200 import models, penalty  # noqa: E402
from mystic_b200.engine import DeviceEngine  # noqa: E402

CASES = [  # (NP, D, strategy id, cost, penalty?, expected family); strategy ids: mystic/strategy.py in file order
    (500, 64, 2, models.corana, True, 'rows_bulk'),      # Rand1Exp
    (300, 128, 4, models.griewangk, False, 'rows_bulk'),  # Best2Exp, 8 lanes per row
    (200, 300, 0, models.rosen, False, 'tma_exp'),        # Best1Exp
    (150, 256, 5, models.corana, False, 'tma_exp'),       # Rand2Exp: a cost without a group reduction
    (180, 256, 1, models.griewangk, False, 'tma_bin'),    # Best1Bin, one-warp groups with a two-stage ring
    (170, 1024, 1, models.rosen, False, 'tma_pool'),      # Best1Bin in place: pooled row buffers behind two producer warps
    (2400, 1024, 1, models.rosen, False, 'tma_pool'),     # ... with rings that revolve (16 rows per CTA)
    (90, 1030, 2, models.rosen, False, 'tma_exp'),        # Rand1Exp, ragged row length
]
for NP, D, sid, cost, pen, family in CASES:
    eng = DeviceEngine(NP, D)
    eng.set_cost(cost.cost_id, cost.params(()))
    lo, hi = np.full(D, -5.), np.full(D, 5.)
    eng.set_bounds(2, lo, hi)
    if pen:
        @penalty.quadratic_inequality(penalty.linear([1.] * D, 10.), k=100, h=5)
        def p(x):
            return 0.0
        eng.set_penalty(*p.device_terms(D))
    eng.set_strategy(sid, 0.8, 0.9)
    eng.set_rng_philox(7)
    eng.init_population(7, lo, hi)
    eng.eval_initial()
    for _ in range(3):
        eng.step()
    r = eng.read_result()
    assert eng.kernel_family() == family, eng.kernel_family()
    print('ok', family, cost.__name__, D, r[0])
    eng.close()

# batch of solvers with device-side termination (stopped islands copy their rows through) and shared bests
from mystic_b200 import termination  # noqa: E402
eng = DeviceEngine(6 * 20, 200)
eng.set_cost(models.rosen.cost_id, ())
eng.set_bounds(2, np.full(200, -2.), np.full(200, 2.))
eng.set_strategy(0, 0.8, 0.9)
eng.set_rng_philox(3)
eng.init_population(3, np.full(200, -2.), np.full(200, 2.))
eng.set_islands(6)
eng.set_islands_termination(termination.device_descriptor(termination.ChangeOverGeneration(1e-1, 3)), 40, None)
recs, done, stopped = eng.run_islands(16)
print('ok islands termination', done.tolist(), stopped.tolist())
eng.close()
eng = DeviceEngine(8 * 12, 64)
eng.set_cost(models.rosen.cost_id, ())
eng.set_bounds(2, np.full(64, -2.), np.full(64, 2.))
eng.set_strategy(1, 0.8, 0.9)
eng.set_rng_philox(3)
eng.init_population(3, np.full(64, -2.), np.full(64, 2.))
eng.set_islands(8, share_best_within=4)
eng.eval_initial()
for _ in range(3):
    eng.step()
print('ok island groups', eng.read_islands()[0].tolist())
eng.close()
