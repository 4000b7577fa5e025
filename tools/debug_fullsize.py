#!/usr/bin/env python
"""Find and explain rows where the device and the C oracle port disagree at full size (BASELINE config 4 by default):
prints the draws, the parent, both trial vectors / energies and the accept decisions of every differing row."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mystic_b200 import models, penalty  # noqa: E402
from mystic_b200.engine import DeviceEngine  # noqa: E402
from oracle import c_port, de_oracle as orc  # noqa: E402


def main():
    NP, D, strategy, cost, lo, hi, CR, F, seed = 1048576, 64, 'Rand1Exp', 'corana', -1000., 1000., 0.9, 0.8, 0x5EED
    gens = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    c_port.load().orc_set_num_threads(len(os.sched_getaffinity(0)))
    lo_v, hi_v = np.full(D, lo), np.full(D, hi)
    pen = dict(G=1.0, h=100., k=100, hpow=5)
    eng = DeviceEngine(NP, D, keep_trials=True)
    eng.set_cost(models.corana.cost_id, models.corana.params(()))
    eng.set_bounds(orc.BOUNDS_CLAMP, lo_v, hi_v)

    def p0(x):
        return 0.0
    pf = penalty.quadratic_inequality(penalty.linear([pen['G']] * D, pen['h']), k=pen['k'], h=pen['hpow'])(p0)
    eng.set_penalty(*pf.device_terms(D))
    terms = [dict(ptype='quadratic_inequality', G=[pen['G']] * D, h=pen['h'], k=pen['k'], hpow=pen['hpow'])]
    eng.set_strategy(orc.STRATEGIES[strategy][0], F, CR)
    eng.set_rng_philox(seed)
    eng.init_population(seed, lo_v, hi_v)
    o = c_port.COracle(c_port.init_population(seed, NP, lo_v, hi_v), strategy, CR, F, cost, (), lo_v, hi_v, orc.BOUNDS_CLAMP, terms)
    for g in range(gens + 1):
        parent = o.population.copy()
        e_old = o.popEnergy.copy()
        if g == 0:
            eng.eval_initial()
            o.step0()
        else:
            eng.step()
            o.step(seed=seed)
        pop = eng.population()
        bad = np.argwhere((pop.view(np.uint64) != o.population.view(np.uint64)).any(axis=1)).ravel()
        print('generation %d: %d differing rows, kernel %s' % (g, len(bad), eng.kernel_family()), flush=True)
        if len(bad) == 0 or g == 0:
            continue
        trial, trial_e = eng.trials()
        for c in bad[:6]:
            donors, nstart, xword = orc.philox_donors(seed, g, np.array([c]), NP, D, 3, with_xword=True)
            L = orc.exp_length(xword, CR, D)
            cols = np.argwhere(pop[c].view(np.uint64) != o.population[c].view(np.uint64)).ravel()
            tcols = np.argwhere(trial[c].view(np.uint64) != o.trial[c].view(np.uint64)).ravel()
            print(' row %d: donors %s nstart %d L %d; differing columns %s; trial differs at %s' % (c, donors[0].tolist(), nstart[0], L[0], cols.tolist(), tcols.tolist()))
            print('   e_old %.17g  device trialE %.17g  oracle trialE %.17g  device acc %s oracle acc %s' % (
                e_old[c], trial_e[c], o.trial_e[c], trial_e[c] < e_old[c], o.trial_e[c] < e_old[c]))
            for j in (tcols.tolist() or cols.tolist())[:4]:
                print('   col %d: parent %.17g device trial %.17g oracle trial %.17g | donors %s' % (
                    j, parent[c, j], trial[c, j], o.trial[c, j], [float(parent[d, j]) for d in donors[0]]))
        break


if __name__ == '__main__':
    main()
