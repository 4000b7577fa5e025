#!/usr/bin/env python
"""Benchmark of the DE generation hot path (BASELINE.json metric: DE candidate-evals/sec,
gen+cost+select).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

A "step" is one generation over the whole (per-GPU) population: trial generation + bounds + cost +
penalty + selection + generation best (+ the best exchange when N > 1).  Prints ONE JSON line.
Workloads:
  c2  Rosenbrock 1024-D, NP=65536 per GPU, Best1Bin, bounds +-5 clamped (tight=True)   [default]
  c4  Corana 64-D, NP=1M per GPU, Rand1Exp, clamp +-1000 + quadratic_inequality(sum(x)-100)
  c5  Griewangk 256-D, NP=2M per GPU, Best1Bin, clamp +-400
  c3  chebyshev8cost M=4096, D=9, NP=1M per GPU, Best1Exp CR=1.0 F=0.9 (Horner path)
N > 1 shards by candidate, one rank per GPU, weak scaling (NP per GPU fixed), per-generation
all-gather of the shard bests over NCCL.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

WORKLOADS = {
    'c2': dict(desc='Rosenbrock 1024-D, NP=65536/GPU, Best1Bin CR=0.9 F=0.8, SetStrictRanges(+-5, tight=True)',
               cost='rosen', extra=(), dim=1024, np=65536, strategy='Best1Bin', CR=0.9, F=0.8, lo=-5., hi=5.,
               bounds='clamp', penalty=None, k=2),
    'c2a': dict(desc='Rosenbrock 1024-D, NP=65536/GPU, Best1Bin, SetStrictRanges(+-5) reject mode (OOB -> inf)',
                cost='rosen', extra=(), dim=1024, np=65536, strategy='Best1Bin', CR=0.9, F=0.8, lo=-5., hi=5.,
                bounds='reject', penalty=None, k=2),
    'c4': dict(desc='Corana 64-D, NP=1048576/GPU, Rand1Exp CR=0.9 F=0.8, clamp +-1000, quadratic_inequality(sum(x)-100)',
               cost='corana', extra=(), dim=64, np=1048576, strategy='Rand1Exp', CR=0.9, F=0.8, lo=-1000., hi=1000.,
               bounds='clamp', penalty=dict(G=1.0, h=100., k=100, hpow=5), k=3),
    'c5': dict(desc='Griewangk 256-D, NP=2097152/GPU, Best1Bin CR=0.9 F=0.8, clamp +-400',
               cost='griewangk', extra=(), dim=256, np=2097152, strategy='Best1Bin', CR=0.9, F=0.8, lo=-400.,
               hi=400., bounds='clamp', penalty=None, k=2),
    # other short-row shapes of the exponential-crossover kernels (K2f / K2d): 8 lanes per row, transcendental cost
    'g128': dict(desc='Griewangk 128-D, NP=1048576/GPU, Rand1Exp CR=0.9 F=0.8, clamp +-400',
                 cost='griewangk', extra=(), dim=128, np=1048576, strategy='Rand1Exp', CR=0.9, F=0.8, lo=-400., hi=400.,
                 bounds='clamp', penalty=None, k=3),
    'r32': dict(desc='Rosenbrock 32-D, NP=2097152/GPU, Best1Exp CR=0.9 F=0.8, clamp +-5',
                cost='rosen', extra=(), dim=32, np=2097152, strategy='Best1Exp', CR=0.9, F=0.8, lo=-5., hi=5.,
                bounds='clamp', penalty=None, k=2),
    # mystic's default strategy on short rows (not a BASELINE config): whole donor rows, binomial crossover
    'b4': dict(desc='Rosenbrock 64-D, NP=1048576/GPU, Best1Bin CR=0.9 F=0.8, clamp +-5 (binomial crossover, short rows)',
               cost='rosen', extra=(), dim=64, np=1048576, strategy='Best1Bin', CR=0.9, F=0.8, lo=-5., hi=5.,
               bounds='clamp', penalty=None, k=2),
    # BASELINE config 5 at its full size: NP = 16 M rows in total (32 GiB per population buffer), split over the
    # ranks (strong scaling); the end-to-end leg is skipped (it would pin 32 GiB of host memory per run)
    'c5s': dict(desc='Griewangk 256-D, NP=16777216 in total (strong scaling), Best1Bin CR=0.9 F=0.8, clamp +-400',
                cost='griewangk', extra=(), dim=256, np=16777216, strategy='Best1Bin', CR=0.9, F=0.8, lo=-400.,
                hi=400., bounds='clamp', penalty=None, k=2, strong=True, no_e2e=True),
    'c3': dict(desc='chebyshev8cost M=4096, 9 coeffs, NP=1048576/GPU, Best1Exp CR=1.0 F=0.9, init +-100 (FP64 DMMA)',
               cost='chebyshev8cost', extra=(4096,), dim=9, np=1048576, strategy='Best1Exp', CR=1.0, F=0.9,
               lo=-100., hi=100., bounds=None, penalty=None, k=2, tensor_cores=True),
    'c3h': dict(desc='chebyshev8cost M=4096, 9 coeffs, NP=1048576/GPU, Best1Exp CR=1.0 F=0.9, init +-100 (Horner)',
                cost='chebyshev8cost', extra=(4096,), dim=9, np=1048576, strategy='Best1Exp', CR=1.0, F=0.9,
                lo=-100., hi=100., bounds=None, penalty=None, k=2),
    'x2': dict(desc='Rosenbrock 1024-D, NP=65536/GPU, Rand1Exp CR=0.9 F=0.8, SetStrictRanges(+-5, tight=True) (exponential crossover, long rows)',
               cost='rosen', extra=(), dim=1024, np=65536, strategy='Rand1Exp', CR=0.9, F=0.8, lo=-5., hi=5.,
               bounds='clamp', penalty=None, k=3),
    # rows of odd length (not 16-byte multiples: no bulk copies) stay on the warp-per-row kernel K2b
    'o2': dict(desc='Rosenbrock 1023-D, NP=65536/GPU, Best1Bin CR=0.9 F=0.8, SetStrictRanges(+-5, tight=True) (odd row length)',
               cost='rosen', extra=(), dim=1023, np=65536, strategy='Best1Bin', CR=0.9, F=0.8, lo=-5., hi=5.,
               bounds='clamp', penalty=None, k=2),
    'ox2': dict(desc='Rosenbrock 1023-D, NP=65536/GPU, Rand1Exp CR=0.9 F=0.8, SetStrictRanges(+-5, tight=True) (odd row length)',
                cost='rosen', extra=(), dim=1023, np=65536, strategy='Rand1Exp', CR=0.9, F=0.8, lo=-5., hi=5.,
                bounds='clamp', penalty=None, k=3),
    # data-fit residual costs of forward_model.CostFactory (extra = points/observations, made in main)
    'f2': dict(desc='polynomial least-squares fit (poly.CostFactory2), 9 coeffs x 4096 points, NP=1048576/GPU, Best1Exp CR=1.0 '
                    'F=0.9, init +-100 (FP64 DMMA)',
               cost='polyfit', extra=(), fit=dict(M=4096), dim=9, np=1048576, strategy='Best1Exp', CR=1.0, F=0.9,
               lo=-100., hi=100., bounds=None, penalty=None, k=2, tensor_cores=True),
    'f2h': dict(desc='polynomial least-squares fit (poly.CostFactory2), 9 coeffs x 4096 points, NP=1048576/GPU, Best1Exp '
                     'CR=1.0 F=0.9, init +-100 (Horner, numpy.polyval order)',
                cost='polyfit', extra=(), fit=dict(M=4096), dim=9, np=1048576, strategy='Best1Exp', CR=1.0, F=0.9,
                lo=-100., hi=100., bounds=None, penalty=None, k=2),
}

BIN_STRATEGIES = ('Best1Bin',)


def fit_data(w):
    """synthetic data of a data-fit workload: Chebyshev-8 values on [-1, 1] plus a smooth perturbation"""
    M = w['fit']['M']
    pts = np.linspace(-1.0, 1.0, M)
    obs = np.polyval([128., 0., -256., 0., 160., 0., -32., 0., 1.], pts) + 0.25 * np.sin(7.0 * pts)
    return pts, obs


def device_cost(w):
    """-> (mystic_b200.models cost object, ExtraArgs)"""
    from mystic_b200 import models
    if w['cost'] == 'polyfit':
        pts, obs = w['extra'][0], w['extra'][1]
        return models.poly.CostFactory2(pts, obs, w['dim']), None
    return getattr(models, w['cost']), (tuple(w['extra']) or None)

# `ncu --set full` captures of the fused generation kernel at each workload's full per-GPU size
# (summaries under profiles/, newest first): roofline.traffic = dram__bytes_read.sum +
# dram__bytes_write.sum of ONE launch, read from the committed summary at run time
NCU_SUMMARIES = {
    'c2': ('profiles/r2_c2_tma_kernel_ncu_full_summary.csv', 'profiles/r1_c2_tma_kernel_ncu_full_summary.csv'),
    'x2': ('profiles/r2_x2_exp_tma_kernel_ncu_full_summary.csv', 'profiles/r1_x2_exp_tma_kernel_ncu_full_summary.csv'),
    'c3': ('profiles/r2_c3_kernel_ncu_full_summary.csv', 'profiles/r1_c3_kernel_ncu_full_summary.csv'),
    'f2': ('profiles/r2_f2_polyfit_dmma_ncu_full_summary.csv', 'profiles/r1_f2_polyfit_dmma_ncu_full_summary.csv'),
    'c4': ('profiles/r2_c4_rows_bulk_kernel_ncu_full_summary.csv', 'profiles/r1_c4_rows_bulk_kernel_ncu_full_summary.csv'),
    'c5': ('profiles/r2_c5_kernel_ncu_full_summary.csv', 'profiles/r1_c5_kernel_ncu_full_summary.csv'),
}
_UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


# the same kernels committing the generation in place (accepted trials only are written)
NCU_SUMMARIES_IN_PLACE = {
    'c2': ('profiles/r2_c2_pool_ncu_full_summary.csv', 'profiles/r2_c2_inplace_ncu_full_summary.csv'),
    'x2': ('profiles/r2_x2_inplace_ncu_full_summary.csv',),
    'c4': ('profiles/r2_c4_inplace_ncu_full_summary.csv',),
    'c5': ('profiles/r2_c5_inplace_ncu_full_summary.csv',),
}


def ncu_traffic(workload, in_place=False):
    """-> (DRAM bytes read + written by one launch of the fused kernel, summary file) or (None, None)"""
    import csv
    for rel in (NCU_SUMMARIES_IN_PLACE if in_place else NCU_SUMMARIES).get(workload, ()):
        path = os.path.join(REPO, rel)
        if not os.path.exists(path):
            continue
        got = {}
        with open(path) as f:
            for row in csv.reader(f):
                if len(row) >= 4 and row[1] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and row[1] not in got:
                    try:
                        got[row[1]] = float(row[3]) * _UNIT.get(row[2], 1.0)
                    except ValueError:
                        pass
        if len(got) == 2:
            return got['dram__bytes_read.sum'] + got['dram__bytes_write.sum'], rel
    return None, None


def algorithmic_bytes_per_candidate(w, in_place=False, accept=0.0):
    """population double buffered; parent + k donor rows read, new row written,
    energy read + written.  Exponential crossover reads only E[L] donor elements.
    in_place (the reference's own population + trialSolution layout, mb2_set_commit_mode): a rejected row is not
    written -- parent + donors read, energy read + written, and the accepted fraction `accept` of the rows written
    once; what de_commit_accepted_kernel moves on top of that (an accepted row read and written once more, the
    energies scanned) is overhead of the implementation and not counted."""
    D, k, CR = w['dim'], w['k'], w['CR']
    row_out = 8.0 * D * (accept if in_place else 1.0)
    if w['strategy'] in BIN_STRATEGIES:
        return 8.0 * D * (k + 1) + row_out + 16.0
    EL = float(D) if CR >= 1.0 else CR * (1.0 - CR ** D) / (1.0 - CR)
    return 8.0 * D + row_out + 8.0 * k * EL + 16.0


def algorithmic_flops_per_candidate(w):
    if w['cost'].startswith('chebyshev'):
        return 2.0 * w['dim'] * (w['extra'][0] + 2)
    if w['cost'] == 'polyfit':
        M = w['fit']['M']
        return (2.0 * w['dim'] + 3.0) * M   # Horner mul+add per coefficient, residual, square, accumulate
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, universal_newlines=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(',')])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def bind_to_gpu_numa_node(device_index):
    """Run this process on the CPUs of the NUMA node the GPU hangs off, so the pinned host buffers of
    the end-to-end leg are first touched (and DMA-read) locally: with 8 ranks uploading at once a
    remote node halves the aggregate H2D rate.  Returns the node, or None when it cannot be found."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        if all(hasattr(pr, a) for a in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        else:
            out = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', str(device_index)],
                                 stdout=subprocess.PIPE, universal_newlines=True, timeout=20).stdout.strip()
            dom, rest = out.split(':', 1)
            bdf = (dom[-4:] + ':' + rest).lower()
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read())
        if node < 0:
            # a VM that hides the topology (`GPU NUMA ID N/A`): with a single node there is nothing to choose
            nodes = [d for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit()]
            return int(nodes[0][4:]) if len(nodes) == 1 else None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the C/OpenMP oracle port on every host core over the arm's whole
# workload, and the reference's own Python solver (baseline/_ref) beside it.  These legs always run
# in their own process (`bench.py --impl reference`): the GPU process never maps oracle code.
# ------------------------------------------------------------------------------------------
def host_threads():
    return len(os.sched_getaffinity(0))


def workload_config(w, NP_local, NP_global):
    """the `config` object both arms print (identical for the same command line)"""
    return {'workload': w['desc'], 'np_per_gpu': NP_local, 'np_total': NP_global, 'dim': w['dim'],
            'l2': 'inputs larger than L2 (population %.0f MiB per GPU per buffer, no flush needed)'
                  % (NP_local * w['dim'] * 8 / 2 ** 20)}


def oracle_setup(w, NP, seed):
    from oracle import de_oracle as orc
    D = w['dim']
    lo, hi = np.full(D, w['lo']), np.full(D, w['hi'])
    mode = {None: orc.BOUNDS_NONE, 'clamp': orc.BOUNDS_CLAMP, 'reject': orc.BOUNDS_REJECT}[w['bounds']]
    pen = ()
    if w['penalty']:
        p = w['penalty']
        pen = [dict(ptype='quadratic_inequality', G=[p['G']] * D, h=p['h'], k=p['k'], hpow=p['hpow'])]
    pop0 = orc.philox_init_population(seed, NP, lo, hi)
    st = orc.DEState(pop0, lo if mode else None, hi if mode else None, mode)
    cost = orc.make_cost(w['cost'], w['extra'])
    orc.step0(st, cost, pen)
    return orc, st, cost, pen


def time_oracle(w, NP, steps, warmup, seed=0x5EED):
    """-> (cand-evals/s, seconds per step, kind, cores, sample description)"""
    try:
        from oracle import c_port
        if c_port.available():
            return c_port.time_workload(w, NP, steps, warmup, seed, threads=host_threads())
    except ImportError:
        pass
    orc, st, cost, pen = oracle_setup(w, NP, seed)
    for _ in range(warmup):
        orc.step(st, w['strategy'], w['CR'], w['F'], cost, pen, philox=dict(seed=seed, offset=0))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.step(st, w['strategy'], w['CR'], w['F'], cost, pen, philox=dict(seed=seed, offset=0))
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return NP / dt, dt, 'port', 1, 'numpy oracle port (oracle/de_oracle.py), NP=%d of %d rows, %d generations' % (
        NP, w['np'], steps)


def reference_python(workload, quick, timeout=600):
    """the UNMODIFIED reference solver (mystic from baseline/_ref) timed by oracle/reference_python.py"""
    cmd = [sys.executable, os.path.join(REPO, 'oracle', 'reference_python.py'), '--workload', workload]
    cmd += (['--np', '1024', '--generations', '2', '--pool-np'] if quick else
            ['--np', '1024', '4096', '--generations', '2', '--pool-np', '64', '--pool-generations', '1', '--procs',
             str(min(4, host_threads()))])
    env = dict(os.environ)
    env.pop('OMP_NUM_THREADS', None)
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, timeout=timeout,
                             env=env)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
        return json.loads(lines[-1]) if lines else {'error': (out.stderr or 'no output')[-300:]}
    except Exception as e:
        return {'error': '%s: %s' % (type(e).__name__, str(e)[:200])}


def run_reference_arm(args, w):
    """`--impl reference`: rank 0 alone; the whole workload of the arm (all N shards), every host core"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = max(1, args.gpus)
    NP_local = w['np'] // world if w.get('strong') else w['np']
    NP_global = NP_local * world
    NP = min(NP_global, args.ref_np) if args.ref_np else NP_global
    steps = max(1, args.steps)
    value, dt, kind, cores, sample = time_oracle(w, NP, steps, max(1, min(args.warmup, 2)))
    cpu = {'value': value, 'unit': 'candidate-evals/s', 'cores': cores, 'kind': kind, 'sample': sample}
    if not args.no_reference_python:
        cpu['reference_python'] = reference_python(args.workload, quick=not args.full_reference_python)
    line = {
        'impl': 'reference', 'metric': 'DE candidate-evals/sec (gen+cost+select)', 'value': value,
        'unit': 'candidate-evals/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong' if w.get('strong') else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(w, NP_local, NP_global),
        'details': {'rows_timed': NP, 'host_threads': cores},
        'cpu_baseline': cpu,
        'e2e': {'value': value, 'unit': 'candidate-evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def cpu_baseline_subprocess(workload, steps=3, warmup=1, full_python=True, timeout=900):
    """our arm's `cpu_baseline`: the reference arm in a child process (so this process maps no oracle code)"""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--workload', workload, '--gpus', '1',
           '--steps', str(steps), '--warmup', str(warmup)] + (['--full-reference-python'] if full_python else [])
    env = dict(os.environ)
    for k in ('OMP_NUM_THREADS', 'RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, env=env), timeout


def collect_cpu_baseline(handle):
    proc, timeout = handle
    try:
        out, err = proc.communicate(timeout=timeout)
        lines = [ln for ln in out.splitlines() if ln.startswith('{')]
        if lines:
            return json.loads(lines[-1])['cpu_baseline']
        return {'value': None, 'error': (err or 'no output')[-300:]}
    except Exception as e:
        proc.kill()
        return {'value': None, 'error': '%s: %s' % (type(e).__name__, str(e)[:200])}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_solver(w, NP_global, distributed, seed):
    from mystic_b200 import DifferentialEvolutionSolver2, penalty
    D = w['dim']
    s = DifferentialEvolutionSolver2(D, NP_global, rng='philox', seed=seed, distributed=distributed,
                                     CrossProbability=w['CR'], ScalingFactor=w['F'], strategy=w['strategy'],
                                     tensor_cores=bool(w.get('tensor_cores')))
    if w['bounds'] == 'clamp':
        s.SetStrictRanges([w['lo']] * D, [w['hi']] * D, tight=True)
    elif w['bounds'] == 'reject':
        s.SetStrictRanges([w['lo']] * D, [w['hi']] * D)
    if w['penalty']:
        p = w['penalty']

        @penalty.quadratic_inequality(penalty.linear([p['G']] * D, p['h']), k=p['k'], h=p['hpow'])
        def pen(x):
            return 0.0
        s.SetPenalty(pen)
    s.SetEvaluationLimits(generations=10 ** 9, evaluations=10 ** 18)
    return s


class Rig(object):
    """process-wide state of our arm: ranks, barrier, peaks"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: the engine has no CPU fallback')
        torch.cuda.set_device(self.local_rank)
        self.distributed = self.world > 1
        if self.distributed:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
        self.warmup = max(3, args.warmup)
        self.steps = max(1, args.steps)
        self.seed = 0x5EED
        self.peaks, self.peaks_src = measured_peaks()
        self.fp64 = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.distributed:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def fp64_peaks(self):
        """FP64 peaks of this GPU measured in this run: DFMA chain and DMMA.8x8x4 (mb2_probe_fp64_peak, our own
        probe kernels) and cuBLAS DGEMM 4096^3 through torch.matmul (a library call, off the hot path)"""
        if self.fp64 is None:
            from mystic_b200.engine import probe_fp64_peak
            torch = self.torch
            dfma, dmma = probe_fp64_peak(self.local_rank)
            n = 4096
            a = torch.randn((n, n), dtype=torch.float64, device='cuda')
            b = torch.randn((n, n), dtype=torch.float64, device='cuda')
            torch.matmul(a, b)
            best = float('inf')
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            del a, b
            torch.cuda.empty_cache()
            self.fp64 = {'dfma_tflops': dfma, 'dmma_m8n8k4_tflops': dmma, 'cublas_dgemm_4096_tflops': 2.0 * n ** 3 / (best * 1e-3) / 1e12}
        return self.fp64


def device_leg(rig, name, w):
    """device-resident throughput of one workload: population generated on the device, K generations timed with
    CUDA events on the launching stream (burst, then again under sustained load), max over ranks"""
    torch, dist = rig.torch, rig.dist
    from mystic_b200 import termination
    world, distributed, steps = rig.world, rig.distributed, rig.steps
    NP_local = w['np'] // world if w.get('strong') else w['np']
    NP_global = NP_local * world
    D = w['dim']
    cost, extra = device_cost(w)
    s = build_solver(w, NP_global, distributed, rig.seed)
    s.SetRandomInitialPoints([w['lo']] * D, [w['hi']] * D)
    s.SetTermination(termination.VTR(-1.0))
    s.Step(cost, ExtraArgs=extra)               # generation 0 (allocates, initialises, evaluates)
    for _ in range(rig.warmup):
        s.Step()
    eng = s._engine
    grid, block, smem = eng.launch_info()
    family = eng.kernel_family()

    def timed(nsteps):
        rig.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            eng.step()                           # K2 + K3 (+ the fused best exchange) on torch's current stream
            if distributed:
                s._exchange_best()               # no-op with the peer exchange; the all-gather with --exchange nccl
        e1.record()
        rig.barrier()
        return rig.max_over_ranks(e0.elapsed_time(e1))

    burst_ms = timed(steps)                      # cold clocks/power state: a burst figure, reported beside `value`
    sampler = ClockSampler(rig.local_rank) if rig.rank == 0 else None
    if sampler:
        sampler.start()
    # un-timed sustain phase: the timed region below is only milliseconds long, too short for nvidia-smi to
    # sample, so clocks / throttle reasons are sampled from here through the end of the timed region, under
    # the same load
    t_sus = time.perf_counter()
    while time.perf_counter() - t_sus < rig.args.sustain:
        for _ in range(10):
            eng.step()
            if distributed:
                s._exchange_best()
        torch.cuda.synchronize()
    launches0 = eng.launch_count()
    dev_ms = timed(steps)                        # the reported value: K generations under sustained load
    launches = eng.launch_count() - launches0
    in_place = eng.last_commit_in_place()
    res = s._engine.read_result()
    clocks = sampler.stop() if sampler else None
    out = dict(name=name, w=w, NP_local=NP_local, NP_global=NP_global, dev_ms=dev_ms, burst_ms=burst_ms, launches=int(launches),
               clocks=clocks, grid=grid, block=block, smem=smem, family=family, in_place=in_place,
               inf_fraction=res[2] / float(NP_global if distributed else NP_local),
               accept_fraction=res[3] / float(NP_global if distributed else NP_local))   # (sharded: the counters are sums over the shards)
    del s, eng
    gc.collect()
    torch.cuda.empty_cache()
    return out


def kernel_name(leg):
    """family of the fused kernel, with the kernel's name where the family has one kernel"""
    if leg['family'] == 'tma_pool':
        return 'tma_pool: de_step_best1bin_pool_kernel'
    return leg['family']


def roofline_of(rig, leg):
    w, name, NP_local = leg['w'], leg['name'], leg['NP_local']
    peaks = rig.peaks
    in_place = bool(leg.get('in_place'))
    bpc = algorithmic_bytes_per_candidate(w, in_place, leg['accept_fraction'])
    bpc_db = algorithmic_bytes_per_candidate(w)   # SURVEY 8(d)'s double-buffer figure (round 1's denominator)
    step_s = leg['dev_ms'] * 1e-3 / rig.steps
    burst_s = leg['burst_ms'] * 1e-3 / rig.steps
    achieved = bpc * NP_local / step_s / 1e9
    base = name.rstrip('s') if name == 'c5s' else name
    traffic, traffic_src = (None, None)
    if NP_local == WORKLOADS.get(base, {}).get('np'):
        traffic, traffic_src = ncu_traffic(base, in_place)
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'], 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': rig.peaks_src, 'frac_burst': bpc * NP_local / burst_s / 1e9 / peaks['hbm_gbs'],
                'algorithmic_bytes_per_candidate': bpc, 'algorithmic_bytes_per_launch': bpc * NP_local,
                'commit': 'in_place' if in_place else 'double_buffer',
                'kernel': 'fused generation kernel (%s)%s + de_finalize_kernel, per generation, per GPU'
                          % (kernel_name(leg), ' + de_commit_accepted_kernel' if in_place else '')}
    if in_place:
        # continuity with round 1 / SURVEY 8(d): the same step time against the bytes a double-buffered generation
        # would move (every row written); above 1 = faster than that design can be
        roofline['double_buffer_model'] = {'algorithmic_bytes_per_candidate': bpc_db,
                                           'frac': bpc_db * NP_local / step_s / 1e9 / peaks['hbm_gbs']}

    flops = algorithmic_flops_per_candidate(w)
    if flops:
        # the Chebyshev / polynomial contraction is FP64-pipe bound; DMMA and DFMA share that pipe on B200, so the
        # denominator is the highest FP64 rate measured in this run, whichever way it was reached
        fp = rig.fp64_peaks()
        which = max(fp, key=fp.get)
        tf = flops * NP_local / step_s / 1e12
        roofline = {'bound': 'tensor', 'achieved': tf, 'peak': fp[which], 'unit': 'TFLOP/s', 'frac': tf / fp[which],
                    'frac_burst': flops * NP_local / burst_s / 1e12 / fp[which],
                    'frac_of_dmma_probe': tf / fp['dmma_m8n8k4_tflops'],
                    'traffic': traffic, 'traffic_source': traffic_src,
                    'peak_source': 'measured in this run: %s (highest of %s)' % (which, json.dumps({k: round(v, 2) for k, v in fp.items()})),
                    'algorithmic_flops_per_candidate': flops, 'hbm_frac': achieved / peaks['hbm_gbs'],
                    'kernel': ('de_step_cheb_mma_kernel' + ('<KS, RESID>' if w['cost'] == 'polyfit' else ''))
                    if w.get('tensor_cores') else 'de_step_kernel (Horner)'}
    return roofline


def leg_record(rig, leg):
    """sub-record of one workload for the `also` object"""
    steps = rig.steps
    return {'workload': leg['w']['desc'], 'np_per_gpu': leg['NP_local'], 'np_total': leg['NP_global'], 'dim': leg['w']['dim'],
            'n_gpus': rig.world, 'scaling': 'strong' if leg['w'].get('strong') else 'weak',
            'value': leg['NP_global'] * steps / (leg['dev_ms'] * 1e-3), 'unit': 'candidate-evals/s',
            'ms_per_step': leg['dev_ms'] / steps, 'burst_ms_per_step': leg['burst_ms'] / steps, 'steps': steps,
            'roofline': roofline_of(rig, leg), 'clocks': leg['clocks'], 'gpu_launches': leg['launches'],
            'kernel_family': leg['family'], 'grid': leg['grid'], 'block': leg['block'], 'smem_bytes': leg['smem'],
            'commit': 'in_place' if leg.get('in_place') else 'double_buffer',
            'inf_fraction': leg['inf_fraction'], 'accept_fraction': leg['accept_fraction']}


def e2e_leg(rig, w):
    """end to end through the public API: pinned host population -> SetPopulation -> Step (H2D + generation 0)
    -> Solve() for K generations, every generation's result read back -> seconds (max over ranks), breakdown"""
    torch = rig.torch
    from mystic_b200 import termination
    world, steps = rig.world, rig.steps
    NP_local = w['np'] // world if w.get('strong') else w['np']
    NP_global = NP_local * world
    D = w['dim']
    cost, extra = device_cost(w)
    numa_node = bind_to_gpu_numa_node(rig.local_rank)
    host_pop = torch.empty((NP_local, D), dtype=torch.float64, pin_memory=True)
    rng = np.random.default_rng(rig.seed + rig.rank)
    hp = host_pop.numpy()
    chunk = max(1, (1 << 24) // D)
    for i in range(0, NP_local, chunk):
        hp[i:i + chunk] = rng.uniform(w['lo'], w['hi'], size=(min(chunk, NP_local - i), D))

    # what the platform allows: a bare copy of the same pinned buffer, all ranks at once (the 8-GPU boxes of this pool
    # are one-socket VMs whose host side delivers ~55 GB/s to one GPU but only ~115-190 GB/s to 4-8 GPUs together:
    # profiles/README.md, tools/h2d_concurrent.py)
    scratch = torch.empty((NP_local, D), dtype=torch.float64, device='cuda')
    floor_ms = float('inf')
    for _ in range(2):
        rig.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scratch.copy_(host_pop, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        floor_ms = min(floor_ms, rig.max_over_ranks(e0.elapsed_time(e1)))
    del scratch
    torch.cuda.empty_cache()

    def e2e_run(breakdown):
        s2 = build_solver(w, NP_global, rig.distributed, rig.seed)
        s2.SetTermination(termination.VTR(-1.0))
        rig.barrier()
        s2.SetEvaluationLimits(generations=steps, evaluations=10 ** 18)
        t0 = time.perf_counter()
        s2.SetPopulation(hp, local=True)         # this rank's rows; pinned, uploaded as is
        s2.Step(cost, ExtraArgs=extra)           # H2D upload + generation 0
        if breakdown:
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        s2.Solve()                               # the user's call: K generations until the generation limit;
        torch.cuda.synchronize()                 # every generation's (bestEnergy, counters, bestSolution) reaches the host
        dt = time.perf_counter() - t0
        assert s2.generations == steps, (s2.generations, steps)
        del s2
        gc.collect()
        return dt, (t1 - t0), (t0 + dt - t1)

    e2e_run(False)                               # warm-up (device allocations come from torch's cache afterwards)
    e2e_s = min(e2e_run(False)[0], e2e_run(False)[0])   # best of two (host-side jitter: PCIe, page faults)
    e2e_s = rig.max_over_ranks(e2e_s)
    _, up_s, st_s = e2e_run(True)                # once more with a synchronize between the parts, for the breakdown
    del host_pop, hp
    torch.cuda.empty_cache()
    return dict(seconds=e2e_s, NP_local=NP_local, NP_global=NP_global, numa_node=numa_node,
                breakdown={'bare_h2d_copy_of_the_inputs_ms': floor_ms,
                           'upload_and_generation0_ms': rig.max_over_ranks(up_s) * 1e3,
                           'solve_%d_generations_ms' % steps: rig.max_over_ranks(st_s) * 1e3,
                           'note': 'the two parts from an extra run with a synchronize between them; the bare copy is the same '
                                   'pinned buffer moved by one cudaMemcpyAsync per rank, all ranks at once: the platform\'s floor '
                                   'for the upload'})


def shard_parity(rig):
    """N > 1: one small sharded run of the drop-in solver against the numpy oracle's emulation of all shards
    (tests/dist_gpu_check.emulate: shard-local donors, best merged by the reference's rule) -> bit-exact?"""
    torch, dist = rig.torch, rig.dist
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    from oracle import de_oracle as orc
    from tests.dist_gpu_check import emulate
    ok = True
    # (populations larger than the dimension: the solver raises NP to max(NP, dim) like the reference does)
    for strategy, cost, D, NP, lo, hi in (('Best1Bin', 'rosen', 256, 64 * rig.world + 259, -5., 5.),
                                          ('Rand1Exp', 'corana', 64, 96 * rig.world + 5, -1000., 1000.)):
        seed, gens = 0xD15C0 + D, 4
        s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=seed, distributed=True, strategy=strategy)
        s.SetRandomInitialPoints([lo] * D, [hi] * D)
        s.SetStrictRanges([lo] * D, [hi] * D, tight=True)
        s.SetTermination(termination.VTR(-1.0))
        s.SetEvaluationLimits(generations=10 ** 6)
        s.Step(getattr(models, cost))
        for _ in range(gens):
            s.Step()
        engs, hist = emulate(rig.world, NP, D, seed, strategy, getattr(models, cost).cost_id,
                             (orc.BOUNDS_CLAMP, np.full(D, lo), np.full(D, hi)), gens)
        mine = engs[rig.rank]
        ok = ok and np.array_equal(s.population, mine.population()) \
            and np.array_equal(np.array(s.bestSolution), mine.state.bestSolution) \
            and np.allclose(np.array(s.energy_history), np.array(hist), rtol=1e-12, atol=0)
        del s
    t = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--np', type=int, default=0, help='override rows per GPU (testing)')
    ap.add_argument('--ref-np', type=int, default=0, help='cap on the rows of the CPU leg (0 = the whole workload)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-python', action='store_true', help='CPU leg: skip the reference\'s own Python solver')
    ap.add_argument('--full-reference-python', action='store_true',
                    help='CPU leg: also NP=4096 serial and the multiprocess.Pool map (about a minute more)')
    ap.add_argument('--also', default=None,
                    help='comma-separated workloads measured after the main one and reported as sub-records of the same '
                         'line (default with the default workload: c5s at every N; c4 and c3 at N=1); "none" disables')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl'],
                    help='multi-GPU best exchange: stores into peer memory fused into the epilogue kernel, or an NCCL all-gather')
    ap.add_argument('--sustain', type=float, default=0.7, help='seconds of un-timed load before the timed region')
    args = ap.parse_args()
    os.environ['MB2_EXCHANGE'] = args.exchange
    if os.environ.get('MB2_BENCH_DEBUG'):
        import faulthandler
        faulthandler.dump_traceback_later(600, exit=True)   # a hung run prints where it hangs

    def workload(name):
        w = dict(WORKLOADS[name])
        if w.get('fit'):
            w['extra'] = fit_data(w)
        return w
    w = workload(args.workload)
    if args.np:
        w['np'] = args.np

    if args.impl == 'reference':
        run_reference_arm(args, w)
        return

    rig = Rig(args)
    torch, dist = rig.torch, rig.dist
    world, rank, steps = rig.world, rig.rank, rig.steps
    if args.also is None:
        also_names = (['c5s'] + (['c4', 'c3'] if world == 1 else [])) if (args.workload == 'c2' and not args.np) else []
    else:
        also_names = [a for a in args.also.split(',') if a and a != 'none']

    leg = device_leg(rig, args.workload, w)
    e2e = None if w.get('no_e2e') else e2e_leg(rig, w)
    # the CPU legs run in a child process beside the device-timed `also` legs (they do not touch the GPU, and the
    # `also` legs are timed on the device); the end-to-end leg above, which is host-timed, ran alone
    cpu_handle = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_handle = cpu_baseline_subprocess(args.workload, full_python=not args.no_reference_python)
    also = {}
    for name in also_names:
        try:
            also[name] = leg_record(rig, device_leg(rig, name, workload(name)))
        except Exception as e:  # one sub-record failing must not cost the line
            also[name] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
            gc.collect()
            torch.cuda.empty_cache()
    parity = shard_parity(rig) if rig.distributed else None

    if rank != 0:
        if rig.distributed:
            dist.destroy_process_group()
        return

    NP_local, NP_global, D = leg['NP_local'], leg['NP_global'], w['dim']
    line = {
        'metric': 'DE candidate-evals/sec (gen+cost+select)', 'value': NP_global * steps / (leg['dev_ms'] * 1e-3),
        'unit': 'candidate-evals/s', 'n_gpus': world, 'steps': steps, 'warmup': rig.warmup,
        'ms_per_step': leg['dev_ms'] / steps, 'higher_is_better': True,
        'scaling': 'strong' if w.get('strong') else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(w, NP_local, NP_global),
        'details': {'rng': 'philox4x32-10 in-kernel', 'grid': leg['grid'], 'block': leg['block'], 'smem_bytes': leg['smem'],
                    'kernel_family': leg['family'], 'inf_fraction': leg['inf_fraction'], 'accept_fraction': leg['accept_fraction'],
                    'commit': 'in_place (accepted trials only are written, de_commit_accepted_kernel moves them over their parents)'
                              if leg.get('in_place') else 'double_buffer (every selected row is written, buffers swapped)',
                    'sustain_s_before_timing': args.sustain, 'burst_ms_per_step': leg['burst_ms'] / steps,
                    'e2e_host_numa_node': e2e['numa_node'] if e2e else None,
                    'parallelism': (('candidate-sharded x%d, ' % world) + (
                        'best exchange by peer-memory stores (NVLink) fused into the argmin epilogue' if args.exchange == 'peer'
                        else 'NCCL all-gather of the bests per generation')) if rig.distributed else 'single GPU'},
        'roofline': roofline_of(rig, leg),
        'gpu_launches': leg['launches'],
        'clocks': leg['clocks'],
    }
    d2h_per_step = 8 * D + 40
    if e2e:
        line['e2e'] = {'value': NP_global * steps / e2e['seconds'], 'unit': 'candidate-evals/s',
                       'h2d_bytes_per_step': NP_local * D * 8.0 / steps, 'd2h_bytes_per_step': d2h_per_step,
                       'what': 'pinned host population -> H2D -> generation 0 -> solver.Solve() for %d generations (every '
                               'generation\'s best energy, counters and bestSolution read back to the host, termination + '
                               'monitor per generation); only the %d generations are counted' % (steps, steps),
                       'breakdown': e2e['breakdown']}
    else:
        line['e2e'] = {'value': None, 'unit': 'candidate-evals/s', 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': d2h_per_step,
                       'what': 'not measured for this workload: the host copy of the population is %.0f GiB per rank'
                               % (NP_local * D * 8 / 2 ** 30)}
    if also:
        line['also'] = also
    if parity is not None:
        line['shard_parity'] = parity
    if cpu_handle is not None:
        line['cpu_baseline'] = collect_cpu_baseline(cpu_handle)
    print(json.dumps(line))
    if rig.distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
