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

# dram__bytes_read.sum + dram__bytes_write.sum of the fused generation kernel, per launch at the
# workload's full per-GPU size, from one `ncu --set full` capture each (summaries under profiles/)
NCU_TRAFFIC = {
    'c2': (1.444322e9 + 0.513861e9, 'profiles/r1_c2_tma_kernel_ncu_full_summary.csv'),
    'x2': (0.567731e9 + 0.493002e9, 'profiles/r1_x2_exp_tma_kernel_ncu_full_summary.csv'),
    'c3': (0.265026e9 + 0.037010e9, 'profiles/r1_c3_kernel_ncu_full_summary.csv'),
    'f2': (0.258753e9 + 0.034031e9, 'profiles/r1_f2_polyfit_dmma_ncu_full_summary.csv'),
    'c4': (1.026244e9 + 0.512175e9, 'profiles/r1_c4_rows_bulk_kernel_ncu_full_summary.csv'),
    'c5': (12.791601e9 + 4.323565e9, 'profiles/r1_c5_kernel_ncu_full_summary.csv'),
}


def algorithmic_bytes_per_candidate(w):
    """population double buffered; parent + k donor rows read, new row written,
    energy read + written.  Exponential crossover reads only E[L] donor elements."""
    D, k, CR = w['dim'], w['k'], w['CR']
    if w['strategy'] in BIN_STRATEGIES:
        return 8.0 * D * (k + 2) + 16.0
    EL = float(D) if CR >= 1.0 else CR * (1.0 - CR ** D) / (1.0 - CR)
    return 16.0 * D + 8.0 * k * EL + 16.0


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
            return None
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
# CPU baseline / reference arm: the oracle port on the host cores, bounded sample
# ------------------------------------------------------------------------------------------
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
            return c_port.time_workload(w, NP, steps, warmup, seed)
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


def run_reference_arm(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    NP = min(w['np'], args.ref_np)
    steps = max(1, args.steps)
    value, dt, kind, cores, sample = time_oracle(w, NP, steps, max(1, min(args.warmup, 2)))
    line = {
        'impl': 'reference', 'metric': 'DE candidate-evals/sec (gen+cost+select)', 'value': value,
        'unit': 'candidate-evals/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': w['desc'], 'sample_rows': NP},
        'cpu_baseline': {'value': value, 'unit': 'candidate-evals/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'candidate-evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--np', type=int, default=0, help='override rows per GPU (testing)')
    ap.add_argument('--ref-np', type=int, default=4096, help='rows of the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl'],
                    help='multi-GPU best exchange: stores into peer memory fused into the epilogue kernel, or an NCCL all-gather')
    ap.add_argument('--sustain', type=float, default=0.7, help='seconds of un-timed load before the timed region')
    args = ap.parse_args()
    os.environ['MB2_EXCHANGE'] = args.exchange
    if os.environ.get('MB2_BENCH_DEBUG'):
        import faulthandler
        faulthandler.dump_traceback_later(90, exit=True)   # a hung run prints where it hangs
    w = dict(WORKLOADS[args.workload])
    if w.get('fit'):
        w['extra'] = fit_data(w)
    if args.np:
        w['np'] = args.np

    if args.impl == 'reference':
        run_reference_arm(args, w)
        return

    import torch
    import torch.distributed as dist
    from mystic_b200 import models, termination

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    NP_local = w['np'] // world if w.get('strong') else w['np']
    NP_global = NP_local * world
    D = w['dim']
    seed = 0x5EED
    cost, extra = device_cost(w)
    never = termination.VTR(-1.0)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: population generated on the device, K generations timed
    s = build_solver(w, NP_global, distributed, seed)
    s.SetRandomInitialPoints([w['lo']] * D, [w['hi']] * D)
    s.SetTermination(never)
    s.Step(cost, ExtraArgs=extra)               # generation 0 (allocates, initialises, evaluates)
    for _ in range(warmup):
        s.Step()
    eng = s._engine
    grid, block, smem = eng.launch_info()
    def timed(nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            eng.step()                           # K2 + K3 on torch's current stream
            if distributed:
                s._exchange_best()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    burst_ms = timed(steps)                      # cold clocks/power state: a burst figure, reported beside `value`
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # un-timed sustain phase (~0.7 s of back-to-back generations): the timed region below is only a
    # few milliseconds, too short for nvidia-smi to sample, so clocks / throttle reasons are sampled
    # from here through the end of the timed region, under the same load.
    t_sus = time.perf_counter()
    while time.perf_counter() - t_sus < args.sustain:
        for _ in range(10):
            eng.step()
            if distributed:
                s._exchange_best()
        torch.cuda.synchronize()
    launches0 = eng.launch_count()
    dev_ms = timed(steps)                        # the reported value: K generations under sustained load
    launches = eng.launch_count() - launches0
    inf_frac = s._last[0] / float(NP_global) if s._last else 0.0
    acc_frac = s._last[1] / float(NP_global) if s._last else 0.0
    clocks = sampler.stop() if sampler else None
    del s, eng
    torch.cuda.empty_cache()

    # ---- end to end: host population (pinned) -> H2D -> generation 0 -> K x Step() with the result
    #      of every step read back to the host
    numa_node = bind_to_gpu_numa_node(local_rank)
    host_rows = 1 if w.get('no_e2e') else NP_local
    host_pop = torch.empty((host_rows, D), dtype=torch.float64, pin_memory=True)
    rng = np.random.default_rng(seed + rank)
    hp = host_pop.numpy()
    chunk = max(1, (1 << 24) // D)
    for i in range(0, host_rows, chunk):
        hp[i:i + chunk] = rng.uniform(w['lo'], w['hi'], size=(min(chunk, host_rows - i), D))
    def e2e_run():
        s2 = build_solver(w, NP_global, distributed, seed)
        s2.SetTermination(never)
        barrier()
        s2.SetEvaluationLimits(generations=steps, evaluations=10 ** 18)
        t0 = time.perf_counter()
        s2.SetPopulation(hp, local=True)         # this rank's rows; pinned, uploaded as is
        s2.Step(cost, ExtraArgs=extra)           # H2D upload + generation 0
        t1 = time.perf_counter()
        s2.Solve()                               # the user's call: K generations until the generation limit;
        torch.cuda.synchronize()                 # every generation's (bestEnergy, counters, bestSolution) reaches the host
        dt = time.perf_counter() - t0
        assert s2.generations == steps, (s2.generations, steps)
        if os.environ.get('MB2_BENCH_DEBUG'):
            print('e2e: upload + generation 0 %.2f ms, %d steps %.2f ms' % ((t1 - t0) * 1e3, steps, (t0 + dt - t1) * 1e3),
                  file=sys.stderr)
        del s2
        gc.collect()
        return dt

    if w.get('no_e2e'):
        e2e_s = float('nan')
    else:
        e2e_run()                                # warm-up: device allocations come from torch's cache afterwards
        e2e_s = min(e2e_run(), e2e_run())        # best of two (host-side jitter: PCIe, page faults)
    t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    d2h_per_step = 8 * D + 40
    h2d_per_step = NP_local * D * 8.0 / steps
    torch.cuda.empty_cache()

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    value = NP_global * steps / (dev_ms * 1e-3)
    peaks, peaks_src = measured_peaks()
    bpc = algorithmic_bytes_per_candidate(w)
    step_s = dev_ms * 1e-3 / steps
    achieved = bpc * NP_local / step_s / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'], 'traffic': None, 'peak_source': peaks_src,
                'frac_burst': bpc * NP_local / (burst_ms * 1e-3 / steps) / 1e9 / peaks['hbm_gbs'],
                'algorithmic_bytes_per_candidate': bpc,
                'kernel': 'de_step_kernel (fused generation) + de_finalize_kernel, per generation, per GPU'}
    if args.workload in NCU_TRAFFIC and NP_local == WORKLOADS[args.workload]['np']:
        roofline['traffic'], roofline['traffic_source'] = NCU_TRAFFIC[args.workload]
        roofline['algorithmic_bytes_per_launch'] = bpc * NP_local
    flops = algorithmic_flops_per_candidate(w)
    if flops:
        # the Chebyshev contraction is FP64-pipe bound: DMMA peak measured on this pool's B200 with
        # tools/fp64_peak.cu (profiles/r1_fp64_peaks_b200.json): 33.02 TFLOP/s (DFMA 34.15, cuBLAS DGEMM 35.4)
        tf = flops * NP_local / step_s / 1e12
        roofline = {'bound': 'tensor', 'achieved': tf, 'peak': 33.02, 'unit': 'TFLOP/s', 'frac': tf / 33.02,
                    'traffic': roofline.get('traffic'), 'traffic_source': roofline.get('traffic_source'),
                    'peak_source': 'measured DMMA.8x8x4 peak (profiles/r1_fp64_peaks_b200.json)',
                    'algorithmic_flops_per_candidate': flops, 'hbm_frac': achieved / peaks['hbm_gbs'],
                    'kernel': ('de_step_cheb_mma_kernel' + ('<KS, RESID>' if w['cost'] == 'polyfit' else ''))
                    if w.get('tensor_cores') else 'de_step_kernel (Horner)'}

    line = {
        'metric': 'DE candidate-evals/sec (gen+cost+select)', 'value': value, 'unit': 'candidate-evals/s',
        'n_gpus': world, 'steps': steps, 'warmup': warmup, 'ms_per_step': dev_ms / steps, 'higher_is_better': True,
        'scaling': 'strong' if w.get('strong') else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': w['desc'], 'np_per_gpu': NP_local, 'np_total': NP_global, 'dim': D,
                   'rng': 'philox4x32-10 in-kernel', 'l2': 'inputs larger than L2 (population %.0f MiB per buffer)'
                   % (NP_local * D * 8 / 2 ** 20), 'grid': grid, 'block': block, 'smem_bytes': smem,
                   'inf_fraction': inf_frac, 'accept_fraction': acc_frac,
                   'sustain_s_before_timing': args.sustain, 'burst_ms_per_step': burst_ms / steps,
                   'e2e_host_numa_node': numa_node,
                   'parallelism': ('candidate-sharded x%d, ' % world) + ('best exchange by peer-memory stores (NVLink) fused into the argmin epilogue'
                                                                          if args.exchange == 'peer' else 'NCCL all-gather of the bests per generation') if distributed
                   else 'single GPU'},
        'roofline': roofline,
        'e2e': {'value': NP_global * steps / e2e_s, 'unit': 'candidate-evals/s',
                'h2d_bytes_per_step': h2d_per_step, 'd2h_bytes_per_step': d2h_per_step,
                'what': 'pinned host population -> H2D -> generation 0 -> solver.Solve() for %d generations (every '
                        'generation\'s best energy, counters and bestSolution read back to the host, termination + '
                        'monitor per generation); only the %d generations are counted' % (steps, steps)},
        'gpu_launches': int(launches),
        'clocks': clocks,
    }
    if w.get('no_e2e'):
        line['e2e'] = {'value': None, 'unit': 'candidate-evals/s', 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': d2h_per_step,
                       'what': 'not measured for this workload: the host copy of the population is %.0f GiB per rank'
                               % (NP_local * D * 8 / 2 ** 30)}
    if world == 1 and not args.no_cpu_baseline:
        v, dt, kind, cores, sample = time_oracle(w, min(w['np'], args.ref_np), 3, 1)
        line['cpu_baseline'] = {'value': v, 'unit': 'candidate-evals/s', 'cores': cores, 'kind': kind, 'sample': sample}
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
