"""B200-native drop-in for mystic.differential_evolution.DifferentialEvolutionSolver2.

Host mirror of the reference's solver interface (mystic/abstract_solver.py:86-1266,
mystic/abstract_map_solver.py:81-136, mystic/differential_evolution.py:436-721): the same
setters, `Step`/`Solve`/`Terminated`, the same attributes that `mystic.termination` and
`mystic.monitors` objects read, the same sticky-keyword handling and error behaviour.  What
differs is `_Step`: trial generation, bounds, cost, penalty, selection and the generation best
(differential_evolution.py:574-613) run as ONE fused CUDA kernel through the C ABI of
include/mystic_b200.h.  The host keeps what the reference keeps on the host: objective
bookkeeping, monitors, termination, callbacks.

There is no CPU path.  Costs, penalties and constraints must be device-expressible
(mystic_b200.models / .penalty / .constraints descriptors, or the mystic.models objects they
mirror); anything else raises TypeError.

Engine-specific constructor keywords (not in the reference):
    rng='philox'     production: donors / crossover masks from the Philox4x32-10 counter RNG
                     inside the kernel; SetRandomInitialPoints runs on the device.
    rng='reference'  validation: the host makes the reference's own stdlib `random` calls in the
                     reference's order (strategy.py:27,42,49) and the kernel consumes those draws,
                     so a run reproduces mystic's trajectory for the same random_seed().
    rng='replay'     validation: draws recorded from the reference are supplied per generation
                     with SetReplayDraws().
    seed=int         Philox key (default: 64 bits drawn from stdlib `random`, so random_seed() still
                     makes runs reproducible)
    device=int       CUDA device ordinal
    keep_trials=bool keep every trial vector / energy of the last generation on the device
                     (`trialSolution`, `trialEnergy`)
    distributed=bool shard the population by candidate over torch.distributed ranks, one GPU each
    tensor_cores=bool evaluate the Chebyshev fitting costs on the FP64 tensor cores (DMMA) instead of the
                     bit-faithful Horner path (default False)
"""
import os
import random
import weakref
from collections.abc import Callable as _Callable

import numpy as np

from . import constraints as _constraints
from . import models as _models
from . import monitors as _monitors
from . import penalty as _penalty
from . import strategy as _strategy
from . import termination as _termination

inf = float('inf')

BOUNDS_NONE, BOUNDS_REJECT, BOUNDS_CLAMP = 0, 1, 2


def _identity(x):
    return x


def _no_penalty(x):
    return 0.0


def _default_termination(inst, *args, **kwds):
    """the AbstractSolver default: never stops on its own (abstract_solver.py:144)"""
    if len(args) < 1 or args[0] is False or kwds.get('info', True) is False:
        return False
    return ''


def _pickler():
    """dill when it is installed (what the reference uses, abstract_solver.py:983): it can also
    serialise mystic's closure-based termination objects; plain pickle otherwise"""
    try:
        import dill
        return dill
    except ImportError:
        import pickle
        return pickle


class _DeviceObjective(object):
    """What `_cost[0]` is in the reference: the decorated objective
    (inf outside the strict range, else cost) + penalty -- evaluated on the device."""

    def __init__(self, solver):
        # weak: the solver holds this object in `_cost`; a strong reference back would make a cycle
        # and keep gigabytes of device buffers alive until the garbage collector runs
        self._solver = weakref.ref(solver)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = self._solver()._engine.eval_cost(np.atleast_2d(x))
        return float(out[0]) if x.ndim == 1 else out


class DifferentialEvolutionSolver2(object):
    """Differential Evolution (Storn & Price) with an invariant current generation; the
    generation step is a fused sm_100a kernel.  Interface of
    mystic.differential_evolution.DifferentialEvolutionSolver2."""

    _engine_factory = None  # tests inject a CPU engine here; the product default is the CUDA engine

    def __init__(self, dim, NP=4, **kwds):
        NP = max(NP, dim, 4)  # differential_evolution.py:456
        self.nDim = dim
        self.nPop = NP
        self._rng = kwds.pop('rng', 'philox')
        if self._rng not in ('philox', 'reference', 'replay'):
            raise ValueError("rng must be 'philox', 'reference' or 'replay'")
        self._seed = kwds.pop('seed', None)
        self._device = kwds.pop('device', None)
        self._keep_trials = bool(kwds.pop('keep_trials', False))
        self._distributed = bool(kwds.pop('distributed', False))
        self._tensor_cores = bool(kwds.pop('tensor_cores', False))
        engine = kwds.pop('engine', None)
        if engine is not None:
            self._engine_factory = engine

        # state of AbstractSolver.__init__ (abstract_solver.py:112-150)
        self._init_popEnergy = inf
        self._host_pop = None       # this rank's rows of the initial population (host side)
        self._host_energy = None
        self._device_init = None       # (lo, hi): generate the initial points on the device
        self._init_solution = [None]
        self._map_solver = True
        self._bestEnergy = None
        self._bestSolution = None
        self._state = None
        self._type = self.__class__.__name__
        self.sigint_callback = None
        self._handle_sigint = False
        self._useStrictRange = False
        self._useTightRange = None
        self._useClipRange = None
        self._defaultMin = [-1e3] * dim
        self._defaultMax = [1e3] * dim
        self._strictMin = []
        self._strictMax = []
        self._maxiter = None
        self._maxfun = None
        self._saveiter = None
        self._evalmon = _monitors.Null()
        self._stepmon = _monitors.Monitor()
        self._fcalls = [0]
        self._energy_history = None
        self._solution_history = None
        self.id = None
        self._strictbounds = _identity
        self._constraints = _identity
        self._penalty = _no_penalty
        self._reducer = None
        self._cost = (None, None, None)
        self._collapse = False
        self._EARLYEXIT = _termination.EARLYEXIT
        self._live = False
        self._map = None
        self._mapconfig = dict()

        # DE2 part (differential_evolution.py:458-465)
        kwds = self._preprocess_inputs(**kwds)
        self.scale = kwds['ScalingFactor'] if 'ScalingFactor' in kwds else 0.8
        self.probability = kwds['CrossProbability'] if 'CrossProbability' in kwds else 0.9
        strategy = kwds['strategy'] if 'strategy' in kwds else 'Best1Bin'
        self.strategy = strategy if isinstance(strategy, str) else getattr(strategy, '__name__', 'Best1Bin')
        self._termination = _termination.VTRChangeOverGeneration(5e-3)

        # engine state
        self._engine = None
        self._bounds_mode = BOUNDS_NONE
        self._replay_source = None
        self._device_cost = None
        self._last = None  # (n_inf, n_accepted, best_index) of the last generation
        self._resume = None  # device state to reinstall after LoadSolver / unpickling
        self._world = 1
        self._rank = 0
        self._np_local = NP
        self._offset = 0
        if self._distributed:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError('distributed=True needs an initialised torch.distributed process group')
            self._world = dist.get_world_size()
            self._rank = dist.get_rank()
            base, rem = divmod(NP, self._world)
            self._np_local = base + (1 if self._rank < rem else 0)
            self._offset = self._rank * base + min(self._rank, rem)

    # ------------------------------------------------------------------------------------
    # properties (abstract_solver.py:168-246, 1260-1265)
    # ------------------------------------------------------------------------------------
    @property
    def generations(self):
        return max(0, len(self._stepmon) - 1)

    @property
    def evaluations(self):
        return self._fcalls[0]

    @property
    def energy_history(self):
        return self._stepmon._y if self._energy_history is None else self._energy_history

    @energy_history.setter
    def energy_history(self, energy):
        self._energy_history = energy

    @property
    def solution_history(self):
        return self._stepmon.x if self._solution_history is None else self._solution_history

    @solution_history.setter
    def solution_history(self, params):
        self._solution_history = params

    @property
    def bestSolution(self):
        if self._bestSolution is None:
            return self.population[0]
        return self._bestSolution

    @bestSolution.setter
    def bestSolution(self, params):
        self._bestSolution = params

    @property
    def bestEnergy(self):
        if self._bestEnergy is None:
            return self.popEnergy[0]
        return self._bestEnergy

    @bestEnergy.setter
    def bestEnergy(self, energy):
        self._bestEnergy = energy

    @property
    def population(self):
        """current generation [nPop, nDim] (this rank's shard when distributed); copied from the device"""
        if self._engine is not None:
            return self._engine.population()
        if self._host_pop is None:
            return np.zeros((self._np_local, self.nDim))
        return self._host_pop

    @property
    def popEnergy(self):
        if self._engine is not None:
            return self._engine.energies()
        return np.full(self._np_local, inf)

    @property
    def trialSolution(self):
        if self._engine is None:
            return np.zeros((self._np_local, self.nDim))
        return self._engine.trials()[0]

    @property
    def trialEnergy(self):
        if self._engine is None:
            return np.full(self._np_local, inf)
        return self._engine.trials()[1]

    @property
    def genealogy(self):
        """The reference appends every accepted child to a host list
        (differential_evolution.py:467-475): unbounded at NP=1M, not kept."""
        return [[] for _ in range(self.nPop)]

    def UpdateGenealogyRecords(self, id, newchild):
        return

    def Solution(self):
        return self.bestSolution

    # ------------------------------------------------------------------------------------
    # setters
    # ------------------------------------------------------------------------------------
    def SetMapper(self, map, **kwds):
        """abstract_map_solver.py:125-136.  The cost is evaluated inside the fused kernel, so the
        map is recorded but never called; as with any non-default map in the reference, the
        evaluation monitor is not fed (differential_evolution.py:510-514)."""
        self._map = map
        self._mapconfig = kwds

    def SetReducer(self, reducer, arraylike=False):
        if not reducer:
            self._reducer = None
            return self._update_objective()
        raise NotImplementedError('reducers apply to vector-valued Python costs; device costs are scalar')

    def SetPenalty(self, penalty):
        """abstract_solver.py:249-266"""
        if not penalty:
            self._penalty = _no_penalty
        elif not isinstance(penalty, _Callable):
            raise TypeError("'%s' is not a callable function" % penalty)
        else:
            self._penalty = _penalty.resolve(penalty)
        return self._update_objective()

    def SetConstraints(self, constraints):
        """differential_evolution.py:477-492"""
        if not constraints:
            self._constraints = _identity
        elif not isinstance(constraints, _Callable):
            raise TypeError("'%s' is not a callable function" % constraints)
        else:
            self._constraints = _constraints.resolve(constraints)
        return self._update_objective()

    def SetGenerationMonitor(self, monitor, new=False):
        """abstract_solver.py:284-307"""
        if monitor is None or monitor is _monitors.Null or isinstance(monitor, _monitors.Null) \
                or type(monitor).__name__ == 'Null' or getattr(monitor, '__name__', '') == 'Null':
            monitor = _monitors.Monitor()  # a Null generation monitor is not allowed (:299)
        elif not (callable(monitor) and hasattr(monitor, '_x') and hasattr(monitor, '_y')):
            raise TypeError("'%s' is not a monitor instance" % monitor)
        current = None if new else self._stepmon
        if current is not None and current is not monitor and len(current):
            monitor._x[:0] = list(current._x)
            monitor._y[:0] = list(current._y)
            monitor._id[:0] = list(current._id)
            if hasattr(monitor, '_info'):
                monitor._info[:0] = list(getattr(current, '_info', []))
        self._stepmon = monitor
        self.energy_history = None
        self.solution_history = None

    def SetEvaluationMonitor(self, monitor, new=False):
        """abstract_solver.py:309-332.  Kept for interface compatibility; never fed (see SetMapper)."""
        if monitor is None:
            monitor = _monitors.Null()
        elif isinstance(monitor, type):
            monitor = monitor()
        if not callable(monitor):
            raise TypeError("'%s' is not a monitor instance" % monitor)
        self._evalmon = monitor

    def SetStrictRanges(self, min=None, max=None, **kwds):
        """abstract_solver.py:334-405.  tight=None,clip=None: trials outside [min,max] get energy
        inf (tools.py:420-426).  tight=True: trials are clamped to the bounds before the cost
        (symbolic.py:1139-1144).  clip=True: clamp (what impose_bounds(clip=True) intends).
        clip=False (random re-mapping into the bounds) is not device-expressible."""
        tight = kwds['tight'] if 'tight' in kwds else None
        clip = kwds['clip'] if 'clip' in kwds else None
        if clip is None:
            mode = BOUNDS_CLAMP if tight else BOUNDS_REJECT
        elif repr(tight) == 'False':
            raise ValueError('can not specify clip when tight is False')
        elif clip:
            mode = BOUNDS_CLAMP
        else:
            raise NotImplementedError('clip=False (impose_bounds re-mapping) is not available on the device')
        self._useTightRange = tight
        self._useClipRange = clip
        if repr(min) == 'False' or repr(max) == 'False':
            self._useStrictRange = False
            self._bounds_mode = BOUNDS_NONE
            return self._update_objective()
        if min is None:
            min = self._defaultMin
        if max is None:
            max = self._defaultMax
        min = [self._defaultMin[0] if v is None else v for v in min]
        max = [self._defaultMax[0] if v is None else v for v in max]
        min = np.asarray(min, dtype=np.float64)
        max = np.asarray(max, dtype=np.float64)
        if min.shape != max.shape or np.any(min > max):
            raise ValueError('each min[i] must be <= the corresponding max[i]')
        if len(min) != self.nDim:
            raise ValueError('bounds array must be length %s' % self.nDim)
        self._useStrictRange = True
        self._strictMin = min
        self._strictMax = max
        self._bounds_mode = mode
        return self._update_objective()

    def _release_engine_population(self):
        """a new initial population invalidates the device state"""
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        self._live = False

    def SetInitialPoints(self, x0, radius=0.05):
        """abstract_solver.py:471-506"""
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.ndim == 0:
            x0 = x0.reshape(1)
        if x0.ndim != 1:
            raise ValueError('Initial guess must be a scalar or rank-1 sequence.')
        if len(x0) != self.nDim:
            raise ValueError('Initial guess must be length %s' % self.nDim)
        radius = np.asarray(radius, dtype=np.float64)
        if radius.ndim and len(radius) != self.nDim:
            raise ValueError('Radius must be length %s' % self.nDim)
        lo = x0 * (1 - radius)
        hi = x0 * (1 + radius)
        lo[lo == 0] = -radius[lo == 0] if radius.ndim else -radius
        hi[hi == 0] = radius[hi == 0] if radius.ndim else radius
        self.SetRandomInitialPoints(lo, hi, _x0=x0)

    def SetRandomInitialPoints(self, min=None, max=None, _x0=None):
        """abstract_solver.py:508-535.  rng='reference': the reference's own
        random.uniform(min[j], max[j]) calls in row-major order.  rng='philox': generated on the
        device when the solver starts (Philox stream 2)."""
        if min is None:
            min = self._defaultMin
        if max is None:
            max = self._defaultMax
        if len(min) != self.nDim or len(max) != self.nDim:
            raise ValueError('bounds array must be length %s' % self.nDim)
        min = [self._defaultMin[0] if v is None else v for v in min]
        max = [self._defaultMax[0] if v is None else v for v in max]
        self._release_engine_population()
        if self._rng == 'philox':
            self._device_init = (np.asarray(min, dtype=np.float64), np.asarray(max, dtype=np.float64),
                                 None if _x0 is None else np.asarray(_x0, dtype=np.float64))
            self._host_pop = None
        else:
            pop = np.empty((self.nPop, self.nDim), dtype=np.float64)
            uniform = random.uniform
            for i in range(self.nPop):
                for j in range(self.nDim):
                    pop[i, j] = uniform(min[j], max[j])
            if _x0 is not None:
                pop[0, :] = _x0
            self._device_init = None
            self._host_pop = pop[self._offset:self._offset + self._np_local]
        self._init_solution = [None if _x0 is None else np.array(_x0)]

    def SetPopulation(self, population, local=False):
        """explicit initial population [nPop, nDim] (all rng modes); with local=True the array is this
        rank's shard [np_local, nDim] (distributed).  A C-contiguous float64 array is used as is
        (a pinned buffer is uploaded by DMA without a staging copy)."""
        pop = np.ascontiguousarray(population, dtype=np.float64)
        rows = self._np_local if local else self.nPop
        if pop.shape != (rows, self.nDim):
            raise ValueError('population must be %d x %d' % (rows, self.nDim))
        self._release_engine_population()
        self._device_init = None
        self._host_pop = pop if local else pop[self._offset:self._offset + self._np_local]

    def SetMultinormalInitialPoints(self, mean, var=None):
        """abstract_solver.py:537-568 (host numpy generator, uploaded)"""
        mean = np.asarray(mean, dtype=np.float64)
        assert len(mean) == self.nDim
        if var is None:
            var = np.eye(self.nDim)
        elif np.ndim(var) == 0:
            var = float(var) * np.eye(self.nDim)
        self.SetPopulation(np.random.multivariate_normal(mean, var, size=self.nPop))

    def SetSampledInitialPoints(self, dist=None):
        """abstract_solver.py:570-590"""
        if dist is None:
            pop = np.random.random_sample((self.nPop, self.nDim))
        else:
            pop = np.array([dist(self.nDim) for _ in range(self.nPop)], dtype=np.float64)
        self.SetPopulation(pop)

    def SetReplayDraws(self, draws):
        """rng='replay': an iterable yielding, per generation, (donors [NP,k], nstart [NP],
        uniforms [NP, nDim+1] with NaN for draws not made) as recorded from the reference."""
        self._replay_source = iter(draws)

    def enable_signal_handler(self):
        self._handle_sigint = True

    def disable_signal_handler(self):
        self._handle_sigint = False

    def SetSaveFrequency(self, generations=None, filename=None):
        """abstract_solver.py:612-624"""
        self._saveiter = generations
        self._state = filename

    def SetEvaluationLimits(self, generations=None, evaluations=None, new=False, **kwds):
        """abstract_solver.py:626-648"""
        self._maxiter = kwds['maxiter'] if 'maxiter' in kwds else generations
        self._maxfun = kwds['maxfun'] if 'maxfun' in kwds else evaluations
        if new:
            if generations is not None:
                self._maxiter += self.generations
            else:
                self._maxiter = '*'
            if evaluations is not None:
                self._maxfun += self.evaluations
            else:
                self._maxfun = '*'

    def _SetEvaluationLimits(self, iterscale=None, evalscale=None):
        """abstract_solver.py:650-674: defaults nDim*nPop*10 iterations, nDim*nPop*1000 evaluations"""
        iterscale = 10 if iterscale is None else iterscale
        evalscale = 1000 if evalscale is None else evalscale
        N = self.nDim
        if self._maxiter is None:
            self._maxiter = N * self.nPop * iterscale
        elif self._maxiter == '*':
            self._maxiter = (N * self.nPop * iterscale) + self.generations
        if self._maxfun is None:
            self._maxfun = N * self.nPop * evalscale
        elif self._maxfun == '*':
            self._maxfun = (N * self.nPop * evalscale) + self.evaluations

    def SetTermination(self, termination):
        """abstract_solver.py:727-748"""
        self._termination = termination
        self._collapse = False
        if termination is None:
            self._termination = _default_termination

    def SetObjective(self, cost, ExtraArgs=None):
        """abstract_solver.py:750-780: remember the raw cost; decoration is delayed"""
        _cost, _raw, _args = self._cost
        if (cost is None or cost is _raw or cost is _cost) and (ExtraArgs is None or ExtraArgs is _args):
            return
        if cost is None:
            cost = _raw
        args = _args if ExtraArgs is None else ExtraArgs
        args = () if args is None else args
        if not isinstance(cost, _DeviceObjective):
            dev = _models.resolve(cost)   # TypeError for host callables
            dev.params(tuple(args))       # validates ExtraArgs
        self._cost = (None, cost, ExtraArgs)
        self._live = False

    # ------------------------------------------------------------------------------------
    # termination (abstract_solver.py:676-725)
    # ------------------------------------------------------------------------------------
    def Terminated(self, disp=False, info=False, termination=None, **kwds):
        if termination is None:
            termination = self._termination
        self._SetEvaluationLimits()
        msg = termination(self, info=True)
        sig = 'SolverInterrupt with %s' % {}
        lim = 'EvaluationLimits with %s' % {'evaluations': self._maxfun, 'generations': self._maxiter}
        if self._maxfun is not None and self._fcalls[0] >= self._maxfun:
            msg = lim
            if disp:
                print('Warning: Maximum number of function evaluations has been exceeded.')
        elif self._maxiter is not None and self.generations >= self._maxiter:
            msg = lim
            if disp:
                print('Warning: Maximum number of iterations has been exceeded')
        elif self._EARLYEXIT:
            msg = sig
            if disp:
                print('Warning: Optimization terminated with signal interrupt.')
        elif msg and disp:
            print('Optimization terminated successfully.')
            print('         Current function value: %s' % self.bestEnergy)
            print('         Iterations: %d' % self.generations)
            print('         Function evaluations: %d' % self._fcalls[0])
        if info:
            return msg
        return bool(msg)

    # ------------------------------------------------------------------------------------
    # objective decoration = engine configuration
    # ------------------------------------------------------------------------------------
    def _update_objective(self):
        """abstract_solver.py:894-901: delay until the next bootstrap"""
        self.Finalize()

    def Finalize(self):
        self._live = False

    # ---- collapse (abstract_solver.py:783-892): the collapse conditions of mystic.termination
    # (CollapseAt, CollapseAs, ...) read the whole solution history on the host and rewrite the
    # constraints; they are outside the data-parallel path, so a device solver
    # never reports a collapse -- Solve() of the reference keeps going exactly like that when no
    # collapse condition is installed
    def Collapsed(self, disp=False, info=False):
        return {} if info else False

    def Collapse(self, disp=False):
        return {}

    def _is_new(self):
        'determine if solver has been run or not (abstract_solver.py:1255-1257)'
        return bool(self.evaluations) or bool(self.generations)

    def _make_engine(self):
        if self._engine_factory is not None:
            return self._engine_factory(self._np_local, self.nDim, self.nPop, self._offset,
                                        device=self._device, keep_trials=self._keep_trials)
        from .engine import DeviceEngine  # raises without the CUDA library / a GPU: no fallback
        return DeviceEngine(self._np_local, self.nDim, self.nPop, self._offset,
                            device=self._device, keep_trials=self._keep_trials)

    def _philox_seed(self):
        if self._seed is None:
            self._seed = random.getrandbits(64)
            if self._distributed and self._world > 1:
                # one key for all shards: the draws depend on the global candidate index only, whatever each
                # rank's private stdlib generator holds (rank 0's draw wins)
                import torch.distributed as dist
                box = [self._seed]
                dist.broadcast_object_list(box, src=0)
                self._seed = int(box[0])
        return self._seed

    def _effective_bounds(self):
        """strict range combined with a clamp constraint -> (mode, lo, hi) for the device"""
        mode = self._bounds_mode if self._useStrictRange else BOUNDS_NONE
        lo = np.asarray(self._strictMin, dtype=np.float64) if self._useStrictRange else None
        hi = np.asarray(self._strictMax, dtype=np.float64) if self._useStrictRange else None
        con = self._constraints
        if con is _identity:
            return mode, lo, hi
        # a BoundsClamp constraint (mystic_b200.constraints.boundsconstrain)
        if mode == BOUNDS_NONE:
            raise NotImplementedError('a clamp constraint needs SetStrictRanges with the same bounds on the device '
                                      '(the kernel carries one bounds box)')
        if not (np.array_equal(con.lo, lo) and np.array_equal(con.hi, hi)):
            raise NotImplementedError('clamp constraint and strict range must use the same bounds on the device')
        return BOUNDS_CLAMP, lo, hi

    def _decorate_objective(self, cost, ExtraArgs=None):
        """differential_evolution.py:494-529: bind bounds, penalty and the cost.  Here that means
        configuring the device engine; the returned object evaluates the decorated objective on
        the device."""
        raw = cost
        dev = _models.resolve(cost)
        args = () if ExtraArgs is None else tuple(ExtraArgs)
        fresh = self._engine is None
        if fresh:
            self._engine = self._make_engine()
            if self._device_init is not None:
                lo, hi, x0 = self._device_init
                self._engine.init_population(self._philox_seed(), lo, hi)
                if x0 is not None and self._offset == 0:
                    self._engine.population_tensor()[0, :] = self._engine.population_tensor().new_tensor(x0)
            else:
                self._engine.upload_population(self._host_pop)
            if self._distributed and self._world > 1 and hasattr(self._engine, 'connect_peers') \
                    and os.environ.get('MB2_EXCHANGE', 'peer') != 'nccl':
                # the shards' bests travel as stores into peer memory (NVLink), fused into the argmin
                # epilogue of every generation; MB2_EXCHANGE=nccl keeps the all-gather of _exchange_best
                import torch.distributed as dist

                import torch

                def gather(obj):
                    out = [None] * self._world
                    dist.all_gather_object(out, obj)
                    return out

                def agree(flag):
                    dev = self._engine.device if dist.get_backend() == 'nccl' else 'cpu'
                    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    return bool(t.item())
                self._engine.connect_peers(self._world, self._rank, gather, agree)
            if self._resume is not None:
                # a solver loaded from a restart file: energies, best member and generation counter
                # are installed as saved, nothing is re-evaluated
                self._engine.restore_state(**self._resume)
                self._resume = None
        eng = self._engine
        mode, lo, hi = self._effective_bounds()
        if mode != BOUNDS_NONE and not fresh and self.generations:
            self._reclip_population(lo, hi)
        eng.set_bounds(mode, lo, hi)
        eng.set_cost(dev.device_id(self._tensor_cores, self.nDim), dev.params(args))
        if self._penalty is _no_penalty:
            eng.set_penalty()
        else:
            eng.set_penalty(*self._penalty.device_terms(self.nDim))
        if self._rng == 'philox':
            eng.set_rng_philox(self._philox_seed())
        self._device_cost = dev
        self._cost = (_DeviceObjective(self), raw, ExtraArgs)
        self._live = True
        return self._cost[0]

    def _reclip_population(self, lo, hi):
        """Re-decoration after generation 0 (differential_evolution.py:516-520 with ngen > 0): the
        best member is clipped at the bounds; in the other members out-of-range coordinates are
        replaced by uniform(min,max) scaled by ONE draw per member
        (abstract_solver.py:450-469).  One-off host operation on the current population."""
        pop_t = self._engine.population_tensor()
        pop = pop_t.cpu().numpy()
        en = self._engine.energies()
        indx = int(np.argmin(en)) if len(en) else 0
        if self._distributed and self._world > 1:
            # the best member is the GLOBAL one of the last exchange; it lives in one shard only
            gbest = self._last[2] if self._last else -1
            indx = gbest - self._offset if self._offset <= gbest < self._offset + pop.shape[0] else -1
        for i in range(pop.shape[0]):
            clipped = pop[i].clip(lo, hi)
            if i == indx:
                pop[i] = clipped
            else:
                bad = clipped != pop[i]
                if bad.any():
                    draw = lo + (hi - lo) * random.random()
                    pop[i][bad] = draw[bad]
        pop_t.copy_(pop_t.new_tensor(pop))

    def _bootstrap_objective(self, cost=None, ExtraArgs=None):
        """abstract_solver.py:939-953"""
        _cost, _raw, _args = self._cost
        if (cost is None or cost is _raw or cost is _cost) and \
           (ExtraArgs is None or ExtraArgs is _args) and self._live:
            return _cost
        self.SetObjective(cost, ExtraArgs)
        if self._cost[1] is None:
            raise TypeError('no cost function was provided')
        return self._decorate_objective(*self._cost[1:])

    # ------------------------------------------------------------------------------------
    # keyword handling (abstract_solver.py:1034-1096, differential_evolution.py:627-693)
    # ------------------------------------------------------------------------------------
    def _preprocess_inputs(self, **kwds):
        if 'evalmon' in kwds:
            if 'EvaluationMonitor' in kwds:
                raise KeyError('cannot specify more than one of %s' % (('EvaluationMonitor', 'evalmon'),))
            kwds['EvaluationMonitor'] = kwds.pop('evalmon')
        if 'itermon' in kwds:
            if 'StepMonitor' in kwds:
                raise KeyError('cannot specify more than one of %s' % (('StepMonitor', 'stepmon', 'itermon'),))
            kwds['StepMonitor'] = kwds.pop('itermon')
        if 'stepmon' in kwds:
            if 'StepMonitor' in kwds:
                raise KeyError('cannot specify more than one of %s' % (('StepMonitor', 'stepmon', 'itermon'),))
            kwds['StepMonitor'] = kwds.pop('stepmon')
        if 'cross' in kwds:
            if 'CrossProbability' in kwds:
                raise KeyError('cannot specify more than one of %s' % (('CrossProbability', 'cross'),))
            kwds['CrossProbability'] = kwds.pop('cross')
        if 'scale' in kwds:
            if 'ScalingFactor' in kwds:
                raise KeyError('cannot specify more than one of %s' % (('ScalingFactor', 'scale'),))
            kwds['ScalingFactor'] = kwds.pop('scale')
        return kwds

    def _process_inputs(self, kwds):
        """Sticky: monitors, penalty, constraints, strategy, CrossProbability, ScalingFactor;
        one-shot: callback, disp.  This mirrors the reference's data flow exactly, including its
        consequence that `CrossProbability=`/`ScalingFactor=` handed to Solve() are overwritten on
        the first Step by the stale 'cross'/'scale' entries that Solve() forwards in `settings`
        (differential_evolution.py:676-692, abstract_solver.py:1171), whereas the `cross=`/`scale=`
        aliases, the constructor and direct Step() keywords do take effect.  Seeded reference
        trajectories (tests/golden/c1_rosen3_best1exp.npz) depend on it."""
        pre = self._preprocess_inputs(**kwds)
        settings = {'callback': None, 'disp': 0}
        settings.update((i, j) for (i, j) in pre.items() if i in settings)
        if 'EvaluationMonitor' in pre:
            self.SetEvaluationMonitor(pre['EvaluationMonitor'])
        if 'StepMonitor' in pre:
            self.SetGenerationMonitor(pre['StepMonitor'])
        if 'penalty' in pre:
            self.SetPenalty(pre['penalty'])
        if 'constraints' in pre:
            self.SetConstraints(pre['constraints'])
        strategy = getattr(_strategy, self.strategy, _strategy.Best1Bin)
        settings.update({'strategy': strategy, 'cross': self.probability, 'scale': self.scale})
        settings.update((i, j) for (i, j) in kwds.items() if i in settings)
        self.probability = kwds['CrossProbability'] if 'CrossProbability' in kwds else settings['cross']
        self.scale = kwds['ScalingFactor'] if 'ScalingFactor' in kwds else settings['scale']
        self.strategy = getattr(settings['strategy'], '__name__', 'Best1Bin')
        return settings

    # ------------------------------------------------------------------------------------
    # the generation step
    # ------------------------------------------------------------------------------------
    def _reference_draws(self, name):
        """rng='reference': make the stdlib calls the reference's strategy would make, in its order
        (strategy.py:27 sample over [0,NP) minus the candidate; :42 randrange(nDim); :49/:80
        random() per crossover decision), and pack them for the kernel."""
        (sid, k, _), name = _strategy.strategy_id(name)
        NP, D = self.nPop, self.nDim
        binomial = (name == 'Best1Bin')
        donors = np.full((NP, 5), -1, dtype=np.int32)
        nstart = np.zeros(NP, dtype=np.int32)
        uni = np.full((NP, D + 1), np.nan, dtype=np.float64)
        sample, randrange, rnd = random.sample, random.randrange, random.random
        cr = self.probability
        for c in range(NP):
            donors[c, :k] = sample(list(range(c)) + list(range(c + 1, NP)), k)
            nstart[c] = randrange(D)
            if binomial:
                for i in range(D):
                    uni[c, i] = rnd()
            else:
                i = 0
                while True:
                    u = rnd()
                    uni[c, i] = u
                    if u >= cr or i == D:
                        break
                    i += 1
        return donors, nstart, uni

    def _exchange_best(self):
        """distributed: all-gather the shards' (energy, global index, counters, bestSolution) records
        and let every shard install the global best (lowest energy, ties to the lowest index)."""
        if getattr(self._engine, 'peer_exchange', False):
            return  # already done by the generation's own epilogue kernel (engine.connect_peers)
        import torch
        import torch.distributed as dist
        rec = self._engine.pack_best()
        gathered = torch.empty((self._world * rec.shape[0],), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(gathered, rec)   # flat output: accepted by both NCCL and gloo
        self._engine.merge_best(gathered.view(self._world, rec.shape[0]))

    def _Step(self, cost=None, ExtraArgs=None, **kwds):
        """One generation (differential_evolution.py:531-625)."""
        settings = self._process_inputs(kwds)
        callback = settings['callback'] if 'callback' in settings else None
        strategy = settings['strategy'] if 'strategy' in settings else self.strategy
        self._bootstrap_objective(cost, ExtraArgs)
        eng = self._engine

        init = not len(self._stepmon)
        if init:
            eng.eval_initial()                       # generation 0 (:559-567, strategy=None)
        else:
            (sid, k, _), name = _strategy.strategy_id(strategy)
            if self._np_local < k + 1:
                raise ValueError('Sample larger than population or is negative')  # what random.sample raises
            eng.set_strategy(sid, self.scale, self.probability)
            if self._rng == 'reference':
                if self._distributed:
                    raise NotImplementedError("rng='reference' is a single-GPU validation mode")
                eng.set_rng_replay(*self._reference_draws(name))
            elif self._rng == 'replay':
                if self._replay_source is None:
                    raise RuntimeError("rng='replay' needs SetReplayDraws()")
                eng.set_rng_replay(*next(self._replay_source))
            eng.step()                               # K2 fused kernel + K3 argmin
        if self._distributed and self._world > 1:
            self._exchange_best()
        best_e, best_idx, n_inf, n_acc, _gen, best_x = eng.read_result()

        # evaluation accounting: the evaluation monitor is never fed, so the reference's
        # "remove skipped evaluations" rule applies (:601)
        self._fcalls[0] += self.nPop - int(n_inf)
        self._last = (int(n_inf), int(n_acc), int(best_idx))
        self.bestEnergy = best_e
        self.bestSolution = best_x
        self._stepmon(best_x[:], best_e, self.id)     # :617
        self._save_state()
        if callback is not None:
            callback(self.bestSolution)
        if init:
            self._termination(self)                   # :624
        return

    def Step(self, cost=None, termination=None, ExtraArgs=None, **kwds):
        """abstract_solver.py:1098-1149"""
        disp = bool(kwds['disp']) if 'disp' in kwds else False
        self._bootstrap_objective(cost, ExtraArgs)
        if termination is not None:
            self.SetTermination(termination)
        if len(self._stepmon):
            msg = self.Terminated(disp=disp, info=True) or None
        else:
            msg = None
        if msg is None:
            self._Step(**kwds)
            if self.Terminated():
                self.Finalize()
            msg = self.Terminated(disp=disp, info=True) or None
            if msg:
                self._stepmon.info('STOP("%s")' % msg)
                self._save_state(force=True)
        return msg

    def _Solve(self, cost, ExtraArgs, **settings):
        """abstract_solver.py:1151-1180 (collapse machinery is out of scope).  When nothing needs the
        host between generations (no callback, no restart files, device-side RNG, a termination the
        device can evaluate) generations run in batches with the termination test inside the
        argmin epilogue; the monitor and the counters are then replayed from the batch log, and the
        solver stops exactly where step-by-step execution would."""
        stop = False
        while not stop:
            desc = self._batch_descriptor(settings)
            if desc is None:
                stop = self.Step(**settings)
            else:
                stop = self._run_batch(desc, settings)
        return

    _batch_generations = 64

    def _batch_descriptor(self, settings):
        if not self._batch_generations or not len(self._stepmon) or not self._live:
            return None                       # generation 0 and (re)decoration go through Step()
        if self._rng != 'philox' or self._saveiter or self._EARLYEXIT:
            return None
        if self._distributed and self._world > 1 and not getattr(self._engine, 'peer_exchange', False):
            return None                       # the all-gather exchange is a host call per generation
        if settings.get('callback') is not None or settings.get('disp'):
            return None
        if not hasattr(self._engine, 'run'):
            return None
        return _termination.device_descriptor(self._termination)

    def _run_batch(self, desc, settings):
        """Step() for up to `_batch_generations` generations without a host round trip."""
        msg = self.Terminated(info=True) or None          # Step() checks before stepping (:1134)
        if msg is None:
            step_settings = self._process_inputs(settings)
            (sid, k, _), name = _strategy.strategy_id(step_settings['strategy'])
            if self._np_local < k + 1:
                raise ValueError('Sample larger than population or is negative')
            self._engine.set_strategy(sid, self.scale, self.probability)
            self._SetEvaluationLimits()
            records, _stopped = self._engine.run(self._batch_generations, desc, list(self.energy_history),
                                                 self.generations, self._fcalls[0], self._maxiter, self._maxfun)
            D = self.nDim
            for rec in records:
                self._fcalls[0] += self.nPop - int(rec[2])
                self._last = (int(rec[2]), int(rec[3]), int(rec[1]))
                self.bestEnergy = float(rec[0])
                self.bestSolution = rec[4:4 + D].copy()
                self._stepmon(self.bestSolution[:], self.bestEnergy, self.id)
            if self.Terminated():
                self.Finalize()
            msg = self.Terminated(info=True) or None
            if msg:
                self._stepmon.info('STOP("%s")' % msg)
                self._save_state(force=True)
        return msg

    def Solve(self, cost=None, termination=None, ExtraArgs=None, **kwds):
        """differential_evolution.py:695-721 -> abstract_solver.py:1182-1229"""
        if 'sigint_callback' in kwds:
            self.sigint_callback = kwds.pop('sigint_callback')
        else:
            self.sigint_callback = None
        settings = self._process_inputs(kwds)
        self._EARLYEXIT = False
        self._bootstrap_objective(cost, ExtraArgs)
        if termination is not None:
            self.SetTermination(termination)
        self._Solve(cost, ExtraArgs, **settings)
        return

    # ------------------------------------------------------------------------------------
    # state saving (abstract_solver.py:975-1018)
    # ------------------------------------------------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        eng = state.pop('_engine', None)
        state['_engine'] = None
        state['_live'] = False
        if eng is not None:
            # device buffers -> host arrays; the engine is rebuilt lazily by the next Step()
            state['_host_pop'] = eng.population()
            state['_device_init'] = None
            if len(self._stepmon):
                state['_resume'] = dict(energies=eng.energies(), best_energy=float(self.bestEnergy),
                                        best_index=(self._last[2] if self._last else -1),
                                        best_x=np.array(self.bestSolution, dtype=np.float64),
                                        generation=len(self._stepmon) - 1)
        state['_engine_factory'] = None if '_engine_factory' not in self.__dict__ else self.__dict__['_engine_factory']
        state['_replay_source'] = None
        cost = state.get('_cost', (None, None, None))
        state['_cost'] = (None, cost[1], cost[2])
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self.__dict__.get('_engine_factory') is None:
            self.__dict__.pop('_engine_factory', None)

    def SaveSolver(self, filename=None, **kwds):
        pickle = _pickler()
        if filename is None:
            if self._state is None:
                import os
                import tempfile
                fd, self._state = tempfile.mkstemp(suffix='.pkl')
                os.close(fd)
            filename = self._state
        self._state = filename
        with open(filename, 'wb') as f:
            pickle.dump(self, f, **kwds)
        self._stepmon.info('DUMPED("%s")' % filename)

    def _save_state(self, force=False):
        if force and bool(self._state):
            self.SaveSolver()
            return
        iters = self._saveiter
        if bool(iters) and not bool(self.generations % iters):
            self.SaveSolver()


B200DifferentialEvolutionSolver2 = DifferentialEvolutionSolver2


def LoadSolver(filename=None, **kwds):
    """mystic/solvers.py:90-125: load a solver saved with SaveSolver / SetSaveFrequency"""
    with open(filename, 'rb') as f:
        solver = _pickler().load(f)
    solver.__dict__.update(kwds)
    return solver


def diffev2(cost, x0, npop=4, args=(), bounds=None, ftol=5e-3, gtol=None, maxiter=None, maxfun=None, cross=0.9,
            scale=0.8, full_output=0, disp=1, retall=0, callback=None, **kwds):
    """The one-call interface of mystic.differential_evolution.diffev2 (differential_evolution.py:721-933,
    scipy.optimize.fmin style) on the B200 engine.  Same keywords (handler, id, strategy, itermon,
    evalmon, constraints, penalty, tightrange, cliprange) and the same return value
    `xopt | (xopt, fopt, iter, funcalls, warnflag[, allvecs])`; `map` is accepted and ignored (the
    whole population is evaluated by one kernel).  Engine options pass through: rng, seed, device,
    distributed, tensor_cores, engine.

    Like the reference it draws the initial population from `bounds` when bounds are given and from
    around `x0` otherwise, and hands `cross` / `scale` to Solve() (where the reference's own quirk
    lets the constructor defaults win, differential_evolution.py:676-692)."""
    import numpy as _np
    strategy = kwds['strategy'] if 'strategy' in kwds else _strategy.Best1Bin
    stepmon = kwds['itermon'] if 'itermon' in kwds else _monitors.Monitor()
    evalmon = kwds['evalmon'] if 'evalmon' in kwds else _monitors.Monitor()
    if gtol:   # a number of generations: ChangeOverGeneration, else the default VTRChangeOverGeneration
        termination = _termination.ChangeOverGeneration(ftol, gtol)
    else:
        termination = _termination.VTRChangeOverGeneration(ftol)
    engine_opts = dict((k, kwds[k]) for k in ('rng', 'seed', 'device', 'distributed', 'tensor_cores', 'engine', 'keep_trials')
                       if k in kwds)
    solver = DifferentialEvolutionSolver2(len(x0), npop, **engine_opts)
    solver.SetEvaluationLimits(maxiter, maxfun)
    solver.SetEvaluationMonitor(evalmon)
    solver.SetGenerationMonitor(stepmon)
    if 'id' in kwds:
        solver.id = int(kwds['id'])
    if 'penalty' in kwds:
        solver.SetPenalty(kwds['penalty'])
    if 'constraints' in kwds:
        solver.SetConstraints(kwds['constraints'])
    if bounds is not None:
        minb, maxb = [list(v) for v in _np.asarray(bounds, dtype=float).T]
        solver.SetStrictRanges(minb, maxb, tight=kwds.get('tightrange'), clip=kwds.get('cliprange'))
        solver.SetRandomInitialPoints(minb, maxb)
    else:
        solver.SetInitialPoints(x0)
    if kwds.get('handler'):
        solver.enable_signal_handler()
    solver.Solve(cost, termination=termination, strategy=strategy, CrossProbability=cross, ScalingFactor=scale,
                 ExtraArgs=args, callback=callback)
    x = solver.bestSolution
    fval = solver.bestEnergy
    warnflag = 0
    fcalls, iterations = solver.evaluations, solver.generations
    if fcalls >= solver._maxfun:
        warnflag = 1
        if disp:
            print('Warning: Maximum number of function evaluations has been exceeded.')
    elif iterations >= solver._maxiter:
        warnflag = 2
        if disp:
            print('Warning: Maximum number of iterations has been exceeded')
    elif disp:
        print('Optimization terminated successfully.')
        print('         Current function value: %s' % fval)
        print('         Iterations: %d' % iterations)
        print('         Function evaluations: %d' % fcalls)
    if full_output:
        out = (x, fval, iterations, fcalls, warnflag)
        return out + (stepmon.x,) if retall else out
    return (x, stepmon.x) if retall else x
