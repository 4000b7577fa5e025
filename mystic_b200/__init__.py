"""mystic_b200: B200-native (sm_100a) differential-evolution generation engine behind the
solver interface of mystic.differential_evolution.DifferentialEvolutionSolver2.

    from mystic_b200 import DifferentialEvolutionSolver2
    from mystic_b200.models import rosen
    from mystic_b200.strategy import Best1Bin
    from mystic_b200.termination import VTR

The CUDA library (mystic_b200/_lib/libmystic_b200.so, built by `python -m mystic_b200.build`)
is loaded when a solver first touches the device; there is no CPU fallback.
"""
from . import constraints, ensemble, models, monitors, penalty, strategy, termination  # noqa: F401
from .solver import B200DifferentialEvolutionSolver2, DifferentialEvolutionSolver2, LoadSolver, diffev2  # noqa: F401

__version__ = '0.1.0'
