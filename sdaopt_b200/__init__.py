"""sdaopt_b200 -- B200-native dual annealing: the hot path of sgubianpm/sdaopt on sm_100a.

Mirrors ``sdaopt/__init__.py:7-13`` of the reference: ``sda`` plus the classes exported "for unit
testing".  All objective evaluation, sampling, acceptance and local search run in the CUDA library
``libsdaopt_b200.so`` (C ABI: include/sdaopt_b200.h); there is no CPU fallback.
"""
import os as _os

# Hardware work queues per process (8 by default, 32 at most): independent batches launched on different
# streams (device.BatchPool: the suite's 250 jobs) share them, and commands of one queue issue in order.
# Must be set before the CUDA context exists; a value the user has set is kept.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._sda import sda
from ._sda import sda_batch

# For unit testing (same surface as the reference)
from ._sda import VisitingDistribution
from ._sda import SDARunner
from ._sda import ObjectiveFunWrapper
from ._sda import EnergyState

from . import functors
from .functors import (DeviceFunctor, Rastrigin, ShiftedRastrigin, Rosenbrock, Ackley01,
                       Schwefel01, Schwefel26, Exponential, SuttonChen, Constant, Sphere, Griewank)

__all__ = ['sda', 'sda_batch', 'VisitingDistribution', 'SDARunner', 'ObjectiveFunWrapper',
           'EnergyState', 'functors', 'DeviceFunctor']
