"""B200-native drop-in for the assembly / Dirichlet / solve hot path of rcapillon/thermoelasticity_fem.

Same module layout as the reference package (``materials``, ``mesh``, ``model``, ``solvers``, ``plots``),
so that ``from thermoelasticity_fem_b200.model import Model`` replaces
``from thermoelasticity_fem.model import Model``.  The numerical work runs in ``libtefem_b200.so``
(hand-written fp64 CUDA for sm_100a, C ABI in ``include/tefem.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"

from .materials import LinearThermoElastic, glass_SG773      # noqa: F401
from .mesh import Mesh                                       # noqa: F401
from .model import Model                                     # noqa: F401
from .solvers import LinearSteadyState, LinearTransient      # noqa: F401
