"""
phasefieldx_b200 — B200-native implementation of PhaseFieldX's staggered phase-field fracture
hot path behind the reference's own driver API (reference package root
`src/phasefieldx/__init__.py:3-39`).

Sub-packages mirror the reference layout for the path only:
`Element.Phase_Field_Fracture.{Input, solver.solver.solve}`, `Boundary`, `Loading`, `Logger`,
`Materials`, `Math`, `PostProcessing`, `files`; plus `mesh`/`fem` (dolfinx stand-ins),
`_lib` (ctypes binding of libpfx_b200.so) and `workloads` (synthetic BASELINE configs).
`phasefieldx_b200.compat.install()` aliases the package as `phasefieldx` so unmodified driver
scripts (`from phasefieldx.Element.Phase_Field_Fracture.solver.solver import solve`) run.
"""
__version__ = "0.3.0+b200.1"

from . import mesh, fem  # noqa: F401,E402
from .Boundary import *  # noqa: F401,F403,E402
from .Loading import *  # noqa: F401,F403,E402
from .Logger import *  # noqa: F401,F403,E402
from .Materials import *  # noqa: F401,F403,E402
from .Math import *  # noqa: F401,F403,E402
from .PostProcessing import *  # noqa: F401,F403,E402
from .files import *  # noqa: F401,F403,E402
from . import Element  # noqa: F401,E402
