"""Phase-field (crack-surface) boundary value problem (reference `src/phasefieldx/Element/Phase_Field/`):
the phase-field half of the staggered step with H = 0."""
from .Input import *  # noqa: F401,F403,E402
from . import solver  # noqa: F401,E402
