"""Linear elasticity (reference `src/phasefieldx/Element/Elasticity/`): the displacement half of the
staggered step with phi = 0."""
from .Input import *  # noqa: F401,F403,E402
from . import solver  # noqa: F401,E402
