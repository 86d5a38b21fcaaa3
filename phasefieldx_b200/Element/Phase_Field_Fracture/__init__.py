from .Input import *  # noqa: F401,F403
from . import solver  # noqa: F401
