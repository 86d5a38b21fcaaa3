from .boundary_conditions import *  # noqa: F401,F403
