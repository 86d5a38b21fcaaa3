from .conversion import *  # noqa: F401,F403
