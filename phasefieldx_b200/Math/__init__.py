from .functions import *  # noqa: F401,F403
