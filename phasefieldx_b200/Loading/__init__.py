from .loading_functions import *  # noqa: F401,F403
