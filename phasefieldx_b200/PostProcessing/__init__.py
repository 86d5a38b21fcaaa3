from .ReferenceResult import *  # noqa: F401,F403
