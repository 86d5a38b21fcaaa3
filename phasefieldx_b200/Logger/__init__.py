from .library_versions import *  # noqa: F401,F403
