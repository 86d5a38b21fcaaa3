from . import solver  # noqa: F401

__all__ = ["solver"]
