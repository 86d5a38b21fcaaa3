from . import solver, solver_anisotropic  # noqa: F401

__all__ = ["solver", "solver_anisotropic"]
