from . import solver, solver_anisotropic, solver_ener_non_variational, solver_ener_variational  # noqa: F401

__all__ = ["solver", "solver_anisotropic", "solver_ener_variational", "solver_ener_non_variational"]
