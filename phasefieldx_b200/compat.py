"""
Drop-in aliasing: `install()` registers this package (and the dolfinx stand-ins) in
`sys.modules` under the names the reference's driver scripts import, so they run unchanged:

    import phasefieldx_b200.compat as compat; compat.install()
    import dolfinx, mpi4py, petsc4py                       # stand-ins (only if the real ones are absent)
    from phasefieldx.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx.Element.Phase_Field_Fracture.solver.solver import solve
    from phasefieldx.Boundary.boundary_conditions import bc_xy, get_ds_bound_from_marker
"""
import importlib
import sys
import types

_SUBMODULES = [
    "Boundary", "Boundary.boundary_conditions", "Loading", "Loading.loading_functions", "Logger",
    "Logger.library_versions", "Materials", "Materials.conversion", "Math", "Math.functions", "PostProcessing",
    "PostProcessing.ReferenceResult", "files", "Element", "Element.Phase_Field_Fracture",
    "Element.Phase_Field_Fracture.Input", "Element.Phase_Field_Fracture.solver",
    "Element.Phase_Field_Fracture.solver.solver",
]


def install(standins=True):
    import phasefieldx_b200 as pkg
    sys.modules.setdefault("phasefieldx", pkg)
    for name in _SUBMODULES:
        sys.modules.setdefault("phasefieldx." + name, importlib.import_module("phasefieldx_b200." + name))
    if not standins:
        return
    from . import fem, mesh
    for real in ("dolfinx", "mpi4py", "petsc4py"):
        try:
            importlib.import_module(real)
        except ImportError:
            sys.modules[real] = _standin(real, fem, mesh)
            if real == "dolfinx":
                for sub in ("mesh", "fem", "io"):
                    sys.modules[f"dolfinx.{sub}"] = getattr(sys.modules[real], sub)
                sys.modules["dolfinx.io.gmsh"] = sys.modules[real].io.gmsh


def _standin(name, fem, mesh):
    m = types.ModuleType(name)
    if name == "dolfinx":
        m.mesh = types.ModuleType("dolfinx.mesh")
        for attr in ("create_rectangle", "create_box", "CellType", "DiagonalType", "locate_entities_boundary",
                     "locate_entities", "meshtags"):
            setattr(m.mesh, attr, getattr(mesh, attr))
        m.fem = types.ModuleType("dolfinx.fem")
        for attr in ("functionspace", "locate_dofs_topological", "dirichletbc", "Constant"):
            setattr(m.fem, attr, getattr(fem, attr))
        m.io = types.ModuleType("dolfinx.io")
        m.io.gmsh = types.ModuleType("dolfinx.io.gmsh")
        m.io.gmsh.read_from_msh = mesh.read_from_msh
        m.__version__ = "stand-in (phasefieldx_b200)"
    elif name == "mpi4py":
        m.MPI = types.SimpleNamespace(COMM_WORLD=mesh.COMM_WORLD, SUM="sum")
    elif name == "petsc4py":
        import numpy as np
        m.PETSc = types.SimpleNamespace(ScalarType=np.float64)
    return m
