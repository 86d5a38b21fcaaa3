import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def hostcheck():
    """Test-only host build of the element headers (tests/native/hostcheck.cpp)."""
    import ctypes
    from phasefieldx_b200 import build
    return ctypes.CDLL(build.build_hostcheck())


@pytest.fixture(scope="session")
def c1_mesh():
    """Config-1 mesh (reference examples/PhaseFieldFracture/mesh/mesh.msh) from the committed fixture."""
    from phasefieldx_b200.mesh import Mesh
    d = np.load(os.path.join(GOLDEN, "c1_mesh.npz"))
    msh = Mesh(d["points"], d["cells"], "triangle", gdim=2)
    return msh, d["bottom_nodes"], d["top_nodes"]


def perturbed_meshes(rng):
    """Small meshes of every supported cell type with perturbed interior nodes."""
    from phasefieldx_b200 import mesh as M
    out = []
    m = M.create_rectangle(None, [[0, 0], [1.0, 0.7]], [9, 7], M.CellType.triangle)
    m.geometry.x[:, :2] += rng.normal(size=(m.num_nodes, 2)) * 0.01
    out.append(("triangle", m))
    m = M.create_rectangle(None, [[0, 0], [1.0, 0.7]], [8, 6], M.CellType.quadrilateral)
    m.geometry.x[:, :2] += rng.normal(size=(m.num_nodes, 2)) * 0.01
    out.append(("quadrilateral", m))
    m = M.create_box(None, [[0, 0, 0], [1.0, 0.7, 0.5]], [5, 4, 3])
    m.geometry.x[:, :3] += rng.normal(size=(m.num_nodes, 3)) * 0.01
    out.append(("tetrahedron", m))
    m = M.create_box(None, [[0, 0, 0], [1.0, 0.7, 0.5]], [4, 3, 2], M.CellType.hexahedron)
    m.geometry.x[:, :3] += rng.normal(size=(m.num_nodes, 3)) * 0.01
    out.append(("hexahedron", m))
    return out


MODELS = [("isotropic", "no"), ("anisotropic", "spectral"), ("anisotropic", "deviatoric"), ("hybrid", "spectral")]


def rel(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return np.abs(x - y).max() / (np.abs(y).max() + 1e-300)


def rel_l2(x, y):
    x, y = np.asarray(x), np.asarray(y)
    den = np.linalg.norm(y)
    return np.linalg.norm(x - y) / den if den > 0 else np.linalg.norm(x - y)
