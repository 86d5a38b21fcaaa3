"""
ORACLE (test infrastructure, not product code) — finite-element tables.

CPU restatement of the FE facts the reference obtains from dolfinx/basix/FFCx (not vendored
under /root/reference; pinned `fenics-dolfinx==0.10.0`, reference `environment.yml:5`):
P1 Lagrange on affine triangles/tetrahedra, Q1 on bilinear quadrilaterals and trilinear hexahedra (tensor vertex
order), Gauss quadrature exact for the polynomial integrands of the staggered solver
(SURVEY §7 H5).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import this package.
"""
import numpy as np
from scipy.special import roots_jacobi

NV = {"triangle": 3, "quadrilateral": 4, "tetrahedron": 4, "hexahedron": 8}
TDIM = {"triangle": 2, "quadrilateral": 2, "tetrahedron": 3, "hexahedron": 3}


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def quadrature(cell_type, degree):
    """Points [Q,tdim] and weights [Q] on the reference cell, exact to `degree`.
    Simplices: collapsed Gauss-Jacobi (Stroud conical product); quads: tensor Gauss."""
    n = degree // 2 + 1
    if cell_type == "quadrilateral":
        x, w = _gauss01(n)
        X, Y = np.meshgrid(x, x, indexing="ij")
        W = np.outer(w, w)
        return np.stack([X.ravel(), Y.ravel()], 1), W.ravel()
    if cell_type == "hexahedron":
        x, w = _gauss01(n)
        X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
        W = w[:, None, None] * w[None, :, None] * w[None, None, :]
        return np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1), W.ravel()
    if cell_type == "triangle":
        x0, w0 = _gauss01(n)
        x1, w1 = roots_jacobi(n, 1, 0)
        x1, w1 = 0.5 * (x1 + 1.0), w1 / 4.0
        pts, wts = [], []
        for a, wa in zip(x1, w1):
            for b, wb in zip(x0, w0):
                pts.append([a, b * (1 - a)])
                wts.append(wa * wb)
        return np.array(pts), np.array(wts)
    if cell_type == "tetrahedron":
        x0, w0 = _gauss01(n)
        x1, w1 = roots_jacobi(n, 1, 0)
        x1, w1 = 0.5 * (x1 + 1.0), w1 / 4.0
        x2, w2 = roots_jacobi(n, 2, 0)
        x2, w2 = 0.5 * (x2 + 1.0), w2 / 8.0
        pts, wts = [], []
        for a, wa in zip(x2, w2):
            for b, wb in zip(x1, w1):
                for c, wc in zip(x0, w0):
                    pts.append([a, b * (1 - a), c * (1 - a) * (1 - b)])
                    wts.append(wa * wb * wc)
        return np.array(pts), np.array(wts)
    raise ValueError(cell_type)


def tabulate(cell_type, pts):
    """Basis values N [Q,nv] and reference gradients dN [Q,nv,tdim]."""
    q = len(pts)
    if cell_type == "triangle":
        x, y = pts[:, 0], pts[:, 1]
        N = np.stack([1 - x - y, x, y], 1)
        dN = np.tile(np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]), (q, 1, 1))
    elif cell_type == "tetrahedron":
        x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
        N = np.stack([1 - x - y - z, x, y, z], 1)
        dN = np.tile(np.array([[-1.0, -1, -1], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]), (q, 1, 1))
    elif cell_type == "quadrilateral":
        x, y = pts[:, 0], pts[:, 1]
        N = np.stack([(1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y], 1)
        dN = np.stack([np.stack([-(1 - y), -(1 - x)], 1), np.stack([(1 - y), -x], 1),
                       np.stack([-y, (1 - x)], 1), np.stack([y, x], 1)], 1)
    elif cell_type == "hexahedron":
        x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
        lx, ly, lz = [1 - x, x], [1 - y, y], [1 - z, z]
        dl = [-np.ones(q), np.ones(q)]
        N = np.stack([lx[a & 1] * ly[(a >> 1) & 1] * lz[(a >> 2) & 1] for a in range(8)], 1)
        dN = np.stack([np.stack([dl[a & 1] * ly[(a >> 1) & 1] * lz[(a >> 2) & 1], lx[a & 1] * dl[(a >> 1) & 1] * lz[(a >> 2) & 1],
                                 lx[a & 1] * ly[(a >> 1) & 1] * dl[(a >> 2) & 1]], 1) for a in range(8)], 1)
    else:
        raise ValueError(cell_type)
    return N, dN


def default_degree(cell_type, fatigue=False):
    """Lowest rule that integrates every form of the staggered solver exactly on affine cells
    (SURVEY H5): simplices degree 3 (4 with the fatigue RHS g(φ)·V_c·w); Q1 3 points/direction."""
    if cell_type in ("quadrilateral", "hexahedron"):
        return 5
    return 4


class Tables:
    """Per-mesh quadrature tables: N[Q,nv], physical gradients G[E,Q,nv,d], weights W[E,Q]."""

    def __init__(self, points, cells, cell_type, degree=None):
        self.cell_type = cell_type
        self.d = TDIM[cell_type]
        self.nv = NV[cell_type]
        self.cells = np.asarray(cells, dtype=np.int64)
        self.degree = default_degree(cell_type) if degree is None else degree
        pts, wts = quadrature(cell_type, self.degree)
        self.N, dN = tabulate(cell_type, pts)
        X = np.asarray(points, dtype=np.float64)[:, : self.d][self.cells]      # [E,nv,d]
        J = np.einsum("eai,qaj->eqij", X, dN)                                   # dx_i/dξ_j
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
        self.G = np.einsum("qaj,eqji->eqai", dN, Jinv)                          # dN_a/dx_i
        self.W = np.abs(det) * wts[None, :]
        self.ne, self.nq = self.W.shape
        self.nn = len(points)
