"""
Host-side mesh container and generators (no dolfinx required).

The reference hands `solve()` live dolfinx objects (`msh`, `V_u`, `V_Φ`, `bc.g.value`,
reference `src/phasefieldx/Element/Phase_Field_Fracture/solver/solver.py:43-60`).  dolfinx is
not importable in this image, so this module supplies the few mesh facilities the reference's
driver scripts use (`create_rectangle`, `create_box`, gmsh `.msh` 4.1 reading,
`locate_entities_boundary`, mesh tags with `.find`), with the same call signatures.  The
arrays it holds (`geometry.x` [N_n,3] f64, `cells` [N_e,n_v] i32) are exactly what the C-ABI
(`include/pfx_b200.h`) consumes.

Vertex ordering conventions (restated from dolfinx/basix): triangles and tetrahedra in any
order (orientation is handled with |det J|); quadrilaterals in tensor order
(x0y0, x1y0, x0y1, x1y1) as written by dolfinx, hexahedra likewise (index x + 2y + 4z).
"""
from __future__ import annotations

import enum
import numpy as np


class CellType(enum.Enum):
    interval = 1
    triangle = 2
    quadrilateral = 3
    tetrahedron = 4
    hexahedron = 5


class DiagonalType(enum.Enum):
    right = 1
    left = 2


_NV = {"triangle": 3, "quadrilateral": 4, "tetrahedron": 4, "hexahedron": 8}
_TDIM = {"triangle": 2, "quadrilateral": 2, "tetrahedron": 3, "hexahedron": 3}


class _Comm:
    """Single-process stand-in for an MPI communicator (`msh.comm`, solver.py:107-108)."""
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        return None

    def allreduce(self, v, op=None):
        return v


COMM_WORLD = _Comm()


class _Geometry:
    def __init__(self, x, dim):
        self.x = x
        self.dim = dim


class _Topology:
    def __init__(self, dim):
        self.dim = dim


class MeshTags:
    """Minimal `dolfinx.mesh.MeshTags`: `.indices`, `.values`, `.find(tag)`."""

    def __init__(self, dim, indices, values):
        order = np.argsort(indices, kind="stable")
        self.dim = dim
        self.indices = np.asarray(indices, dtype=np.int32)[order]
        self.values = np.asarray(values, dtype=np.int32)[order]

    def find(self, tag):
        return self.indices[self.values == tag]


class Mesh:
    """Unstructured P1/Q1 mesh: points [N_n,3] float64, cells [N_e,n_v] int32."""

    def __init__(self, points, cells, cell_type, gdim=None, comm=None):
        if isinstance(cell_type, CellType):
            cell_type = cell_type.name
        if cell_type not in _NV:
            raise ValueError(f"unsupported cell type {cell_type!r}")
        pts = np.zeros((len(points), 3), dtype=np.float64)
        p = np.asarray(points, dtype=np.float64)
        pts[:, : p.shape[1]] = p
        self.cell_type = cell_type
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        if self.cells.ndim != 2 or self.cells.shape[1] != _NV[cell_type]:
            raise ValueError("cells must be [N_e, n_v]")
        tdim = _TDIM[cell_type]
        self.geometry = _Geometry(pts, gdim if gdim is not None else tdim)
        self.topology = _Topology(tdim)
        self.comm = comm if comm is not None else COMM_WORLD
        self._facets = None

    # ------------------------------------------------------------------ sizes
    @property
    def num_nodes(self):
        return self.geometry.x.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    # ----------------------------------------------------------------- facets
    def _cell_facets_local(self):
        if self.cell_type == "triangle":
            return np.array([[1, 2], [0, 2], [0, 1]])
        if self.cell_type == "quadrilateral":
            return np.array([[0, 1], [0, 2], [1, 3], [2, 3]])
        if self.cell_type == "hexahedron":          # tensor vertex order x + 2y + 4z: the six quadrilateral faces
            return np.array([[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 4, 5], [2, 3, 6, 7], [0, 2, 4, 6], [1, 3, 5, 7]])
        return np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])

    def facets(self):
        """Returns (facet_vertices [N_f, k] sorted per row, is_boundary [N_f], facet_cell [N_f])."""
        if self._facets is None:
            loc = self._cell_facets_local()
            fv = self.cells[:, loc]                      # [N_e, n_f, k]
            ne, nf, k = fv.shape
            flat = np.sort(fv.reshape(ne * nf, k), axis=1)
            uniq, inv, counts = np.unique(flat, axis=0, return_inverse=True, return_counts=True)
            inv = inv.reshape(-1)
            first_cell = np.full(len(uniq), -1, dtype=np.int64)
            cell_of = np.repeat(np.arange(ne), nf)
            first_cell[inv[::-1]] = cell_of[::-1]
            self._facets = (uniq.astype(np.int32), counts == 1, first_cell)
        return self._facets


# ------------------------------------------------------------------ generators
def create_rectangle(comm, points, n, cell_type=CellType.triangle, diagonal=DiagonalType.right):
    """Structured rectangle, signature of `dolfinx.mesh.create_rectangle`
    (used by reference `examples/PhaseFieldFracture/plot_1713.py:120-124`)."""
    p0, p1 = np.asarray(points[0], float), np.asarray(points[1], float)
    nx, ny = int(n[0]), int(n[1])
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")            # node id = j*(nx+1)+i
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (j * (nx + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    name = cell_type.name if isinstance(cell_type, CellType) else cell_type
    if name == "quadrilateral":
        cells = np.stack([v00, v10, v01, v11], axis=1)
    elif name == "triangle":
        if diagonal in (DiagonalType.left, "left"):
            t1 = np.stack([v00, v10, v01], axis=1)
            t2 = np.stack([v10, v11, v01], axis=1)
        else:
            t1 = np.stack([v00, v10, v11], axis=1)
            t2 = np.stack([v00, v11, v01], axis=1)
        cells = np.empty((2 * len(v00), 3), dtype=np.int64)
        cells[0::2], cells[1::2] = t1, t2
    else:
        raise ValueError(name)
    return Mesh(pts, cells, name, gdim=2, comm=comm)


# Kuhn (Freudenthal) split of the unit cube into 6 tetrahedra sharing the main diagonal 0-7.
_KUHN = np.array([[0, 1, 3, 7], [0, 1, 5, 7], [0, 2, 3, 7], [0, 2, 6, 7], [0, 4, 5, 7], [0, 4, 6, 7]])


def create_box(comm, points, n, cell_type=CellType.tetrahedron):
    """Structured box of tetrahedra (Kuhn 6-split) or Q1 hexahedra (tensor vertex order x + 2y + 4z, as dolfinx writes
    them), signature of `dolfinx.mesh.create_box`."""
    p0, p1 = np.asarray(points[0], float), np.asarray(points[1], float)
    nx, ny, nz = (int(v) for v in n)
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")      # node id = (k*(ny+1)+j)*(nx+1)+i
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = ((k * (ny + 1) + j) * (nx + 1) + i).ravel()
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    corner = np.stack([base + (c & 1) * sx + ((c >> 1) & 1) * sy + ((c >> 2) & 1) * sz
                       for c in range(8)], axis=1)      # [n_cells, 8]
    name = cell_type.name if isinstance(cell_type, CellType) else cell_type
    if name == "hexahedron":
        return Mesh(pts, corner, "hexahedron", gdim=3, comm=comm)
    if name != "tetrahedron":
        raise ValueError("3D cells: tetrahedron or hexahedron")
    cells = corner[:, _KUHN].reshape(-1, 4)
    return Mesh(pts, cells, "tetrahedron", gdim=3, comm=comm)


def cut_slit(msh, on_slit, side_of_cell):
    """Open a crack (notch) in a conforming mesh by node duplication.

    on_slit(x[3,N]) -> bool mask of nodes lying on the slit *excluding the tip/front*;
    side_of_cell(centroids[3,N_e]) -> bool mask of the cells that must be re-attached to the
    duplicated nodes (the cells on the "upper" side).  Synthetic stand-in for the geometric
    notch of `examples/PhaseFieldFracture/mesh/mesh.msh` (SURVEY §8d, C2/C5 inputs).
    """
    x = msh.geometry.x
    mask = np.asarray(on_slit(x.T), dtype=bool)
    ids = np.nonzero(mask)[0]
    new_id = np.full(len(x), -1, dtype=np.int64)
    new_id[ids] = len(x) + np.arange(len(ids))
    cent = x[msh.cells].mean(axis=1)
    upper = np.asarray(side_of_cell(cent.T), dtype=bool)
    cells = msh.cells.astype(np.int64).copy()
    sel = cells[upper]
    dup = new_id[sel]
    sel = np.where(dup >= 0, dup, sel)
    cells[upper] = sel
    pts = np.vstack([x, x[ids]])
    return Mesh(pts, cells, msh.cell_type, gdim=msh.geometry.dim, comm=msh.comm)


# ------------------------------------------------------------------ entity location
def locate_entities_boundary(msh, dim, marker):
    """Boundary facets whose vertices all satisfy `marker(x)` (x is [3,N], dolfinx convention)."""
    if dim != msh.topology.dim - 1:
        raise ValueError("only facets (tdim-1) are supported")
    fv, is_bnd, _ = msh.facets()
    ok = np.asarray(marker(msh.geometry.x.T), dtype=bool)
    sel = is_bnd & ok[fv].all(axis=1)
    return np.nonzero(sel)[0].astype(np.int32)


def locate_entities(msh, dim, marker):
    if dim == msh.topology.dim:
        ok = np.asarray(marker(msh.geometry.x.T), dtype=bool)
        return np.nonzero(ok[msh.cells].all(axis=1))[0].astype(np.int32)
    fv, _, _ = msh.facets()
    ok = np.asarray(marker(msh.geometry.x.T), dtype=bool)
    return np.nonzero(ok[fv].all(axis=1))[0].astype(np.int32)


def meshtags(msh, dim, indices, values):
    return MeshTags(dim, indices, values)


def facet_nodes(msh, facets):
    fv, _, _ = msh.facets()
    return np.unique(fv[np.asarray(facets, dtype=np.int64)].ravel()).astype(np.int32)


# ------------------------------------------------------------------ gmsh 4.1 ASCII reader
class MeshData:
    def __init__(self, mesh, cell_tags, facet_tags, physical_groups):
        self.mesh = mesh
        self.cell_tags = cell_tags
        self.facet_tags = facet_tags
        self.physical_groups = physical_groups

    def __iter__(self):                                   # old tuple-unpacking style
        return iter((self.mesh, self.cell_tags, self.facet_tags))


def read_from_msh(filename, comm=None, rank=0, gdim=3):
    """Reader for gmsh MSH 4.1 ASCII files, signature of `dolfinx.io.gmsh.read_from_msh`
    (reference `examples/PhaseFieldFracture/plot_1711.py:125`).  Only physical-tagged
    entities are kept (gmsh/dolfinx convention); nodes are renumbered densely in file order.
    """
    with open(filename) as fh:
        lines = fh.read().split("\n")
    pos = {ln.strip(): i for i, ln in enumerate(lines) if ln.startswith("$")}
    ver = lines[pos["$MeshFormat"] + 1].split()
    if not ver[0].startswith("4") or ver[1] != "0":
        raise ValueError("only MSH 4.x ASCII is supported")
    phys_names = {}
    if "$PhysicalNames" in pos:
        i0 = pos["$PhysicalNames"] + 1
        for ln in lines[i0 + 1: i0 + 1 + int(lines[i0])]:
            d, t, name = ln.split(maxsplit=2)
            phys_names[name.strip('"')] = (int(d), int(t))
    # entities -> physical tag
    ent_phys = {}
    i = pos["$Entities"] + 1
    npnt, ncur, nsur, nvol = (int(v) for v in lines[i].split())
    i += 1
    for _ in range(npnt):
        t = lines[i].split(); i += 1
        nphys = int(t[4])
        if nphys:
            ent_phys[(0, int(t[0]))] = int(t[5])
    for dim, cnt in ((1, ncur), (2, nsur), (3, nvol)):
        for _ in range(cnt):
            t = lines[i].split(); i += 1
            nphys = int(t[7])
            if nphys:
                ent_phys[(dim, int(t[0]))] = int(t[8])
    # nodes
    i = pos["$Nodes"] + 1
    nblocks, nnodes = (int(v) for v in lines[i].split()[:2])
    i += 1
    tags = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3))
    k = 0
    for _ in range(nblocks):
        _, _, _, nb = (int(v) for v in lines[i].split()); i += 1
        tags[k:k + nb] = [int(v) for v in lines[i:i + nb]]; i += nb
        xyz[k:k + nb] = [[float(v) for v in ln.split()] for ln in lines[i:i + nb]]; i += nb
        k += nb
    tag2id = np.full(tags.max() + 1, -1, dtype=np.int64)
    tag2id[tags] = np.arange(nnodes)
    # elements
    i = pos["$Elements"] + 1
    nblocks = int(lines[i].split()[0]); i += 1
    by_type = {}
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(v) for v in lines[i].split()); i += 1
        conn = np.array([[int(v) for v in ln.split()[1:]] for ln in lines[i:i + nb]], dtype=np.int64)
        i += nb
        ptag = ent_phys.get((edim, etag))
        if ptag is None:
            continue
        by_type.setdefault(etype, []).append((tag2id[conn], np.full(nb, ptag, dtype=np.int32)))
    tdim = 3 if 4 in by_type else 2
    ctype, cname, ftype = (4, "tetrahedron", 2) if tdim == 3 else (2, "triangle", 1)
    if ctype not in by_type:
        if tdim == 2 and 3 in by_type:
            ctype, cname = 3, "quadrilateral"
        else:
            raise ValueError("no supported cells with a physical tag in the file")
    cells = np.vstack([c for c, _ in by_type[ctype]])
    ctags = np.concatenate([t for _, t in by_type[ctype]])
    if cname == "quadrilateral":                          # gmsh cyclic -> tensor order
        cells = cells[:, [0, 1, 3, 2]]
    used = np.unique(cells)
    remap = np.full(nnodes, -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    msh = Mesh(xyz[used][:, :3], remap[cells], cname, gdim=gdim, comm=comm)
    cell_tags = MeshTags(tdim, np.arange(len(cells)), ctags)
    facet_tags = MeshTags(tdim - 1, np.zeros(0, np.int32), np.zeros(0, np.int32))
    if ftype in by_type:
        fconn = remap[np.vstack([c for c, _ in by_type[ftype]])]
        ftag = np.concatenate([t for _, t in by_type[ftype]])
        fv, _, _ = msh.facets()
        key = {tuple(r): n for n, r in enumerate(fv.tolist())}
        idx = np.array([key[tuple(sorted(r))] for r in fconn.tolist()], dtype=np.int32)
        facet_tags = MeshTags(tdim - 1, idx, ftag)
    return MeshData(msh, cell_tags, facet_tags, phys_names)
