"""
Minimal function-space / boundary-condition objects with the dolfinx call signatures the
reference's driver scripts use (`dolfinx.fem.functionspace`, `locate_dofs_topological`,
`dirichletbc`, `Constant`), so `solve()` receives what it needs when dolfinx is absent
(SURVEY §7 H6).  `extract_*` also accept real dolfinx objects by duck typing.

P1/Q1 only: dofs coincide with mesh nodes; vector spaces are blocked and interleaved
(dof = node*bs + comp, the layout `Reactions/reactions_forces.py:57-63` assumes).
"""
import numpy as np

from . import mesh as _mesh

ScalarType = np.float64


class FunctionSpace:
    def __init__(self, msh, element, _parent=None, _component=None):
        family, degree = element[0], element[1]
        shape = element[2] if len(element) > 2 else ()
        if str(family) not in ("Lagrange", "P", "CG", "Q") or degree != 1:
            raise NotImplementedError("only first-order Lagrange spaces are supported (P1/Q1)")
        self.mesh = msh
        self.element = element
        self.bs = int(shape[0]) if len(shape) else 1
        self.parent, self.component = _parent, _component

    def sub(self, i):
        if self.bs == 1:
            raise ValueError("scalar space has no sub-spaces")
        return FunctionSpace(self.mesh, ("Lagrange", 1), _parent=self, _component=int(i))

    @property
    def num_nodes(self):
        return self.mesh.num_nodes


def functionspace(msh, element):
    return FunctionSpace(msh, element)


def locate_dofs_topological(V, entity_dim, entities):
    """Dofs on the closure of mesh entities.  Blocked space -> block (node) indices; sub-space ->
    unrolled indices into the parent (dolfinx convention)."""
    msh = V.mesh
    if entity_dim != msh.topology.dim - 1:
        raise NotImplementedError("only facets are supported")
    nodes = _mesh.facet_nodes(msh, np.asarray(entities, dtype=np.int64)) if len(entities) else np.zeros(0, np.int32)
    if V.parent is not None:
        return (nodes.astype(np.int64) * V.parent.bs + V.component).astype(np.int32)
    return nodes


class _Value:
    """`bc.g.value`: a mutable numpy array (0-d for scalars), as dolfinx Constant.value."""

    def __init__(self, v):
        self.value = np.array(v, dtype=np.float64)


class DirichletBC:
    def __init__(self, value, dofs, V):
        self.g = value if isinstance(value, _Value) else _Value(value)
        self.function_space = V
        dofs = np.asarray(dofs, dtype=np.int64)
        if V.parent is not None:                      # sub-space: dofs already unrolled, scalar value
            self._dofs = dofs
            self._comp = None
        elif V.bs > 1:                                # vector space: node indices, vector value
            self._dofs = (dofs[:, None] * V.bs + np.arange(V.bs)[None, :]).ravel()
            self._comp = np.tile(np.arange(V.bs), len(dofs))
        else:
            self._dofs = dofs
            self._comp = None
        self._dofs = self._dofs.astype(np.int32)

    def dof_indices(self):
        return self._dofs, len(self._dofs)

    def values_per_dof(self):
        """Current prescribed value of every constrained dof (re-read each load step,
        reference solver.py:306-309 mutates `bc.g.value` in place)."""
        v = np.asarray(self.g.value, dtype=np.float64)
        if self._comp is None:
            return np.full(len(self._dofs), float(v.reshape(-1)[0]) if v.size == 1 else np.nan)
        return v.reshape(-1)[self._comp]


def dirichletbc(value, dofs, V=None):
    return DirichletBC(value, dofs, V)


class Constant(_Value):
    """`dolfinx.fem.Constant(msh, value)` stand-in used for tractions."""

    def __init__(self, msh, value):
        super().__init__(value)
        self.mesh = msh


class Measure:
    """`ds` restricted to tagged facets, product of `get_ds_bound_from_marker`."""

    def __init__(self, msh, facets):
        self.mesh = msh
        self.facets = np.asarray(facets, dtype=np.int64)

    def nodal_weights(self):
        """nodes, w_i = ∫ N_i ds over the tagged facets (P1/Q1 trace: facet measure / #facet vertices)."""
        fv, _, _ = self.mesh.facets()
        x = self.mesh.geometry.x
        acc = {}
        for f in self.facets:
            v = fv[f]
            p = x[v]
            if len(v) == 4:
                # bilinear quadrilateral face (hexahedra); the sorted facet vertices are re-ordered into the tensor order
                # of the owning cell face: w_a = int N_a ds by 2 x 2 Gauss
                cell = self.mesh.cells[self.mesh.facets()[2][f]]
                v = np.array([n for n in cell if n in set(v.tolist())])
                p = x[v]
                gp = [0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)]
                for s_ in gp:
                    for t_ in gp:
                        N = np.array([(1 - s_) * (1 - t_), s_ * (1 - t_), (1 - s_) * t_, s_ * t_])
                        ds_ = (p[1] - p[0]) * (1 - t_) + (p[3] - p[2]) * t_
                        dt_ = (p[2] - p[0]) * (1 - s_) + (p[3] - p[1]) * s_
                        jac = np.linalg.norm(np.cross(ds_, dt_)) * 0.25
                        for a, n in enumerate(v):
                            acc[int(n)] = acc.get(int(n), 0.0) + N[a] * jac
                continue
            if len(v) == 2:
                meas = np.linalg.norm(p[1] - p[0])
            else:
                meas = 0.5 * np.linalg.norm(np.cross(p[1] - p[0], p[2] - p[0]))
            for n in v:
                acc[int(n)] = acc.get(int(n), 0.0) + meas / len(v)
        nodes = np.array(sorted(acc), dtype=np.int32)
        return nodes, np.array([acc[int(n)] for n in nodes])


# ---------------------------------------------------------------------------- extraction
def extract_mesh(msh):
    """(points [N,3] f64, cells [E,nv] i32, cell_type str) from our Mesh or a dolfinx mesh."""
    if isinstance(msh, _mesh.Mesh):
        return msh.geometry.x, msh.cells, msh.cell_type
    x = np.asarray(msh.geometry.x, dtype=np.float64)                      # dolfinx
    cells = np.asarray(msh.geometry.dofmap, dtype=np.int32)
    ctype = msh.topology.cell_type.name if hasattr(msh.topology, "cell_type") else msh.topology.cell_name()
    return x, cells, str(ctype)


def extract_bc(bc):
    """(unrolled dofs int32, per-dof values f64) from our DirichletBC or a dolfinx one."""
    if isinstance(bc, DirichletBC):
        return bc.dof_indices()[0], bc.values_per_dof()
    dofs = np.asarray(bc.dof_indices()[0], dtype=np.int32)                # dolfinx
    val = np.asarray(bc.g.value, dtype=np.float64).reshape(-1)
    bs = bc.function_space.dofmap.index_map_bs if hasattr(bc.function_space, "dofmap") else 1
    if val.size == 1:
        return dofs, np.full(len(dofs), val[0])
    return dofs, val[dofs % bs]
