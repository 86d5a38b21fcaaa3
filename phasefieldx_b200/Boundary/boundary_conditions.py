"""Dirichlet-condition factories with the API of reference
`src/phasefieldx/Boundary/boundary_conditions.py:23-279` (bc_phi/bc_x/bc_y/bc_z/bc_xy/bc_xyz,
get_ds_bound_from_marker).  The returned objects expose `.g.value` (mutable) exactly as the
user callbacks of the examples expect (e.g. `bcs[0].g.value[1] = val`, plot_1711.py:236-243)."""
import numpy as np

from phasefieldx_b200 import fem


def bc_phi(facet, V_phi, fdim, value=0.0):
    return fem.dirichletbc(fem.ScalarType(value), fem.locate_dofs_topological(V_phi, fdim, facet), V_phi)


def _component_bc(i, facet, V_u, fdim, value):
    sub = V_u.sub(i)
    return fem.dirichletbc(fem.ScalarType(value), fem.locate_dofs_topological(sub, fdim, facet), sub)


def bc_x(facet, V_u, fdim, value=0.0):
    return _component_bc(0, facet, V_u, fdim, value)


def bc_y(facet, V_u, fdim, value=0.0):
    return _component_bc(1, facet, V_u, fdim, value)


def bc_z(facet, V_u, fdim, value=0.0):
    return _component_bc(2, facet, V_u, fdim, value)


def bc_xy(facet, V_u, fdim, value_x=0.0, value_y=0.0):
    vals = np.array([value_x, value_y], dtype=fem.ScalarType)
    return fem.dirichletbc(vals, fem.locate_dofs_topological(V_u, fdim, facet), V_u)


def bc_xyz(facet, V_u, fdim, value_x=0.0, value_y=0.0, value_z=0.0):
    vals = np.array([value_x, value_y, value_z], dtype=fem.ScalarType)
    return fem.dirichletbc(vals, fem.locate_dofs_topological(V_u, fdim, facet), V_u)


def get_ds_bound_from_marker(facet_marker, msh, fdim):
    """Boundary measure over the marked facets (used for tractions)."""
    return fem.Measure(msh, np.sort(np.hstack([facet_marker]).astype(np.int64)))
