"""
ASCII VTU/PVD writer reproducing the layout of dolfinx `VTKFile.write_function([phi, u], t)`
as stored by the reference (layout visible in reference
examples/PhaseFieldFracture/1711_*/paraview-solutions_vtu/phasefieldx_p0_000149.vtu):
`VTKFile version="2.2"`, Points (3 comps), Cells connectivity/offsets/types (VTK Lagrange
family: 69 triangle, 70 quadrilateral, 71 tetrahedron; Q1 cells in VTK cyclic order),
CellData vtkGhostType/vtkOriginalCellIds, PointData vtkOriginalPointIds/vtkGhostType/phi/u
(u padded to 3 components).
"""
import os

import numpy as np

_VTK_TYPE = {"triangle": 69, "quadrilateral": 70, "tetrahedron": 71}
_VTK_PERM = {"triangle": [0, 1, 2], "quadrilateral": [0, 1, 3, 2], "tetrahedron": [0, 1, 2, 3]}


def _arr(a, fmt):
    return " ".join(fmt % v for v in np.asarray(a).ravel())


class VTKFile:
    def __init__(self, comm, filename, mode="w"):
        self.pvd = filename
        self.dir = os.path.dirname(filename)
        self.stem = os.path.splitext(os.path.basename(filename))[0]
        self.records = []
        os.makedirs(self.dir, exist_ok=True)

    def write_function(self, msh, fields, t):
        """fields: dict name -> array [N] or [N, ncomp<=3] in mesh node order."""
        x = msh.geometry.x
        cells = msh.cells[:, _VTK_PERM[msh.cell_type]]
        nn, ne, nv = len(x), len(cells), cells.shape[1]
        step = int(t)
        name = f"{self.stem}_p0_{step:06d}.vtu"
        L = ['<?xml version="1.0"?>', '<VTKFile type="UnstructuredGrid" version="2.2">', "  <UnstructuredGrid>",
             f'    <Piece NumberOfPoints="{nn}" NumberOfCells="{ne}">', "      <Points>",
             '        <DataArray type="Float64" NumberOfComponents="3" format="ascii">' + _arr(x, "%.16g") + "</DataArray>",
             "      </Points>", "      <Cells>",
             '        <DataArray type="Int32" Name="connectivity" format="ascii">' + _arr(cells, "%d") + "</DataArray>",
             '        <DataArray type="Int32" Name="offsets" format="ascii">' + _arr(np.arange(1, ne + 1) * nv, "%d") + "</DataArray>",
             '        <DataArray type="Int8" Name="types" format="ascii">' + _arr(np.full(ne, _VTK_TYPE[msh.cell_type]), "%d") + "</DataArray>",
             "      </Cells>", "      <CellData>",
             '        <DataArray type="UInt8" Name="vtkGhostType" format="ascii">' + _arr(np.zeros(ne, int), "%d") + "</DataArray>",
             '        <DataArray type="Int64" Name="vtkOriginalCellIds" format="ascii">' + _arr(np.arange(ne), "%d") + "</DataArray>",
             "      </CellData>", "      <PointData>",
             '        <DataArray type="Int64" Name="vtkOriginalPointIds" format="ascii">' + _arr(np.arange(nn), "%d") + "</DataArray>",
             '        <DataArray type="UInt8" Name="vtkGhostType" format="ascii">' + _arr(np.zeros(nn, int), "%d") + "</DataArray>"]
        for fname, vals in fields.items():
            v = np.asarray(vals, dtype=np.float64)
            if v.ndim == 2:
                pad = np.zeros((nn, 3))
                pad[:, : v.shape[1]] = v
                L.append(f'        <DataArray type="Float64" Name="{fname}" NumberOfComponents="3" format="ascii">' + _arr(pad, "%.16g") + "</DataArray>")
            else:
                L.append(f'        <DataArray type="Float64" Name="{fname}" format="ascii">' + _arr(v, "%.16g") + "</DataArray>")
        L += ["      </PointData>", "    </Piece>", "  </UnstructuredGrid>", "</VTKFile>"]
        with open(os.path.join(self.dir, name), "w") as fh:
            fh.write("\n".join(L) + "\n")
        self.records.append((t, name))
        self._write_pvd()

    def _write_pvd(self):
        L = ['<?xml version="1.0"?>', '<VTKFile type="Collection" version="0.1">', "  <Collection>"]
        L += [f'    <DataSet timestep="{t}" part="0" file="{n}" />' for t, n in self.records]
        L += ["  </Collection>", "</VTKFile>"]
        with open(self.pvd, "w") as fh:
            fh.write("\n".join(L) + "\n")

    def close(self):
        self._write_pvd()


def read_vtu_points_cells(path):
    """Tiny reader for the ASCII VTUs above (and the reference's stored ones): points, cells."""
    import re
    txt = open(path).read()
    npts = int(re.search(r'NumberOfPoints="(\d+)"', txt).group(1))
    ncell = int(re.search(r'NumberOfCells="(\d+)"', txt).group(1))

    def data(pattern):
        m = re.search(pattern + r'[^>]*>([^<]*)</DataArray>', txt)
        return np.array(m.group(1).split(), dtype=np.float64)
    pts = data(r'<Points>\s*<DataArray').reshape(npts, 3)
    conn = data(r'Name="connectivity"').astype(np.int64).reshape(ncell, -1)
    out = {"points": pts, "cells": conn}
    for nm in ("phi", "u"):
        m = re.search(rf'Name="{nm}"[^>]*>([^<]*)</DataArray>', txt)
        if m:
            a = np.array(m.group(1).split(), dtype=np.float64)
            out[nm] = a.reshape(npts, -1) if a.size != npts else a
    return out
