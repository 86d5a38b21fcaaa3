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


def _f64(a, ncomp_out=None):
    from phasefieldx_b200._lib import format_f64
    return format_f64(a, ncomp_out)


def _i64(a):
    from phasefieldx_b200._lib import format_i64
    return format_i64(a)


def _da(attrs, payload):
    return b'        <DataArray ' + attrs.encode() + b' format="ascii">' + payload + b"</DataArray>"


class VTKFile:
    """Numbers are formatted natively (pfx_format_f64/_i64 in libpfx_b200.so: shortest text
    that reads back bit-exactly); everything that does not change between load steps —
    points, cells, id arrays — is formatted once and reused."""

    def __init__(self, comm, filename, mode="w"):
        self.pvd = filename
        self.dir = os.path.dirname(filename)
        self.stem = os.path.splitext(os.path.basename(filename))[0]
        self.records = []
        self._static = None
        self._pending = None
        self._error = None
        os.makedirs(self.dir, exist_ok=True)

    def _static_part(self, msh):
        key = (id(msh), msh.geometry.x.shape, msh.cells.shape)
        if self._static is not None and self._static[0] == key:
            return self._static[1]
        x = np.asarray(msh.geometry.x, dtype=np.float64)
        if x.shape[1] < 3:
            x = np.hstack([x, np.zeros((len(x), 3 - x.shape[1]))])
        cells = msh.cells[:, _VTK_PERM[msh.cell_type]]
        nn, ne, nv = len(x), len(cells), cells.shape[1]
        L = [b'<?xml version="1.0"?>', b'<VTKFile type="UnstructuredGrid" version="2.2">', b"  <UnstructuredGrid>",
             f'    <Piece NumberOfPoints="{nn}" NumberOfCells="{ne}">'.encode(), b"      <Points>",
             _da('type="Float64" NumberOfComponents="3"', _f64(x)),
             b"      </Points>", b"      <Cells>",
             _da('type="Int32" Name="connectivity"', _i64(cells)),
             _da('type="Int32" Name="offsets"', _i64(np.arange(1, ne + 1) * nv)),
             _da('type="Int8" Name="types"', _i64(np.full(ne, _VTK_TYPE[msh.cell_type]))),
             b"      </Cells>", b"      <CellData>",
             _da('type="UInt8" Name="vtkGhostType"', _i64(np.zeros(ne, int))),
             _da('type="Int64" Name="vtkOriginalCellIds"', _i64(np.arange(ne))),
             b"      </CellData>", b"      <PointData>",
             _da('type="Int64" Name="vtkOriginalPointIds"', _i64(np.arange(nn))),
             _da('type="UInt8" Name="vtkGhostType"', _i64(np.zeros(nn, int)))]
        head = b"\n".join(L) + b"\n"
        self._static = (key, head)
        return head

    def write_function(self, msh, fields, t, background=False):
        """fields: dict name -> array [N] or [N, ncomp<=3] in mesh node order.  With
        background=True the file is formatted and written by a helper thread while the caller
        goes on with the next load step (at most one file in flight; the arrays must not be
        modified by the caller afterwards; close() waits for the last one)."""
        self._join()
        if background:
            import threading

            def run():
                try:
                    self._write(msh, fields, t)
                except BaseException as exc:      # surfaced by the next call / close()
                    self._error = exc
            self._pending = threading.Thread(target=run, daemon=True)
            self._pending.start()
        else:
            self._write(msh, fields, t)

    def _join(self):
        if self._pending is not None:
            self._pending.join()
            self._pending = None
        if self._error is not None:
            err, self._error = self._error, None
            raise err

    def _write(self, msh, fields, t):
        head = self._static_part(msh)
        step = int(t)
        name = f"{self.stem}_p0_{step:06d}.vtu"
        L = []
        for fname, vals in fields.items():
            v = np.asarray(vals, dtype=np.float64)
            if v.ndim == 2:
                L.append(_da(f'type="Float64" Name="{fname}" NumberOfComponents="3"', _f64(v, 3)))
            else:
                L.append(_da(f'type="Float64" Name="{fname}"', _f64(v)))
        L += [b"      </PointData>", b"    </Piece>", b"  </UnstructuredGrid>", b"</VTKFile>"]
        with open(os.path.join(self.dir, name), "wb") as fh:
            fh.write(head)
            fh.write(b"\n".join(L) + b"\n")
        self.records.append((t, name))
        self._write_pvd()

    def _write_pvd(self):
        L = ['<?xml version="1.0"?>', '<VTKFile type="Collection" version="0.1">', "  <Collection>"]
        L += [f'    <DataSet timestep="{t}" part="0" file="{n}" />' for t, n in self.records]
        L += ["  </Collection>", "</VTKFile>"]
        with open(self.pvd, "w") as fh:
            fh.write("\n".join(L) + "\n")

    def close(self):
        self._join()
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
