"""
ASCII VTU/PVD writer reproducing the layout of dolfinx `VTKFile.write_function([phi, u], t)`
as stored by the reference (layout visible in reference
examples/PhaseFieldFracture/1711_*/paraview-solutions_vtu/phasefieldx_p0_000149.vtu):
`VTKFile version="2.2"`, Points (3 comps), Cells connectivity/offsets/types (VTK Lagrange
family: 69 triangle, 70 quadrilateral, 71 tetrahedron; Q1 cells in VTK cyclic order),
CellData vtkGhostType/vtkOriginalCellIds, PointData vtkOriginalPointIds/vtkGhostType/phi/u
(u padded to 3 components).  Multi-rank runs (rank, world given) follow dolfinx's parallel layout
(reference examples/PhaseField/2003_example_mpi/paraview-solutions_vtu/): one piece per rank
`<stem>_p<rank>_<step>.vtu` holding the rank's owned + ghost nodes and its local cells with
vtkGhostType = 1 on ghosts and the global ids in vtkOriginal*Ids, plus `<stem><step>.pvtu`
written by rank 0.
"""
import os

import numpy as np

_VTK_TYPE = {"triangle": 69, "quadrilateral": 70, "tetrahedron": 71, "hexahedron": 72}
_VTK_PERM = {"triangle": [0, 1, 2], "quadrilateral": [0, 1, 3, 2], "tetrahedron": [0, 1, 2, 3],
             "hexahedron": [0, 1, 3, 2, 4, 5, 7, 6]}


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

    def __init__(self, comm, filename, mode="w", fmt="ascii", rank=0, world=1, keep_last=None):
        """fmt "ascii": the reference's layout; "binary": same arrays as raw appended data (VTK
        `format="appended"`, UInt64 block headers) — a third of the bytes and no number formatting,
        for meshes where the ASCII file is what a load step waits for."""
        if fmt not in ("ascii", "binary"):
            raise ValueError("fmt must be 'ascii' or 'binary'")
        self.fmt = fmt
        # keep_last = k: only the k most recent step files of this rank stay on disk (bounded disk use for long
        # runs on large meshes); None (default, the reference's behaviour): every step file is kept
        self.keep_last = None if keep_last is None else max(1, int(keep_last))
        self._written = []
        self.rank, self.world = int(rank), int(world)
        self.pvd = filename
        self.dir = os.path.dirname(filename)
        self.stem = os.path.splitext(os.path.basename(filename))[0]
        self.records = []
        self._static = None
        self._pending = None
        self._error = None
        os.makedirs(self.dir, exist_ok=True)

    @staticmethod
    def _ids(msh, nn, ne):
        """(point ids, cell ids, point ghost flags, cell ghost flags): global numbering and ghost marks of a
        rank-local piece (attributes point_gid / cell_gid / point_ghost / cell_ghost), else the identity."""
        pg = getattr(msh, "point_gid", None)
        cg = getattr(msh, "cell_gid", None)
        pgh = getattr(msh, "point_ghost", None)
        cgh = getattr(msh, "cell_ghost", None)
        return (np.arange(nn) if pg is None else np.asarray(pg, dtype=np.int64),
                np.arange(ne) if cg is None else np.asarray(cg, dtype=np.int64),
                np.zeros(nn, np.uint8) if pgh is None else np.asarray(pgh, dtype=np.uint8),
                np.zeros(ne, np.uint8) if cgh is None else np.asarray(cgh, dtype=np.uint8))

    def _static_part(self, msh):
        key = (id(msh), msh.geometry.x.shape, msh.cells.shape)
        if self._static is not None and self._static[0] == key:
            return self._static[1]
        x = np.asarray(msh.geometry.x, dtype=np.float64)
        if x.shape[1] < 3:
            x = np.hstack([x, np.zeros((len(x), 3 - x.shape[1]))])
        cells = msh.cells[:, _VTK_PERM[msh.cell_type]]
        nn, ne, nv = len(x), len(cells), cells.shape[1]
        pid, cid, pgh, cgh = self._ids(msh, nn, ne)
        L = [b'<?xml version="1.0"?>', b'<VTKFile type="UnstructuredGrid" version="2.2">', b"  <UnstructuredGrid>",
             f'    <Piece NumberOfPoints="{nn}" NumberOfCells="{ne}">'.encode(), b"      <Points>",
             _da('type="Float64" NumberOfComponents="3"', _f64(x)),
             b"      </Points>", b"      <Cells>",
             _da('type="Int32" Name="connectivity"', _i64(cells)),
             _da('type="Int32" Name="offsets"', _i64(np.arange(1, ne + 1) * nv)),
             _da('type="Int8" Name="types"', _i64(np.full(ne, _VTK_TYPE[msh.cell_type]))),
             b"      </Cells>", b"      <CellData>",
             _da('type="UInt8" Name="vtkGhostType"', _i64(cgh)),
             _da('type="Int64" Name="vtkOriginalCellIds"', _i64(cid)),
             b"      </CellData>", b"      <PointData>",
             _da('type="Int64" Name="vtkOriginalPointIds"', _i64(pid)),
             _da('type="UInt8" Name="vtkGhostType"', _i64(pgh))]
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

    def _write_binary(self, msh, fields, t):
        x = np.asarray(msh.geometry.x, dtype=np.float64)
        if x.shape[1] < 3:
            x = np.hstack([x, np.zeros((len(x), 3 - x.shape[1]))])
        cells = np.ascontiguousarray(msh.cells[:, _VTK_PERM[msh.cell_type]], dtype=np.int32)
        nn, ne, nv = len(x), len(cells), cells.shape[1]
        key = (id(msh), x.shape, cells.shape)
        if self._static is None or self._static[0] != key:
            pid, cid, pgh, cgh = self._ids(msh, nn, ne)
            self._static = (key, [("Points", None, "Float64", 3, np.ascontiguousarray(x)),
                                  ("Cells", "connectivity", "Int32", 0, cells),
                                  ("Cells", "offsets", "Int32", 0, (np.arange(1, ne + 1) * nv).astype(np.int32)),
                                  ("Cells", "types", "Int8", 0, np.full(ne, _VTK_TYPE[msh.cell_type], np.int8)),
                                  ("CellData", "vtkGhostType", "UInt8", 0, np.ascontiguousarray(cgh)),
                                  ("CellData", "vtkOriginalCellIds", "Int64", 0, np.ascontiguousarray(cid)),
                                  ("PointData", "vtkOriginalPointIds", "Int64", 0, np.ascontiguousarray(pid)),
                                  ("PointData", "vtkGhostType", "UInt8", 0, np.ascontiguousarray(pgh))])
        arrays = list(self._static[1])
        for fname, vals in fields.items():
            v = np.asarray(vals, dtype=np.float64)
            if v.ndim == 2:
                pad = np.zeros((nn, 3))
                pad[:, : v.shape[1]] = v
                arrays.append(("PointData", fname, "Float64", 3, pad))
            else:
                arrays.append(("PointData", fname, "Float64", 0, np.ascontiguousarray(v)))
        L = ['<?xml version="1.0"?>',
             '<VTKFile type="UnstructuredGrid" version="2.2" byte_order="LittleEndian" header_type="UInt64">',
             "  <UnstructuredGrid>", f'    <Piece NumberOfPoints="{nn}" NumberOfCells="{ne}">']
        off, section = 0, None
        for sec, name, typ, ncomp, arr in arrays:
            if sec != section:
                if section is not None:
                    L.append(f"      </{section}>")
                L.append(f"      <{sec}>")
                section = sec
            attrs = f'type="{typ}"' + (f' Name="{name}"' if name else "") + (f' NumberOfComponents="{ncomp}"' if ncomp else "")
            L.append(f'        <DataArray {attrs} format="appended" offset="{off}"/>')
            off += 8 + arr.nbytes
        L += [f"      </{section}>", "    </Piece>", "  </UnstructuredGrid>", '  <AppendedData encoding="raw">']
        step = int(t)
        fname_out = f"{self.stem}_p{self.rank}_{step:06d}.vtu"
        with open(os.path.join(self.dir, fname_out), "wb") as fh:
            fh.write(("\n".join(L) + "\n_").encode())
            for _, _, _, _, arr in arrays:
                fh.write(np.uint64(arr.nbytes).tobytes())
                fh.write(arr.tobytes())
            fh.write(b"\n  </AppendedData>\n</VTKFile>\n")
        self._record(t, step, fname_out, fields)

    def _record(self, t, step, piece_name, fields):
        """Book-keeping after a piece: rank 0 writes the .pvtu of a multi-rank step and the .pvd index."""
        self._written.append(os.path.join(self.dir, piece_name))
        while self.keep_last is not None and len(self._written) > self.keep_last:
            old = self._written.pop(0)
            if os.path.exists(old):
                os.remove(old)
        if self.rank != 0:
            return
        name = piece_name
        if self.world > 1:
            name = f"{self.stem}{step:06d}.pvtu"
            L = ['<?xml version="1.0"?>', '<VTKFile type="PUnstructuredGrid" version="1.0">', '  <PUnstructuredGrid GhostLevel="1">',
                 "    <PPointData>", '      <PDataArray type="Int64" Name="vtkOriginalPointIds" IdType="1" />',
                 '      <PDataArray type="UInt8" Name="vtkGhostType" />']
            for fname, vals in fields.items():
                nc = 3 if np.ndim(vals) == 2 else 1
                L.append(f'      <PDataArray type="Float64" Name="{fname}" NumberOfComponents="{nc}" />')
            L += ["    </PPointData>", "    <PCellData>", '      <PDataArray type="UInt8" Name="vtkGhostType" />',
                  '      <PDataArray type="Int64" Name="vtkOriginalCellIds" IdType="1" />', "    </PCellData>", "    <PPoints>",
                  '      <PDataArray type="Float64" NumberOfComponents="3" />', "    </PPoints>"]
            L += [f'    <Piece Source="{self.stem}_p{r}_{step:06d}.vtu" />' for r in range(self.world)]
            L += ["  </PUnstructuredGrid>", "</VTKFile>"]
            with open(os.path.join(self.dir, name), "w") as fh:
                fh.write("\n".join(L) + "\n")
        self.records.append((t, name))
        if self.keep_last is not None:
            self.records = self.records[-self.keep_last:]
        self._write_pvd()

    def _write(self, msh, fields, t):
        if self.fmt == "binary":
            return self._write_binary(msh, fields, t)
        head = self._static_part(msh)
        step = int(t)
        name = f"{self.stem}_p{self.rank}_{step:06d}.vtu"
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
        self._record(t, step, name, fields)

    def _write_pvd(self):
        if self.rank != 0:
            return
        L = ['<?xml version="1.0"?>', '<VTKFile type="Collection" version="0.1">', "  <Collection>"]
        L += [f'    <DataSet timestep="{t}" part="0" file="{n}" />' for t, n in self.records]
        L += ["  </Collection>", "</VTKFile>"]
        with open(self.pvd, "w") as fh:
            fh.write("\n".join(L) + "\n")

    def close(self):
        self._join()
        self._write_pvd()


def read_vtu_points_cells(path):
    """Tiny reader for the VTUs above (ASCII, also the reference's stored ones, or raw appended): points,
    cells, phi, u."""
    import re
    raw = open(path, "rb").read()
    cut = raw.find(b"<AppendedData")
    txt = (raw if cut < 0 else raw[:cut]).decode()
    npts = int(re.search(r'NumberOfPoints="(\d+)"', txt).group(1))
    ncell = int(re.search(r'NumberOfCells="(\d+)"', txt).group(1))
    if cut >= 0:
        base = raw.index(b"_", cut) + 1
        dt = {"Float64": np.float64, "Int32": np.int32, "Int8": np.int8, "UInt8": np.uint8, "Int64": np.int64}

        def block(tag):
            typ = re.search(r'type="(\w+)"', tag).group(1)
            off = base + int(re.search(r'offset="(\d+)"', tag).group(1))
            nbytes = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=off)[0])
            return np.frombuffer(raw, dtype=dt[typ], count=nbytes // np.dtype(dt[typ]).itemsize, offset=off + 8)
        tags = re.findall(r"<DataArray[^>]*/>", txt)
        named = {re.search(r'Name="(\w+)"', t).group(1): t for t in tags if 'Name="' in t}
        pts = block(next(t for t in tags if 'Name="' not in t)).astype(np.float64).reshape(npts, 3)
        out = {"points": pts, "cells": block(named["connectivity"]).astype(np.int64).reshape(ncell, -1)}
        for nm in ("phi", "u"):
            if nm in named:
                a = block(named[nm]).astype(np.float64)
                out[nm] = a.reshape(npts, -1) if a.size != npts else a
        ghost = [block(t) for t in tags if 'Name="vtkGhostType"' in t]
        out.update(cell_ghost=ghost[0].astype(np.int64), point_ghost=ghost[1].astype(np.int64),
                   cell_gid=block(named["vtkOriginalCellIds"]).astype(np.int64),
                   point_gid=block(named["vtkOriginalPointIds"]).astype(np.int64))
        return out

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
    ghost = re.findall(r'Name="vtkGhostType"[^>]*>([^<]*)</DataArray>', txt)
    if len(ghost) == 2:
        out["cell_ghost"], out["point_ghost"] = (np.array(g.split(), dtype=np.int64) for g in ghost)
    for key, nm in (("cell_gid", "vtkOriginalCellIds"), ("point_gid", "vtkOriginalPointIds")):
        m = re.search(rf'Name="{nm}"[^>]*>([^<]*)</DataArray>', txt)
        if m:
            out[key] = np.array(m.group(1).split(), dtype=np.int64)
    return out
