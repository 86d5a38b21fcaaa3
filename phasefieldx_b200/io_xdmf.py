"""
XDMF time-series writer standing in for `dolfinx.io.XDMFFile` as used by the reference solvers
(`src/phasefieldx/Element/Phase_Field_Fracture/solver/solver.py:259-278, 460-462`: `phi.xdmf`, `u.xdmf` in
`paraview-solutions_xdmf/`, `write_mesh` once, `write_function(f, step)` every step).

The light data is standard XDMF 3 XML (a temporal grid collection; ParaView's "XDMF Reader" opens it); the heavy data
goes to a raw little-endian side file `<stem>.bin` referenced with `Format="Binary"` + `Seek` instead of the HDF5 file
dolfinx writes — no HDF5 library is needed and nothing is re-encoded.  Mesh (geometry, topology) is written once and
referenced by every time step.
"""
import os

import numpy as np

_TOPOLOGY = {"triangle": ("Triangle", 3, [0, 1, 2]), "quadrilateral": ("Quadrilateral", 4, [0, 1, 3, 2]),
             "tetrahedron": ("Tetrahedron", 4, [0, 1, 2, 3]), "hexahedron": ("Hexahedron", 8, [0, 1, 3, 2, 4, 5, 7, 6])}


class XDMFFile:
    def __init__(self, comm, filename, mode="w"):
        if mode != "w":
            raise ValueError("XDMFFile: write mode only")
        self.path = filename
        self.dir = os.path.dirname(filename)
        self.stem = os.path.splitext(os.path.basename(filename))[0]
        os.makedirs(self.dir, exist_ok=True)
        self.bin = open(os.path.join(self.dir, self.stem + ".bin"), "wb")
        self.offset = 0
        self.mesh = None
        self.steps = []            # (t, name, seek, shape)

    def _append(self, arr):
        arr = np.ascontiguousarray(arr)
        seek = self.offset
        self.bin.write(arr.tobytes())
        self.offset += arr.nbytes
        return seek

    def write_mesh(self, msh):
        name, nv, perm = _TOPOLOGY[msh.cell_type]
        x = np.asarray(msh.geometry.x, dtype="<f8")
        gdim = 3 if msh.cell_type in ("tetrahedron", "hexahedron") else 2
        cells = np.asarray(msh.cells, dtype="<i8")[:, perm]
        self.mesh = dict(topology=name, nv=nv, ne=len(cells), nn=len(x), gdim=gdim,
                         geo_seek=self._append(x[:, :gdim]), topo_seek=self._append(cells))
        self._write_xml()

    def write_function(self, name, values, t):
        """values: [N] scalar or [N, ncomp] vector nodal field in mesh node order (vectors are padded to 3 components
        as dolfinx does)."""
        if self.mesh is None:
            raise RuntimeError("XDMFFile.write_function: call write_mesh first")
        v = np.asarray(values, dtype="<f8")
        if v.ndim == 2:
            pad = np.zeros((len(v), 3), dtype="<f8")
            pad[:, : v.shape[1]] = v
            v = pad
        self.steps.append((float(t), name, self._append(v), v.shape))
        self.bin.flush()
        self._write_xml()

    def _item(self, dtype, prec, dims, seek):
        return (f'<DataItem Format="Binary" DataType="{dtype}" Precision="{prec}" Endian="Little" Dimensions="{dims}" '
                f'Seek="{seek}">{self.stem}.bin</DataItem>')

    def _write_xml(self):
        m = self.mesh
        geo = "XY" if m["gdim"] == 2 else "XYZ"
        L = ['<?xml version="1.0"?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>', '<Xdmf Version="3.0" xmlns:xi="http://www.w3.org/2001/XInclude">',
             "  <Domain>", '    <Grid Name="mesh" GridType="Uniform">',
             f'      <Topology TopologyType="{m["topology"]}" NumberOfElements="{m["ne"]}" NodesPerElement="{m["nv"]}">',
             "        " + self._item("Int", 8, f'{m["ne"]} {m["nv"]}', m["topo_seek"]), "      </Topology>",
             f'      <Geometry GeometryType="{geo}">', "        " + self._item("Float", 8, f'{m["nn"]} {m["gdim"]}', m["geo_seek"]),
             "      </Geometry>", "    </Grid>"]
        names = []
        for _, name, _, _ in self.steps:
            if name not in names:
                names.append(name)
        for name in names:
            L.append(f'    <Grid Name="{name}" GridType="Collection" CollectionType="Temporal">')
            for t, nm, seek, shape in self.steps:
                if nm != name:
                    continue
                kind = "Scalar" if len(shape) == 1 else "Vector"
                dims = f"{shape[0]} 1" if len(shape) == 1 else f"{shape[0]} {shape[1]}"
                L += [f'      <Grid Name="{name}" GridType="Uniform">',
                      '        <xi:include xpointer="xpointer(/Xdmf/Domain/Grid[@GridType=\'Uniform\'][1]/*[self::Topology or self::Geometry])" />',
                      f'        <Time Value="{t}" />', f'        <Attribute Name="{name}" AttributeType="{kind}" Center="Node">',
                      "          " + self._item("Float", 8, dims, seek), "        </Attribute>", "      </Grid>"]
            L.append("    </Grid>")
        L += ["  </Domain>", "</Xdmf>"]
        with open(self.path, "w") as fh:
            fh.write("\n".join(L) + "\n")

    def close(self):
        if not self.bin.closed:
            self.bin.close()


def read_xdmf(path):
    """Reader for the files above (tests): {'points', 'cells', 'steps': [(t, name, array)]}."""
    import re
    txt = open(path).read()
    base = os.path.dirname(path)

    def load(tag):
        a = dict(re.findall(r'(\w+)="([^"]*)"', tag[0]))
        dims = [int(d) for d in a["Dimensions"].split()]
        dt = np.dtype("<f8") if a["DataType"] == "Float" else np.dtype("<i8")
        arr = np.fromfile(os.path.join(base, tag[1]), dtype=dt, count=int(np.prod(dims)), offset=int(a["Seek"]))
        return arr.reshape(dims)
    items = re.findall(r"<DataItem ([^>]*)>([^<]*)</DataItem>", txt)
    out = {"cells": load(items[0]), "points": load(items[1]), "steps": []}
    for m in re.finditer(r'<Time Value="([^"]*)" />\s*<Attribute Name="(\w+)"[^>]*>\s*<DataItem ([^>]*)>([^<]*)</DataItem>', txt):
        out["steps"].append((float(m.group(1)), m.group(2), load((m.group(3), m.group(4)))))
    return out
