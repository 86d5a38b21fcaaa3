"""Host-side drop-in surface: file utilities (mirrors reference test/test_file_utilities.py),
Input, AllResults, boundary-condition factories on the dolfinx stand-ins, gmsh reader, VTU
writer, and the `phasefieldx` import aliasing."""
import os

import numpy as np
import pytest


@pytest.fixture
def setup_directory(tmp_path):
    folder = tmp_path / "test_folder"
    folder.mkdir()
    (folder / "file1.txt").write_text("This is file 1.")
    (folder / "file2.txt").write_text("This is file 2.")
    (folder / "file3.log").write_text("This is a log file.")
    (folder / "subfolder").mkdir()
    return str(folder)


def test_file_utilities(setup_directory, tmp_path):
    import phasefieldx_b200.compat as compat
    compat.install()
    from phasefieldx import (append_results_to_file, delete_folder, delete_folders_with_contents, delete_specific_files,
                             find_files_with_extension, get_filenames_in_folder, get_foldernames_in_folder,
                             prepare_simulation, read_last_line, read_last_lines, replace_text)
    d = setup_directory
    assert sorted(get_filenames_in_folder(d)) == ["file1.txt", "file2.txt", "file3.log"]
    assert get_foldernames_in_folder(d) == ["subfolder"]
    assert sorted(find_files_with_extension(d, ".txt")) == sorted([os.path.join(d, "file1.txt"), os.path.join(d, "file2.txt")])
    assert read_last_line(os.path.join(d, "file1.txt")) == "This is file 1."
    assert read_last_lines(os.path.join(d, "file1.txt"), 1) == ["This is file 1."]
    assert replace_text(os.path.join(d, "file1.txt"), os.path.join(d, "r.txt"), [("file 1", "replaced file 1")])
    assert "replaced file 1." in open(os.path.join(d, "r.txt")).read()
    delete_specific_files(d, [".log"])
    assert not os.path.exists(os.path.join(d, "file3.log"))
    delete_folders_with_contents(d, ["subfolder"])
    assert not os.path.exists(os.path.join(d, "subfolder"))
    prepare_simulation(d, "results")
    assert os.path.exists(os.path.join(d, "results"))
    os.mkdir(os.path.join(d, "temp_folder"))
    open(os.path.join(d, "temp_folder", "t.txt"), "w").write("x")
    delete_folder(os.path.join(d, "temp_folder"))
    assert not os.path.exists(os.path.join(d, "temp_folder"))
    out = str(tmp_path / "results.txt")
    append_results_to_file(out, "Step\tValue1\tValue2", 1, 0.5, 1.2)
    append_results_to_file(out, "Step\tValue1\tValue2", 2, 0.6, 1.3)
    lines = open(out).read().splitlines()
    assert lines == ["Step\tValue1\tValue2", "1\t0.5\t1.2", "2\t0.6\t1.3"]


def test_input_and_allresults(tmp_path):
    from phasefieldx_b200.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx_b200.PostProcessing.ReferenceResult import AllResults
    from phasefieldx_b200.files import append_results_to_file
    D = Input(E=210.0, nu=0.3, degradation="anisotropic", split_energy="spectral", results_folder_name=str(tmp_path))
    assert abs(D.lambda_ - 121.15384615384615) < 1e-12 and abs(D.mu - 80.76923076923077) < 1e-12
    D.save_parameters_to_csv(str(tmp_path / "parameters.input"))
    rows = open(tmp_path / "parameters.input").read().splitlines()
    assert rows[0] == "E\t210.0" and rows[5].startswith("lambda\t121.15") and rows[-1].startswith("results_folder_name\t")
    append_results_to_file(str(tmp_path / "top.reaction"), "#step\tRx\tRy\tRz", 0, 0.0, -1.5, 0.0)
    S = AllResults(str(tmp_path))
    assert S.reaction_files["top.reaction"]["Ry"][0] == -1.5
    assert S.input["parameters.input"]["degradation"] == "anisotropic"
    with pytest.raises(ValueError):
        Input(nu=0.5)


def test_boundary_factories_and_value_semantics():
    """`bc.g.value[...]` mutation as in the reference callbacks (plot_1711.py:236-243,
    plot_1718.py:286-290) is re-read per dof; later BCs win on shared dofs downstream."""
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    import mpi4py
    import petsc4py
    from phasefieldx.Boundary.boundary_conditions import bc_x, bc_xy, bc_y, bc_phi, get_ds_bound_from_marker
    from phasefieldx_b200 import fem
    msh = dolfinx.mesh.create_rectangle(mpi4py.MPI.COMM_WORLD, [np.array([0, 0]), np.array([1, 1])], [4, 3],
                                        cell_type=dolfinx.mesh.CellType.quadrilateral)
    fdim = msh.topology.dim - 1
    top = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 1))
    assert len(top) == 4
    V_u = dolfinx.fem.functionspace(msh, ("Lagrange", 1, (msh.geometry.dim,)))
    V_phi = dolfinx.fem.functionspace(msh, ("Lagrange", 1))
    b = bc_xy(top, V_u, fdim)
    b.g.value[1] = petsc4py.PETSc.ScalarType(0.3)
    dofs, vals = fem.extract_bc(b)
    assert len(dofs) == 10 and np.all(vals[1::2] == 0.3) and np.all(vals[0::2] == 0.0)
    by = bc_y(top, V_u, fdim)
    by.g.value[...] = -0.7
    dofs, vals = fem.extract_bc(by)
    assert np.all(dofs % 2 == 1) and np.all(vals == -0.7)
    assert np.all(fem.extract_bc(bc_x(top, V_u, fdim, 2.0))[1] == 2.0)
    assert len(fem.extract_bc(bc_phi(top, V_phi, fdim, 1.0))[0]) == 5
    nodes, w = get_ds_bound_from_marker(top, msh, fdim).nodal_weights()
    assert abs(w.sum() - 1.0) < 1e-14 and len(nodes) == 5


MSH = """$MeshFormat
4.1 0 8
$EndMeshFormat
$PhysicalNames
2
1 9 "bottom"
2 8 "surface"
$EndPhysicalNames
$Entities
4 4 1 0
1 0 0 0 0
2 1 0 0 0
3 1 1 0 0
4 0 1 0 0
1 0 0 0 1 0 0 1 9 2 1 -2
2 1 0 0 1 1 0 0 2 2 -3
3 0 1 0 1 1 0 0 2 3 -4
4 0 0 0 0 1 0 0 2 4 -1
1 0 0 0 1 1 0 1 8 4 1 2 3 4
$EndEntities
$Nodes
5 5 1 5
0 1 0 1
1
0 0 0
0 2 0 1
2
1 0 0
0 3 0 1
3
1 1 0
0 4 0 1
4
0 1 0
2 1 0 1
5
0.5 0.5 0
$EndNodes
$Elements
2 5 1 5
1 1 1 1
1 1 2
2 1 2 4
2 1 2 5
3 2 3 5
4 3 4 5
5 4 1 5
$EndElements
"""


def test_gmsh_reader_and_vtu_writer(tmp_path):
    from phasefieldx_b200 import io_vtk, mesh as M
    p = tmp_path / "m.msh"
    p.write_text(MSH)
    md = M.read_from_msh(str(p), None, 0, 2)
    msh = md.mesh
    assert msh.num_nodes == 5 and msh.num_cells == 4 and msh.cell_type == "triangle"
    assert sorted(M.facet_nodes(msh, md.facet_tags.find(9)).tolist()) == [0, 1]
    assert md.facet_tags.find(11).size == 0                      # plot_1711.py queries absent tags
    vt = io_vtk.VTKFile(None, str(tmp_path / "out" / "phasefieldx.pvd"), "w")
    phi = np.linspace(0, 1, 5)
    u = np.arange(10.0).reshape(5, 2)
    vt.write_function(msh, {"phi": phi, "u": u}, 3)
    vt.close()
    f = tmp_path / "out" / "phasefieldx_p0_000003.vtu"
    txt = f.read_text()
    assert 'VTKFile type="UnstructuredGrid" version="2.2"' in txt and 'Name="vtkOriginalPointIds"' in txt
    back = io_vtk.read_vtu_points_cells(str(f))
    assert np.allclose(back["phi"], phi) and np.allclose(back["u"][:, :2], u) and np.all(back["u"][:, 2] == 0)
    assert np.array_equal(back["cells"], msh.cells)
    assert "phasefieldx_p0_000003.vtu" in (tmp_path / "out" / "phasefieldx.pvd").read_text()


def test_workloads_are_deterministic():
    from phasefieldx_b200 import workloads as W
    a, b = W.sen_shear(16), W.sen_shear(16)
    assert np.array_equal(a.msh.cells, b.msh.cells) and a.msh.num_nodes == 17 * 17 + 8
    w = W.notched_beam_3d(12, 4, 2)
    assert w.msh.cell_type == "tetrahedron" and all(len(d) > 0 for _, d in w.bc_sets)
    assert w.bc_values(3.0)[0][0] == -0.03


def test_native_number_formatting_round_trips_bit_exactly():
    """pfx_format_f64/_i64 (the VTU writer's ASCII conversion): text parses back to the same
    bits, vectors are zero-padded like dolfinx pads u to 3 components, and the threaded path
    (> 65536 rows) gives the same bytes as formatting the pieces one by one."""
    from phasefieldx_b200 import _lib
    rng = np.random.default_rng(5)
    special = np.array([0.0, -0.0, 1.0, -1.0, 5e-324, 1.7976931348623157e308, 1e-5, 0.1, 1 / 3, 2.5e-7, 123456789.125])
    txt = _lib.format_f64(special).decode()
    back = np.array(txt.split(), dtype=np.float64)
    assert np.array_equal(back.view(np.int64), special.view(np.int64))
    assert txt.split()[:3] == ["0", "-0", "1"]
    v = rng.standard_normal((200001, 2)) * 10.0 ** rng.integers(-12, 6, size=(200001, 1))
    txt = _lib.format_f64(v, 3)
    back = np.array(txt.split(), dtype=np.float64).reshape(-1, 3)
    assert np.array_equal(back[:, :2], v) and not back[:, 2].any()
    pieces = b" ".join(_lib.format_f64(v[a:a + 50000], 3) for a in range(0, len(v), 50000))
    assert pieces == txt
    ints = rng.integers(-2**62, 2**62, size=150000)
    assert np.array_equal(np.array(_lib.format_i64(ints).split(), dtype=np.int64), ints)
    assert _lib.format_f64(np.zeros((0, 2)), 3) == b"" and _lib.format_i64(np.zeros(0, int)) == b""


def test_single_field_inputs_match_reference_format(tmp_path):
    """Input classes of the Elasticity / Phase_Field elements: reference keyword names, defaults and the
    two-column `parameters.input` layout (reference Element/Elasticity/Input.py:46-111,
    Element/Phase_Field/Input.py:33-90)."""
    from phasefieldx_b200.Element.Elasticity.Input import Input as EInput
    from phasefieldx_b200.Element.Phase_Field.Input import Input as PInput
    e = EInput(E=100.0, nu=0.25)
    assert abs(e.lambda_ - 40.0) < 1e-12 and abs(e.mu - 40.0) < 1e-12
    f = tmp_path / "parameters.input"
    e.save_parameters_to_csv(str(f))
    rows = [l.split("\t") for l in f.read_text().splitlines()]
    assert [r[0] for r in rows] == ["E", "nu", "lambda", "mu", "save_solution_xdmf", "save_solution_vtu", "results_folder_name"]
    assert rows[0][1] == "100.0" and rows[-1][1] == "results"
    assert str(e).splitlines()[0] == "Material parameters:"
    p = PInput(l=4.0, results_folder_name="r")
    p.save_parameters_to_csv(str(f))
    assert f.read_text().splitlines() == ["l\t4.0", "save_solution_xdmf\tFalse", "save_solution_vtu\tTrue", "results_folder_name\tr"]
    assert str(p) == "Parameters:\n  l: 4.0"


def test_binary_vtu_round_trips_bit_exactly(tmp_path):
    """`VTKFile(fmt="binary")`: raw appended VTU with the same arrays and cell types as
    the ASCII layout; fields, points and connectivity read back bit-exactly, for every cell type."""
    from phasefieldx_b200 import io_vtk, mesh as M
    rng = np.random.default_rng(11)
    meshes = [M.create_rectangle(None, [[0, 0], [1, 2]], [7, 5]),
              M.create_rectangle(None, [[0, 0], [1, 2]], [4, 3], M.CellType.quadrilateral),
              M.create_box(None, [[0, 0, 0], [1, 1, 1]], [3, 2, 2])]
    for k, msh in enumerate(meshes):
        d = msh.geometry.dim
        phi, u = rng.uniform(size=msh.num_nodes), rng.standard_normal((msh.num_nodes, d)) * 1e-3
        out = {}
        for fmt in ("ascii", "binary"):
            vt = io_vtk.VTKFile(None, str(tmp_path / f"{fmt}{k}" / "phasefieldx.pvd"), "w", fmt=fmt)
            vt.write_function(msh, {"phi": phi, "u": u}, 2, background=(fmt == "binary"))
            vt.close()
            out[fmt] = io_vtk.read_vtu_points_cells(str(tmp_path / f"{fmt}{k}" / "phasefieldx_p0_000002.vtu"))
            assert "phasefieldx_p0_000002.vtu" in (tmp_path / f"{fmt}{k}" / "phasefieldx.pvd").read_text()
        for key in ("points", "cells", "phi", "u"):
            assert np.array_equal(out["ascii"][key], out["binary"][key]), key
        assert np.array_equal(out["binary"]["phi"], phi) and np.array_equal(out["binary"]["u"][:, :d], u)
        head = (tmp_path / f"binary{k}" / "phasefieldx_p0_000002.vtu").read_bytes()[:400].decode(errors="ignore")
        assert 'header_type="UInt64"' in head and 'format="appended"' in head
    with pytest.raises(ValueError):
        io_vtk.VTKFile(None, str(tmp_path / "x" / "p.pvd"), "w", fmt="hdf5")


@pytest.mark.parametrize("fmt", ["ascii", "binary"])
def test_parallel_vtu_pieces_merge_to_the_global_fields(tmp_path, fmt):
    """Multi-rank output (layout of reference examples/PhaseField/2003_example_mpi/
    paraview-solutions_vtu/): one piece per rank with global ids and ghost marks + a .pvtu by rank 0.  Merging the
    non-ghost nodes of all pieces gives back the global fields; every global cell is non-ghost in exactly one piece."""
    from phasefieldx_b200 import io_vtk, mesh as M
    from phasefieldx_b200.distributed import RankMesh
    msh = M.create_box(None, [[0, 0, 0], [2, 1, 1]], [7, 4, 3])
    x, cells = msh.geometry.x, msh.cells
    rng = np.random.default_rng(2)
    phi_g, u_g = rng.uniform(size=len(x)), rng.standard_normal((len(x), 3))
    world = 3
    for r in range(world):
        rm = RankMesh(x, cells, "tetrahedron", world, r)
        vt = io_vtk.VTKFile(None, str(tmp_path / "out" / "phasefieldx.pvd"), "w", fmt=fmt, rank=r, world=world)
        vt.write_function(rm.piece("tetrahedron"), {"phi": phi_g[rm.l2g], "u": u_g[rm.l2g]}, 7, background=True)
        vt.close()
    pvtu = (tmp_path / "out" / "phasefieldx000007.pvtu").read_text()
    assert all(f'Piece Source="phasefieldx_p{r}_000007.vtu"' in pvtu for r in range(world)) and 'Name="u" NumberOfComponents="3"' in pvtu
    assert 'file="phasefieldx000007.pvtu"' in (tmp_path / "out" / "phasefieldx.pvd").read_text()
    phi, u, seen_pt, seen_cell = np.zeros(len(x)), np.zeros((len(x), 3)), np.zeros(len(x), int), np.zeros(len(cells), int)
    for r in range(world):
        d = io_vtk.read_vtu_points_cells(str(tmp_path / "out" / f"phasefieldx_p{r}_000007.vtu"))
        own = d["point_ghost"] == 0
        gid = d["point_gid"][own]
        phi[gid], u[gid] = d["phi"][own], d["u"][own]
        seen_pt[gid] += 1
        seen_cell[d["cell_gid"][d["cell_ghost"] == 0]] += 1
        assert np.array_equal(d["points"], x[d["point_gid"]])                       # local points are the global ones
        assert np.array_equal(np.sort(d["point_gid"][d["cells"]], 1), np.sort(cells[d["cell_gid"]], 1))
    assert np.all(seen_pt == 1) and np.all(seen_cell == 1)
    assert np.array_equal(phi, phi_g) and np.array_equal(u, u_g)


def test_reference_stored_vtu_parses_and_writer_reproduces_its_layout(tmp_path):
    """O2 pinned on the reference's own file: tests/golden/ref_1711_final.vtu.gz is the stored final VTU of reference
    example 1711 (examples/PhaseFieldFracture/1711_*/paraview-solutions_vtu/phasefieldx_p0_000149.vtu).  The reader
    parses it; writing the same mesh and fields with VTKFile gives the same sequence of XML elements and DataArrays
    (type, Name, NumberOfComponents) and the same numbers."""
    import gzip
    import re
    from conftest import GOLDEN
    from phasefieldx_b200 import io_vtk, mesh as M
    ref_txt = gzip.open(os.path.join(GOLDEN, "ref_1711_final.vtu.gz")).read().decode()
    src = tmp_path / "ref.vtu"
    src.write_text(ref_txt)
    ref = io_vtk.read_vtu_points_cells(str(src))
    assert ref["points"].shape == (5089, 3) and ref["cells"].shape == (9990, 3)
    assert ref["phi"].shape == (5089,) and ref["u"].shape == (5089, 3) and np.all(ref["u"][:, 2] == 0.0)
    assert 0.99 < ref["phi"].max() < 1.05 and ref["cells"].min() == 0 and ref["cells"].max() == 5088
    assert np.array_equal(ref["point_gid"], np.arange(5089)) and np.array_equal(ref["cell_gid"], np.arange(9990))
    msh = M.Mesh(ref["points"], ref["cells"].astype(np.int32), "triangle")
    vt = io_vtk.VTKFile(None, str(tmp_path / "out" / "phasefieldx.pvd"), "w")
    vt.write_function(msh, {"phi": ref["phi"], "u": ref["u"][:, :2]}, 149)
    vt.close()
    mine_txt = (tmp_path / "out" / "phasefieldx_p0_000149.vtu").read_text()

    def skeleton(txt):
        out = []
        for tag in re.findall(r"<(/?\w+)([^>]*)>", txt):
            name, attrs = tag
            if name == "DataArray":
                a = dict(re.findall(r'(\w+)="([^"]*)"', attrs))
                out.append(("DataArray", a.get("type"), a.get("Name"), a.get("NumberOfComponents"), a.get("format")))
            elif name == "Piece":
                out.append((name, attrs.strip()))
            elif not name.startswith("?"):
                out.append((name,))
        return out
    assert skeleton(mine_txt) == skeleton(ref_txt)
    back = io_vtk.read_vtu_points_cells(str(tmp_path / "out" / "phasefieldx_p0_000149.vtu"))
    for key in ("points", "cells", "phi", "u"):
        assert np.array_equal(back[key], ref[key]), key
    # the reference stores VTK_LAGRANGE_TRIANGLE (69) cells, offsets 3, 6, ...
    types = re.search(r'Name="types"[^>]*>([^<]*)<', mine_txt).group(1).split()
    assert set(types) == {"69"} == set(re.search(r'Name="types"[^>]*>([^<]*)<', ref_txt).group(1).split())


@pytest.mark.parametrize("ctype", ["triangle", "quadrilateral", "tetrahedron"])
def test_xdmf_time_series_round_trip(tmp_path, ctype):
    """the XDMF output of `save_solution_xdmf=True` (reference solver.py:259-278, 460-462): mesh once,
    one temporal grid per step and field; light data is XDMF 3 XML, heavy data raw binary."""
    import xml.dom.minidom
    from phasefieldx_b200 import io_xdmf, mesh as M
    if ctype == "tetrahedron":
        msh = M.create_box(None, [[0, 0, 0], [1, 1, 1]], [3, 2, 2])
    else:
        msh = M.create_rectangle(None, [[0, 0], [2, 1]], [4, 3], M.CellType.triangle if ctype == "triangle" else M.CellType.quadrilateral)
    d = msh.geometry.dim
    rng = np.random.default_rng(3)
    xf = io_xdmf.XDMFFile(None, str(tmp_path / "paraview-solutions_xdmf" / "phi.xdmf"), "w")
    xf.write_mesh(msh)
    fields = []
    for step in range(3):
        phi, u = rng.uniform(size=msh.num_nodes), rng.normal(size=(msh.num_nodes, d))
        xf.write_function("phi", phi, step)
        xf.write_function("u", u, step)
        fields.append((phi, u))
    xf.close()
    xml.dom.minidom.parse(str(tmp_path / "paraview-solutions_xdmf" / "phi.xdmf"))          # well-formed
    back = io_xdmf.read_xdmf(str(tmp_path / "paraview-solutions_xdmf" / "phi.xdmf"))
    assert np.array_equal(back["points"], msh.geometry.x[:, :d])
    perm = [0, 1, 3, 2] if ctype == "quadrilateral" else list(range(msh.cells.shape[1]))
    assert np.array_equal(back["cells"], msh.cells[:, perm])
    phis = [a for t, n, a in back["steps"] if n == "phi"]
    us = [a for t, n, a in back["steps"] if n == "u"]
    assert len(phis) == 3 and len(us) == 3 and sorted({t for t, _, _ in back["steps"]}) == [0.0, 1.0, 2.0]
    for k, (phi, u) in enumerate(fields):
        assert np.array_equal(phis[k].ravel(), phi)
        assert np.array_equal(us[k][:, :d], u) and (d == 3 or np.all(us[k][:, 2] == 0.0))


def test_energy_controlled_dropins_validate_their_arguments_before_touching_the_gpu():
    """`solver_ener_variational.solve` / `solver_ener_non_variational.solve` keep the reference signature
    (solver_ener_variational.py:41-58) and refuse what the GPU path does not cover before any device work."""
    import inspect
    from phasefieldx_b200 import fem, mesh as M
    from phasefieldx_b200.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx_b200.Element.Phase_Field_Fracture.solver import solver_ener_non_variational, solver_ener_variational
    ref_names = ["Data", "msh", "final_gamma", "V_u", "V_Φ", "bc_list_u", "bc_list_phi", "f_list_u", "T_list_u", "ds_bound", "dtau",
                 "dtau_min", "dtau_max", "path", "bcs_list_u_names", "c1", "c2", "threshold_gamma_save"]
    for mod in (solver_ener_variational, solver_ener_non_variational):
        sig = inspect.signature(mod.solve)
        assert list(sig.parameters)[: len(ref_names)] == ref_names
        assert sig.parameters["dtau"].default == 0.0005 and sig.parameters["dtau_min"].default == 1e-12
        assert sig.parameters["dtau_max"].default == 0.1 and sig.parameters["c1"].default == 1.0
    msh = M.create_rectangle(None, [[0, 0], [1, 1]], [2, 2], M.CellType.quadrilateral)
    V_u, V_phi = fem.functionspace(msh, ("Lagrange", 1, (2,))), fem.functionspace(msh, ("Lagrange", 1))
    base = dict(E=1.0, nu=0.3, Gc=1.0, l=0.1, degradation="isotropic", split_energy="no", degradation_function="quadratic",
                irreversibility="no", fatigue=False, k=0.0, save_solution_vtu=False, results_folder_name="unused")
    with pytest.raises(ValueError, match="T_list_u"):
        solver_ener_variational.solve(Input(**base), msh, 1.0, V_u, V_phi, [], [], None, None)
    top = M.locate_entities_boundary(msh, 1, lambda x: np.isclose(x[1], 1.0))
    T = [[fem.Constant(msh, np.array([0.0, 1.0])), fem.Measure(msh, top)]]
    with pytest.raises(NotImplementedError, match="hybrid"):
        solver_ener_non_variational.solve(Input(**dict(base, degradation="hybrid", split_energy="spectral")), msh, 1.0, V_u, V_phi, [], [], None, T)
    with pytest.raises(OverflowError):
        solver_ener_variational.solve(Input(**dict(base, degradation_function="sargado")), msh, 1.0, V_u, V_phi, [], [], None, T)


def test_vtu_keep_last_bounds_the_files_on_disk(tmp_path):
    """`VTKFile(keep_last=k)` / `solve(..., vtu_keep=k)`: only the k most recent step files stay on disk (what bench.py
    uses for the 1.5 GB-per-step files of the 25 M-dof beam); the .pvd index lists exactly those; default keeps all."""
    from phasefieldx_b200 import io_vtk, mesh as M
    msh = M.create_rectangle(None, [[0, 0], [1, 1]], [3, 3])
    for keep, expect in ((2, [4, 5]), (None, [0, 1, 2, 3, 4, 5])):
        d = tmp_path / f"k{keep}"
        vt = io_vtk.VTKFile(None, str(d / "phasefieldx.pvd"), "w", fmt="binary", keep_last=keep)
        for step in range(6):
            vt.write_function(msh, {"phi": np.full(msh.num_nodes, float(step))}, step, background=True)
        vt.close()
        files = sorted(f for f in os.listdir(d) if f.endswith(".vtu"))
        assert files == [f"phasefieldx_p0_{s:06d}.vtu" for s in expect]
        pvd = (d / "phasefieldx.pvd").read_text()
        assert [f in pvd for f in files] == [True] * len(files) and pvd.count("<DataSet") == len(expect)
        last = io_vtk.read_vtu_points_cells(str(d / files[-1]))
        assert np.all(last["phi"] == 5.0)
