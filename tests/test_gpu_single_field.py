"""The single-field drop-in solvers: the reference's own phase-field profile test run
unchanged in structure through `phasefieldx.Element.Phase_Field.solver.solver.solve` (reference
test/Element/Phase_Field/test_phase_field_element.py:15-124, the 2D case; intervals and hexahedra are
not built), and the elasticity solver against the CPU oracle's direct solve."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_phase_field_profile_reference_test(tmp_path):
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    import mpi4py
    from phasefieldx.Element.Phase_Field.Input import Input
    from phasefieldx.Element.Phase_Field.solver.solver import solve
    from phasefieldx.Boundary.boundary_conditions import bc_phi, get_ds_bound_from_marker
    from phasefieldx.PostProcessing.ReferenceResult import AllResults
    from phasefieldx_b200.io_vtk import read_vtu_points_cells

    lx, ly, divx, divy = 5.0, 1.0, 200, 1
    msh = dolfinx.mesh.create_rectangle(mpi4py.MPI.COMM_WORLD, [np.array([0.0, 0.0]), np.array([lx, ly])], [divx, divy],
                                        cell_type=dolfinx.mesh.CellType.quadrilateral)
    Data = Input(l=4.0, save_solution_xdmf=False, save_solution_vtu=True, results_folder_name=str(tmp_path / "test_results"))

    def left(x):
        return np.equal(x[0], 0)
    fdim = msh.topology.dim - 1
    left_facet_marker = dolfinx.mesh.locate_entities_boundary(msh, fdim, left)
    ds_left = get_ds_bound_from_marker(left_facet_marker, msh, fdim)
    ds_list = np.array([[ds_left, "left"]])
    V_phi = dolfinx.fem.functionspace(msh, ("Lagrange", 1))
    bcs_list_phi = [bc_phi(left_facet_marker, V_phi, fdim, value=1.0)]
    solve(Data, msh, 1.0, V_phi, bcs_list_phi, None, None, ds_list, 1.0, path=str(tmp_path), quadrature_degree=2)

    out = read_vtu_points_cells(str(tmp_path / "test_results" / "paraview-solutions_vtu" / "phasefieldx_p0_000000.vtu"))
    xt = out["points"][:, 0]
    phi_theory = np.exp(-abs(xt) / Data.l) + 1 / (np.exp(2 * lx / Data.l) + 1) * 2 * np.sinh(np.abs(xt) / Data.l)
    assert np.allclose(out["phi"], phi_theory, atol=1e-8)          # the reference's tolerance (rtol 1e-5)
    S = AllResults(Data.results_folder_name)
    a = lx / Data.l
    th = np.tanh(a)
    W_phi, W_grad = 0.5 * th + 0.5 * a * (1.0 - th**2), 0.5 * th - 0.5 * a * (1.0 - th**2)
    en = S.energy_files["total.energy"]
    np.testing.assert_allclose(en["gamma_phi"][0], 0.5 * W_phi, rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(en["gamma_gradphi"][0], 0.5 * W_grad, rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(en["gamma"][0], 0.5 * th, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("ctype", ["triangle", "tetrahedron"])
def test_elasticity_solver_vs_oracle(tmp_path, ctype):
    """Drop-in elasticity solve() (reference Element/Elasticity/solver/solver.py) against the oracle's
    direct solve of the same problem: u to 1e-8, reactions to 1e-6, elastic energy to 1e-7."""
    from oracle.solver import BC, Oracle, Params
    from phasefieldx_b200 import fem, mesh as M
    from phasefieldx_b200.Boundary.boundary_conditions import bc_xy, bc_xyz
    from phasefieldx_b200.Element.Elasticity.Input import Input
    from phasefieldx_b200.Element.Elasticity.solver.solver import solve
    from phasefieldx_b200.PostProcessing.ReferenceResult import AllResults
    from phasefieldx_b200.io_vtk import read_vtu_points_cells

    if ctype == "triangle":
        msh = M.create_rectangle(None, [[0, 0], [2, 1]], [24, 12])
        d, bc_fix = 2, bc_xy
    else:
        msh = M.create_box(None, [[0, 0, 0], [2, 1, 0.5]], [10, 5, 3])
        d, bc_fix = 3, bc_xyz
    fdim = msh.topology.dim - 1
    V_u = fem.functionspace(msh, ("Lagrange", 1, (d,)))
    left = M.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[0], 0.0))
    right = M.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[0], 2.0))
    bcl, bcr = bc_fix(left, V_u, fdim), bc_fix(right, V_u, fdim)

    def update(bcs, t):
        bcs[1].g.value[...] = 0.0
        bcs[1].g.value[0] = 1e-3 * t
        bcs[1].g.value[1] = 5e-4 * t
        return 1e-3 * t, 5e-4 * t, 0
    Data = Input(E=210.0, nu=0.3, save_solution_vtu=True, results_folder_name=str(tmp_path / "res"))
    solve(Data, msh, 3.0, V_u, [bcl, bcr], update, None, None, None, None, 1.0, path=str(tmp_path),
          bcs_list_u_names=["left", "right"])

    P = Params(E=210.0, nu=0.3, Gc=1.0, l=1.0, k=0.0)
    o = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P)
    dl, dr = fem.extract_bc(bcl)[0], fem.extract_bc(bcr)[0]
    vr = np.zeros(len(dr))
    vr[dr % d == 0] = 2e-3
    vr[dr % d == 1] = 1e-3
    o.solve_u([BC(dl), BC(dr, vr)])
    out = read_vtu_points_cells(str(tmp_path / "res" / "paraview-solutions_vtu" / "phasefieldx_p0_000002.vtu"))
    u = out["u"][:, :d].ravel()
    assert np.linalg.norm(u - o.u) <= 1e-8 * np.linalg.norm(o.u)
    S = AllResults(Data.results_folder_name)
    Ro = o.reaction(BC(dr, vr))
    Rg = np.array([S.reaction_files["right.reaction"][c][2] for c in ("Rx", "Ry", "Rz")])
    assert np.all(np.abs(Rg[:d] - Ro[:d]) <= 1e-6 * np.abs(Ro).max())
    E = S.energy_files["total.energy"]["E"][2]
    assert abs(E - o.energies()["E"]) <= 1e-7 * abs(o.energies()["E"])
    assert list(S.convergence_files["phasefieldx.conv"]["iterations"]) == [0, 1, 1]


@pytest.mark.parametrize("split_energy_i", ["spectral", "deviatoric"])
def test_anisotropic_decomposition_reference_test(tmp_path, split_energy_i):
    """Reference test/Element/Phase_Field_Fracture/test_anisotropic_decomposition.py:15-197 run unchanged in
    structure through the drop-in solver_anisotropic.solve: traction-loaded 10x10 quads through a load
    reversal; PSI_a + PSI_b must equal the elastic energy of the unsplit law at every step (rtol 1e-3)."""
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    import mpi4py
    import petsc4py
    from phasefieldx.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx.Element.Phase_Field_Fracture.solver.solver_anisotropic import solve
    from phasefieldx.Boundary.boundary_conditions import bc_xy, get_ds_bound_from_marker
    from phasefieldx.Loading.loading_functions import loading_Txy
    from phasefieldx.PostProcessing.ReferenceResult import AllResults
    from phasefieldx_b200.io_vtk import read_vtu_points_cells

    Data = Input(E=210.0, nu=0.3, Gc=0.0027, l=0.06, degradation="anisotropic", split_energy=split_energy_i,
                 degradation_function="quadratic", irreversibility="miehe", fatigue=False,
                 fatigue_degradation_function="asymptotic", fatigue_val=0.0, k=0.0, save_solution_xdmf=False,
                 save_solution_vtu=True, results_folder_name=str(tmp_path / "check"))
    divx, divy, lx, ly = 10, 10, 1.0, 1.0
    msh = dolfinx.mesh.create_rectangle(mpi4py.MPI.COMM_WORLD, [np.array([0, 0]), np.array([lx, ly])], [divx, divy],
                                        cell_type=dolfinx.mesh.CellType.quadrilateral)
    fdim = msh.topology.dim - 1
    bottom_facet_marker = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 0))
    top_facet_marker = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], ly))
    ds_bottom = get_ds_bound_from_marker(bottom_facet_marker, msh, fdim)
    ds_top = get_ds_bound_from_marker(top_facet_marker, msh, fdim)
    ds_list = np.array([[ds_bottom, "bottom"]])
    V_u = dolfinx.fem.functionspace(msh, ("Lagrange", 1, (msh.geometry.dim,)))
    bcs_list_u = [bc_xy(bottom_facet_marker, V_u, fdim)]
    T_list_u = [[loading_Txy(msh), ds_top]]

    def update_boundary_conditions(bcs, time):
        return 0, 0, 0

    def update_loading(T_list_u, time):
        if time <= 50:
            val = 0.3 * time
        elif time <= 100:
            val = -0.3 * (time - 50) + 0.3 * 50
        else:
            val = 0.3 * (time - 100)
        T_list_u[0][0].value[0] = petsc4py.PETSc.ScalarType(val)
        T_list_u[0][0].value[1] = petsc4py.PETSc.ScalarType(val)
        return val, val, 0
    solve(Data, msh, 200.0, V_u, bcs_list_u, update_boundary_conditions, None, T_list_u, update_loading, ds_list, 10.0,
          path=str(tmp_path), quadrature_degree=2, bcs_list_u_names=["bottom"])
    S = AllResults(Data.results_folder_name)
    en = S.energy_files["total.energy"]
    np.testing.assert_allclose(np.asarray(en["PSI_a"]) + np.asarray(en["PSI_b"]), np.asarray(en["E"]), rtol=1e-3, atol=1e-8)
    assert np.asarray(en["E"])[5] > 0 and len(en["E"]) == 20
    # reactions balance the applied traction (total force = T * lx) at the peak load
    R = S.reaction_files["bottom.reaction"]
    assert abs(R["Rx"][5] + 15.0 * lx) < 1e-6 * 15.0 and abs(R["Ry"][5] + 15.0 * lx) < 1e-6 * 15.0
    # projected energy fields are written next to u and integrate to the reported totals (mass-lumped check)
    out = read_vtu_points_cells(str(tmp_path / "check" / "paraview-solutions_vtu" / "phasefieldx_p0_000005.vtu"))
    assert out["u"].shape == ((divx + 1) * (divy + 1), 3)
    txt = (tmp_path / "check" / "paraview-solutions_vtu" / "phasefieldx_p0_000005.vtu").read_text()
    assert 'Name="PSI_a"' in txt and 'Name="PSI_b"' in txt


def test_reaction_forces_2d_quadrilateral_equals_3d_hexahedron_reference_test(tmp_path):
    """reference test/Reactions/test_reaction_forces.py:16-178: the same uniaxial stretch on one Q1 quadrilateral and
    on one Q1 hexahedron (Elasticity solver, bc_xy / bc_xyz on bottom and top) gives the same bottom reaction R_y
    (rtol 1e-5 there)."""
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    import mpi4py
    import petsc4py
    from phasefieldx.Element.Elasticity.Input import Input
    from phasefieldx.Element.Elasticity.solver.solver import solve
    from phasefieldx.Boundary.boundary_conditions import bc_xy, bc_xyz, get_ds_bound_from_marker
    from phasefieldx.PostProcessing.ReferenceResult import AllResults

    def run(dim):
        Data = Input(E=210.0, nu=0.3, save_solution_xdmf=False, save_solution_vtu=True, results_folder_name=str(tmp_path / f"{dim}_dimension"))
        if dim == 2:
            msh = dolfinx.mesh.create_rectangle(mpi4py.MPI.COMM_WORLD, [np.array([0, 0]), np.array([1.0, 1.0])], [1, 1],
                                                cell_type=dolfinx.mesh.CellType.quadrilateral)
        else:
            msh = dolfinx.mesh.create_box(mpi4py.MPI.COMM_WORLD, [np.array([0, 0, 0]), np.array([1.0, 1.0, 1.0])], [1, 1, 1],
                                          cell_type=dolfinx.mesh.CellType.hexahedron)
        fdim = msh.topology.dim - 1
        bottom = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 0))
        top = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 1.0))
        ds_list = np.array([[get_ds_bound_from_marker(top, msh, fdim), "top"], [get_ds_bound_from_marker(bottom, msh, fdim), "bottom"]])
        V_u = dolfinx.fem.functionspace(msh, ("Lagrange", 1, (msh.geometry.dim,)))
        if dim == 2:
            bcs = [bc_xy(top, V_u, fdim, value_x=0.0, value_y=0.0), bc_xy(bottom, V_u, fdim, value_x=0.0, value_y=0.0)]
        else:
            bcs = [bc_xyz(top, V_u, fdim, value_x=0.0, value_y=0.0, value_z=0.0), bc_xyz(bottom, V_u, fdim, value_x=0.0, value_y=0.0, value_z=0.0)]

        def update_boundary_conditions(bcs_, time):
            val = 0.0003 * time
            bcs_[0].g.value[1] = petsc4py.PETSc.ScalarType(val)
            return 0, val, 0
        solve(Data, msh, 10.0, V_u, bcs, update_boundary_conditions, None, None, None, ds_list, 1.0, path=str(tmp_path),
              quadrature_degree=2, bcs_list_u_names=["top", "bottom"])
        return AllResults(Data.results_folder_name)
    S2, S3 = run(2), run(3)
    r2, r3 = np.asarray(S2.reaction_files["bottom.reaction"]["Ry"]), np.asarray(S3.reaction_files["bottom.reaction"]["Ry"])
    assert len(r2) == 10 and abs(r2[-1]) > 0
    np.testing.assert_allclose(r2, r3, rtol=1e-5, atol=1e-8)
    # plane strain == fully clamped lateral faces: R_y = (lambda + 2 mu) * strain on the unit face
    lam, mu = 210.0 * 0.3 / (1.3 * 0.4), 210.0 / 2.6
    np.testing.assert_allclose(np.abs(r3), (lam + 2 * mu) * 0.0003 * np.arange(10), rtol=1e-9, atol=1e-12)
    import re
    from phasefieldx_b200.io_vtk import read_vtu_points_cells
    txt = open(str(tmp_path / "3_dimension" / "paraview-solutions_vtu" / "phasefieldx_p0_000009.vtu")).read()
    assert 'NumberOfCells="1"' in txt
    assert re.search(r'Name="types"[^>]*>([^<]*)<', txt).group(1).split() == ["72"]          # VTK_LAGRANGE_HEXAHEDRON
    back = read_vtu_points_cells(str(tmp_path / "3_dimension" / "paraview-solutions_vtu" / "phasefieldx_p0_000009.vtu"))
    assert back["cells"].shape == (1, 8) and sorted(back["cells"][0].tolist()) == list(range(8))


def test_phase_field_AT1_profile(tmp_path):
    """`case='AT1'` of the Phase_Field solver (reference Element/Phase_Field/solver/solver.py:116-117 with
    geometric_crack.py:22-23, 47-48, 72-73: w = phi, c0 = 8/3): the linear boundary value problem
    (2 l / c0) phi'' = 1 / (c0 l) on a strip with phi(0) = 1 and a natural boundary at x = L has the parabola
    phi = 1 - L x / (2 l^2) + x^2 / (4 l^2) as its solution (nodally exact for linear elements on a uniform strip);
    gamma = (1/(c0 l)) int phi + (l/c0) int phi'^2 by the same case.  WU / DOUBLE are refused."""
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    import mpi4py
    from phasefieldx.Element.Phase_Field.Input import Input
    from phasefieldx.Element.Phase_Field.solver.solver import solve
    from phasefieldx.Boundary.boundary_conditions import bc_phi, get_ds_bound_from_marker
    from phasefieldx.PostProcessing.ReferenceResult import AllResults
    from phasefieldx_b200.io_vtk import read_vtu_points_cells
    lx, ly, divx, divy, l = 2.0, 0.5, 160, 2, 1.5
    msh = dolfinx.mesh.create_rectangle(mpi4py.MPI.COMM_WORLD, [np.array([0.0, 0.0]), np.array([lx, ly])], [divx, divy],
                                        cell_type=dolfinx.mesh.CellType.quadrilateral)
    Data = Input(l=l, save_solution_xdmf=False, save_solution_vtu=True, results_folder_name=str(tmp_path / "at1"))
    fdim = msh.topology.dim - 1
    left = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.equal(x[0], 0))
    ds_list = np.array([[get_ds_bound_from_marker(left, msh, fdim), "left"]])
    V_phi = dolfinx.fem.functionspace(msh, ("Lagrange", 1))
    bcs = [bc_phi(left, V_phi, fdim, value=1.0)]
    solve(Data, msh, 1.0, V_phi, bcs, None, None, ds_list, 1.0, path=str(tmp_path), quadrature_degree=2, case="AT1")
    out = read_vtu_points_cells(str(tmp_path / "at1" / "paraview-solutions_vtu" / "phasefieldx_p0_000000.vtu"))
    x = out["points"][:, 0]
    phi = 1.0 - lx * x / (2 * l * l) + x * x / (4 * l * l)
    assert np.abs(out["phi"] - phi).max() < 1e-8
    en = AllResults(Data.results_folder_name).energy_files["total.energy"]
    c0 = 8.0 / 3.0
    g_phi = ly / (c0 * l) * (lx - lx**3 / (4 * l * l) + lx**3 / (12 * l * l))
    g_grad = ly * l / c0 * (lx**3 / (12 * l**4))                  # int (x - L)^2 / (4 l^4)
    np.testing.assert_allclose(en["gamma_phi"][0], g_phi, rtol=1e-4)
    np.testing.assert_allclose(en["gamma_gradphi"][0], g_grad, rtol=1e-3)
    np.testing.assert_allclose(en["gamma"][0], g_phi + g_grad, rtol=1e-4)
    with pytest.raises(NotImplementedError, match="WU"):
        solve(Data, msh, 1.0, V_phi, bcs, None, None, ds_list, 1.0, path=str(tmp_path), quadrature_degree=2, case="WU")
