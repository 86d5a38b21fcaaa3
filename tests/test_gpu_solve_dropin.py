"""The reference's own end-to-end tests, run UNCHANGED in structure through the drop-in
`phasefieldx.Element.Phase_Field_Fracture.solver.solver.solve` (aliased onto phasefieldx_b200):
build a mesh, call solve(), read the result files back through AllResults, compare with the
closed form at rtol 1e-3 (reference test/Element/Phase_Field_Fracture/
test_phase_field_fracture_element.py:16-136 and ..._anisotropic.py:17-157)."""
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("split_energy_i", ["no", "spectral"])
def test_phase_field_simulation(tmp_path, split_energy_i):
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    import mpi4py
    import petsc4py
    from phasefieldx.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx.Element.Phase_Field_Fracture.solver.solver import solve
    from phasefieldx.Boundary.boundary_conditions import bc_xy, get_ds_bound_from_marker
    from phasefieldx.PostProcessing.ReferenceResult import AllResults
    from phasefieldx.Math import macaulay_bracket_positive, macaulay_bracket_negative

    aniso = split_energy_i != "no"
    Data = Input(E=210.0, nu=0.3, Gc=0.005, l=0.1, degradation="anisotropic" if aniso else "isotropic",
                 split_energy=split_energy_i, degradation_function="quadratic", irreversibility="no", fatigue=False,
                 fatigue_degradation_function="asymptotic", fatigue_val=0.05625, k=0.0, save_solution_xdmf=False,
                 save_solution_vtu=True, results_folder_name=str(tmp_path / "1700_One_element_test"))
    msh = dolfinx.mesh.create_rectangle(mpi4py.MPI.COMM_WORLD, [np.array([0, 0]), np.array([1, 1])], [1, 1],
                                        cell_type=dolfinx.mesh.CellType.quadrilateral)

    def bottom(x):
        return np.isclose(x[1], 0)

    def top(x):
        return np.isclose(x[1], 1)
    fdim = msh.topology.dim - 1
    bottom_facet_marker = dolfinx.mesh.locate_entities_boundary(msh, fdim, bottom)
    top_facet_marker = dolfinx.mesh.locate_entities_boundary(msh, fdim, top)
    ds_bottom = get_ds_bound_from_marker(top_facet_marker, msh, fdim)
    ds_top = get_ds_bound_from_marker(top_facet_marker, msh, fdim)
    ds_list = np.array([[ds_top, "top"], [ds_bottom, "bottom"]])
    V_u = dolfinx.fem.functionspace(msh, ("Lagrange", 1, (msh.geometry.dim,)))
    V_phi = dolfinx.fem.functionspace(msh, ("Lagrange", 1))
    bc_bottom = bc_xy(bottom_facet_marker, V_u, fdim)
    bc_top = bc_xy(top_facet_marker, V_u, fdim)
    bcs_list_u = [bc_top, bc_bottom]
    bcs_list_u_names = ["top", "bottom"]
    t_mid = 150 if aniso else 100

    def update_boundary_conditions(bcs, time):
        if time <= 50:
            val = 0.0003 * time
        elif time <= t_mid:
            val = -0.0003 * (time - 50) + 0.015
        else:
            val = 0.0003 * (time - t_mid) + (-0.015 if aniso else 0.0)
        bcs[0].g.value[1] = petsc4py.PETSc.ScalarType(val)
        return 0, val, 0
    solve(Data, msh, 200.0, V_u, V_phi, bcs_list_u, [], update_boundary_conditions, None, None, None, ds_list, 1.0,
          path=str(tmp_path), bcs_list_u_names=bcs_list_u_names, min_stagger_iter=2, max_stagger_iter=500,
          stagger_error_tol=1e-8)

    S = AllResults(Data.results_folder_name)
    displacement = S.dof_files["top.dof"]["Uy"]
    ep, en = macaulay_bracket_positive(displacement), macaulay_bracket_negative(displacement)
    if not aniso:
        ep, en = displacement, 0 * displacement
    psi_t = 0.5 * ep**2 * (Data.lambda_ + 2 * Data.mu)
    phi_t = 2 * psi_t / (Data.Gc / Data.l + 2 * psi_t)
    sigma_t = ep * (Data.lambda_ + 2 * Data.mu) * (1 - phi_t) ** 2 + en * (Data.lambda_ + 2 * Data.mu)
    np.testing.assert_allclose(S.reaction_files["top.reaction"]["Ry"], sigma_t, rtol=1e-3, atol=1e-8)
    # result-file contract of the reference (solver.py:419-457)
    assert list(S.convergence_files["phasefieldx.conv"].columns) == ["#step", "stagger", "iterPhi", "iterU"]
    assert list(S.energy_files["total.energy"].columns) == ["#step", "EplusW", "E", "W", "W_phi", "W_gradphi", "gamma",
                                                            "gamma_phi", "gamma_gradphi", "PSI_a", "PSI_b", "alpha_acum"]
    assert len(S.paraview_filenames) == 201 and "phasefieldx.pvd" in S.paraview_filenames
    assert os.path.exists(os.path.join(Data.results_folder_name, "simulation.log"))
    shutil.rmtree(Data.results_folder_name)


def test_config1_through_solve_matches_stored_reference_files(tmp_path, c1_mesh):
    """Config 1 driven exactly like reference examples/PhaseFieldFracture/plot_1711.py:189-320
    (bottom bc_xy = 0, top bc_y = 1e-4 t, stagger 2/500/1e-8): the result FILES written by the
    drop-in solve() reproduce the reference's stored bottom.reaction / total.energy /
    phasefieldx.conv rows (first load steps; tolerances = oracle-vs-stored agreement)."""
    import phasefieldx_b200.compat as compat
    compat.install()
    import dolfinx
    from phasefieldx.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx.Element.Phase_Field_Fracture.solver.solver import solve
    from phasefieldx.Boundary.boundary_conditions import bc_xy, bc_y
    from phasefieldx.PostProcessing.ReferenceResult import AllResults
    from conftest import GOLDEN
    msh, _, _ = c1_mesh
    Data = Input(E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no",
                 degradation_function="quadratic", irreversibility="miehe", fatigue=False, k=0.0,
                 save_solution_xdmf=True, save_solution_vtu=True, results_folder_name=str(tmp_path / "1711"))
    fdim = msh.topology.dim - 1
    bottom = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], -0.5))
    top = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 0.5))
    V_u = dolfinx.fem.functionspace(msh, ("Lagrange", 1, (msh.geometry.dim,)))
    V_phi = dolfinx.fem.functionspace(msh, ("Lagrange", 1))
    bcs = [bc_y(top, V_u, fdim), bc_xy(bottom, V_u, fdim)]

    def update_boundary_conditions(bcs_, time):
        val = 1e-4 * time if time <= 50 else 5e-3 + 1e-5 * (time - 50)
        bcs_[0].g.value[...] = val
        return 0, val, 0
    solve(Data, msh, 5.0, V_u, V_phi, bcs, [], update_boundary_conditions, None, None, None, None, 1.0,
          path=str(tmp_path), bcs_list_u_names=["top", "bottom"], min_stagger_iter=2, max_stagger_iter=500,
          stagger_error_tol=1e-8, vtu_every=4)
    S = AllResults(Data.results_folder_name)
    ref = np.load(os.path.join(GOLDEN, "ref_1711.npz"))
    Ry = S.reaction_files["bottom.reaction"]["Ry"].to_numpy()
    np.testing.assert_allclose(Ry[1:5], ref["bottom_reaction"][1:5, 2], rtol=2e-9)
    np.testing.assert_allclose(S.reaction_files["top.reaction"]["Ry"].to_numpy()[1:5], -ref["bottom_reaction"][1:5, 2], rtol=2e-9)
    en = S.energy_files["total.energy"]
    np.testing.assert_allclose(en["PSI_a"].to_numpy()[1:5], ref["energy"][1:5, 9], rtol=1e-11)
    np.testing.assert_allclose(en["E"].to_numpy()[1:5], ref["energy"][1:5, 2], rtol=1e-7)
    np.testing.assert_allclose(en["gamma_phi"].to_numpy()[1:5], ref["energy"][1:5, 7], rtol=1e-5)   # reference KSP noise (H1)
    conv = S.convergence_files["phasefieldx.conv"]
    assert conv["stagger"].tolist() == ref["conv"][:5, 1].astype(int).tolist()
    assert conv["iterU"].tolist() == [0] * 5 and conv["iterPhi"].tolist() == [0] * 5
    assert S.dof_files["top.dof"]["Uy"].tolist()[:3] == [0.0, 0.0001, 0.0002]
    assert sorted(f for f in S.paraview_filenames if f.endswith(".vtu")) == ["phasefieldx_p0_000000.vtu", "phasefieldx_p0_000004.vtu"]
    # save_solution_xdmf=True (reference solver.py:259-278, 460-462): phi.xdmf / u.xdmf, one temporal grid per load step
    from phasefieldx_b200 import io_vtk, io_xdmf
    xd = os.path.join(Data.results_folder_name, "paraview-solutions_xdmf")
    xphi, xu = io_xdmf.read_xdmf(os.path.join(xd, "phi.xdmf")), io_xdmf.read_xdmf(os.path.join(xd, "u.xdmf"))
    assert [t for t, _, _ in xphi["steps"]] == [0.0, 1.0, 2.0, 3.0, 4.0] and len(xu["steps"]) == 5
    last = io_vtk.read_vtu_points_cells(os.path.join(Data.results_folder_name, "paraview-solutions_vtu", "phasefieldx_p0_000004.vtu"))
    assert np.array_equal(xphi["steps"][4][2].ravel(), last["phi"]) and np.array_equal(xu["steps"][4][2], last["u"])
