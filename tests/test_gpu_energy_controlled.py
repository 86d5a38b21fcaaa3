"""GPU parity of the energy-controlled monolithic solvers (pfx_ec_solve through the C-ABI, SURVEY §8f N1 / BASELINE
config 4) against the CPU oracle (oracle/energy_controlled.py, itself pinned on the reference's stored three-point
results): same Newton iteration counts (they drive the d-tau path), u and phi within 1e-8 relative L2, lambda, gamma
and the support reaction within 1e-8 / 1e-6 relative, on Q1 quadrilaterals (the reference example), P1 triangles with
the spectral split, and tetrahedra; then the drop-in `solve()` modules against the oracle's d-tau loop."""
import os

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _setup(msh, ctype, params, sets, traction, variational, c1, c2, **opts):
    from oracle.energy_controlled import EnergyControlled
    from oracle.solver import BC, Params
    from phasefieldx_b200 import _lib
    P = Params(**params)
    o = EnergyControlled(msh.geometry.x, msh.cells, ctype, P, variational=variational, c1=c1, c2=c2)
    split = None if params["split_energy"] == "no" else params["split_energy"]
    e = _lib.Engine(msh.geometry.x, msh.cells, ctype, P.lambda_, P.mu, P.Gc, P.l, k=P.k, model=_lib.model_code(params["degradation"], split), **opts)
    bcs = [BC(d) for _, d in sets]
    for i, bc in enumerate(bcs):
        e.set_bc_u(i, bc.dofs)
        e.set_bc_values_u(i, bc.values)
    nodes, w, T = traction
    o.set_traction(nodes, w, T)
    e.set_traction(0, nodes, w)
    e.set_traction_value(0, T)
    return o, e, bcs


def _march(o, e, bcs, taus, c1, c2, variational, **ec):
    lam = 0.0
    for k, tau in enumerate(taus):
        its_o, hist = o.newton(tau, bcs)
        rep = e.ec_solve(tau, lam, c1=c1, c2=c2, variational=variational, **ec)
        assert rep["converged"], rep
        lam = rep["lam"]
        assert rep["its"] == its_o, (k, rep["its"], its_o, hist)
        assert rel_l2(e.get_field("u"), o.u) < 1e-8, (k, rel_l2(e.get_field("u"), o.u))
        assert np.linalg.norm(e.get_field("phi") - o.phi) <= 1e-8 * np.linalg.norm(o.phi) + 1e-14, k
        assert abs(lam - o.lam) <= 1e-8 * max(abs(o.lam), 1e-3), (k, lam, o.lam)
        assert abs(rep["gamma"] - o.gamma()) <= 1e-8 * abs(o.gamma()) + 1e-20
        assert abs(rep["external"] - o.external_energy()) <= 1e-8 * abs(o.external_energy())
        Rg, Ro = e.reaction(0), o.reaction(bcs[0])
        assert np.all(np.abs(Rg - Ro) <= 1e-6 * np.abs(Ro).max() + 1e-14), (k, Rg, Ro)
        eg, eo = e.energies(), o.energies()
        for key in ("E", "gamma", "PSI_a", "W"):
            assert abs(eg[key] - eo[key]) <= 1e-7 * abs(eo[key]) + 1e-20, (k, key)


@pytest.mark.parametrize("variational", [True, False])
def test_three_point_bending_quads_vs_oracle(variational):
    """Config 4 geometry (reference plot_three_point_[non_]var.py) at oracle size, marching tau far enough for the
    crack to grow from the notch: in the variational run the load multiplier lambda falls from 1 to 0.16 (the
    snap-back branch a displacement- or load-controlled solver cannot follow)."""
    from phasefieldx_b200 import workloads as W
    msh, params, sets, traction, kw = W.three_point_bending(80, 20, variational)
    params = dict(params, l=0.12)
    o, e, bcs = _setup(msh, "quadrilateral", params, sets, traction, variational, kw["c1"], kw["c2"])
    taus = [0.002, 0.05, 0.5, 1.3, 2.5, 4.0, 6.0, 8.0] if variational else [0.002, 0.03, 0.06, 0.09, 0.12]
    _march(o, e, bcs, taus, kw["c1"], kw["c2"], variational)
    assert o.phi.max() > 0.2 and (not variational or o.lam < 0.2)
    e.close()


def test_triangles_spectral_two_level_vs_oracle(monkeypatch):
    """P1 triangles, spectral split (nonlinear u-block), mesh large enough for the two-level PCG inside the
    preconditioner."""
    from phasefieldx_b200 import fem, mesh as M
    monkeypatch.setenv("PFX_NO_COOP", "1")
    msh = M.create_rectangle(None, [[0, 0], [2.0, 1.0]], [48, 24])
    x = msh.geometry.x
    bot = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
    sets = [("bottom", (bot[:, None] * 2 + np.arange(2)).ravel())]
    nodes, w = fem.Measure(msh, M.locate_entities_boundary(msh, 1, lambda p: np.isclose(p[1], 1.0) & (p[0] > 0.5) & (p[0] < 1.5))).nodal_weights()
    params = dict(E=210.0, nu=0.3, Gc=0.0027, l=0.08, degradation="anisotropic", split_energy="spectral",
                  degradation_function="quadratic", irreversibility="no", fatigue=False, k=0.0)
    o, e, bcs = _setup(msh, "triangle", params, sets, (nodes, w, np.array([0.3, 1.0])), True, 2.0, 1.0, block_nodes=32)
    _march(o, e, bcs, [1e-3, 4e-3, 1e-2, 2e-2], 2.0, 1.0, True)
    assert e.coarse_stats()["active"] == 1
    e.close()


def test_tetrahedra_vs_oracle():
    from phasefieldx_b200 import fem, mesh as M
    msh = M.create_box(None, [[0, 0, 0], [2.0, 1.0, 0.5]], [8, 4, 2])
    x = msh.geometry.x
    bot = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
    sets = [("bottom", (bot[:, None] * 3 + np.arange(3)).ravel())]
    nodes, w = fem.Measure(msh, M.locate_entities_boundary(msh, 2, lambda p: np.isclose(p[1], 1.0))).nodal_weights()
    params = dict(E=210.0, nu=0.3, Gc=0.0027, l=0.3, degradation="anisotropic", split_energy="deviatoric",
                  degradation_function="quadratic", irreversibility="no", fatigue=False, k=0.0)
    o, e, bcs = _setup(msh, "tetrahedron", params, sets, (nodes, w, np.array([0.0, 1.0, 0.2])), False, 1.0, 1.0)
    _march(o, e, bcs, [1e-3, 5e-3, 1.5e-2], 1.0, 1.0, False)
    e.close()


@pytest.mark.parametrize("scheme", ["variational", "non_variational"])
def test_dropin_solve_matches_the_oracle_dtau_loop(scheme, tmp_path):
    """`solver_ener_[non_]variational.solve` with the reference signature: the files it writes (total.energy with the
    `external` column, lambda.dof, tau.input, <bc>.reaction, phasefieldx.conv) against the oracle's d-tau loop,
    including one rejected step (tiny max_it makes a solve fail -> rollback, d-tau shrinks by the golden ratio)."""
    import importlib
    from oracle.energy_controlled import EnergyControlled
    from oracle.solver import BC, Params
    from phasefieldx_b200 import fem, workloads as W
    from phasefieldx_b200.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx_b200.PostProcessing.ReferenceResult import AllResults
    var = scheme == "variational"
    mod = importlib.import_module(f"phasefieldx_b200.Element.Phase_Field_Fracture.solver.solver_ener_{scheme}")
    msh, params, sets, (nodes, w, T), kw = W.three_point_bending(60, 15, var)
    params = dict(params, l=0.15)
    Data = Input(save_solution_vtu=True, save_solution_xdmf=False, results_folder_name=str(tmp_path / "res"), **params)
    V_u = fem.functionspace(msh, ("Lagrange", 1, (2,)))
    V_phi = fem.functionspace(msh, ("Lagrange", 1))
    bcs = []
    for _, dofs in sets:
        bc = fem.DirichletBC(0.0, np.zeros(0, np.int64), V_phi)
        bc._dofs, bc._comp, bc._vals = dofs.astype(np.int32), None, np.zeros(len(dofs))
        bc.values_per_dof = (lambda b=bc: b._vals)
        bcs.append(bc)
    facets = fem  # noqa: F841
    from phasefieldx_b200 import mesh as M
    ft = M.locate_entities_boundary(msh, 1, lambda x: np.logical_and(np.isclose(x[1], 1.0), np.greater_equal(x[0], -0.075)))
    ds_top = fem.Measure(msh, ft)
    T_top = fem.Constant(msh, T)
    nsteps = 9
    mod.solve(Data, msh, kw["final_gamma"], V_u, V_phi, bcs, [], None, [[T_top, ds_top]], None, dtau=kw["dtau"], dtau_min=kw["dtau_min"],
              dtau_max=kw["dtau_max"], path=str(tmp_path), bcs_list_u_names=[n for n, _ in sets], c1=kw["c1"], c2=kw["c2"],
              threshold_gamma_save=0.0, max_steps=nsteps)
    o = EnergyControlled(msh.geometry.x, msh.cells, "quadrilateral", Params(**params), variational=var, c1=kw["c1"], c2=kw["c2"])
    o.set_traction(nodes, w, T)
    obcs = [BC(d) for _, d in sets]
    recs, reac = [], []
    o.run(kw["final_gamma"], obcs, dtau=kw["dtau"], dtau_min=kw["dtau_min"], dtau_max=kw["dtau_max"], max_steps=nsteps,
          callback=lambda r: (recs.append(r), reac.append(o.reaction(obcs[0])[1])))
    S = AllResults(str(tmp_path / "res"))
    en = S.energy_files["total.energy"]
    assert len(en["gamma"]) == nsteps
    np.testing.assert_allclose(np.asarray(en["external"]), [r["external"] for r in recs], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(en["E"]), [r["energies"]["E"] for r in recs], rtol=1e-7)
    np.testing.assert_allclose(np.asarray(en["gamma"]), [r["gamma"] for r in recs], rtol=1e-7, atol=1e-20)
    np.testing.assert_allclose(np.asarray(S.dof_files["lambda.dof"]["lambda"]), [r["lam"] for r in recs], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(S.reaction_files["bottom_left.reaction"]["Ry"]), reac, rtol=1e-6)
    conv = np.loadtxt(os.path.join(str(tmp_path / "res"), "phasefieldx.conv"), skiprows=1)
    assert [int(v) for v in conv[:, 1]] == [r["iterations"] for r in recs]
    assert os.path.exists(os.path.join(str(tmp_path / "res"), "paraview-solutions_vtu", f"phasefieldx_p0_{nsteps:06d}.vtu"))
