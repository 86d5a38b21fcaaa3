"""The energy-controlled oracle (oracle/energy_controlled.py) pinned on the reference: its one-element analytic tests
(test/Element/Phase_Field_Fracture/test_phase_field_fracture_element_ener_[non_]variational.py) and the stored outputs
of the three-point bending examples (tests/golden/ref_energy_controlled.npz, from
examples/PhaseFieldFractureEnergyControlled/results_three_point_[non_]var)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _one_element(variational):
    from oracle.energy_controlled import EnergyControlled
    from oracle.solver import BC, Params
    from phasefieldx_b200 import fem, mesh as M
    msh = M.create_rectangle(None, [np.array([0, 0]), np.array([1, 1])], [1, 1], M.CellType.quadrilateral)
    P = Params(E=210.0, nu=0.3, Gc=0.005, l=0.1)
    o = EnergyControlled(msh.geometry.x, msh.cells, "quadrilateral", P, variational=variational, c1=1.0, c2=1.0)
    x = msh.geometry.x
    top = np.nonzero(np.isclose(x[:, 1], 1))[0]
    bot = np.nonzero(np.isclose(x[:, 1], 0))[0]
    bcs = [BC(top * 2 + 0), BC((bot[:, None] * 2 + np.arange(2)).ravel())]
    nodes, w = fem.Measure(msh, M.locate_entities_boundary(msh, 1, lambda p: np.isclose(p[1], 1.0))).nodal_weights()
    o.set_traction(nodes, w, np.array([0.0, 1.0]))
    return o, bcs, P


@pytest.mark.parametrize("variational", [True, False])
def test_one_element_reference_test(variational):
    """reference test_phase_field_fracture_element_ener_variational.py:113-127 (and the non-variational twin):
    reaction * alpha == analytic stress of the homogeneous AT2 solution at every step, rtol 1e-3 there."""
    o, bcs, P = _one_element(variational)
    rows = []

    def cb(rec):
        rows.append((rec["lam"], rec["energies"]["E"], o.reaction(bcs[1])[1], rec["energies"]["gamma"]))
    recs = o.run(3.5, bcs, dtau=1e-4, dtau_min=1e-6, dtau_max=0.1, callback=cb)
    assert recs[-1]["gamma"] >= 3.5 and len(recs) > 100
    lam, E, Ry, gamma = (np.array(c) for c in zip(*rows))
    alpha = 1 / np.sqrt(1 + o.c1 * lam / P.Gc) if variational else np.ones_like(lam)
    disp = np.abs(2 * E / Ry) * alpha
    psi = 0.5 * disp**2 * (P.lambda_ + 2 * P.mu)
    phi = 2 * psi / (P.Gc / P.l + 2 * psi)
    sigma = disp * (P.lambda_ + 2 * P.mu) * (1 - phi) ** 2
    np.testing.assert_allclose(np.abs(Ry * alpha), sigma, rtol=1e-3, atol=1e-8)
    # the constraint c1 gamma + c2 external = tau holds at every converged step
    for r in recs:
        assert abs(o.c1 * r["gamma"] + o.c2 * r["external"] - r["tau"]) < 1e-9


@pytest.mark.parametrize("scheme,steps", [("var", 3), ("nonvar", 1)])
def test_three_point_bending_reproduces_the_stored_reference_rows(scheme, steps):
    """400 x 100 Q1 cells as in the reference example: external work (= tau path, i.e. the Newton iteration counts
    that drive the dtau growth), E, gamma, the support reaction and lambda of the first steps agree with the stored
    files to 1e-10 relative (measured: 1e-13)."""
    from oracle.energy_controlled import EnergyControlled
    from oracle.solver import BC, Params
    from phasefieldx_b200 import workloads as W
    var = scheme == "var"
    msh, params, sets, (nodes, w, T), kw = W.three_point_bending(400, 100, var)
    o = EnergyControlled(msh.geometry.x, msh.cells, "quadrilateral", Params(**params), variational=var, c1=kw["c1"], c2=kw["c2"])
    bcs = [BC(d) for _, d in sets]
    o.set_traction(nodes, w, T)
    ref = np.load(os.path.join(GOLDEN, "ref_energy_controlled.npz"))
    en, re = ref[scheme + "_energy"], ref[scheme + "_reaction"]
    seen = []

    def cb(rec):
        k = rec["step"] - 1
        e = rec["energies"]
        assert abs(rec["external"] - en[k, 11]) <= 1e-10 * abs(en[k, 11])
        assert abs(e["E"] - en[k, 2]) <= 1e-10 * abs(en[k, 2])
        assert abs(e["gamma"] - en[k, 6]) <= 1e-10 * abs(en[k, 6])
        assert abs(e["PSI_a"] - en[k, 9]) <= 1e-10 * abs(en[k, 9])
        assert abs(o.reaction(bcs[0])[1] - re[k, 2]) <= 1e-10 * abs(re[k, 2])
        if var:
            assert abs(rec["lam"] - ref["var_lambda"][k, 1]) <= 1e-12
        seen.append(k)
    o.run(kw["final_gamma"], bcs, dtau=kw["dtau"], dtau_min=kw["dtau_min"], dtau_max=kw["dtau_max"], max_steps=steps, callback=cb)
    assert seen == list(range(steps))


def test_central_cracked_plate_reproduces_the_stored_reference_row():
    """Third stored example (plot_central_cracked_non_var.py:82-131, 208-245: 100 x 300 Q1 cells, the crack is the free
    part x < a0 of the bottom edge, non-variational scheme, c1 = c2 = 1, dtau = 1e-4): external work, E, gamma and the
    bottom reaction of the first row to 1e-10 relative (measured: 1e-14) — gamma is already 5e-9 there, so the
    phase-field side of the blocked system is pinned too."""
    from oracle.energy_controlled import EnergyControlled
    from oracle.solver import BC, Params
    from phasefieldx_b200 import fem, mesh as M
    divx, divy, lx, ly, a0 = 100, 300, 1.0, 3.0, 0.3
    msh = M.create_rectangle(None, [np.array([0.0, 0.0]), np.array([lx, ly])], [divx, divy], M.CellType.quadrilateral)
    o = EnergyControlled(msh.geometry.x, msh.cells, "quadrilateral", Params(E=210.0, nu=0.3, Gc=0.0027, l=0.025), variational=False)
    fb = M.locate_entities_boundary(msh, 1, lambda x: np.logical_and(np.isclose(x[1], 0), np.greater_equal(x[0], a0)))
    fl = M.locate_entities_boundary(msh, 1, lambda x: np.isclose(x[0], 0.0))
    ft = M.locate_entities_boundary(msh, 1, lambda x: np.isclose(x[1], ly))
    bcs = [BC(M.facet_nodes(msh, fb) * 2 + 1), BC(M.facet_nodes(msh, fl) * 2 + 0)]
    nodes, w = fem.Measure(msh, ft).nodal_weights()
    o.set_traction(nodes, w, np.array([0.0, 1.0]))
    ref = np.load(os.path.join(GOLDEN, "ref_energy_controlled.npz"))
    en, re = ref["cc_energy"], ref["cc_reaction"]
    rec = o.run(0.5, bcs, dtau=0.0001, dtau_min=1e-12, dtau_max=1.0, max_steps=1)[0]
    e = rec["energies"]
    assert abs(rec["external"] - en[0, 11]) <= 1e-10 * abs(en[0, 11])
    assert abs(e["E"] - en[0, 2]) <= 1e-10 * abs(en[0, 2])
    assert abs(e["gamma"] - en[0, 6]) <= 1e-10 * abs(en[0, 6])
    assert abs(e["gamma_gradphi"] - en[0, 8]) <= 1e-10 * abs(en[0, 8])
    assert abs(o.reaction(bcs[0])[1] - re[0, 2]) <= 1e-10 * abs(re[0, 2])
