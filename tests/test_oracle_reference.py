"""Pins the oracle (CPU restatement of the staggered solver) against the reference's own
known-answer tests and stored example outputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.solver import BC, Oracle, Params
from phasefieldx_b200 import mesh as M


def _one_element(split, deg, cycle, nsteps=200):
    msh = M.create_rectangle(None, [np.array([0, 0]), np.array([1, 1])], [1, 1], M.CellType.quadrilateral)
    P = Params(E=210.0, nu=0.3, Gc=0.005, l=0.1, degradation=deg, split_energy=split, irreversibility="no")
    o = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P)
    x = msh.geometry.x
    top = np.nonzero(np.isclose(x[:, 1], 1))[0]
    bot = np.nonzero(np.isclose(x[:, 1], 0))[0]
    bt, bb = BC((top[:, None] * 2 + np.arange(2)).ravel()), BC((bot[:, None] * 2 + np.arange(2)).ravel())
    out = []
    for step in range(nsteps):
        val = cycle(float(step))
        bt.values = np.tile([0.0, val], len(top))
        o.load_step([bt, bb], [], max_stagger_iter=500)
        out.append((val, o.reaction(bt)[1]))
    return P, np.array(out)


def test_reference_one_element_isotropic():
    """reference test/Element/Phase_Field_Fracture/test_phase_field_fracture_element.py:124-131"""
    cyc = lambda t: 0.0003 * t if t <= 50 else (-0.0003 * (t - 50) + 0.015 if t <= 100 else 0.0003 * (t - 100))
    P, res = _one_element("no", "isotropic", cyc)
    d = res[:, 0]
    psi_t = 0.5 * d**2 * (P.lambda_ + 2 * P.mu)
    phi_t = 2 * psi_t / (P.Gc / P.l + 2 * psi_t)
    sigma_t = d * (P.lambda_ + 2 * P.mu) * (1 - phi_t) ** 2
    np.testing.assert_allclose(res[:, 1], sigma_t, rtol=1e-3, atol=1e-8)
    np.testing.assert_allclose(res[:, 1], sigma_t, rtol=1e-9, atol=1e-12)     # the oracle is far tighter


def test_reference_one_element_spectral():
    """reference test_phase_field_fracture_element_anisotropic.py:128-152"""
    cyc = lambda t: 0.0003 * t if t <= 50 else (-0.0003 * (t - 50) + 0.015 if t <= 150 else 0.0003 * (t - 150) - 0.015)
    P, res = _one_element("spectral", "anisotropic", cyc)
    d = res[:, 0]
    ep, en = 0.5 * (d + np.abs(d)), 0.5 * (d - np.abs(d))
    psi_t = 0.5 * ep**2 * (P.lambda_ + 2 * P.mu)
    phi_t = 2 * psi_t / (P.Gc / P.l + 2 * psi_t)
    sigma_t = ep * (P.lambda_ + 2 * P.mu) * (1 - phi_t) ** 2 + en * (P.lambda_ + 2 * P.mu)
    np.testing.assert_allclose(res[:, 1], sigma_t, rtol=1e-3, atol=1e-8)


def test_stored_example_1711_first_steps(c1_mesh):
    """Config 1: stored outputs of reference examples/PhaseFieldFracture/1711_* — reaction to 1e-9,
    PSI_a to 1e-12, stagger counts and iterU = 0 exactly, gamma_phi at the reference's own
    projection-solver noise (KSP rtol 1e-5)."""
    msh, bot, top = c1_mesh
    P = Params(E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no", irreversibility="miehe")
    o = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P)
    bb, bt = BC((bot[:, None] * 2 + np.arange(2)).ravel()), BC(top * 2 + 1)
    ref = np.load(os.path.join(GOLDEN, "ref_1711.npz"))
    for step in range(4):
        bt.values[:] = 1e-4 * step
        r = o.load_step([bt, bb], [], max_stagger_iter=500)
        R, e = o.reaction(bb), o.energies()
        assert r["stagger"] == int(ref["conv"][step, 1])
        assert r["u_its"] == int(ref["conv"][step, 3]) and r["phi_its"] == int(ref["conv"][step, 2])
        if step:
            assert abs(R[1] - ref["bottom_reaction"][step, 2]) < 1e-9 * abs(ref["bottom_reaction"][step, 2])
            assert abs(e["PSI_a"] - ref["energy"][step, 9]) < 1e-12 * ref["energy"][step, 9]
            assert abs(e["E"] - ref["energy"][step, 2]) < 1e-8 * ref["energy"][step, 2]
            assert abs(e["gamma_phi"] - ref["energy"][step, 7]) < 1e-5 * ref["energy"][step, 7]
        assert abs(R[0]) < 1e-15


def test_stored_example_1713_q1_first_step():
    """Structured 200x100 Q1 mesh of reference plot_1713.py:93-124 (no mesh file needed): step-1
    reaction of the stored bottom.reaction."""
    msh = M.create_rectangle(None, [np.array([0, 0]), np.array([1.0, 0.5])], [200, 100], M.CellType.quadrilateral)
    P = Params(E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no", irreversibility="miehe")
    o = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P)
    fdim = 1
    top_f = M.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 0.5))
    bot_f = M.locate_entities_boundary(msh, fdim, lambda x: np.logical_and(np.isclose(x[1], 0), np.greater(x[0], 0.5)))
    top, bot = M.facet_nodes(msh, top_f), M.facet_nodes(msh, bot_f)
    bt, bb = BC((top[:, None] * 2 + np.arange(2)).ravel()), BC(bot * 2 + 1)       # bc_xy(top), bc_y(bottom)
    ref = np.load(os.path.join(GOLDEN, "ref_1713.npz"))
    bt.values = np.tile([0.0, float(ref["dof"][1, 2])], len(top))
    assert float(ref["dof"][1, 2]) == 0.5e-4
    o.load_step([bt, bb], [], max_stagger_iter=500)
    Ry = o.reaction(bb)[1]
    stored = ref["bottom_reaction"][1, 2]
    assert abs(Ry - stored) < 1e-8 * abs(stored), (Ry, stored)


def test_oracle_l2_projection_modes_agree_to_solver_noise(c1_mesh):
    """H1: the reference's KSP-default projection (rtol 1e-5) differs from the exact one by ~1e-6."""
    msh, bot, top = c1_mesh
    P = Params(irreversibility="miehe")
    o1 = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P, projection="exact")
    o2 = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P, projection="petsc_default")
    rng = np.random.default_rng(0)
    u = rng.normal(size=o1.nn * 2) * 1e-4
    a, b = o1.project(o1.split(u, False)["psi_a"]), o2.project(o2.split(u, False)["psi_a"])
    err = np.linalg.norm(a - b) / np.linalg.norm(a)
    assert 0 < err < 1e-3


def test_stored_fatigue_example_1800_first_steps(c1_mesh):
    """Config 3 (reference examples/Fatigue/plot_1800.py:102-116, 223-264): the fatigue bookkeeping
    (alpha_c projection, heaviside(.,1), per-step carry-over, Fatigue field inside the phi form)
    against the stored total.energy: alpha_acum to ~2e-6 (the reference's own
    projection-solver noise), PSI_a to 1e-9.  (The 1800 *.reaction files predate the current
    reaction routine and are not pins.)"""
    from phasefieldx_b200 import workloads as W
    msh, bot, top = c1_mesh
    w = W.sen_tension(msh, bot, top, fatigue=True)
    o = Oracle(msh.geometry.x, msh.cells, msh.cell_type, Params(**w.params))
    bcs = [BC(d) for _, d in w.bc_sets]
    ref = np.load(os.path.join(GOLDEN, "ref_1800.npz"))
    for step in range(4):
        for bc, v in zip(bcs, w.bc_values(float(step))):
            bc.values = v
        assert abs(bcs[0].values[0] - ref["dof"][step, 2]) < 1e-15
        r = o.load_step(bcs, [], dt=1.0, max_stagger_iter=500)
        e = o.energies()
        assert r["stagger"] == int(ref["conv"][step, 1])
        if step:
            assert abs(e["alpha_acum"] - ref["energy"][step, 11]) < 5e-6 * ref["energy"][step, 11]
            assert abs(e["PSI_a"] - ref["energy"][step, 9]) < 1e-9 * ref["energy"][step, 9]
