"""Pins the oracle's constitutive layer against (a) the reference's own code executed unmodified
under a torch `ufl` stand-in (committed golden vectors, tests/golden/make_golden.py), and (b)
mathematical identities the reference's tests use (psi_a + psi_b = psi,
reference test_anisotropic_decomposition.py:167-182)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel
from oracle import constitutive as cons

CASES = (("isotropic", "no"), ("anisotropic", "spectral"), ("anisotropic", "deviatoric"), ("hybrid", "spectral"),
         ("hybrid", "deviatoric"))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_constitutive.npz"))


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("deg,sp", CASES)
def test_oracle_matches_reference_constitutive_code(golden, d, deg, sp):
    lam, mu = golden["lam_mu"]
    G, dG = golden[f"G{d}"], golden[f"dG{d}"]
    A, dA = 0.5 * (G + G.transpose(0, 2, 1)), 0.5 * (dG + dG.transpose(0, 2, 1))
    key = f"{d}_{deg}_{sp}"
    r = cons.split_numpy(A, lam, mu, deg, sp)
    tol = 1e-13
    assert rel(r["psi_a"], golden[f"psi_a_{key}"]) < tol
    assert rel(r["sig_a"], golden[f"sig_a_{key}"]) < tol
    if np.abs(golden[f"psi_b_{key}"]).max() > 0:
        assert rel(r["psi_b"], golden[f"psi_b_{key}"]) < tol
    if np.abs(golden[f"sig_b_{key}"]).max() > 0:
        assert rel(r["sig_b"], golden[f"sig_b_{key}"]) < tol
    nd = slice(3, None)      # rows 0-2 are exactly degenerate states: tangent is rounding noise there
    dsa = np.einsum("nijkl,nkl->nij", r["C_a"], dA)
    assert rel(dsa[nd], golden[f"dsig_a_{key}"][nd]) < 1e-11
    if np.abs(golden[f"dsig_b_{key}"]).max() > 0:
        dsb = np.einsum("nijkl,nkl->nij", r["C_b"], dA)
        assert rel(dsb[nd], golden[f"dsig_b_{key}"][nd]) < 1e-11


@pytest.mark.parametrize("d", [2, 3])
def test_torch_ad_matches_closed_forms(d):
    rng = np.random.default_rng(0)
    G = rng.normal(size=(30, d, d)) * 1e-3
    A = 0.5 * (G + G.transpose(0, 2, 1))
    for deg, sp in CASES:
        a = cons.split_numpy(A, 121.15, 80.77, deg, sp)
        b = cons.split_torch(A, 121.15, 80.77, deg, sp)
        for k in a:
            assert rel(a[k], b[k]) < 1e-12 or np.abs(b[k]).max() == 0, (deg, sp, k)


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("sp", ["spectral", "deviatoric"])
def test_energy_split_is_complete(d, sp):
    """psi_a + psi_b = psi and sigma_a + sigma_b = sigma."""
    rng = np.random.default_rng(1)
    G = rng.normal(size=(40, d, d)) * 1e-3
    A = 0.5 * (G + G.transpose(0, 2, 1))
    lam, mu = 121.15, 80.77
    r = cons.split_numpy(A, lam, mu, "anisotropic", sp)
    iso = cons.split_numpy(A, lam, mu, "isotropic", "no")
    assert rel(r["psi_a"] + r["psi_b"], iso["psi_a"]) < 1e-12
    assert rel(r["sig_a"] + r["sig_b"], iso["sig_a"]) < 1e-12


def test_discriminant_table_is_the_true_discriminant():
    """Sum-of-products table (reference Math/invariants.py:105-144) == prod (λi-λj)^2, also for
    non-symmetric matrices; and the code generator carries the same table."""
    import torch
    import importlib.util
    rng = np.random.default_rng(2)
    G = rng.normal(size=(25, 3, 3))
    A = torch.tensor(G)
    Dl = 0
    for spec, w in zip(cons._DX3, cons._DD3):
        Dl = Dl + cons._poly(A, spec, False) * w * cons._poly(A, spec, True)
    ev = np.linalg.eigvals(G)
    disc = np.array([np.prod([(e[i] - e[j]) ** 2 for i in range(3) for j in range(i + 1, 3)]) for e in ev]).real
    assert np.abs(Dl.numpy() - disc).max() < 1e-9 * np.abs(disc).max()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_eig3", os.path.join(root, "tools", "gen_eig3.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert gen.DX3 == cons._DX3 and gen.DD3 == cons._DD3


def test_regulariser_pins_from_stored_step0_rows():
    """Stored step-0 rows of the spectral examples hold PSI_a = PSI_b = |Ω|·ψ_a(ε=0)
    = area·2μ(0.75e-16)²: 1712 on the unit square."""
    ref = np.load(os.path.join(GOLDEN, "ref_1712.npz"))
    mu = 210.0 / (2 * 1.3)
    r = cons.split_numpy(np.zeros((1, 2, 2)), 121.15384615384615, mu, "anisotropic", "spectral", False)
    stored = ref["energy"][0, 9]
    assert stored > 0
    area = stored / r["psi_a"][0]
    assert abs(area - 1.0) < 2e-3          # unit square minus the thin geometric notch
    assert ref["energy"][0, 9] == ref["energy"][0, 10]
