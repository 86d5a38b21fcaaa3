"""GPU parity (through the C-ABI): every operator kernel against the CPU oracle on seeded
random states, all cell types x energy splits x fatigue.  Tolerance: fp64 round-off (1e-11)."""
import numpy as np
import pytest

from conftest import MODELS, perturbed_meshes, rel

pytestmark = pytest.mark.gpu


def _setup(ctype, m, deg, sp, fat, rng, block_nodes=0, k=1e-3):
    from oracle.solver import Oracle, Params
    from phasefieldx_b200 import _lib
    P = Params(E=210.0, nu=0.3, Gc=0.0027, l=0.1, degradation=deg, split_energy=sp, irreversibility="miehe",
               fatigue=fat, k=k)
    o = Oracle(m.geometry.x, m.cells, m.cell_type, P)
    nn, d = o.nn, o.d
    o.u = rng.normal(size=nn * d) * 1e-3
    o.phi = rng.uniform(0, 1, nn)
    o.H = rng.uniform(0, 2, nn)
    o.Fatigue = rng.uniform(0.2, 1, nn) if fat else np.ones(nn)
    o.abar_c = rng.uniform(0, 1, nn)
    e = _lib.Engine(m.geometry.x, m.cells, ctype, P.lambda_, P.mu, P.Gc, P.l, k=P.k,
                    model=_lib.model_code(deg, None if sp == "no" else sp), irreversibility=True, fatigue=fat,
                    block_nodes=block_nodes)
    e.set_field("u", o.u)
    e.set_field("phi", o.phi)
    e.set_field("H", o.H)
    if fat:
        e.set_field("fatigue", o.Fatigue)
        e.set_field("alpha_cum_bar_c", o.abar_c)
    return o, e, P


@pytest.mark.parametrize("deg,sp", MODELS)
@pytest.mark.parametrize("fat", [False, True])
@pytest.mark.parametrize("path", ["assembled", "matrix_free"])
def test_operators_match_oracle(deg, sp, fat, path, monkeypatch):
    """Every operator of the load step against the assembled oracle matrices, on every cell type — once with the
    operators assembled per linearisation where the engine does that (triangles, quadrilaterals, tetrahedra: sliced-ELL
    SpMV) and once matrix-free everywhere (PFX_NO_SPARSE=1: the pipeline / block-scatter kernels)."""
    if path == "matrix_free":
        monkeypatch.setenv("PFX_NO_SPARSE", "1")
    else:
        monkeypatch.setenv("PFX_TRI_SPARSE", "1")          # triangles are matrix-free by default
    rng = np.random.default_rng(7)
    for ctype, m in perturbed_meshes(rng):
        o, e, P = _setup(ctype, m, deg, sp, fat, rng, block_nodes=64)
        nn, d = o.nn, o.d
        v = rng.normal(size=nn * d)
        p = rng.normal(size=nn)
        J = o.J_u(o.u, o.phi)
        K = o.K_phi(o.H, o.Fatigue)
        tol = 1e-11
        assert rel(e.apply_u(v), J @ v) < tol, (ctype, "apply_u")
        assert rel(e.residual_u(), o.F_int_u(o.u, o.phi)) < tol, (ctype, "residual_u")
        assert rel(e.diag_u(), J.diagonal()) < tol, (ctype, "diag_u")
        assert rel(e.apply_phi(p), K @ p) < tol, (ctype, "apply_phi")
        assert rel(e.apply_mass(p), o.M @ p) < tol, (ctype, "apply_mass")
        assert rel(e.rhs_psi(), o.rhs_psi(o.u)) < tol, (ctype, "rhs_psi")
        # Dirichlet rows/cols = identity
        bc_nodes = np.nonzero(m.geometry.x[:, 1] < 0.05)[0]
        dofs = (bc_nodes[:, None] * d + np.arange(d)).ravel()
        e.set_bc_u(0, dofs)
        mask = np.zeros(nn * d, bool)
        mask[dofs] = True
        assert rel(e.apply_u(v, use_bc=True), o.apply_u(v, mask)) < tol, (ctype, "apply_u bc")
        # norms and energies
        ub = o.u + rng.normal(size=nn * d) * 1e-5
        e.set_field("u_old", ub)
        assert abs(e.l2_error("u", "u_old") - o.l2_error_normalized(o.u, ub, True)) < 1e-12
        pb = o.phi + rng.normal(size=nn) * 1e-3
        e.set_field("phi_old", pb)
        assert abs(e.l2_error("phi", "phi_old") - o.l2_error_normalized(o.phi, pb, False)) < 1e-12
        en_g, en_o = e.energies(), o.energies()
        for key, val in en_o.items():
            assert abs(en_g[key] - val) <= 1e-11 * max(abs(val), 1e-30), (ctype, key, en_g[key], val)
        e.close()


def test_apply_is_deterministic_and_block_size_independent():
    """Atomic-free scatter: bit-identical results run to run; round-off-level across tilings."""
    rng = np.random.default_rng(3)
    from phasefieldx_b200 import mesh as M
    m = M.create_rectangle(None, [[0, 0], [1, 1]], [60, 50])
    outs = []
    for bn in (64, 64, 128, 240):
        rng2 = np.random.default_rng(11)
        o, e, P = _setup("triangle", m, "anisotropic", "spectral", False, rng2, block_nodes=bn)
        v = np.random.default_rng(5).normal(size=o.nn * 2)
        outs.append(e.apply_u(v))
        e.close()
    assert np.array_equal(outs[0], outs[1])
    assert rel(outs[2], outs[0]) < 1e-13 and rel(outs[3], outs[0]) < 1e-13


def test_large_mesh_linearity_and_symmetry():
    """Size-independent properties at a size the oracle cannot assemble quickly:
    J(u,phi) is linear in v and symmetric (v.Jw = w.Jv) for the spectral split."""
    from phasefieldx_b200 import mesh as M, _lib
    n = 400
    m = M.create_rectangle(None, [[-0.5, -0.5], [0.5, 0.5]], [n, n])
    rng = np.random.default_rng(0)
    nn = m.num_nodes
    e = _lib.Engine(m.geometry.x, m.cells, "triangle", 121.15, 80.77, 0.0027, 0.015, model=1)
    x = m.geometry.x
    e.set_field("u", np.stack([1e-3 * x[:, 1] * (1 + x[:, 0]), 5e-4 * x[:, 0] * x[:, 1]], 1).ravel())
    e.set_field("phi", np.clip(np.exp(-np.abs(x[:, 1]) / 0.015), 0, 1))
    v, w = rng.uniform(-1, 1, nn * 2), rng.uniform(-1, 1, nn * 2)
    Jv, Jw = e.apply_u(v), e.apply_u(w)
    Jvw = e.apply_u(2.0 * v - 3.0 * w)
    assert rel(Jvw, 2.0 * Jv - 3.0 * Jw) < 1e-12
    assert abs(w @ Jv - v @ Jw) < 1e-11 * abs(w @ Jv)
    e.close()


@pytest.mark.parametrize("deg,sp", [("anisotropic", "spectral"), ("anisotropic", "deviatoric"), ("hybrid", "spectral")])
def test_projected_energy_parts_match_oracle(deg, sp):
    """pfx_project_psi: nodal L2 projections of psi_a and psi_b (solver_anisotropic.py:236-237 output fields)
    against the oracle's exact mass solve, every cell type."""
    rng = np.random.default_rng(13)
    for ctype, m in perturbed_meshes(rng):
        o, e, P = _setup(ctype, m, deg, sp, False, rng, block_nodes=64)
        s = o.split(o.u, False)
        for which in ("a", "b"):
            ref = o.project(s["psi_" + which])
            got = e.project_psi(which)
            assert np.linalg.norm(got - ref) <= 1e-9 * max(np.linalg.norm(ref), 1e-30), (ctype, which)
        e.close()
