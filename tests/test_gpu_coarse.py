"""GPU tests of the aggregate coarse space of the PCG (pfx_coarse.cuh): it only changes how fast
the linear solves converge, never what they converge to — load steps with it agree with the CPU
oracle (direct solves) and with the Jacobi-only PCG to the same 1e-8 / 1e-6 bars as every other
parity test, on every cell type, while needing several times fewer iterations."""
import numpy as np
import pytest

from conftest import rel_l2
from test_gpu_load_step import _compare_step, _pair

pytestmark = pytest.mark.gpu


@pytest.fixture
def no_coop(monkeypatch):
    # small test meshes would otherwise take the single-launch cooperative PCG (Jacobi)
    monkeypatch.setenv("PFX_NO_COOP", "1")
    return monkeypatch


def _shear_mesh(n):
    from phasefieldx_b200 import mesh as M
    msh = M.create_rectangle(None, [[-0.5, -0.5], [0.5, 0.5]], [n, n])
    return M.cut_slit(msh, lambda x: np.isclose(x[1], 0.0) & (x[0] < -1e-9), lambda c: c[1] > 0)


@pytest.mark.parametrize("tri_tangent_cache", [False, True])
def test_coarse_space_load_steps_vs_oracle(no_coop, tri_tangent_cache):
    """Config-2-like slit square (reference plot_1712.py), spectral split, coarse space active; also with
    the opt-in cached-element-tangent variant of the persistent apply kernel (PFX_TRI_TANGENT_CACHE=1)."""
    from oracle.solver import BC
    no_coop.setenv("PFX_TRI_TANGENT_CACHE", "1" if tri_tangent_cache else "0")
    msh = _shear_mesh(56)
    o, e, P = _pair(msh, "triangle", "anisotropic", "spectral", l=0.06, block_nodes=64)
    st = e.coarse_stats()
    assert st["active"] == 1 and st["aggregates"] >= 32
    x = msh.geometry.x
    top = np.nonzero(np.isclose(x[:, 1], 0.5))[0]
    bot = np.nonzero(np.isclose(x[:, 1], -0.5))[0]
    side = np.nonzero((np.isclose(x[:, 0], -0.5) | np.isclose(x[:, 0], 0.5)) & (np.abs(x[:, 1]) < 0.5 - 1e-9))[0]
    bcs = [BC((top[:, None] * 2 + np.arange(2)).ravel()), BC((bot[:, None] * 2 + np.arange(2)).ravel()), BC(side * 2 + 1)]
    for i, bc in enumerate(bcs):
        e.set_bc_u(i, bc.dofs)
        e.set_bc_values_u(i, bc.values)
    for step in range(1, 4):
        bcs[0].values = np.tile([2e-3 * step, 0.0], len(top))
        e.set_bc_values_u(0, bcs[0].values)
        ro = o.load_step(bcs, [], max_stagger_iter=60)
        rg = e.load_step(max_stagger_iter=60)
        assert rg["stagger"] == ro["stagger"], (step, rg["stagger"], ro["stagger"])
        _compare_step(o, e, bcs, step)
    assert e.coarse_stats()["setups"] >= 1
    e.close()


def _run(msh, ctype, deg, sp, bc_sets, load, steps, monkeypatch, coarse, **kw):
    from phasefieldx_b200 import _lib
    from oracle.solver import Params
    monkeypatch.setenv("PFX_NO_COARSE", "0" if coarse else "1")
    P = Params(degradation=deg, split_energy=sp, **kw)
    e = _lib.Engine(msh.geometry.x, msh.cells, ctype, P.lambda_, P.mu, P.Gc, P.l, k=P.k,
                    model=_lib.model_code(deg, None if sp == "no" else sp), block_nodes=64)
    assert e.coarse_stats()["active"] == (1 if coarse else 0)
    for i, dofs in enumerate(bc_sets):
        e.set_bc_u(i, dofs)
    its = 0
    for s in range(1, steps + 1):
        for i, vals in enumerate(load(s)):
            e.set_bc_values_u(i, vals)
        r = e.load_step(max_stagger_iter=60)
        its += r["cg_its_u"] + r["cg_its_phi"]
    out = dict(u=e.get_field("u"), phi=e.get_field("phi"), H=e.get_field("H"), R=e.reaction(0), its=its)
    e.close()
    return out


@pytest.mark.parametrize("ctype", ["quadrilateral", "tetrahedron", "triangle_isotropic"])
def test_coarse_space_matches_jacobi_on_every_cell_type(ctype, no_coop):
    from phasefieldx_b200 import mesh as M
    if ctype == "tetrahedron":
        msh = M.create_box(None, [[0, 0, 0], [3.0, 1.0, 0.8]], [20, 10, 8])
        x = msh.geometry.x
        left = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
        right = np.nonzero(np.isclose(x[:, 0], 3.0))[0]
        sets = [right * 3 + 1, (left[:, None] * 3 + np.arange(3)).ravel()]
        load = lambda s: [np.full(len(right), 4e-3 * s), np.zeros(3 * len(left))]
        args = ("tetrahedron", "anisotropic", "spectral")
        kw = dict(E=0.6, nu=0.2, Gc=0.00013, l=0.3)
    else:
        if ctype == "quadrilateral":
            msh = M.create_rectangle(None, [[0, 0], [1, 1]], [60, 60], M.CellType.quadrilateral)
            args = ("quadrilateral", "anisotropic", "deviatoric")
        else:
            msh = _shear_mesh(60)
            args = ("triangle", "isotropic", "no")
        x = msh.geometry.x
        top = np.nonzero(np.isclose(x[:, 1], x[:, 1].max()))[0]
        bot = np.nonzero(np.isclose(x[:, 1], x[:, 1].min()))[0]
        sets = [(top[:, None] * 2 + np.arange(2)).ravel(), (bot[:, None] * 2 + np.arange(2)).ravel()]
        load = lambda s: [np.tile([1e-3 * s, 1.5e-3 * s], len(top)), np.zeros(2 * len(bot))]
        kw = dict(l=0.05)
    a = _run(msh, *args, sets, load, 2, no_coop, True, **kw)
    b = _run(msh, *args, sets, load, 2, no_coop, False, **kw)
    for f in ("u", "phi", "H"):
        assert rel_l2(a[f], b[f]) < 1e-8, (f, rel_l2(a[f], b[f]))
    assert np.all(np.abs(a["R"] - b["R"]) <= 1e-6 * np.abs(b["R"]).max())
    assert a["its"] * 2 < b["its"], (a["its"], b["its"])


def test_coarse_space_with_tractions_phi_dirichlet_and_fatigue(no_coop):
    """The less common ingredients together — Neumann load, a phase-field Dirichlet seed (masked
    constants in the phi coarse space), hybrid split, fatigue degradation in J_phi — against the CPU
    oracle with the coarse space active (mesh large enough for >= 32 blocks of 64 nodes)."""
    from phasefieldx_b200 import mesh as M, fem
    from oracle.solver import BC
    msh = M.create_rectangle(None, [[0, 0], [1.0, 1.0]], [52, 52])
    o, e, P = _pair(msh, "triangle", "hybrid", "spectral", l=0.05, Gc=0.0027, fatigue=True, fatigue_val=0.002, block_nodes=64)
    assert e.coarse_stats()["active"] == 1
    x = msh.geometry.x
    bot = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
    seed = np.nonzero(np.isclose(x[:, 1], 0.5) & (x[:, 0] < 0.3))[0]
    top_facets = M.locate_entities_boundary(msh, 1, lambda p: np.isclose(p[1], 1.0))
    nodes, w = fem.Measure(msh, top_facets).nodal_weights()
    bcs_u = [BC((bot[:, None] * 2 + np.arange(2)).ravel())]
    bcs_phi = [BC(seed, np.ones(len(seed)))]
    e.set_bc_u(0, bcs_u[0].dofs)
    e.set_bc_values_u(0, bcs_u[0].values)
    e.set_bc_phi(0, seed)
    e.set_bc_values_phi(0, bcs_phi[0].values)
    e.set_traction(0, nodes, w)
    for step in range(1, 4):
        T = np.array([0.02 * step, 0.05 * step]) * (1 if step != 2 else 0.5)      # load, partial unload, reload
        e.set_traction_value(0, T)
        fext = o.traction_vector(nodes, w, T)
        ro = o.load_step(bcs_u, bcs_phi, max_stagger_iter=80, fext=fext)
        rg = e.load_step(max_stagger_iter=80)
        assert rg["stagger"] == ro["stagger"], (step, rg["stagger"], ro["stagger"])
        _compare_step(o, e, bcs_u, step, names=("u", "phi", "H", "fatigue", "alpha_cum_bar_c"))
        assert np.allclose(e.get_field("phi")[seed], 1.0)
    assert np.allclose(e.reaction(0)[:2], -T, rtol=1e-6, atol=1e-10)
    e.close()


@pytest.mark.parametrize("n", [37, 64, 333, 1500])
def test_in_tree_dense_spd_inverse_matches_lapack(n):
    """pfx_dense.cuh (blocked Cholesky + L^-1 + Y^T Y: the inverse of the coarse matrices, in place of the cuSOLVER
    potrf / potri of round 1) against numpy on random SPD matrices with a 1e6 condition number; a matrix that is not
    positive definite raises the info flag."""
    from phasefieldx_b200 import _lib
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    A = (Q * np.logspace(0, 6, n)) @ Q.T
    A = 0.5 * (A + A.T)
    Ainv, info = _lib.dense_spd_inverse(A)
    assert info == 0
    ref = np.linalg.inv(A)
    assert np.abs(Ainv - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(Ainv @ A - np.eye(n)).max() < 1e-7
    B = A.copy()
    B[n // 2, n // 2] = -1.0
    _, info = _lib.dense_spd_inverse(B)
    assert info == 1
