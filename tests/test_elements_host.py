"""The exact C++ formulas the CUDA kernels evaluate (pfx_physics.cuh / pfx_elements.cuh, compiled
for the host by tests/native/hostcheck.cpp) against the oracle: pointwise splits + tangent
action, every element integral, and the block (CTA tile) builder driving a CPU mirror of the
scatter kernel's index logic."""
import ctypes

import numpy as np
import pytest

from conftest import MODELS, perturbed_meshes, rel
from oracle import constitutive as cons
from oracle.solver import Oracle, Params

PD = ctypes.POINTER(ctypes.c_double)
PI = ctypes.POINTER(ctypes.c_int32)
CODES = {("isotropic", "no"): (0, 0), ("anisotropic", "spectral"): (1, 1), ("anisotropic", "deviatoric"): (2, 2),
         ("hybrid", "spectral"): (0, 1)}
CT = {"triangle": 0, "quadrilateral": 1, "tetrahedron": 2, "hexahedron": 3}


def dp(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(PD)


def sym_store(A):
    if A.shape[-1] == 2:
        return np.stack([A[..., 0, 0], A[..., 1, 1], A[..., 0, 1]], -1)
    return np.stack([A[..., 0, 0], A[..., 1, 1], A[..., 2, 2], A[..., 0, 1], A[..., 0, 2], A[..., 1, 2]], -1)


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("scale", [1e-2, 1e-4, 1e-6])
@pytest.mark.parametrize("deg,sp", MODELS)
def test_pointwise_split_and_tangent(hostcheck, d, scale, deg, sp):
    rng = np.random.default_rng(10 * d)
    n = 40
    G = rng.normal(size=(n, d, d)) * scale
    A = 0.5 * (G + G.transpose(0, 2, 1))
    A[0] = np.diag([1.0, 0.0, 0.0][:d]) * scale
    if d == 3:
        A[1] = np.diag([1.0, -0.3, -0.3001]) * scale
    G = rng.normal(size=(n, d, d))
    dA = 0.5 * (G + G.transpose(0, 2, 1))
    lam, mu = 121.15, 80.77
    r = cons.split_torch(A, lam, mu, deg, sp)
    ns = 3 if d == 2 else 6
    e, de = np.ascontiguousarray(sym_store(A)), np.ascontiguousarray(sym_store(dA))
    mat = np.array([lam, mu, 0.0027, 0.015, 0.0])
    sm, em = CODES[(deg, sp)]
    for fn in (hostcheck.hc_split, hostcheck.hc_split_dual):
        sa, sb, dsa, dsb = (np.zeros((n, ns)) for _ in range(4))
        pa, pb = np.zeros(n), np.zeros(n)
        fn(d, dp(mat), sm, em, ctypes.c_int64(n), dp(e), dp(de), dp(sa), dp(sb), dp(pa), dp(pb), dp(dsa), dp(dsb))
        tol = 1e-11
        assert rel(sa, sym_store(r["sig_a"])) < tol and rel(pa, r["psi_a"]) < tol
        if np.abs(r["psi_b"]).max() > 0:
            assert rel(pb, r["psi_b"]) < tol and rel(sb, sym_store(r["sig_b"])) < tol
        assert rel(dsa, sym_store(np.einsum("nijkl,nkl->nij", r["C_a"], dA))) < 1e-9
        ref_b = np.einsum("nijkl,nkl->nij", r["C_b"], dA)
        if np.abs(ref_b).max() > 0:
            assert rel(dsb, sym_store(ref_b)) < 1e-9


@pytest.mark.parametrize("deg,sp", MODELS)
@pytest.mark.parametrize("fat", [False, True])
def test_element_integrals_match_oracle(hostcheck, deg, sp, fat):
    rng = np.random.default_rng(1)
    sm, em = CODES[(deg, sp)]
    for ctype, m in perturbed_meshes(rng):
        Pm = Params(E=210.0, nu=0.3, Gc=0.0027, l=0.1, degradation=deg, split_energy=sp, irreversibility="miehe",
                    fatigue=fat, k=1e-3)
        o = Oracle(m.geometry.x, m.cells, m.cell_type, Pm)
        nn, d = o.nn, o.d
        o.u = rng.normal(size=nn * d) * 1e-3
        o.phi, o.H = rng.uniform(0, 1, nn), rng.uniform(0, 2, nn)
        o.Fatigue = rng.uniform(0.2, 1, nn) if fat else np.ones(nn)
        o.V_c, o.abar_c, o.phi_old = rng.uniform(0, 1, nn), rng.uniform(0, 1, nn), rng.uniform(0, 1, nn)
        v, p = rng.normal(size=nn * d), rng.normal(size=nn)
        mat = np.array([Pm.lambda_, Pm.mu, Pm.Gc, Pm.l, Pm.k])
        cells = np.ascontiguousarray(m.cells, dtype=np.int32)
        X = np.ascontiguousarray(m.geometry.x)

        def asm(op, f0=None, f1=None, f2=None, f3=None, n=nn):
            out = np.zeros(n)
            hostcheck.hc_assemble(op, CT[ctype], dp(mat), sm, em, int(fat), 3, ctypes.c_int64(len(cells)), dp(X),
                                  cells.ctypes.data_as(PI), dp(f0), dp(f1), dp(f2), dp(f3), dp(out))
            return out
        F = o.Fatigue
        cm, cg = 2 * o.H + F * Pm.Gc / Pm.l, F * Pm.Gc * Pm.l
        J, K = o.J_u(o.u, o.phi), o.K_phi(o.H, o.Fatigue)
        tol = 1e-11
        assert rel(asm(0, o.phi, o.u, v, n=nn * d), J @ v) < tol
        assert rel(asm(1, o.phi, o.u, n=nn * d), o.F_int_u(o.u, o.phi)) < tol
        assert rel(asm(2, o.phi, o.u, n=nn * d), J.diagonal()) < tol
        # nodal d x d diagonal blocks (block-Jacobi smoother), stored (00, 11[, 22], 01[, 02, 12])
        Jd = J.tocsr()
        pairs = [(i, i) for i in range(d)] + [(0, 1)] + ([(0, 2), (1, 2)] if d == 3 else [])
        nodes = np.arange(nn)
        blk = np.stack([np.asarray(Jd[nodes * d + i, nodes * d + j]).ravel() for i, j in pairs], 1)
        assert rel(asm(14, o.phi, o.u, n=nn * len(pairs)), blk.ravel()) < tol
        assert rel(asm(3, cm, cg, p), K @ p) < tol
        assert rel(asm(4, cm, cg, o.H, o.phi), K @ o.phi - o.b_phi(o.H)) < tol
        assert rel(asm(5, cm, cg), K.diagonal()) < tol
        assert rel(asm(6, p), o.M @ p) < tol
        assert rel(asm(7), o.M.diagonal()) < tol
        assert rel(asm(8, o.u), o.rhs_psi(o.u)) < tol
        gq = cons.g_quadratic(o.at_qp(o.phi_old))
        fe = np.einsum("eq,eq,qa->ea", o.t.W, gq * o.at_qp(o.V_c), o.t.N)
        assert rel(asm(9, o.phi_old, o.V_c), o._vec(fe, o.edof_s, nn)) < tol
        a, b = rng.normal(size=nn * d), rng.normal(size=nn * d)
        s = asm(11, a, b, n=2)
        assert abs(np.sqrt(s[0] / s[1]) - o.l2_error_normalized(a, b, True)) < 1e-12
        s = asm(12, o.u, o.phi, F, o.abar_c, n=8)
        en = o.energies()
        assert abs(s[0] + s[2] - en["E"]) < tol * abs(en["E"]) and abs(s[1] - en["PSI_a"]) < tol * en["PSI_a"]
        assert abs(s[3] / (2 * Pm.l) + s[4] * Pm.l / 2 - en["gamma"]) < tol * en["gamma"]
        if fat:
            assert abs(s[5] * Pm.Gc / (2 * Pm.l) - en["W_phi"]) < tol * en["W_phi"]
            assert abs(s[7] - en["alpha_acum"]) < tol * en["alpha_acum"]


def test_block_builder_drives_scatter_logic(hostcheck):
    """pfx_blocks.cpp (RCB tiles, halo lists, 16-bit local connectivity, padded incidence lists)
    through a CPU mirror of scatter_kernel: result equals the assembled operator, every cell is
    primary in exactly one block."""
    from phasefieldx_b200 import mesh as M
    rng = np.random.default_rng(2)
    cases = [("triangle", M.create_rectangle(None, [[0, 0], [1, 1]], [31, 27])),
             ("quadrilateral", M.create_rectangle(None, [[0, 0], [1, 1]], [23, 19], M.CellType.quadrilateral)),
             ("tetrahedron", M.create_box(None, [[0, 0, 0], [1, 1, 1]], [9, 8, 7])),
             ("hexahedron", M.create_box(None, [[0, 0, 0], [1, 1, 1]], [8, 7, 6], M.CellType.hexahedron))]
    m0 = cases[0][1]
    cases.append(("triangle", M.cut_slit(m0, lambda x: np.isclose(x[1], 13 / 27) & (x[0] < 0.5), lambda c: c[1] > 13 / 27)))
    for ctype, m in cases:
        Pm = Params(degradation="anisotropic", split_energy="spectral", k=1e-4)
        o = Oracle(m.geometry.x, m.cells, m.cell_type, Pm)
        nn, d = o.nn, o.d
        o.u, o.phi = rng.normal(size=nn * d) * 1e-3, rng.uniform(0, 1, nn)
        v = rng.normal(size=nn * d)
        ref = o.J_u(o.u, o.phi) @ v
        mat = np.array([Pm.lambda_, Pm.mu, Pm.Gc, Pm.l, Pm.k])
        cells = np.ascontiguousarray(m.cells, dtype=np.int32)
        X = np.ascontiguousarray(m.geometry.x)
        for max_own, chunk in [(64, 32), (100, 64), (240, 512)]:
            y, dot, st = np.zeros(nn * d), ctypes.c_double(), np.zeros(6, dtype=np.int64)
            rc = hostcheck.hc_blocks_apply_u(CT[ctype], dp(mat), 1, 3, ctypes.c_int64(nn), ctypes.c_int64(nn),
                                             ctypes.c_int64(len(cells)), dp(X), cells.ctypes.data_as(PI), None, max_own,
                                             chunk, dp(o.phi), dp(o.u), dp(v), dp(y), ctypes.byref(dot),
                                             st.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
            assert rc == 0
            assert rel(y, ref) < 1e-12
            assert abs(dot.value - v @ ref) < 1e-11 * abs(v @ ref)
            assert st[1] <= max_own and st[5] == len(cells)


@pytest.mark.parametrize("sp", ["spectral", "deviatoric"])
def test_cached_tangent_apply_matches_direct(hostcheck, sp):
    """SURVEY H4 path: element tangent evaluated once per linearisation, then contracted."""
    rng = np.random.default_rng(4)
    sm = CODES[("anisotropic", sp)][0]
    for ctype, m in perturbed_meshes(rng):
        if ctype in ("quadrilateral", "hexahedron"):
            continue
        Pm = Params(degradation="anisotropic", split_energy=sp, k=1e-3)
        o = Oracle(m.geometry.x, m.cells, m.cell_type, Pm)
        nn, d = o.nn, o.d
        o.u, o.phi = rng.normal(size=nn * d) * 1e-3, rng.uniform(0, 1, nn)
        v = rng.normal(size=nn * d)
        mat = np.array([Pm.lambda_, Pm.mu, Pm.Gc, Pm.l, Pm.k])
        cells = np.ascontiguousarray(m.cells, dtype=np.int32)
        out = np.zeros(nn * d)
        hostcheck.hc_assemble(13, CT[ctype], dp(mat), sm, sm, 0, 3, ctypes.c_int64(len(cells)), dp(m.geometry.x),
                              cells.ctypes.data_as(PI), dp(o.phi), dp(o.u), dp(v), None, dp(out))
        assert rel(out, o.J_u(o.u, o.phi) @ v) < 1e-10


def test_coarse_space_aggregates_colouring_and_regions(hostcheck):
    """pfx_blocks.cpp::build_aggregates (host side of the two-level PCG): aggregates are runs of blocks
    with contiguous, even-aligned chunks; regions are contiguous; the colouring is a distance-2 colouring of
    the aggregate graph (checked against a graph built independently from the cells), which is what makes
    one probing apply per (colour, mode) valid; mode coordinates are centred and scaled by the aggregate size."""
    from phasefieldx_b200 import mesh as M
    m0 = M.create_rectangle(None, [[0, 0], [1, 1]], [60, 45])
    cases = [("triangle", M.cut_slit(m0, lambda x: np.isclose(x[1], 15 / 45) & (x[0] < 0.5), lambda c: c[1] > 15 / 45), 64, 1, 9),
             ("quadrilateral", M.create_rectangle(None, [[0, 0], [2, 1]], [50, 30], M.CellType.quadrilateral), 48, 2, 4),
             ("tetrahedron", M.create_box(None, [[0, 0, 0], [2, 1, 1]], [14, 9, 8]), 64, 1, 7)]
    PI64 = ctypes.POINTER(ctypes.c_int64)
    PF = ctypes.POINTER(ctypes.c_float)
    for ctype, m, max_own, g, reg_aggs in cases:
        nn, d = m.num_nodes, m.geometry.dim
        cells = np.ascontiguousarray(m.cells, dtype=np.int32)
        X = np.ascontiguousarray(m.geometry.x)
        stats = np.zeros(5, dtype=np.int64)
        node_agg, colour, agg_reg = (np.zeros(nn, np.int32) for _ in range(3))
        chunk0, nbrcol, rel = np.zeros(8 * nn, np.int32), np.zeros(nn * 64, np.int32), np.zeros(nn * d, np.float32)
        rc = hostcheck.hc_aggregates(CT[ctype], ctypes.c_int64(nn), ctypes.c_int64(nn), ctypes.c_int64(len(cells)), dp(X),
                                     cells.ctypes.data_as(PI), max_own, g, reg_aggs, 32, stats.ctypes.data_as(PI64),
                                     node_agg.ctypes.data_as(PI), chunk0.ctypes.data_as(PI), colour.ctypes.data_as(PI),
                                     agg_reg.ctypes.data_as(PI), nbrcol.ctypes.data_as(PI), ctypes.c_int64(nbrcol.size),
                                     rel.ctypes.data_as(PF))
        assert rc == 0
        n_agg, S, n_chunks, n_col, n_reg = (int(v) for v in stats)
        colour, agg_reg, chunk0 = colour[:n_agg], agg_reg[:n_agg], chunk0[: n_chunks + 1]
        nbrcol = nbrcol[: n_agg * n_col].reshape(n_agg, n_col)
        assert n_chunks == n_agg * S and node_agg.min() == 0 and node_agg.max() == n_agg - 1
        # chunks: monotone cover of the owned nodes, even starts, at most 32 (+1 for the even rounding) nodes
        assert chunk0[0] == 0 and chunk0[-1] == nn and np.all(np.diff(chunk0) >= 0)
        assert np.all(chunk0[:-1] % 2 == 0) and np.diff(chunk0).max() <= 33
        sizes = np.bincount(node_agg, minlength=n_agg)
        assert np.array_equal(chunk0[::S][:n_agg], np.r_[0, np.cumsum(sizes)[:-1]])
        # regions: contiguous, non-decreasing, at most reg_aggs aggregates each
        assert agg_reg[0] == 0 and agg_reg[-1] == n_reg - 1 and np.all(np.diff(agg_reg) >= 0)
        assert np.bincount(agg_reg).max() <= reg_aggs
        # independent aggregate graph from the cells
        A = np.zeros((n_agg, n_agg), bool)
        ca = node_agg[cells]
        for i in range(ca.shape[1]):
            for j in range(ca.shape[1]):
                A[ca[:, i], ca[:, j]] = True
        np.fill_diagonal(A, False)
        A2 = (A.astype(np.int32) @ A.astype(np.int32) > 0) | A
        np.fill_diagonal(A2, False)
        same = colour[:, None] == colour[None, :]
        assert not np.any(A2 & same), "not a distance-2 colouring"
        assert colour.max() == n_col - 1
        for a in range(n_agg):
            expect = np.full(n_col, -1)
            expect[colour[a]] = a
            nb = np.nonzero(A[a])[0]
            expect[colour[nb]] = nb
            assert np.array_equal(nbrcol[a], expect)
        # mode geometry: centred per aggregate, scaled by the half extent (so within [-2, 2])
        r = rel.reshape(nn, d)
        assert np.abs(r).max() <= 2.0 + 1e-6
        for k in range(d):
            means = np.bincount(node_agg, weights=r[:, k], minlength=n_agg) / sizes
            assert np.abs(means).max() < 1e-5


@pytest.mark.parametrize("deg,sp", [("isotropic", "no"), ("anisotropic", "spectral"), ("hybrid", "spectral")])
@pytest.mark.parametrize("dfun", ["quadratic", "borden"])
def test_general_degradation_element_integrals_match_oracle(hostcheck, deg, sp, dfun):
    """the borden degradation function (reference g_degradation_functions.py:38-64 as coded,
    s = 0) in the element integrals — closed-form simplex mean of g(phi) in the u-forms, quadrature-based
    F_phi / J_phi / diag(J_phi) with dg, ddg — against the oracle (exact on P1 simplices); the general phi
    forms also reproduce the quadratic ones."""
    rng = np.random.default_rng(21)
    sm, em = CODES[(deg, sp)]
    code = {"quadratic": 0, "borden": 1}[dfun]
    for ctype, m in perturbed_meshes(rng):
        if ctype == "quadrilateral":
            continue                              # Q1: the reference's quadrature degree is not recoverable
        Pm = Params(E=210.0, nu=0.3, Gc=0.0027, l=0.1, degradation=deg, split_energy=sp, irreversibility="miehe", k=1e-3,
                    degradation_function=dfun)
        o = Oracle(m.geometry.x, m.cells, m.cell_type, Pm, degree=5)
        nn, d = o.nn, o.d
        o.u = rng.normal(size=nn * d) * 1e-3
        o.phi, o.H = rng.uniform(0, 1, nn), rng.uniform(0, 2, nn)
        o.Fatigue = np.ones(nn)
        v, p = rng.normal(size=nn * d), rng.normal(size=nn)
        mat = np.array([Pm.lambda_, Pm.mu, Pm.Gc, Pm.l, Pm.k])
        cells = np.ascontiguousarray(m.cells, dtype=np.int32)
        X = np.ascontiguousarray(m.geometry.x)

        def asm(op, f0=None, f1=None, f2=None, f3=None, n=nn):
            out = np.zeros(n)
            hostcheck.hc_assemble(op, CT[ctype], dp(mat), sm, em, code << 8, 3, ctypes.c_int64(len(cells)), dp(X),
                                  cells.ctypes.data_as(PI), dp(f0), dp(f1), dp(f2), dp(f3), dp(out))
            return out
        tol = 1e-11
        J = o.J_u(o.u, o.phi)
        assert rel(asm(0, o.phi, o.u, v, n=nn * d), J @ v) < tol
        assert rel(asm(1, o.phi, o.u, n=nn * d), o.F_int_u(o.u, o.phi)) < tol
        assert rel(asm(2, o.phi, o.u, n=nn * d), J.diagonal()) < tol
        Jp = o.J_phi(o.phi, o.H, o.Fatigue)
        assert rel(asm(15, o.phi, o.H, o.Fatigue, p), o.F_phi(o.phi, o.H, o.Fatigue)) < tol
        assert rel(asm(16, o.phi, o.H, o.Fatigue, p), Jp @ p) < tol
        assert rel(asm(17, o.phi, o.H, o.Fatigue, p), Jp.diagonal()) < tol
        s = asm(12, o.u, o.phi, o.Fatigue, o.abar_c, n=8)
        en = o.energies()
        assert abs(s[0] + s[2] - en["E"]) < tol * abs(en["E"])
        # fatigue projection RHS  ∫ g(phi_old) V_c N  (solver.py:363) with the general g
        phi_old, Vc = rng.uniform(0, 1, nn), rng.uniform(0, 1, nn)
        fe = np.einsum("eq,eq,qa->ea", o.t.W, o.g(o.at_qp(phi_old)) * o.at_qp(Vc), o.t.N)
        assert rel(asm(9, phi_old, Vc), o._vec(fe, o.edof_s, nn)) < tol


@pytest.mark.parametrize("deg,sp", [("isotropic", "no"), ("anisotropic", "spectral"), ("anisotropic", "deviatoric")])
@pytest.mark.parametrize("variational", [True, False])
def test_coupled_forms_of_the_energy_controlled_solvers_match_the_oracle(hostcheck, deg, sp, variational):
    """pfx_elements_coupled.cuh compiled on the host (the code the `OpResidualPhiE` / `OpApplyUP` kernels run) against
    the blocked residual and Jacobian of oracle/energy_controlled.py (reference solver_ener_variational.py:167-203),
    on every cell type: F_phi, K_gamma phi and the action of [[J_uu, J_uphi], [J_phiu, J_phiphi]]."""
    from oracle.energy_controlled import EnergyControlled
    rng = np.random.default_rng(11)
    sm, em = CODES[(deg, sp)]
    for ctype, m in perturbed_meshes(rng):
        Pm = Params(E=210.0, nu=0.3, Gc=0.0027, l=0.1, degradation=deg, split_energy=sp, k=1e-3)
        o = EnergyControlled(m.geometry.x, m.cells, m.cell_type, Pm, variational=variational, c1=3.0, c2=1.0)
        nn, d = o.nn, o.d
        u, phi, lam = rng.normal(size=nn * d) * 1e-3, rng.uniform(0, 0.9, nn), 0.37
        vu, vp = rng.normal(size=nn * d), rng.normal(size=nn)
        coef = Pm.Gc + (lam * o.c1 if variational else 0.0)
        mat = np.array([Pm.lambda_, Pm.mu, Pm.Gc, Pm.l, Pm.k])
        cells = np.ascontiguousarray(m.cells, dtype=np.int32)
        X = np.ascontiguousarray(m.geometry.x)

        def run(which, with_v):
            ou, op = np.zeros(nn * d), np.zeros(nn)
            hostcheck.hc_coupled(which, CT[ctype], dp(mat), sm, em, 0, 3, ctypes.c_int64(len(cells)), dp(X), cells.ctypes.data_as(PI),
                                 dp(u), dp(phi), dp(vu) if with_v else None, dp(vp) if with_v else None, ctypes.c_double(coef), dp(ou), dp(op))
            return ou, op
        Fu, Fp, _ = o.residual(u, phi, lam, 0.0)
        assert rel(run(0, False)[1], Fp) < 1e-11
        assert rel(run(1, False)[1], o.Kg @ phi) < 1e-11
        J = o.jacobian(u, phi, lam).tocsr()
        nu = nn * d
        y = J[: nu + nn, : nu + nn] @ np.concatenate([vu, vp])
        ou, op = run(2, True)
        assert rel(ou, y[:nu]) < 1e-10 and rel(op, y[nu:]) < 1e-10
