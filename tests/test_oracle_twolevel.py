"""The two-level preconditioner that stands in for the reference's MUMPS solves (pfx_coarse.cuh),
restated in numpy (oracle/twolevel.py) and checked on the CPU: it is symmetric positive definite —
also with the coarse matrix cut into regions and with an fp32 inverse —, the PCG built on it
reproduces the direct solution of the oracle's tangent systems, and it needs several times fewer
iterations than Jacobi.  (The GPU implementation is compared with this behaviour in
tests/test_gpu_coarse.py.)"""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

from oracle import twolevel as T
from oracle.solver import Oracle, Params


def _system(kind):
    from phasefieldx_b200 import mesh as M
    rng = np.random.default_rng(3)
    if kind == "tri":
        msh = M.create_rectangle(None, [[-0.5, -0.5], [0.5, 0.5]], [48, 48])
        msh = M.cut_slit(msh, lambda x: np.isclose(x[1], 0.0) & (x[0] < -1e-9), lambda c: c[1] > 0)
        P = Params(degradation="anisotropic", split_energy="spectral", l=0.06, k=1e-6)
    else:
        msh = M.create_box(None, [[0, 0, 0], [2, 1, 0.5]], [12, 6, 4])
        P = Params(degradation="anisotropic", split_energy="deviatoric", l=0.2, k=1e-6)
    o = Oracle(msh.geometry.x, msh.cells, msh.cell_type, P)
    x = msh.geometry.x[:, : o.d]
    u = rng.normal(size=o.nn * o.d) * 1e-4
    phi = np.clip(np.exp(-np.abs(x[:, 1]) / P.l) * (x[:, 0] < 0), 0, 0.95)
    A = o.J_u(u, phi).tocsr()
    mask = np.zeros(o.nn * o.d, bool)
    fixed = np.nonzero(np.isclose(x[:, 1], x[:, 1].min()) | np.isclose(x[:, 1], x[:, 1].max()))[0]
    mask[(fixed[:, None] * o.d + np.arange(o.d)).ravel()] = True
    b = np.where(mask, 0.0, rng.normal(size=o.nn * o.d))
    return o, x, A, mask, b


@pytest.mark.parametrize("kind", ["tri", "tet"])
def test_two_level_pcg_matches_direct_solve_with_fewer_iterations(kind):
    o, x, A, mask, b = _system(kind)
    free = np.nonzero(~mask)[0]
    ref = np.zeros_like(b)
    ref[free] = spl.spsolve(A[free][:, free].tocsc(), b[free])
    dinv = 1.0 / A.diagonal()
    _, it_j = T.pcg(A, b, lambda r: np.where(mask, 0.0, dinv * r), mask)
    agg = T.rcb_aggregates(x, 48 if kind == "tri" else 24)
    results = {}
    for name, kw in (("one region", dict(n_regions=1)), ("four regions", dict(n_regions=4)),
                     ("diagonal smoother", dict(n_regions=1, block=False))):
        minv = T.TwoLevel(A, x, agg, o.d, mask, **kw)
        sol, it = T.pcg(A, b, minv, mask)
        assert np.linalg.norm(sol - ref) <= 1e-8 * np.linalg.norm(ref), name
        results[name] = it
        # symmetric positive definite: <a, M^-1 b> = <b, M^-1 a>, <a, M^-1 a> > 0
        rng = np.random.default_rng(5)
        a1, a2 = (np.where(mask, 0.0, rng.normal(size=len(b))) for _ in range(2))
        assert abs(a1 @ minv(a2) - a2 @ minv(a1)) <= 1e-10 * abs(a1 @ minv(a2))
        assert a1 @ minv(a1) > 0
    gain = 3 if kind == "tri" else 2                  # the small tet mesh is well conditioned to begin with
    assert results["one region"] * gain < it_j, (results, it_j)
    assert results["four regions"] * (gain - 1) < it_j
    assert results["one region"] <= results["diagonal smoother"]


def test_modes_are_rigid_body_motions():
    """K_e annihilates the aggregate modes away from aggregate boundaries: Z^T A Z only couples aggregates
    that share a cell (what makes the distance-2 colouring of the probing set-up sufficient)."""
    o, x, A, mask, b = _system("tri")
    agg = T.rcb_aggregates(x, 30)
    nomask = np.zeros_like(mask)
    Z = T.modes(x, agg, o.d, nomask)
    Ac = (Z.T @ A @ Z).toarray()
    n_agg = agg.max() + 1
    adj = np.eye(n_agg, dtype=bool)
    ca = agg[o.t.cells]
    for i in range(ca.shape[1]):
        for j in range(ca.shape[1]):
            adj[ca[:, i], ca[:, j]] = True
    blocks = np.abs(Ac).reshape(n_agg, 3, n_agg, 3).max(axis=(1, 3))
    assert np.all(blocks[~adj] <= 1e-9 * blocks.max())
    # rows of a rigid body motion of the WHOLE body vanish: sum over aggregates of the translation columns
    t = np.zeros(Z.shape[1]); t[0::3] = 1.0
    assert np.abs(A @ (Z @ t)).max() <= 1e-9 * np.abs(A).max()


def test_region_level_third_coarse_space_restores_the_modes_lost_to_block_diagonal_regions():
    """pfx_coarse.cuh `Coarse3View` (the rank-level coarse space of multi-GPU runs) restated in numpy: a slender 3D
    beam cut into 8 regions (= 8 ranks: block-diagonal aggregate-level coarse matrix) loses the modes spanning several
    regions and needs ~2.5x the PCG iterations of the one-region preconditioner; the additive term with the 8 x 6
    rigid-body modes of whole regions brings it back to within ~40 % — and the preconditioner stays symmetric positive
    definite, the solution unchanged."""
    from oracle import twolevel as TL
    from oracle.solver import Oracle, Params
    from phasefieldx_b200 import workloads as W
    w = W.notched_beam_3d(36, 12, 6)
    P = Params(**dict(w.params, degradation="isotropic", split_energy="no"))
    o = Oracle(w.msh.geometry.x, w.msh.cells, w.cell_type, P)
    A = o.J_u(o.u, o.phi).tocsr()
    N = A.shape[0]
    mask, g = np.zeros(N, bool), np.zeros(N)
    for (_, d), v in zip(w.bc_sets, w.bc_values(1.0)):
        mask[d] = True
        g[d] = v
    b = -(A @ g)
    b[mask] = 0
    x = w.msh.geometry.x[:, :3]
    agg = TL.rcb_aggregates(x, 128)
    one = TL.TwoLevel(A, x, agg, 3, mask, n_regions=1)
    cut = TL.TwoLevel(A, x, agg, 3, mask, n_regions=8)
    three = TL.ThirdLevel(cut, A, x, agg, 3, mask, 8)
    sols, its = {}, {}
    for name, m in (("one", one), ("cut", cut), ("three", three)):
        sols[name], its[name] = TL.pcg(A, b, m, mask, rtol=1e-11)
    assert its["cut"] > 1.8 * its["one"]
    assert its["three"] < 0.65 * its["cut"] and its["three"] < 1.6 * its["one"], its
    for name in ("cut", "three"):
        assert np.linalg.norm(sols[name] - sols["one"]) < 1e-8 * np.linalg.norm(sols["one"])
    # symmetric positive definite on the free dofs
    rng = np.random.default_rng(0)
    free = ~mask
    u, v = rng.normal(size=N) * free, rng.normal(size=N) * free
    assert abs(u @ three(v) - v @ three(u)) < 1e-10 * abs(u @ three(v)) and u @ three(u) > 0
