"""CPU restatement of the two-level PCG preconditioner of phasefieldx_b200/csrc/pfx_coarse.cuh.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): imported by tests/ to check the algorithm that
replaces the reference's MUMPS factorisation (reference solver.py:198-205) independently of the GPU:

    M^-1 = B^-1 + Z (Z^T A Z restricted to regions)^-1 Z^T

B = nodal d x d diagonal blocks of A (Dirichlet dofs decoupled), Z = rigid-body modes of aggregates of
nodes (2D: two translations + rotation about the aggregate centroid, scaled by the aggregate size; 3D:
three + three; scalar fields: the constant), masked on Dirichlet dofs; couplings between different
regions of aggregates are dropped from Z^T A Z.  Nothing here is on the product path.
"""
import numpy as np
import scipy.sparse as sp


def rcb_aggregates(x, n_agg):
    """Recursive coordinate bisection of the nodes into n_agg aggregates, numbered in tree order."""
    out = np.zeros(len(x), dtype=np.int64)

    def rec(idx, k, base):
        if k == 1:
            out[idx] = base
            return
        c = x[idx]
        ax = int(np.argmax(c.max(0) - c.min(0)))
        order = np.argsort(c[:, ax], kind="stable")
        k1 = k // 2
        m = len(idx) * k1 // k
        rec(idx[order[:m]], k1, base)
        rec(idx[order[m:]], k - k1, base + k1)
    rec(np.arange(len(x)), n_agg, 0)
    return out


def modes(x, agg, d_field, mask=None):
    """Z [N*d_field, n_agg*M] (csr).  x [N, gdim]; d_field = gdim (vector field) or 1 (scalar)."""
    n, gd = x.shape
    n_agg = int(agg.max()) + 1
    if d_field == 1:
        Z = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, n_agg))
    else:
        cen = np.stack([np.bincount(agg, weights=x[:, k], minlength=n_agg) for k in range(gd)], 1) / np.bincount(agg, minlength=n_agg)[:, None]
        half = np.zeros(n_agg)
        for k in range(gd):
            lo = np.full(n_agg, np.inf)
            hi = np.full(n_agg, -np.inf)
            np.minimum.at(lo, agg, x[:, k])
            np.maximum.at(hi, agg, x[:, k])
            half = np.maximum(half, 0.5 * (hi - lo))
        rel = (x - cen[agg]) / np.where(half > 0, half, 1.0)[agg][:, None]
        rel = rel.astype(np.float32).astype(np.float64)            # the kernels keep these in fp32
        M = 3 if gd == 2 else 6
        rows, cols, vals = [], [], []

        def add(comp, mode, v):
            rows.append(np.arange(n) * gd + comp)
            cols.append(agg * M + mode)
            vals.append(v)
        one = np.ones(n)
        for k in range(gd):
            add(k, k, one)
        if gd == 2:
            add(0, 2, -rel[:, 1]); add(1, 2, rel[:, 0])
        else:
            add(1, 3, -rel[:, 2]); add(2, 3, rel[:, 1])
            add(0, 4, rel[:, 2]); add(2, 4, -rel[:, 0])
            add(0, 5, -rel[:, 1]); add(1, 5, rel[:, 0])
        Z = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n * gd, n_agg * M))
    if mask is not None:
        Z = sp.diags((~mask).astype(np.float64)) @ Z
    return Z.tocsr()


class TwoLevel:
    """Preconditioner object: callable r -> M^-1 r (r zero on Dirichlet dofs)."""

    def __init__(self, A, x, agg, d_field, mask, n_regions=1, block=True):
        N = A.shape[0]
        self.mask = mask
        free = ~mask
        Af = sp.diags(free.astype(float)) @ A @ sp.diags(free.astype(float)) + sp.diags(mask.astype(float))
        self.Z = modes(x, agg, d_field, mask)
        n_agg = int(agg.max()) + 1
        M = self.Z.shape[1] // n_agg
        Ac = (self.Z.T @ Af @ self.Z).toarray()
        reg = (np.arange(n_agg) * n_regions) // n_agg                 # contiguous runs of aggregates
        creg = np.repeat(reg, M)
        Ac[creg[:, None] != creg[None, :]] = 0.0
        zero = np.diag(Ac) <= 0
        Ac[zero, zero] = 1.0
        Ainv = np.linalg.inv(Ac)
        self.Ainv = (0.5 * (Ainv + Ainv.T)).astype(np.float32).astype(np.float64)   # explicit fp32 inverse
        d = d_field
        if block and d > 1:
            nn = N // d
            blk = np.zeros((nn, d, d))
            Acsr = Af.tocsr()
            nodes = np.arange(nn)
            for i in range(d):
                for j in range(d):
                    blk[:, i, j] = np.asarray(Acsr[nodes * d + i, nodes * d + j]).ravel()
            self.Binv = np.linalg.inv(blk)
            self.dinv = None
        else:
            self.Binv = None
            self.dinv = 1.0 / Af.diagonal()
        self.d = d

    def __call__(self, r):
        if self.Binv is not None:
            z = np.einsum("nij,nj->ni", self.Binv, r.reshape(-1, self.d)).ravel()
        else:
            z = self.dinv * r
        z = z + self.Z @ (self.Ainv @ (self.Z.T @ r))
        z[self.mask] = 0.0
        return z


class ThirdLevel:
    """Additive third level over regions (restating pfx_coarse.cuh `Coarse3View`: the rank-level coarse space of
    multi-GPU runs, where every rank is one region): rigid-body modes of whole regions, E = Z_g^T A Z_g WITH the
    couplings across regions, M^-1 = two_level + Z_g E^-1 Z_g^T.  The region modes lie in the span of the aggregate
    modes (up to the fp32 rounding of the mode geometry), which is how the kernels apply them (Z T)."""

    def __init__(self, two_level, A, x, agg, d_field, mask, n_regions):
        self.base = two_level
        n_agg = int(agg.max()) + 1
        reg = ((np.arange(n_agg) * n_regions) // n_agg)[agg]
        free = ~mask
        Af = sp.diags(free.astype(float)) @ A @ sp.diags(free.astype(float)) + sp.diags(mask.astype(float))
        self.Zg = modes(x, reg, d_field, mask)
        E = (self.Zg.T @ Af @ self.Zg).toarray()
        E = 0.5 * (E + E.T)
        zero = np.diag(E) <= 0
        E[zero, :] = 0.0
        E[:, zero] = 0.0
        E[zero, zero] = 1.0
        np.linalg.cholesky(E)                       # must be positive definite (raises otherwise)
        self.Einv = np.linalg.inv(E)
        self.mask = mask

    def __call__(self, r):
        z = self.base(r) + self.Zg @ (self.Einv @ (self.Zg.T @ r))
        z[self.mask] = 0.0
        return z


def pcg(A, b, minv, mask, rtol=1e-11, max_it=20000):
    """Preconditioned CG on the free dofs (b, iterates zero on Dirichlet dofs); returns x, iterations."""
    free = ~mask
    x = np.zeros_like(b)
    r = np.where(free, b, 0.0)
    z = minv(r)
    p = z.copy()
    rz = r @ z
    b0 = np.linalg.norm(r)
    for it in range(1, max_it + 1):
        Ap = np.where(free, A @ p, 0.0)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= rtol * b0:
            return x, it
        z = minv(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, max_it
