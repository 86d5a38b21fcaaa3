"""
ORACLE (test infrastructure, not product code) — assembled CPU restatement of the staggered
phase-field fracture / fatigue load step.

Follows reference `src/phasefieldx/Element/Phase_Field_Fracture/solver/solver.py:148-468`:
  u-residual / Jacobian                 :173-205  (dolfinx NonlinearProblem + SNES + MUMPS)
  φ-residual / Jacobian (AT2, c0 = 2)   :214-245  (Element/Phase_Field/geometric_crack.py:53,82)
  stagger loop + stop rule              :319-398
  L2 projection of ψ_a → V_c            :355      (Math/projection.py:19-59)
  history H = max(V_c, V_n)             :357-360, :401-402
  fatigue block                         :362-375, :404-411
  normalised L2 error                   errors_functions.py:190-228
  reactions                             Reactions/reactions_forces.py:12-65
  energies                              Element/Phase_Field_Fracture/energy.py:13-117,
                                        Element/Phase_Field/energy.py:11-50, solver.py:437-447
Third-party pieces restated (dolfinx 0.10 / PETSc, not vendored): Dirichlet lifting idiom
(in-tree mirror solver_ener_variational.py:238-247), SNES default convergence test
(rtol 1e-8, atol 1e-9, stol 1e-8; reasons 2/3/4), sparse direct solve (scipy SuperLU ↔ MUMPS).
Projection modes: "exact" (sparse LU; the parity gate) or "petsc_default"
(BiCGStab + ILU, rtol 1e-5; only to quantify the reference's own solver noise).

parity unpinned at 1e-8 by the reference itself (no dolfinx here): pinned instead against the
reference's analytic one-element tests and the stored example outputs (tests/test_oracle_*.py).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import constitutive as cons
from .fe import Tables


class Params:
    """Subset of the reference `Input` (Element/Phase_Field_Fracture/Input.py:66-102)."""

    def __init__(self, E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no",
                 degradation_function="quadratic", irreversibility="no", fatigue=False,
                 fatigue_degradation_function="asymptotic", fatigue_val=0.05625, k=0.0, **_):
        self.E, self.nu, self.Gc, self.l, self.k = E, nu, Gc, l, k
        self.lambda_ = E * nu / ((1.0 - 2.0 * nu) * (1.0 + nu))
        self.mu = E / (2.0 * (1.0 + nu))
        self.degradation, self.split_energy = degradation, split_energy
        self.degradation_function = degradation_function
        self.irreversibility = irreversibility
        self.fatigue = fatigue
        self.fatigue_degradation_function = fatigue_degradation_function
        self.fatigue_val = fatigue_val
        if degradation_function not in ("quadratic", "borden"):
            raise NotImplementedError("oracle covers the quadratic and borden degradation functions (SURVEY §8f N2)")

    @classmethod
    def from_input(cls, Data):
        return cls(**{k: getattr(Data, k) for k in ("E", "nu", "Gc", "l", "degradation", "split_energy",
                                                    "degradation_function", "irreversibility", "fatigue",
                                                    "fatigue_degradation_function", "fatigue_val", "k")})


class BC:
    """Unrolled Dirichlet data: `dofs` int array, `values` float array of equal length."""

    def __init__(self, dofs, values=None):
        self.dofs = np.asarray(dofs, dtype=np.int64)
        self.values = np.zeros(len(self.dofs)) if values is None else np.asarray(values, float)


def _bc_vectors(n, bcs):
    """mask[n] bool and g[n]; later BCs win on shared dofs (dolfinx set_bc order)."""
    mask = np.zeros(n, dtype=bool)
    g = np.zeros(n)
    for bc in bcs:
        mask[bc.dofs] = True
        g[bc.dofs] = bc.values
    return mask, g


class Oracle:
    def __init__(self, points, cells, cell_type, params, degree=None, projection="exact"):
        self.p = params
        self.g, self.dg, self.ddg = cons.degradation(params.degradation_function)
        if degree is None and params.degradation_function != "quadratic":
            degree = 5          # cubic g: g(φ)·V_c·w and g''(φ)·H·N·N are of degree 5 / 4 on P1 simplices
        self.t = Tables(points, cells, cell_type, degree)
        t = self.t
        self.d, self.nv, self.nn, self.ne = t.d, t.nv, t.nn, t.ne
        self.projection = projection
        d, nv = self.d, self.nv
        # element dof maps
        self.edof_s = t.cells                                                    # [E,nv]
        self.edof_u = (t.cells[:, :, None] * d + np.arange(d)[None, None, :]).reshape(self.ne, nv * d)
        # B-operator: ε_ij = Σ_{a,c} Bm[e,q,i,j,a*d+c] u_{a,c}
        Bm = np.zeros((t.ne, t.nq, d, d, nv, d))
        for i in range(d):
            for j in range(d):
                Bm[:, :, i, j, :, i] += 0.5 * t.G[:, :, :, j]
                Bm[:, :, i, j, :, j] += 0.5 * t.G[:, :, :, i]
        self.Bm = Bm.reshape(t.ne, t.nq, d, d, nv * d)
        # scalar mass matrix and its LU (projection "exact")
        Me = np.einsum("eq,qa,qb->eab", t.W, t.N, t.N)
        self.M = self._csr(Me, self.edof_s, self.nn)
        self._Mlu = None
        # state (reference solver.py:148-170)
        self.u = np.zeros(self.nn * d)
        self.u_old = np.zeros(self.nn * d)
        self.phi = np.zeros(self.nn)
        self.phi_old = np.zeros(self.nn)
        self.H = np.zeros(self.nn)
        self.V_c = np.zeros(self.nn)
        self.V_n = np.zeros(self.nn)
        self.Fatigue = np.ones(self.nn) * 0.0       # dolfinx Functions start at zero
        self.alpha_c = np.zeros(self.nn)
        self.alpha_n = np.zeros(self.nn)
        self.abar_c = np.zeros(self.nn)
        self.abar_n = np.zeros(self.nn)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _csr(Ke, edof, n):
        ne, m = edof.shape
        rows = np.repeat(edof, m, axis=1).ravel()
        cols = np.tile(edof, (1, m)).ravel()
        return sp.csr_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n))

    @staticmethod
    def _vec(fe, edof, n):
        return np.bincount(edof.ravel(), weights=fe.ravel(), minlength=n)

    def strain(self, u):
        return np.einsum("eqijm,em->eqij", self.Bm, u[self.edof_u])

    def at_qp(self, f):
        return np.einsum("qa,ea->eq", self.t.N, f[self.edof_s])

    def grad_qp(self, f):
        return np.einsum("eqai,ea->eqi", self.t.G, f[self.edof_s])

    def split(self, u, want_tangent=True):
        p = self.p
        return cons.split_numpy(self.strain(u), p.lambda_, p.mu, p.degradation, p.split_energy, want_tangent)

    # ------------------------------------------------------------------ u-problem (solver.py:173-187)
    def F_int_u(self, u, phi, s=None):
        s = self.split(u, False) if s is None else s
        gq = self.g(self.at_qp(phi)) + self.p.k
        sig = gq[..., None, None] * s["sig_a"] + s["sig_b"]
        fe = np.einsum("eq,eqij,eqijm->em", self.t.W, sig, self.Bm)
        return self._vec(fe, self.edof_u, self.nn * self.d)

    def J_u(self, u, phi, s=None):
        s = self.split(u, True) if s is None else s
        gq = self.g(self.at_qp(phi)) + self.p.k
        C = gq[..., None, None, None, None] * s["C_a"] + s["C_b"]
        Ke = np.einsum("eq,eqijm,eqijkl,eqkln->emn", self.t.W, self.Bm, C, self.Bm, optimize=True)
        return self._csr(Ke, self.edof_u, self.nn * self.d)

    # ------------------------------------------------------------------ φ-problem (solver.py:214-223)
    def K_phi(self, H, Fat=None):
        p, t = self.p, self.t
        Hq = self.at_qp(H)
        if p.fatigue:
            Fq = self.at_qp(Fat)
            cm = 2.0 * Hq + Fq * p.Gc / p.l
            cg = Fq * p.Gc * p.l
        else:
            cm = 2.0 * Hq + p.Gc / p.l * np.ones_like(Hq)      # Gc·(1/(c0 l))·2φ with c0 = 2
            cg = p.Gc * p.l * np.ones_like(Hq)                 # Gc·l·2/c0
        Ke = (np.einsum("eq,eq,qa,qb->eab", t.W, cm, t.N, t.N)
              + np.einsum("eq,eq,eqai,eqbi->eab", t.W, cg, t.G, t.G))
        return self._csr(Ke, self.edof_s, self.nn)

    def b_phi(self, H):
        """∫ 2 H N_i  (−g'(0)·H; F_φ(φ) = K_φ φ − b_φ for the quadratic degradation)."""
        fe = np.einsum("eq,eq,qa->ea", self.t.W, 2.0 * self.at_qp(H), self.t.N)
        return self._vec(fe, self.edof_s, self.nn)

    # ------------------------------------------------------------------ projection (Math/projection.py:19-59)
    def project(self, fq):
        """L2-project a quadrature-point field fq[E,Q] onto the nodal P1/Q1 space."""
        fe = np.einsum("eq,eq,qa->ea", self.t.W, fq, self.t.N)
        b = self._vec(fe, self.edof_s, self.nn)
        if self.projection == "exact":
            if self._Mlu is None:
                self._Mlu = spla.splu(self.M.tocsc())
            return self._Mlu.solve(b)
        ilu = spla.spilu(self.M.tocsc(), drop_tol=0.0, fill_factor=1.0)
        x, info = spla.bicgstab(self.M, b, rtol=1e-5, atol=0.0, M=spla.LinearOperator(self.M.shape, ilu.solve))
        assert info == 0
        return x

    def rhs_psi(self, u):
        s = self.split(u, False)
        fe = np.einsum("eq,eq,qa->ea", self.t.W, s["psi_a"], self.t.N)
        return self._vec(fe, self.edof_s, self.nn)

    # ------------------------------------------------------------------ norms (errors_functions.py:190-228)
    def l2_error_normalized(self, a, b, vector):
        def sq(f):
            if vector:
                fq = np.einsum("qa,eac->eqc", self.t.N, f.reshape(-1, self.d)[self.edof_s])
                return np.einsum("eq,eqc,eqc->", self.t.W, fq, fq)
            fq = self.at_qp(f)
            return np.einsum("eq,eq,eq->", self.t.W, fq, fq)
        num, den = np.sqrt(sq(a - b)), np.sqrt(sq(a))
        return 0.0 if den == 0 else num / den

    # ------------------------------------------------------------------ SNES emulation
    @staticmethod
    def _snes(residual, jacobian, x, mask, g, rtol=1e-8, atol=1e-9, stol=1e-8, max_it=50):
        """PETSc SNES newtonls, linesearch none, default convergence test; dolfinx BC lifting."""
        def F(x):
            r, J = residual(x), None
            dbc = np.where(mask, g - x, 0.0)
            if np.any(dbc != 0.0):
                J = jacobian(x)
                r = r + J @ dbc                       # apply_lifting(alpha=-1, x0=x)
            r[mask] = (x - g)[mask]                   # set_bc(alpha=-1, x0=x)
            return r, J
        hist = []
        r, J = F(x)
        fnorm = np.linalg.norm(r)
        hist.append(fnorm)
        if fnorm < atol:
            return x, 0, 2, hist
        ttol = fnorm * rtol
        free = ~mask
        for it in range(1, max_it + 1):
            J = jacobian(x) if J is None else J
            Jff = J[free][:, free].tocsc()
            dx = np.zeros_like(x)
            dx[free] = spla.splu(Jff).solve(r[free])
            dx[mask] = r[mask]
            x = x - dx
            r, J = F(x)
            fnorm = np.linalg.norm(r)
            hist.append(fnorm)
            if not np.isfinite(fnorm):
                return x, it, -4, hist
            if fnorm < atol:
                return x, it, 2, hist
            if fnorm <= ttol:
                return x, it, 3, hist
            if np.linalg.norm(dx) < stol * np.linalg.norm(x):
                return x, it, 4, hist
        return x, max_it, -5, hist

    def traction_vector(self, nodes, weights, T):
        """Load vector of reference solver.py:179-184 (∫ T·δu ds) for pre-integrated nodal weights
        w_i = ∫ N_i ds of the tagged boundary."""
        f = np.zeros(self.nn * self.d)
        for c in range(self.d):
            np.add.at(f, np.asarray(nodes) * self.d + c, np.asarray(weights) * T[c])
        return f

    def solve_u(self, bcs_u, fext=None):
        mask, g = _bc_vectors(self.nn * self.d, bcs_u)
        cache = {}

        def split_at(x):
            key = x.tobytes()
            if cache.get("k") != key:
                cache["k"], cache["s"] = key, self.split(x, True)
            return cache["s"]
        if fext is None:
            res = lambda x: self.F_int_u(x, self.phi, split_at(x))
        else:
            res = lambda x: self.F_int_u(x, self.phi, split_at(x)) - fext       # F_u -= L  (solver.py:184)
        jac = lambda x: self.J_u(x, self.phi, split_at(x))
        self.u, its, reason, hist = self._snes(res, jac, self.u, mask, g)
        return its, reason, hist

    def F_phi(self, phi, H, Fat=None):
        """F_phi = dg(phi)·H·w + (F·)Gc(phi·w/l + l grad phi·grad w), general degradation (solver.py:214-222)"""
        K0 = self.K_phi(np.zeros_like(H), Fat)                       # the crack-surface part alone
        fe = np.einsum("eq,eq,qa->ea", self.t.W, self.dg(self.at_qp(phi)) * self.at_qp(H), self.t.N)
        return K0 @ phi + self._vec(fe, self.edof_s, self.nn)

    def J_phi(self, phi, H, Fat=None):
        K0 = self.K_phi(np.zeros_like(H), Fat)
        Ke = np.einsum("eq,eq,qa,qb->eab", self.t.W, self.ddg(self.at_qp(phi)) * self.at_qp(H), self.t.N, self.t.N)
        return K0 + self._csr(Ke, self.edof_s, self.nn)

    def solve_phi(self, bcs_phi):
        mask, g = _bc_vectors(self.nn, bcs_phi)
        if self.p.degradation_function == "quadratic":
            K = self.K_phi(self.H, self.Fatigue)
            b = self.b_phi(self.H)
            self.phi, its, reason, hist = self._snes(lambda x: K @ x - b, lambda x: K, self.phi, mask, g)
        else:
            self.phi, its, reason, hist = self._snes(lambda x: self.F_phi(x, self.H, self.Fatigue),
                                                     lambda x: self.J_phi(x, self.H, self.Fatigue), self.phi, mask, g)
        return its, reason, hist

    # ------------------------------------------------------------------ one load step (solver.py:319-411)
    def load_step(self, bcs_u, bcs_phi, dt=1.0, min_stagger_iter=2, max_stagger_iter=10000, tol=1e-8, fext=None):
        p = self.p
        err_phi = err_u = 1.0
        it = 0
        log = []
        u_its = phi_its = 0
        while (err_phi > tol or err_u > tol or it < min_stagger_iter) and it < max_stagger_iter:
            it += 1
            u_its, reason, hist_u = self.solve_u(bcs_u, fext)
            err_u = self.l2_error_normalized(self.u, self.u_old, True)
            if reason <= 0:
                raise RuntimeError(f"Solver did not converge, got {reason}.")
            self.u_old = self.u.copy()
            s = self.split(self.u, False)
            self.V_c = self.project(s["psi_a"])
            self.H = np.maximum(self.V_c, self.V_n) if p.irreversibility == "miehe" else self.V_c.copy()
            if p.fatigue:
                gq = self.g(self.at_qp(self.phi_old))
                self.alpha_c = self.project(gq * self.at_qp(self.V_c))
                da = self.alpha_c - self.alpha_n
                delta = np.abs(da) * np.heaviside(da / dt, 1)
                self.abar_c = self.abar_n + delta
                if p.fatigue_degradation_function != "asymptotic":
                    raise NotImplementedError(p.fatigue_degradation_function)
                self.Fatigue = cons.fatigue_asymptotic(self.abar_c, p.fatigue_val)
            phi_its, reason, hist_p = self.solve_phi(bcs_phi)
            err_phi = self.l2_error_normalized(self.phi, self.phi_old, False)
            if reason <= 0:
                raise RuntimeError(f"Solver did not converge, got {reason}.")
            self.phi_old = self.phi.copy()
            log.append(dict(u_its=u_its, phi_its=phi_its, err_u=err_u, err_phi=err_phi,
                            fnorm_u=hist_u[-1], fnorm_phi=hist_p[-1]))
        if p.irreversibility == "miehe":
            self.V_n = np.maximum(self.V_c, self.V_n)
        if p.fatigue:
            self.abar_n = self.abar_c.copy()
            self.alpha_n = self.alpha_c.copy()
        return dict(stagger=it, u_its=u_its, phi_its=phi_its, err_u=err_u, err_phi=err_phi, log=log)

    # ------------------------------------------------------------------ outputs
    def reaction(self, bc):
        """−Σ of the residual with this BC's rows zeroed, per component
        (Reactions/reactions_forces.py:34-63; lifting term vanishes because u satisfies the BC)."""
        r = self.F_int_u(self.u, self.phi)
        r[bc.dofs] = 0.0
        r = -r
        R = np.zeros(3)
        for c in range(self.d):
            R[c] = np.sum(r[c:: self.d])
        return R

    def energies(self):
        """The 11 data columns of total.energy (solver.py:434-457)."""
        p, t = self.p, self.t
        s = self.split(self.u, False)
        gq = self.g(self.at_qp(self.phi))
        PSI_a = np.einsum("eq,eq->", t.W, s["psi_a"])
        PSI_b = np.einsum("eq,eq->", t.W, s["psi_b"])
        E = np.einsum("eq,eq,eq->", t.W, gq, s["psi_a"]) + PSI_b      # note: k omitted (energy.py:81-82)
        phq, gph = self.at_qp(self.phi), self.grad_qp(self.phi)
        gamma_phi = np.einsum("eq,eq->", t.W, phq**2) / (2.0 * p.l)
        gamma_grad = np.einsum("eq,eqi,eqi->", t.W, gph, gph) * p.l / 2.0
        gamma = gamma_phi + gamma_grad
        if p.fatigue:
            Fq = self.at_qp(self.Fatigue)
            W_phi = np.einsum("eq,eq,eq->", t.W, Fq, phq**2) * p.Gc / (2 * p.l)
            W_grad = np.einsum("eq,eq,eqi,eqi->", t.W, Fq, gph, gph) * p.Gc * p.l / 2
            alpha_acum = np.einsum("eq,eq->", t.W, self.at_qp(self.abar_c))
        else:
            W_phi, W_grad, alpha_acum = p.Gc * gamma_phi, p.Gc * gamma_grad, 0.0
        W = W_phi + W_grad
        return dict(EplusW=E + W, E=E, W=W, W_phi=W_phi, W_gradphi=W_grad, gamma=gamma, gamma_phi=gamma_phi,
                    gamma_gradphi=gamma_grad, PSI_a=PSI_a, PSI_b=PSI_b, alpha_acum=alpha_acum)

    # ------------------------------------------------------------------ operator-level hooks (kernel parity)
    def apply_u(self, v, mask=None):
        """y = J_u(u,φ)·v with Dirichlet rows/cols replaced by identity (dolfinx assemble_matrix(bcs))."""
        J = self.J_u(self.u, self.phi)
        if mask is None:
            return J @ v
        vm = np.where(mask, 0.0, v)
        y = J @ vm
        y[mask] = v[mask]
        return y

    def apply_phi(self, v, mask=None):
        K = self.K_phi(self.H, self.Fatigue)
        if mask is None:
            return K @ v
        vm = np.where(mask, 0.0, v)
        y = K @ vm
        y[mask] = v[mask]
        return y

    def apply_mass(self, v):
        return self.M @ v
