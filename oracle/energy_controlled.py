"""
ORACLE (test infrastructure, not product code) — energy-controlled phase-field fracture solvers.

CPU restatement of reference
  `src/phasefieldx/Element/Phase_Field_Fracture/solver/solver_ener_variational.py:41-527`      (variational scheme)
  `src/phasefieldx/Element/Phase_Field_Fracture/solver/solver_ener_non_variational.py:40-532`  (non-variational)
Monolithic Newton on (u, phi, lambda) (forms :167-203 / :166-201):
  F_u   = int (g(phi)+k) sigma_a(u):eps(w) + sigma_b(u):eps(w) dx - (1 - lambda c2) int T.w ds
  F_phi = int g'(phi) psi_a(u) v dx + (Gc + lambda c1 [variational only]) int (phi v / l + l grad phi.grad v) dx   (AT2, c0 = 2)
  F_lam = c1 gamma(phi) + c2 int T.u ds - tau          (the -tau shift: RealSpaceNewtonSolver._assemble_residual :206-259)
  J     = d(F_u, F_phi, F_lam)/d(u, phi, lambda) with J_phi,lambda = None in the non-variational scheme (:198-201)
psi_a(u) enters F_phi directly: no history field, no projection.  Dirichlet data: bc_list_u only (rows = identity,
lifting as in oracle/solver.py).  Third-party pieces restated (not vendored under /root/reference): scifem
`BlockedNewtonSolver` = dolfinx `nls::petsc::NewtonSolver` (fenics-dolfinx 0.10.0 defaults: rtol 1e-9, atol 1e-10,
max_it 50, convergence_criterion "residual", relaxation 1, error_on_nonconvergence -> RuntimeError), sparse direct solve
(scipy SuperLU <-> MUMPS; the non-variational script's gmres+lu is an exact solve too).
Adaptive dtau with rollback: :339-383, :486-514 (golden-ratio growth when Newton needed <= 2 iterations, shrink and
restore the last converged state when the solve raised).

Pinned (tests/test_oracle_energy_controlled.py) against the reference's one-element analytic tests
(test/Element/Phase_Field_Fracture/test_phase_field_fracture_element_ener_[non_]variational.py) and the stored outputs
examples/PhaseFieldFractureEnergyControlled/results_three_point_[non_]var/{total.energy,lambda.dof,bottom_left.reaction}.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .solver import Oracle, _bc_vectors

GOLDEN_RATIO = 2 / (1 + 5 ** 0.5)


class NewtonFailure(RuntimeError):
    pass


class EnergyControlled(Oracle):
    """State: u, phi (inherited) and the scalar lam."""

    def __init__(self, points, cells, cell_type, params, variational=True, c1=1.0, c2=1.0, degree=None):
        super().__init__(points, cells, cell_type, params, degree=degree)
        self.variational, self.c1, self.c2 = bool(variational), float(c1), float(c2)
        self.lam = 0.0
        p, t = self.p, self.t
        Ke = (np.einsum("eq,qa,qb->eab", t.W, t.N, t.N) / p.l + np.einsum("eq,eqai,eqbi->eab", t.W, t.G, t.G) * p.l)
        self.Kg = self._csr(Ke, self.edof_s, self.nn)              # gamma(phi) = 1/2 phi^T Kg phi
        self.tvec = np.zeros(self.nn * self.d)                     # int T.w ds  (set_traction)

    def set_traction(self, nodes, weights, T):
        self.tvec = self.traction_vector(nodes, weights, T)

    # ------------------------------------------------------------------ forms
    def gamma(self, phi=None):
        phi = self.phi if phi is None else phi
        return 0.5 * float(phi @ (self.Kg @ phi))

    def residual(self, u, phi, lam, tau, s=None):
        s = self.split(u, True) if s is None else s
        Fu = self.F_int_u(u, phi, s) - (1.0 - lam * self.c2) * self.tvec
        fe = np.einsum("eq,eq,qa->ea", self.t.W, self.dg(self.at_qp(phi)) * s["psi_a"], self.t.N)
        coef = self.p.Gc + (lam * self.c1 if self.variational else 0.0)
        Fp = self._vec(fe, self.edof_s, self.nn) + coef * (self.Kg @ phi)
        Fl = self.c1 * self.gamma(phi) + self.c2 * float(self.tvec @ u) - tau
        return Fu, Fp, Fl

    def jacobian(self, u, phi, lam, s=None):
        """Sparse (n_u + n_phi + 1)^2 matrix in block order (u, phi, lambda)."""
        s = self.split(u, True) if s is None else s
        t, d, nv = self.t, self.d, self.nv
        nu, npf = self.nn * d, self.nn
        Juu = self.J_u(u, phi, s)
        dgq, ddgq = self.dg(self.at_qp(phi)), self.ddg(self.at_qp(phi))
        # J_u,phi[(a,c), b] = int g'(phi) N_b sigma_a : eps(N_a e_c)
        Se = np.einsum("eq,eq,eqij,eqijm->eqm", t.W, dgq, s["sig_a"], self.Bm)
        Kup = np.einsum("eqm,qb->emb", Se, t.N)
        rows = np.repeat(self.edof_u, nv, axis=1).ravel()
        cols = np.tile(self.edof_s, (1, nv * d)).ravel()
        Jup = sp.csr_matrix((Kup.reshape(-1), (rows, cols)), shape=(nu, npf))
        Kpp = np.einsum("eq,eq,qa,qb->eab", t.W, ddgq * s["psi_a"], t.N, t.N)
        coef = self.p.Gc + (lam * self.c1 if self.variational else 0.0)
        Jpp = self._csr(Kpp, self.edof_s, npf) + coef * self.Kg
        kgphi = self.Kg @ phi
        col_u = sp.csr_matrix((self.c2 * self.tvec).reshape(-1, 1))
        col_p = sp.csr_matrix(((self.c1 * kgphi) if self.variational else np.zeros(npf)).reshape(-1, 1))
        row_u = sp.csr_matrix((self.c2 * self.tvec).reshape(1, -1))
        row_p = sp.csr_matrix((self.c1 * kgphi).reshape(1, -1))
        return sp.bmat([[Juu, Jup, col_u], [Jup.T, Jpp, col_p], [row_u, row_p, None]], format="csr")

    # ------------------------------------------------------------------ dolfinx NewtonSolver on the blocked system
    def newton(self, tau, bcs_u, rtol=1e-9, atol=1e-10, max_it=50):
        nu, npf = self.nn * self.d, self.nn
        n = nu + npf + 1
        mask = np.zeros(n, dtype=bool)
        g = np.zeros(n)
        mu, gu = _bc_vectors(nu, bcs_u)
        mask[:nu], g[:nu] = mu, gu
        free = ~mask
        x = np.concatenate([self.u, self.phi, [self.lam]])

        def F(x):
            u, phi, lam = x[:nu], x[nu:nu + npf], x[-1]
            s = self.split(u, True)
            Fu, Fp, Fl = self.residual(u, phi, lam, tau, s)
            r = np.concatenate([Fu, Fp, [Fl]])
            dbc = np.where(mask, g - x, 0.0)
            J = None
            if np.any(dbc != 0.0):
                J = self.jacobian(u, phi, lam, s)
                r = r + J @ dbc
            r[mask] = (x - g)[mask]
            return r, J, s

        r, J, s = F(x)
        res0 = res = np.linalg.norm(r)
        hist = [res]
        it = 0
        conv = res < atol                      # relative residual is 1 at iteration 0
        while not conv and it < max_it:
            if J is None:
                J = self.jacobian(x[:nu], x[nu:nu + npf], x[-1], s)
            dx = np.zeros(n)
            try:
                dx[free] = spla.splu(J[free][:, free].tocsc()).solve(r[free])
            except RuntimeError as exc:           # singular factorisation
                raise NewtonFailure(f"linear solve failed: {exc}")
            dx[mask] = r[mask]
            x = x - dx
            it += 1
            r, J, s = F(x)
            res = np.linalg.norm(r)
            hist.append(res)
            conv = np.isfinite(res) and (res / res0 < rtol or res < atol)
        if not conv:
            raise NewtonFailure("Newton solver did not converge because maximum number of iterations reached"
                                if it == max_it else "Newton solver did not converge")
        self.u, self.phi, self.lam = x[:nu].copy(), x[nu:nu + npf].copy(), float(x[-1])
        return it, hist

    def initial_crack(self, bcs_phi):
        """Optional phase-field Dirichlet pre-solve (:276-288): Gc Kg phi = 0 with phi = g on the seed; returns gamma0."""
        mask, g = _bc_vectors(self.nn, bcs_phi)
        K = (self.p.Gc * self.Kg).tocsr()
        phi = np.where(mask, g, 0.0)
        rhs = -(K @ phi)
        free = ~mask
        phi[free] += spla.splu(K[free][:, free].tocsc()).solve(rhs[free])
        self.phi = phi
        return self.gamma()

    def external_energy(self):
        return float(self.tvec @ self.u)

    # ------------------------------------------------------------------ the dtau loop (:339-514)
    def run(self, final_gamma, bcs_u, dtau=0.0005, dtau_min=1e-12, dtau_max=0.1, gamma0=0.0, max_steps=10 ** 9, callback=None):
        """Returns the list of per-step records (step, iterations, lam, tau, gamma, energies, external)."""
        tau = gamma0 + dtau
        step, gamma, out = 0, 0.0, []
        saved = (self.u.copy(), self.phi.copy(), self.lam)
        while gamma < final_gamma and step < max_steps:
            tau += dtau
            try:
                its, hist = self.newton(tau, bcs_u)
            except NewtonFailure:
                self.u, self.phi, self.lam = saved[0].copy(), saved[1].copy(), saved[2]
                tau -= dtau
                dtau = max(dtau * GOLDEN_RATIO, dtau_min)
                if dtau <= dtau_min:
                    break
                continue
            saved = (self.u.copy(), self.phi.copy(), self.lam)
            if its <= 2:
                dtau = min(dtau * (1 + GOLDEN_RATIO), dtau_max)
            step += 1
            en = self.energies()
            gamma = en["gamma"]
            rec = dict(step=step, iterations=its, lam=self.lam, tau=tau, dtau=dtau, gamma=gamma, energies=en,
                       external=self.external_energy(), residual=hist[-1], hist=hist)
            out.append(rec)
            if callback is not None:
                callback(rec)
        return out
