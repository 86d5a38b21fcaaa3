// pfx_elements_coupled.cuh — element forms of the energy-controlled monolithic solvers (host + device; no CUDA-only
// constructs, so tests/native/hostcheck.cpp compiles the very same code on the CPU).
//
// Replaces the blocked forms of reference solver_ener_variational.py:167-203 / solver_ener_non_variational.py:166-201:
// F_phi with psi_a(u) directly (no history field) and the action of J = d(F_u, F_phi)/d(u, phi).
#pragma once
#include "pfx_elements.cuh"

namespace pfx {

// mode 0: out[a] = int (g'(phi) psi_a(u) + coef/l phi) N_a + coef l grad phi . grad N_a        (F_phi)
// mode 1: out[a] = int (phi N_a / l + l grad phi . grad N_a)                                   (K_gamma phi: border of J)
template <class Cell>
PFX_HD void elem_residual_phi_e(int mode, const Material& m, int n1d, const double* X, const double* U, const double* phi,
                                double coef, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D], e[NS], pa = 0.0, pb = 0.0;
    if (mode == 1) coef = 1.0;
    if constexpr (Cell::simplex) {
        if (mode == 0) { geo.eval(0, N, G); sym_grad<Cell>(G, U, e); energy_split<D>(m, m.emodel, e, pa, pb); }
    }
    for (int q = 0; q < geo.nq; ++q) {
        const double w = geo.eval(q, N, G);
        if constexpr (!Cell::simplex) {
            if (mode == 0) { sym_grad<Cell>(G, U, e); energy_split<D>(m, m.emodel, e, pa, pb); }
        }
        const double ph = interp<NV>(N, phi);
        const double src = (mode == 0 ? dg_fun(m.dfun, ph) * pa : 0.0) + coef / m.l * ph;
        double gp[D];
        for (int i = 0; i < D; ++i) { gp[i] = 0.0; for (int a = 0; a < NV; ++a) gp[i] += G[a][i] * phi[a]; }
        for (int a = 0; a < NV; ++a) {
            double gg = 0.0;
            for (int i = 0; i < D; ++i) gg += G[a][i] * gp[i];
            out[a] += w * (src * N[a] + coef * m.l * gg);
        }
    }
}

// (y_u, y_phi) = [[J_uu, J_uphi], [J_phiu, J_phiphi]] (v_u, v_phi) at (u, phi):
//   y_u[a,i] = int ((g+k) dsigma_a(v_u) + dsigma_b(v_u) + g'(phi) v_phi sigma_a(u)) : eps(N_a e_i)
//   y_phi[a] = int (g'(phi) sigma_a(u):eps(v_u) + (g''(phi) psi_a(u) + coef/l) v_phi) N_a + coef l grad v_phi . grad N_a
// out: NV rows of D + 1 entries (u-components, then phi).  Needs smodel == emodel (sigma_a = d psi_a / d eps).
template <class Cell>
PFX_HD void elem_apply_up(const Material& m, int n1d, const double* X, const double* phi, const double* U, const double* Vu,
                          const double* Vp, double coef, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    double ou[NV * D], op[NV];
    for (int i = 0; i < NV * D; ++i) ou[i] = 0.0;
    for (int a = 0; a < NV; ++a) op[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D], e[NS], de[NS], sa[NS], sb[NS], dsa[NS], dsb[NS], S[NS], pa = 0.0, pb = 0.0, sade = 0.0;
    auto point = [&]() {
        sym_grad<Cell>(G, U, e);
        sym_grad<Cell>(G, Vu, de);
        stress_split<D, double>(m, m.smodel, e, sa, sb);
        dstress_split<D>(m, m.smodel, e, de, dsa, dsb);
        energy_split<D>(m, m.emodel, e, pa, pb);
        sade = ddot<D, double>(sa, de);
    };
    if constexpr (Cell::simplex) { geo.eval(0, N, G); point(); }
    for (int q = 0; q < geo.nq; ++q) {
        const double w = geo.eval(q, N, G);
        if constexpr (!Cell::simplex) point();
        const double ph = interp<NV>(N, phi), vp = interp<NV>(N, Vp);
        const double gq = g_fun(m.dfun, ph) + m.k, dgq = dg_fun(m.dfun, ph), ddgq = ddg_fun(m.dfun, ph);
        for (int i = 0; i < NS; ++i) S[i] = gq * dsa[i] + dsb[i] + dgq * vp * sa[i];
        add_div<Cell>(G, S, w, ou);
        const double src = dgq * sade + (ddgq * pa + coef / m.l) * vp;
        double gv[D];
        for (int i = 0; i < D; ++i) { gv[i] = 0.0; for (int a = 0; a < NV; ++a) gv[i] += G[a][i] * Vp[a]; }
        for (int a = 0; a < NV; ++a) {
            double gg = 0.0;
            for (int i = 0; i < D; ++i) gg += G[a][i] * gv[i];
            op[a] += w * (src * N[a] + coef * m.l * gg);
        }
    }
    for (int a = 0; a < NV; ++a) {
        for (int i = 0; i < D; ++i) out[a * (D + 1) + i] = ou[a * D + i];
        out[a * (D + 1) + D] = op[a];
    }
}

}  // namespace pfx
