// TEST-ONLY host build of the element/physics headers (g++, no CUDA): lets the CPU test-suite
// check the exact formulas the CUDA kernels evaluate against the oracle before any GPU run.
// Not part of libpfx_b200.so and never used by the product path.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../phasefieldx_b200/csrc/pfx_elements.cuh"

using namespace pfx;

template <int D>
static void split_pts(const Material& m, int64_t n, const double* e, const double* de, double* sa, double* sb,
                      double* pa, double* pb, double* dsa, double* dsb) {
    constexpr int NS = SymN<D>::N;
    for (int64_t i = 0; i < n; ++i) {
        stress_split<D, double>(m, m.smodel, e + i * NS, sa + i * NS, sb + i * NS);
        energy_split<D>(m, m.emodel, e + i * NS, pa[i], pb[i]);
        dstress_split<D>(m, m.smodel, e + i * NS, de + i * NS, dsa + i * NS, dsb + i * NS);
    }
}

template <class Cell>
static void assemble(int op, const Material& m, int n1d, int64_t ne, const double* coords, const int32_t* cells,
                     const double* f0, const double* f1, const double* f2, const double* f3, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int64_t e = 0; e < ne; ++e) {
        const int32_t* c = cells + e * NV;
        double X[NV * D], a0[NV * D], a1[NV * D], a2[NV * D], a3[NV * D], o[NV * D];
        auto gs = [&](const double* f, double* dst) { if (f) for (int a = 0; a < NV; ++a) dst[a] = f[c[a]]; };
        auto gv = [&](const double* f, double* dst) { if (f) for (int a = 0; a < NV; ++a) for (int i = 0; i < D; ++i) dst[a * D + i] = f[c[a] * D + i]; };
        for (int a = 0; a < NV; ++a) for (int i = 0; i < D; ++i) X[a * D + i] = coords[c[a] * 3 + i];
        int nout = 1;
        switch (op) {
            case 0: gs(f0, a0); gv(f1, a1); gv(f2, a2); elem_apply_u<Cell>(m, n1d, X, a0, a1, a2, o); nout = D; break;      // phi,u,v
            case 1: gs(f0, a0); gv(f1, a1); elem_residual_u<Cell>(m, n1d, X, a0, a1, o); nout = D; break;                  // phi,u
            case 2: gs(f0, a0); gv(f1, a1); elem_diag_u<Cell>(m, n1d, X, a0, a1, o); nout = D; break;
            case 3: gs(f0, a0); gs(f1, a1); gs(f2, a2); elem_apply_phi<Cell>(n1d, X, a0, a1, a2, o); break;                // cm,cg,p
            case 4: gs(f0, a0); gs(f1, a1); gs(f2, a2); gs(f3, a3); elem_residual_phi<Cell>(n1d, X, a0, a1, a2, a3, o); break; // cm,cg,H,phi
            case 5: gs(f0, a0); gs(f1, a1); elem_diag_phi<Cell>(n1d, X, a0, a1, o); break;
            case 6: gs(f0, a0); elem_apply_mass<Cell>(n1d, X, a0, o); break;
            case 7: elem_diag_mass<Cell>(n1d, X, o); break;
            case 8: gv(f0, a0); elem_rhs_psi<Cell>(m, n1d, X, a0, o); break;
            case 9: gs(f0, a0); gs(f1, a1); elem_rhs_fatigue<Cell>(n1d, X, a0, a1, o); break;
            case 10: gs(f0, a0); gs(f1, a1); elem_l2<Cell, 1>(n1d, X, a0, a1, out); continue;
            case 11: gv(f0, a0); gv(f1, a1); elem_l2<Cell, D>(n1d, X, a0, a1, out); continue;
            case 12: gv(f0, a0); gs(f1, a1); gs(f2, a2); gs(f3, a3); elem_energies<Cell>(m, n1d, X, a0, a1, a2, a3, out); continue;
            default: return;
        }
        for (int a = 0; a < NV; ++a) for (int i = 0; i < nout; ++i) out[c[a] * nout + i] += o[a * nout + i];
    }
}

extern "C" {
void hc_split(int d, const double* mat /*lam,mu,Gc,l,k*/, int smodel, int emodel, int64_t n, const double* e,
              const double* de, double* sa, double* sb, double* pa, double* pb, double* dsa, double* dsb) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, emodel, 0, 0.0};
    if (d == 2) split_pts<2>(m, n, e, de, sa, sb, pa, pb, dsa, dsb);
    else split_pts<3>(m, n, e, de, sa, sb, pa, pb, dsa, dsb);
}
void hc_assemble(int op, int cell_type, const double* mat, int smodel, int emodel, int fatigue, int n1d, int64_t ne,
                 const double* coords, const int32_t* cells, const double* f0, const double* f1, const double* f2,
                 const double* f3, double* out) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, emodel, fatigue, 0.0};
    if (cell_type == 0) assemble<Tri>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
    else if (cell_type == 1) assemble<Quad>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
    else assemble<Tet>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
}
}
