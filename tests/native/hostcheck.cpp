// TEST-ONLY host build of the element/physics headers (g++, no CUDA): lets the CPU test-suite
// check the exact formulas the CUDA kernels evaluate against the oracle before any GPU run.
// Not part of libpfx_b200.so and never used by the product path.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../phasefieldx_b200/csrc/pfx_elements.cuh"
#include "../../phasefieldx_b200/csrc/pfx_elements_coupled.cuh"

using namespace pfx;

template <int D>
static void split_pts(const Material& m, int64_t n, const double* e, const double* de, double* sa, double* sb,
                      double* pa, double* pb, double* dsa, double* dsb, int dual = 0) {
    constexpr int NS = SymN<D>::N;
    for (int64_t i = 0; i < n; ++i) {
        stress_split<D, double>(m, m.smodel, e + i * NS, sa + i * NS, sb + i * NS);
        energy_split<D>(m, m.emodel, e + i * NS, pa[i], pb[i]);
        if (dual) dstress_split_dual<D>(m, m.smodel, e + i * NS, de + i * NS, dsa + i * NS, dsb + i * NS);
        else dstress_split<D>(m, m.smodel, e + i * NS, de + i * NS, dsa + i * NS, dsb + i * NS);
    }
}

template <class Cell>
static void assemble(int op, const Material& m, int n1d, int64_t ne, const double* coords, const int32_t* cells,
                     const double* f0, const double* f1, const double* f2, const double* f3, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int64_t e = 0; e < ne; ++e) {
        const int32_t* c = cells + e * NV;
        double X[NV * D], a0[NV * D], a1[NV * D], a2[NV * D], a3[NV * D], o[NV * 6];
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
            case 9: gs(f0, a0); gs(f1, a1); elem_rhs_fatigue<Cell>(n1d, X, a0, a1, o, m.dfun); break;
            case 10: gs(f0, a0); gs(f1, a1); elem_l2<Cell, 1>(n1d, X, a0, a1, out); continue;
            case 11: gv(f0, a0); gv(f1, a1); elem_l2<Cell, D>(n1d, X, a0, a1, out); continue;
            case 12: gv(f0, a0); gs(f1, a1); gs(f2, a2); gs(f3, a3); elem_energies<Cell>(m, n1d, X, a0, a1, a2, a3, out); continue;
            case 13:        // cached-tangent apply (simplices): phi,u,v
                if constexpr (Cell::simplex) {
                    double T[21];
                    gs(f0, a0); gv(f1, a1); gv(f2, a2);
                    elem_tangent_u<Cell>(m, X, a0, a1, T);
                    elem_apply_u_tan<Cell>(X, a2, T, o);
                    nout = D;
                    break;
                } else return;
            case 14: gs(f0, a0); gv(f1, a1); elem_diagblk_u<Cell>(m, n1d, X, a0, a1, o); nout = D * (D + 1) / 2; break;  // nodal blocks
            case 15: case 16: case 17:          // general-g phase-field forms: phi, H, Fat, p
                gs(f0, a0); gs(f1, a1); gs(f2, a2); gs(f3, a3);
                elem_phi_general<Cell>(op - 15, m, n1d, X, a0, a1, a2, a3, o);
                break;
            default: return;
        }
        for (int a = 0; a < NV; ++a) for (int i = 0; i < nout; ++i) out[c[a] * nout + i] += o[a * nout + i];
    }
}

// coupled (u, phi) forms of the energy-controlled solvers (pfx_elements_coupled.cuh):
// which = 0: out_p = F_phi(u, phi; coef);  1: out_p = K_gamma phi;  2: (out_u, out_p) = J (vu, vp) at (u, phi)
template <class Cell>
static void assemble_coupled(int which, const Material& m, int n1d, int64_t ne, const double* coords, const int32_t* cells,
                             const double* u, const double* phi, const double* vu, const double* vp, double coef, double* out_u,
                             double* out_p) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int64_t e = 0; e < ne; ++e) {
        const int32_t* c = cells + e * NV;
        double X[NV * D], U[NV * D], P[NV], VU[NV * D], VP[NV], o[NV * (D + 1)];
        for (int a = 0; a < NV; ++a) {
            for (int i = 0; i < D; ++i) {
                X[a * D + i] = coords[c[a] * 3 + i]; U[a * D + i] = u[c[a] * D + i];
                VU[a * D + i] = vu ? vu[c[a] * D + i] : 0.0;
            }
            P[a] = phi[c[a]]; VP[a] = vp ? vp[c[a]] : 0.0;
        }
        if (which < 2) {
            elem_residual_phi_e<Cell>(which, m, n1d, X, U, P, coef, o);
            for (int a = 0; a < NV; ++a) out_p[c[a]] += o[a];
        } else {
            elem_apply_up<Cell>(m, n1d, X, P, U, VU, VP, coef, o);
            for (int a = 0; a < NV; ++a) {
                for (int i = 0; i < D; ++i) out_u[c[a] * D + i] += o[a * (D + 1) + i];
                out_p[c[a]] += o[a * (D + 1) + D];
            }
        }
    }
}

extern "C" {
void hc_coupled(int which, int cell_type, const double* mat, int smodel, int emodel, int dfun, int n1d, int64_t ne,
                const double* coords, const int32_t* cells, const double* u, const double* phi, const double* vu, const double* vp,
                double coef, double* out_u, double* out_p) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, emodel, 0, 0.0, dfun};
    if (cell_type == 0) assemble_coupled<Tri>(which, m, n1d, ne, coords, cells, u, phi, vu, vp, coef, out_u, out_p);
    else if (cell_type == 1) assemble_coupled<Quad>(which, m, n1d, ne, coords, cells, u, phi, vu, vp, coef, out_u, out_p);
    else if (cell_type == 3) assemble_coupled<Hex>(which, m, n1d, ne, coords, cells, u, phi, vu, vp, coef, out_u, out_p);
    else assemble_coupled<Tet>(which, m, n1d, ne, coords, cells, u, phi, vu, vp, coef, out_u, out_p);
}

void hc_split(int d, const double* mat /*lam,mu,Gc,l,k*/, int smodel, int emodel, int64_t n, const double* e,
              const double* de, double* sa, double* sb, double* pa, double* pb, double* dsa, double* dsb) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, emodel, 0, 0.0};
    if (d == 2) split_pts<2>(m, n, e, de, sa, sb, pa, pb, dsa, dsb);
    else split_pts<3>(m, n, e, de, sa, sb, pa, pb, dsa, dsb);
}
// same, tangent action by the reference-faithful dual-number evaluation
void hc_split_dual(int d, const double* mat, int smodel, int emodel, int64_t n, const double* e,
                   const double* de, double* sa, double* sb, double* pa, double* pb, double* dsa, double* dsb) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, emodel, 0, 0.0};
    if (d == 2) split_pts<2>(m, n, e, de, sa, sb, pa, pb, dsa, dsb, 1);
    else split_pts<3>(m, n, e, de, sa, sb, pa, pb, dsa, dsb, 1);
}
void hc_assemble(int op, int cell_type, const double* mat, int smodel, int emodel, int fatigue, int n1d, int64_t ne,
                 const double* coords, const int32_t* cells, const double* f0, const double* f1, const double* f2,
                 const double* f3, double* out) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, emodel, fatigue & 1, 0.0, fatigue >> 8};   // bits 8..: degradation function
    if (cell_type == 0) assemble<Tri>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
    else if (cell_type == 1) assemble<Quad>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
    else if (cell_type == 3) assemble<Hex>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
    else assemble<Tet>(op, m, n1d, ne, coords, cells, f0, f1, f2, f3, out);
}
}

// ---------------------------------------------------------------------------------------------
// CPU mirror of scatter_kernel's index logic (phase A/B/C over the BlockMesh arrays) for the
// u-apply, used to validate pfx_blocks.cpp before any GPU run.  Same chunking, same cursors.
#include "../../phasefieldx_b200/csrc/pfx_blocks.h"

template <class Cell>
static void emulate_apply_u(const BlockMesh& bm, const Material& m, int n1d, const double* coords, const double* phi,
                            const double* u, const double* v, int chunk, double* y, double* dot) {
    constexpr int NV = Cell::NV, D = Cell::D, NF = 3 * D + 1;
    std::vector<double> s_nodes, s_ebuf((size_t)NV * D * chunk);
    *dot = 0.0;
    for (int b = 0; b < bm.nb; ++b) {
        int n0 = bm.node0[b], n_own = bm.node0[b + 1] - n0, h0 = bm.halo_off[b], n_loc = n_own + bm.halo_off[b + 1] - h0;
        s_nodes.assign((size_t)n_loc * NF, 0.0);
        for (int i = 0; i < n_loc; ++i) {
            int g = i < n_own ? n0 + i : bm.halo[h0 + i - n_own];
            int old = bm.perm[g];
            double* s = &s_nodes[(size_t)i * NF];
            for (int k = 0; k < D; ++k) { s[k] = coords[3 * (size_t)old + k]; s[D + 1 + k] = v[(size_t)old * D + k]; s[2 * D + 1 + k] = u[(size_t)old * D + k]; }
            s[D] = phi[old];
        }
        int e0 = bm.elem_off[b], ne = bm.elem_off[b + 1] - e0;
        std::vector<double> acc((size_t)n_own * D, 0.0);
        std::vector<int> cur(n_own), end(n_own);
        for (int t = 0; t < n_own; ++t) { cur[t] = bm.inc_ptr[n0 + t]; end[t] = bm.inc_ptr[n0 + t + 1]; }
        for (int c0 = 0; c0 < ne; c0 += chunk) {
            int nc = std::min(chunk, ne - c0);
            for (int j = 0; j < nc; ++j) {
                const uint16_t* lc = &bm.lconn[(size_t)bm.lstride * (size_t)(e0 + c0 + j)];
                double X[NV * D], ph[NV], V[NV * D], U[NV * D], out[NV * D];
                for (int k = 0; k < NV; ++k) {
                    const double* nd = &s_nodes[(size_t)lc[k] * NF];
                    for (int i = 0; i < D; ++i) { X[k * D + i] = nd[i]; V[k * D + i] = nd[D + 1 + i]; U[k * D + i] = nd[2 * D + 1 + i]; }
                    ph[k] = nd[D];
                }
                elem_apply_u<Cell>(m, n1d, X, ph, U, V, out);
                for (int k = 0; k < NV * D; ++k) s_ebuf[(size_t)k * chunk + j] = out[k];
            }
            for (int t = 0; t < n_own; ++t)
                while (cur[t] < end[t]) {
                    int code = bm.inc[cur[t]], j = (code >> bm.vbits) - c0;
                    if (j >= nc) break;
                    int vtx = code & ((1 << bm.vbits) - 1);
                    for (int c = 0; c < D; ++c) acc[(size_t)t * D + c] += s_ebuf[(size_t)(vtx * D + c) * chunk + j];
                    ++cur[t];
                }
        }
        for (int t = 0; t < n_own; ++t) {
            int old = bm.perm[n0 + t];
            for (int c = 0; c < D; ++c) { y[(size_t)old * D + c] = acc[(size_t)t * D + c]; *dot += v[(size_t)old * D + c] * acc[(size_t)t * D + c]; }
        }
    }
}

extern "C" {
// returns 0 on success; stats = {nb, max_own, max_loc, max_elems, total_elems, sum_nprim}
int hc_blocks_apply_u(int cell_type, const double* mat, int smodel, int n1d, int64_t n_nodes, int64_t n_owned,
                      int64_t n_cells, const double* coords, const int32_t* cells, const uint8_t* cell_primary,
                      int max_own, int chunk, const double* phi, const double* u, const double* v, double* y,
                      double* dot, int64_t* stats) {
    Material m{mat[0], mat[1], mat[2], mat[3], mat[4], smodel, smodel, 0, 0.0};
    int nv = cell_type == 0 ? 3 : (cell_type == 3 ? 8 : 4), dim = cell_type >= 2 ? 3 : 2;
    BlockMesh bm;
    std::string er = build_blocks(nv, dim, n_nodes, n_owned, n_cells, coords, cells, cell_primary, max_own, bm);
    if (!er.empty()) return 1;
    int64_t np = 0;
    for (int b = 0; b < bm.nb; ++b) np += bm.nprim[b];
    stats[0] = bm.nb; stats[1] = bm.max_own; stats[2] = bm.max_loc; stats[3] = bm.max_elems; stats[4] = bm.total_elems; stats[5] = np;
    if (cell_type == 0) emulate_apply_u<Tri>(bm, m, n1d, coords, phi, u, v, chunk, y, dot);
    else if (cell_type == 1) emulate_apply_u<Quad>(bm, m, n1d, coords, phi, u, v, chunk, y, dot);
    else if (cell_type == 3) emulate_apply_u<Hex>(bm, m, n1d, coords, phi, u, v, chunk, y, dot);
    else emulate_apply_u<Tet>(bm, m, n1d, coords, phi, u, v, chunk, y, dot);
    return 0;
}
// aggregates of the PCG coarse space: returns sizes in stats = {n_agg, S, n_chunks, n_col, n_reg}; arrays are
// filled up to the given capacities: node_agg [n_nodes] in CALLER node ids (-1 for ghosts), chunk_node0,
// colour, agg_reg, nbrcol [n_agg*n_col], rel [n_nodes*dim] in caller node order
int hc_aggregates(int cell_type, int64_t n_nodes, int64_t n_owned, int64_t n_cells, const double* coords,
                  const int32_t* cells, int max_own, int g, int reg_aggs, int chunk_nodes, int64_t* stats, int32_t* node_agg,
                  int32_t* chunk_node0, int32_t* colour, int32_t* agg_reg, int32_t* nbrcol, int64_t nbrcol_cap, float* rel) {
    int nv = cell_type == 0 ? 3 : (cell_type == 3 ? 8 : 4), dim = cell_type >= 2 ? 3 : 2;
    BlockMesh bm;
    std::string er = build_blocks(nv, dim, n_nodes, n_owned, n_cells, coords, cells, nullptr, max_own, bm);
    if (!er.empty()) return 1;
    AggregateSet as;
    er = build_aggregates(bm, n_cells, coords, cells, g, reg_aggs, chunk_nodes, as);
    if (!er.empty()) return 2;
    stats[0] = as.n_agg; stats[1] = as.S; stats[2] = as.n_chunks; stats[3] = as.n_col; stats[4] = as.n_reg;
    if ((int64_t)as.nbrcol.size() > nbrcol_cap) return 3;
    for (int64_t i = 0; i < n_nodes; ++i) node_agg[i] = -1;
    for (int a = 0; a < as.n_agg; ++a)
        for (int i = as.agg_node0[a]; i < as.agg_node0[a + 1]; ++i) {
            node_agg[bm.perm[i]] = a;
            for (int k = 0; k < dim; ++k) rel[(int64_t)bm.perm[i] * dim + k] = as.rel[(int64_t)i * dim + k];
        }
    for (size_t i = 0; i < as.chunk_node0.size(); ++i) chunk_node0[i] = as.chunk_node0[i];
    for (int a = 0; a < as.n_agg; ++a) { colour[a] = as.colour[a]; agg_reg[a] = as.agg_reg[a]; }
    for (size_t i = 0; i < as.nbrcol.size(); ++i) nbrcol[i] = as.nbrcol[i];
    return 0;
}
}
