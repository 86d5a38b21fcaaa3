// pfx_elements.cuh — per-element integrals of the staggered solver's forms, host+device.
//
// What FFCx-generated tabulate_tensor kernels do for dolfinx (not vendored in the reference):
// P1 on affine triangles / tetrahedra, Q1 on bilinear quadrilaterals (tensor vertex order).
// Forms follow pfx/Element/Phase_Field_Fracture/solver/solver.py:
//   F_u, J_u  :173-187     F_phi, J_phi :214-223 (AT2, c0 = 2)     projection RHS :355, :363
//   energies  pfx/Element/Phase_Field_Fracture/energy.py:13-117, pfx/Element/Phase_Field/energy.py:11-50
//   L2 norms  pfx/errors_functions.py:190-228
// Generic path = quadrature (simplices: degree-5 conical rule, exact for all integrands; Q1:
// n1d×n1d Gauss); simplex fast paths (closed-form P1 integrals) for the CG-hot applies.
#pragma once
#include "pfx_physics.cuh"
#include "pfx_quad_gen.cuh"

namespace pfx {

struct Tri  { static constexpr int NV = 3, D = 2; static constexpr bool simplex = true; };
struct Quad { static constexpr int NV = 4, D = 2; static constexpr bool simplex = false; };
struct Tet  { static constexpr int NV = 4, D = 3; static constexpr bool simplex = true; };
struct Hex  { static constexpr int NV = 8, D = 3; static constexpr bool simplex = false; };   // Q1, tensor vertex order x + 2y + 4z

#if defined(__CUDACC__)
static __constant__ double d_triq[9][3] = PFX_TRIQ_DATA;
static __constant__ double d_tetq[27][4] = PFX_TETQ_DATA;
static __constant__ double d_gauss[4][8] = PFX_GAUSS_DATA;
#endif
static const double h_triq[9][3] = PFX_TRIQ_DATA;
static const double h_tetq[27][4] = PFX_TETQ_DATA;
static const double h_gauss[4][8] = PFX_GAUSS_DATA;

#if defined(__CUDA_ARCH__)
#define PFX_TAB(name) d_##name
#else
#define PFX_TAB(name) h_##name
#endif

// ----------------------------------------------------------------------------- geometry
template <class Cell> struct Geo;

template <> struct Geo<Tri> {
    double G[3][2], vol;
    int nq;
    PFX_HD void init(const double* X, int) {
        double ax = X[2] - X[0], ay = X[3] - X[1], bx = X[4] - X[0], by = X[5] - X[1];
        double det = ax * by - ay * bx, inv = 1.0 / det;
        G[1][0] = by * inv;  G[1][1] = -bx * inv;
        G[2][0] = -ay * inv; G[2][1] = ax * inv;
        G[0][0] = -(G[1][0] + G[2][0]); G[0][1] = -(G[1][1] + G[2][1]);
        vol = 0.5 * ::fabs(det);
        nq = 9;
    }
    PFX_HD double eval(int q, double* N, double (*Gq)[2]) const {
        double x = PFX_TAB(triq)[q][0], y = PFX_TAB(triq)[q][1];
        N[0] = 1.0 - x - y; N[1] = x; N[2] = y;
        for (int a = 0; a < 3; ++a) { Gq[a][0] = G[a][0]; Gq[a][1] = G[a][1]; }
        return 2.0 * vol * PFX_TAB(triq)[q][2];
    }
};

template <> struct Geo<Tet> {
    double G[4][3], vol;
    int nq;
    PFX_HD void init(const double* X, int) {
        double J[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) J[i][j] = X[3 * (j + 1) + i] - X[i];   // dx_i/dxi_j
        double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, inv = 1.0 / det;
        // inverse Jinv[j][i] = dxi_j/dx_i = cof(J)[i][j]/det
        double Ji[3][3];
        Ji[0][0] = c00 * inv; Ji[1][0] = c01 * inv; Ji[2][0] = c02 * inv;
        Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv;
        Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv;
        Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv;
        Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
        Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv;
        Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
        for (int i = 0; i < 3; ++i) {
            G[1][i] = Ji[0][i]; G[2][i] = Ji[1][i]; G[3][i] = Ji[2][i];
            G[0][i] = -(Ji[0][i] + Ji[1][i] + Ji[2][i]);
        }
        vol = ::fabs(det) / 6.0;
        nq = 27;
    }
    PFX_HD double eval(int q, double* N, double (*Gq)[3]) const {
        double x = PFX_TAB(tetq)[q][0], y = PFX_TAB(tetq)[q][1], z = PFX_TAB(tetq)[q][2];
        N[0] = 1.0 - x - y - z; N[1] = x; N[2] = y; N[3] = z;
        for (int a = 0; a < 4; ++a) for (int i = 0; i < 3; ++i) Gq[a][i] = G[a][i];
        return 6.0 * vol * PFX_TAB(tetq)[q][3];
    }
};

template <> struct Geo<Quad> {
    double X[8];
    int n1d, nq;
    PFX_HD void init(const double* X_, int n1d_) {
        for (int i = 0; i < 8; ++i) X[i] = X_[i];
        n1d = n1d_; nq = n1d_ * n1d_;
    }
    PFX_HD double eval(int q, double* N, double (*Gq)[2]) const {
        int qi = q / n1d, qj = q - qi * n1d;
        double x = PFX_TAB(gauss)[n1d - 1][qi], y = PFX_TAB(gauss)[n1d - 1][qj];
        double w = PFX_TAB(gauss)[n1d - 1][4 + qi] * PFX_TAB(gauss)[n1d - 1][4 + qj];
        N[0] = (1 - x) * (1 - y); N[1] = x * (1 - y); N[2] = (1 - x) * y; N[3] = x * y;
        double dr[4][2] = {{-(1 - y), -(1 - x)}, {(1 - y), -x}, {-y, (1 - x)}, {y, x}};
        double J00 = 0, J01 = 0, J10 = 0, J11 = 0;          // J_ij = dx_i/dxi_j
        for (int a = 0; a < 4; ++a) {
            J00 += X[2 * a] * dr[a][0];     J01 += X[2 * a] * dr[a][1];
            J10 += X[2 * a + 1] * dr[a][0]; J11 += X[2 * a + 1] * dr[a][1];
        }
        double det = J00 * J11 - J01 * J10, inv = 1.0 / det;
        for (int a = 0; a < 4; ++a) {
            Gq[a][0] = (dr[a][0] * J11 - dr[a][1] * J10) * inv;
            Gq[a][1] = (-dr[a][0] * J01 + dr[a][1] * J00) * inv;
        }
        return ::fabs(det) * w;
    }
};

template <> struct Geo<Hex> {
    double X[24];
    int n1d, nq;
    PFX_HD void init(const double* X_, int n1d_) {
        for (int i = 0; i < 24; ++i) X[i] = X_[i];
        n1d = n1d_; nq = n1d_ * n1d_ * n1d_;
    }
    PFX_HD double eval(int q, double* N, double (*Gq)[3]) const {
        const int qk = q / (n1d * n1d), qj = (q / n1d) % n1d, qi = q % n1d;
        const double xi[3] = {PFX_TAB(gauss)[n1d - 1][qi], PFX_TAB(gauss)[n1d - 1][qj], PFX_TAB(gauss)[n1d - 1][qk]};
        const double w = PFX_TAB(gauss)[n1d - 1][4 + qi] * PFX_TAB(gauss)[n1d - 1][4 + qj] * PFX_TAB(gauss)[n1d - 1][4 + qk];
        double l[3][2], dl[3][2] = {{-1.0, 1.0}, {-1.0, 1.0}, {-1.0, 1.0}}, dr[8][3];
        for (int d = 0; d < 3; ++d) { l[d][0] = 1.0 - xi[d]; l[d][1] = xi[d]; }
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};             // J_ij = dx_i/dxi_j
        for (int a = 0; a < 8; ++a) {
            const int bx = a & 1, by = (a >> 1) & 1, bz = (a >> 2) & 1;
            N[a] = l[0][bx] * l[1][by] * l[2][bz];
            dr[a][0] = dl[0][bx] * l[1][by] * l[2][bz];
            dr[a][1] = l[0][bx] * dl[1][by] * l[2][bz];
            dr[a][2] = l[0][bx] * l[1][by] * dl[2][bz];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) J[i][j] += X[3 * a + i] * dr[a][j];
        }
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                     c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, inv = 1.0 / det;
        double Ji[3][3];                                                // Ji[j][i] = dxi_j/dx_i
        Ji[0][0] = c00 * inv; Ji[1][0] = c01 * inv; Ji[2][0] = c02 * inv;
        Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv;
        Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv;
        Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv;
        Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
        Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv;
        Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
        for (int a = 0; a < 8; ++a)
            for (int i = 0; i < 3; ++i) Gq[a][i] = dr[a][0] * Ji[0][i] + dr[a][1] * Ji[1][i] + dr[a][2] * Ji[2][i];
        return ::fabs(det) * w;
    }
};

// symmetric gradient of a nodal vector field W[NV][D] (interleaved) -> sym storage
template <class Cell>
PFX_HD void sym_grad(const double (*G)[Cell::D], const double* W, double* e) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int i = 0; i < SymN<D>::N; ++i) e[i] = 0.0;
    for (int a = 0; a < NV; ++a) {
        const double* w = W + a * D;
        if (D == 2) {
            e[0] += G[a][0] * w[0]; e[1] += G[a][1] * w[1];
            e[2] += 0.5 * (G[a][1] * w[0] + G[a][0] * w[1]);
        } else {
            e[0] += G[a][0] * w[0]; e[1] += G[a][1] * w[1]; e[2] += G[a][2] * w[2];
            e[3] += 0.5 * (G[a][1] * w[0] + G[a][0] * w[1]);
            e[4] += 0.5 * (G[a][2] * w[0] + G[a][0] * w[2]);
            e[5] += 0.5 * (G[a][2] * w[1] + G[a][1] * w[2]);
        }
    }
}

// out[a][i] += w * Σ_j S_ij G[a][j]
template <class Cell>
PFX_HD void add_div(const double (*G)[Cell::D], const double* S, double w, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int a = 0; a < NV; ++a) {
        if (D == 2) {
            out[2 * a]     += w * (S[0] * G[a][0] + S[2] * G[a][1]);
            out[2 * a + 1] += w * (S[2] * G[a][0] + S[1] * G[a][1]);
        } else {
            out[3 * a]     += w * (S[0] * G[a][0] + S[3] * G[a][1] + S[4] * G[a][2]);
            out[3 * a + 1] += w * (S[3] * G[a][0] + S[1] * G[a][1] + S[5] * G[a][2]);
            out[3 * a + 2] += w * (S[4] * G[a][0] + S[5] * G[a][1] + S[2] * G[a][2]);
        }
    }
}

template <int NV> PFX_HD double interp(const double* N, const double* f) {
    double s = 0.0;
    for (int a = 0; a < NV; ++a) s += N[a] * f[a];
    return s;
}

// P1 simplex closed forms: ∫N_a N_b = vol*m2(a,b), ∫N_a N_b N_c = vol*m3(a,b,c)
template <class Cell> PFX_HD double m2(int a, int b) {
    return (Cell::D == 2) ? ((a == b) ? 1.0 / 6.0 : 1.0 / 12.0) : ((a == b) ? 1.0 / 10.0 : 1.0 / 20.0);
}
template <class Cell> PFX_HD double m3(int a, int b, int c) {
    int same = (a == b) + (b == c) + (a == c);        // 0 distinct, 1 two equal, 3 all equal
    if (Cell::D == 2) return same == 3 ? 1.0 / 10.0 : (same == 1 ? 1.0 / 30.0 : 1.0 / 60.0);
    return same == 3 ? 1.0 / 20.0 : (same == 1 ? 1.0 / 60.0 : 1.0 / 120.0);
}
// mean over the simplex of g(φ) for P1 φ: with t = 1-φ and power sums p_k = Σ t_a^k,
// mean t² = (p1² + p2)·2 d!/(d+2)!/2,  mean t³ = (p1³ + 3 p1 p2 + 2 p3)·d!/(d+3)!  (complete homogeneous sums)
template <class Cell> PFX_HD double gbar_simplex(const double* phi, int dfun = 0) {
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int a = 0; a < Cell::NV; ++a) { double t = 1.0 - phi[a]; s1 += t; s2 += t * t; s3 += t * t * t; }
    const double m2 = (s2 + s1 * s1) * ((Cell::D == 2) ? 1.0 / 12.0 : 1.0 / 20.0);
    if (dfun == 0) return m2;
    const double m3 = (s1 * s1 * s1 + 3.0 * s1 * s2 + 2.0 * s3) * ((Cell::D == 2) ? 1.0 / 60.0 : 1.0 / 120.0);
    return 3.0 * m2 - 2.0 * m3;
}

// ============================================================================= u-space
// out[NV*D] = ∫ ((g(φ)+k) dσ_a + dσ_b) : ε(N_a e_i)   with dσ the tangent action on ε(v)
// (J_u·v, solver.py:186); nl = stress law nonlinear in u (needs U)
template <class Cell, int SM = -1>
PFX_HD void elem_apply_u(const Material& m_, int n1d, const double* X, const double* phi,
                         const double* U, const double* V, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    struct { double lam, mu, k; int smodel; } m = {m_.lam, m_.mu, m_.k, SM >= 0 ? SM : m_.smodel};
    Material mm = m_;
    mm.smodel = m.smodel;
    for (int i = 0; i < NV * D; ++i) out[i] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double e[NS] = {}, de[NS], dsa[NS], dsb[NS], S[NS];
    if constexpr (Cell::simplex) {
        double gb = gbar_simplex<Cell>(phi, m_.dfun) + m.k;
        if (m.smodel != M_ISO) sym_grad<Cell>(geo.G, U, e);
        sym_grad<Cell>(geo.G, V, de);
        dstress_split<D>(mm, m.smodel, e, de, dsa, dsb);
        for (int i = 0; i < NS; ++i) S[i] = gb * dsa[i] + dsb[i];
        add_div<Cell>(geo.G, S, geo.vol, out);
    } else {
        double N[NV], G[NV][D];
        for (int q = 0; q < geo.nq; ++q) {
            double w = geo.eval(q, N, G);
            double ph = interp<NV>(N, phi), gq = g_fun(m_.dfun, ph) + m.k;
            if (m.smodel != M_ISO) sym_grad<Cell>(G, U, e);
            sym_grad<Cell>(G, V, de);
            dstress_split<D>(mm, m.smodel, e, de, dsa, dsb);
            for (int i = 0; i < NS; ++i) S[i] = gq * dsa[i] + dsb[i];
            add_div<Cell>(G, S, w, out);
        }
    }
}

// ---- cached-tangent path for P1 simplices with a nonlinear split: the element
// tangent is constant over a simplex, so it is evaluated ONCE per Newton linearisation and the
// CG-hot apply only contracts it.  T[21] = upper triangle of A[i][j] = w_i·∂s_i/∂e_j
// (w = 1 for normal, 2 for shear components: A is the symmetric matrix of the bilinear form
// dε':C:dε in sym storage), already multiplied by vol and with g(φ)+k folded in.
template <class Cell>
PFX_HD void elem_tangent_u(const Material& m, const double* X, const double* phi, const double* U, double* T) {
    constexpr int D = Cell::D, NS = SymN<D>::N;
    Geo<Cell> geo; geo.init(X, 1);
    const double gb = gbar_simplex<Cell>(phi, m.dfun) + m.k;
    double e[NS], de[NS], dsa[NS], dsb[NS];
    sym_grad<Cell>(geo.G, U, e);
    int t = 0;
    double A[NS][NS];
    for (int j = 0; j < NS; ++j) {
        for (int i = 0; i < NS; ++i) de[i] = 0.0;
        de[j] = 1.0;
        dstress_split<D>(m, m.smodel, e, de, dsa, dsb);
        for (int i = 0; i < NS; ++i) A[i][j] = ((i < D) ? 1.0 : 2.0) * geo.vol * (gb * dsa[i] + dsb[i]);
    }
    for (int i = 0; i < NS; ++i)
        for (int j = i; j < NS; ++j) T[t++] = 0.5 * (A[i][j] + A[j][i]);
}

template <class Cell>
PFX_HD void elem_apply_u_tan(const double* X, const double* V, const double* T, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    for (int i = 0; i < NV * D; ++i) out[i] = 0.0;
    Geo<Cell> geo; geo.init(X, 1);
    double de[NS], S[NS];
    sym_grad<Cell>(geo.G, V, de);
    for (int i = 0; i < NS; ++i) S[i] = 0.0;
    int t = 0;
    for (int i = 0; i < NS; ++i)
        for (int j = i; j < NS; ++j) {
            const double a = T[t++];
            S[i] += a * de[j];
            if (j != i) S[j] += a * de[i];
        }
    for (int i = D; i < NS; ++i) S[i] *= 0.5;          // undo the shear weight
    add_div<Cell>(geo.G, S, 1.0, out);
}

// strain (sym storage) of the unit displacement N_a e_i on a simplex with shape gradients G
template <int D>
PFX_HD void unit_strain(const double* Ga, int i, double* de) {
    for (int s = 0; s < SymN<D>::N; ++s) de[s] = 0.0;
    de[i] = Ga[i];
    if (D == 2) { de[2] = 0.5 * Ga[1 - i]; }
    else {
        if (i == 0) { de[3] = 0.5 * Ga[1]; de[4] = 0.5 * Ga[2]; }
        if (i == 1) { de[3] = 0.5 * Ga[0]; de[5] = 0.5 * Ga[2]; }
        if (i == 2) { de[4] = 0.5 * Ga[0]; de[5] = 0.5 * Ga[1]; }
    }
}

// y = A x for the cached tangent (packed upper triangle T of the symmetric NS x NS matrix A of the bilinear form)
template <int NS>
PFX_HD void tan_mv(const double* T, const double* x, double* y) {
    for (int i = 0; i < NS; ++i) y[i] = 0.0;
    int t = 0;
    for (int i = 0; i < NS; ++i)
        for (int j = i; j < NS; ++j) {
            const double a = T[t++];
            y[i] += a * x[j];
            if (j != i) y[j] += a * x[i];
        }
}

// row a of the element matrix from the cached tangent: K[b][i][j] = de_(a,i)' A de_(b,j)  (NV blocks of D x D)
template <class Cell>
PFX_HD void elem_matrix_row_tan(const double (*G)[Cell::D], const double* T, int a, double* K) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    double dea[D][NS], de[NS], Ad[NS];
    for (int i = 0; i < D; ++i) unit_strain<D>(G[a], i, dea[i]);
    for (int b = 0; b < NV; ++b)
        for (int j = 0; j < D; ++j) {
            unit_strain<D>(G[b], j, de);
            tan_mv<NS>(T, de, Ad);
            for (int i = 0; i < D; ++i) {
                double q = 0.0;
                for (int s = 0; s < NS; ++s) q += dea[i][s] * Ad[s];
                K[(b * D + i) * D + j] = q;
            }
        }
}

// out[NV*D] = ∫ ((g(φ)+k) σ_a(u) + σ_b(u)) : ε(N_a e_i)     (F_u, solver.py:173-177)
template <class Cell>
PFX_HD void elem_residual_u(const Material& m, int n1d, const double* X, const double* phi,
                            const double* U, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    for (int i = 0; i < NV * D; ++i) out[i] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double e[NS], sa[NS], sb[NS], S[NS];
    if constexpr (Cell::simplex) {
        double gb = gbar_simplex<Cell>(phi, m.dfun) + m.k;
        sym_grad<Cell>(geo.G, U, e);
        stress_split<D, double>(m, m.smodel, e, sa, sb);
        for (int i = 0; i < NS; ++i) S[i] = gb * sa[i] + sb[i];
        add_div<Cell>(geo.G, S, geo.vol, out);
    } else {
        double N[NV], G[NV][D];
        for (int q = 0; q < geo.nq; ++q) {
            double w = geo.eval(q, N, G);
            double ph = interp<NV>(N, phi), gq = g_fun(m.dfun, ph) + m.k;
            sym_grad<Cell>(G, U, e);
            stress_split<D, double>(m, m.smodel, e, sa, sb);
            for (int i = 0; i < NS; ++i) S[i] = gq * sa[i] + sb[i];
            add_div<Cell>(G, S, w, out);
        }
    }
}

// out[NV*D] = diagonal of the element tangent
template <class Cell>
PFX_HD void elem_diag_u(const Material& m, int n1d, const double* X, const double* phi,
                        const double* U, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    for (int i = 0; i < NV * D; ++i) out[i] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double e[NS], de[NS], dsa[NS], dsb[NS], S[NS];
    double N[NV], G[NV][D];
    int nq = Cell::simplex ? 1 : geo.nq;
    for (int q = 0; q < nq; ++q) {
        double w, gq;
        if constexpr (Cell::simplex) {
            geo.eval(0, N, G); w = geo.vol; gq = gbar_simplex<Cell>(phi, m.dfun) + m.k;
        } else {
            w = geo.eval(q, N, G);
            double ph = interp<NV>(N, phi); gq = g_fun(m.dfun, ph) + m.k;
        }
        if (m.smodel != M_ISO) sym_grad<Cell>(G, U, e);
        for (int a = 0; a < NV; ++a)
            for (int i = 0; i < D; ++i) {
                // ε of the unit displacement N_a e_i
                for (int s = 0; s < NS; ++s) de[s] = 0.0;
                de[i] = G[a][i];
                if (D == 2) { de[2] = 0.5 * G[a][1 - i]; }
                else {
                    if (i == 0) { de[3] = 0.5 * G[a][1]; de[4] = 0.5 * G[a][2]; }
                    if (i == 1) { de[3] = 0.5 * G[a][0]; de[5] = 0.5 * G[a][2]; }
                    if (i == 2) { de[4] = 0.5 * G[a][0]; de[5] = 0.5 * G[a][1]; }
                }
                dstress_split<D>(m, m.smodel, e, de, dsa, dsb);
                for (int s = 0; s < NS; ++s) S[s] = gq * dsa[s] + dsb[s];
                out[a * D + i] += w * ddot<D, double>(S, de);
            }
    }
}

// out[NV*NB] = nodal diagonal blocks K_aa of the element tangent, NB = D(D+1)/2 values per vertex in
// the order (00, 11[, 22], 01[, 02, 12]) — block-Jacobi smoother of the two-level PCG
template <class Cell>
PFX_HD void elem_diagblk_u(const Material& m, int n1d, const double* X, const double* phi,
                           const double* U, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N, NB = D * (D + 1) / 2;
    for (int i = 0; i < NV * NB; ++i) out[i] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double e[NS], de[D][NS], dsa[NS], dsb[NS], S[D][NS];
    double N[NV], G[NV][D];
    int nq = Cell::simplex ? 1 : geo.nq;
    for (int q = 0; q < nq; ++q) {
        double w, gq;
        if constexpr (Cell::simplex) {
            geo.eval(0, N, G); w = geo.vol; gq = gbar_simplex<Cell>(phi, m.dfun) + m.k;
        } else {
            w = geo.eval(q, N, G);
            double ph = interp<NV>(N, phi); gq = g_fun(m.dfun, ph) + m.k;
        }
        if (m.smodel != M_ISO) sym_grad<Cell>(G, U, e);
        for (int a = 0; a < NV; ++a) {
            for (int i = 0; i < D; ++i) {
                for (int s = 0; s < NS; ++s) de[i][s] = 0.0;
                de[i][i] = G[a][i];
                if (D == 2) { de[i][2] = 0.5 * G[a][1 - i]; }
                else {
                    if (i == 0) { de[i][3] = 0.5 * G[a][1]; de[i][4] = 0.5 * G[a][2]; }
                    if (i == 1) { de[i][3] = 0.5 * G[a][0]; de[i][5] = 0.5 * G[a][2]; }
                    if (i == 2) { de[i][4] = 0.5 * G[a][0]; de[i][5] = 0.5 * G[a][1]; }
                }
                dstress_split<D>(m, m.smodel, e, de[i], dsa, dsb);
                for (int s = 0; s < NS; ++s) S[i][s] = gq * dsa[s] + dsb[s];
            }
            for (int i = 0; i < D; ++i) out[a * NB + i] += w * ddot<D, double>(S[i], de[i]);
            out[a * NB + D] += w * ddot<D, double>(S[0], de[1]);
            if (D == 3) {
                out[a * NB + 4] += w * ddot<D, double>(S[0], de[2]);
                out[a * NB + 5] += w * ddot<D, double>(S[1], de[2]);
            }
        }
    }
}

// ============================================================================= φ-space
// cm = 2H + F·Gc/l, cg = F·Gc·l as nodal (P1/Q1) coefficients.
// out[a] = Σ_b ∫ (cm N_a N_b + cg ∇N_a·∇N_b) p_b   (J_phi·p, solver.py:214-223)
template <class Cell>
PFX_HD void elem_apply_phi(int n1d, const double* X, const double* cm, const double* cg,
                           const double* p, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    if constexpr (Cell::simplex) {
        // ∫N_aN_bN_c = vol·base·(1 + δab + δac + δbc + 2δabc), base = 1/60 (tri), 1/120 (tet)
        double cgm = 0.0, gp[D], C = 0.0, P = 0.0, Q = 0.0;
        for (int i = 0; i < D; ++i) gp[i] = 0.0;
        for (int a = 0; a < NV; ++a) {
            cgm += cg[a]; C += cm[a]; P += p[a]; Q += cm[a] * p[a];
            for (int i = 0; i < D; ++i) gp[i] += geo.G[a][i] * p[a];
        }
        cgm *= geo.vol / NV;
        const double vb = geo.vol * ((D == 2) ? 1.0 / 60.0 : 1.0 / 120.0), cpq = C * P + Q;
        for (int a = 0; a < NV; ++a) {
            double gg = 0.0;
            for (int i = 0; i < D; ++i) gg += geo.G[a][i] * gp[i];
            out[a] = vb * (cpq + C * p[a] + cm[a] * P + 2.0 * cm[a] * p[a]) + cgm * gg;
        }
    } else {
        double N[NV], G[NV][D];
        for (int q = 0; q < geo.nq; ++q) {
            double w = geo.eval(q, N, G);
            double cmq = interp<NV>(N, cm), cgq = interp<NV>(N, cg), pq = interp<NV>(N, p);
            double gp[D];
            for (int i = 0; i < D; ++i) { gp[i] = 0.0; for (int a = 0; a < NV; ++a) gp[i] += G[a][i] * p[a]; }
            for (int a = 0; a < NV; ++a) {
                double gg = 0.0;
                for (int i = 0; i < D; ++i) gg += G[a][i] * gp[i];
                out[a] += w * (cmq * pq * N[a] + cgq * gg);
            }
        }
    }
}

// out[a] = (K_phi φ)_a − ∫ 2H N_a      (F_phi with g'(φ) = −2(1−φ), solver.py:214-221)
template <class Cell>
PFX_HD void elem_residual_phi(int n1d, const double* X, const double* cm, const double* cg,
                              const double* H, const double* phi, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    elem_apply_phi<Cell>(n1d, X, cm, cg, phi, out);
    Geo<Cell> geo; geo.init(X, n1d);
    if constexpr (Cell::simplex) {
        for (int a = 0; a < NV; ++a) {
            double s = 0.0;
            for (int c = 0; c < NV; ++c) s += H[c] * m2<Cell>(a, c);
            out[a] -= 2.0 * geo.vol * s;
        }
    } else {
        double N[NV], G[NV][D];
        for (int q = 0; q < geo.nq; ++q) {
            double w = geo.eval(q, N, G);
            double hq = interp<NV>(N, H);
            for (int a = 0; a < NV; ++a) out[a] -= 2.0 * w * hq * N[a];
        }
    }
}

template <class Cell>
PFX_HD void elem_diag_phi(int n1d, const double* X, const double* cm, const double* cg, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D];
    for (int q = 0; q < geo.nq; ++q) {
        double w = geo.eval(q, N, G);
        double cmq = interp<NV>(N, cm), cgq = interp<NV>(N, cg);
        for (int a = 0; a < NV; ++a) {
            double gg = 0.0;
            for (int i = 0; i < D; ++i) gg += G[a][i] * G[a][i];
            out[a] += w * (cmq * N[a] * N[a] + cgq * gg);
        }
    }
}

// consistent mass: out[a] = Σ_b ∫N_a N_b p_b   (Math/projection.py:35)
template <class Cell>
PFX_HD void elem_apply_mass(int n1d, const double* X, const double* p, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    Geo<Cell> geo; geo.init(X, n1d);
    if constexpr (Cell::simplex) {
        double sum = 0.0;
        for (int a = 0; a < NV; ++a) sum += p[a];
        double off = m2<Cell>(0, 1), dg = m2<Cell>(0, 0) - off;
        for (int a = 0; a < NV; ++a) out[a] = geo.vol * (off * sum + dg * p[a]);
    } else {
        for (int a = 0; a < NV; ++a) out[a] = 0.0;
        double N[NV], G[NV][D];
        for (int q = 0; q < geo.nq; ++q) {
            double w = geo.eval(q, N, G);
            double pq = interp<NV>(N, p);
            for (int a = 0; a < NV; ++a) out[a] += w * pq * N[a];
        }
    }
}

template <class Cell>
PFX_HD void elem_diag_mass(int n1d, const double* X, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D];
    for (int q = 0; q < geo.nq; ++q) {
        double w = geo.eval(q, N, G);
        for (int a = 0; a < NV; ++a) out[a] += w * N[a] * N[a];
    }
}

// Phase-field forms for a general degradation function g (solver.py:214-224), by quadrature (degree-5 rule on
// simplices: exact for the cubic borden function with P1 fields; n1d x n1d Gauss on Q1):
//   mode 0: out[a] = F_phi_a   = ∫ (dg(φ) H + F Gc/l φ) N_a + F Gc l ∇φ·∇N_a
//   mode 1: out[a] = (J_phi p)_a = ∫ (ddg(φ) H + F Gc/l) p N_a + F Gc l ∇p·∇N_a
//   mode 2: out[a] = diag(J_phi)_a
// F = fatigue degradation field (P1, interpolated) or 1.
template <class Cell>
PFX_HD void elem_phi_general(int mode, const Material& m, int n1d, const double* X, const double* phi, const double* H,
                             const double* Fat, const double* p, double* out) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D];
    const double* v = mode == 0 ? phi : p;
    for (int q = 0; q < geo.nq; ++q) {
        double w = geo.eval(q, N, G);
        const double ph = interp<NV>(N, phi), hq = interp<NV>(N, H), fq = m.fatigue ? interp<NV>(N, Fat) : 1.0;
        const double cg = fq * m.Gc * m.l;
        if (mode == 2) {
            const double cmq = ddg_fun(m.dfun, ph) * hq + fq * m.Gc / m.l;
            for (int a = 0; a < NV; ++a) {
                double gg = 0.0;
                for (int i = 0; i < D; ++i) gg += G[a][i] * G[a][i];
                out[a] += w * (cmq * N[a] * N[a] + cg * gg);
            }
            continue;
        }
        const double vq = interp<NV>(N, v);
        const double src = mode == 0 ? dg_fun(m.dfun, ph) * hq + fq * m.Gc / m.l * ph
                                     : (ddg_fun(m.dfun, ph) * hq + fq * m.Gc / m.l) * vq;
        double gv[D];
        for (int i = 0; i < D; ++i) { gv[i] = 0.0; for (int a = 0; a < NV; ++a) gv[i] += G[a][i] * v[a]; }
        for (int a = 0; a < NV; ++a) {
            double gg = 0.0;
            for (int i = 0; i < D; ++i) gg += G[a][i] * gv[i];
            out[a] += w * (src * N[a] + cg * gg);
        }
    }
}

// out[a] = ∫ ψ_a(u) N_a     (RHS of project(psi_a(u_new), V_c), solver.py:355); which = 1: ψ_b instead
// (solver_anisotropic.py:236-237 projects both parts for output)
template <class Cell>
PFX_HD void elem_rhs_psi(const Material& m, int n1d, const double* X, const double* U, double* out, int which = 0) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double e[NS], pa, pb;
    if constexpr (Cell::simplex) {
        sym_grad<Cell>(geo.G, U, e);
        energy_split<D>(m, m.emodel, e, pa, pb);
        for (int a = 0; a < NV; ++a) out[a] = geo.vol * (which ? pb : pa) / NV;
    } else {
        double N[NV], G[NV][D];
        for (int q = 0; q < geo.nq; ++q) {
            double w = geo.eval(q, N, G);
            sym_grad<Cell>(G, U, e);
            energy_split<D>(m, m.emodel, e, pa, pb);
            for (int a = 0; a < NV; ++a) out[a] += w * (which ? pb : pa) * N[a];
        }
    }
}

// out[a] = ∫ g(φ_old) V_c N_a     (RHS of project(g(Φ_old)*V_c, alpha_c), solver.py:363)
template <class Cell>
PFX_HD void elem_rhs_fatigue(int n1d, const double* X, const double* phi, const double* Vc, double* out, int dfun = 0) {
    constexpr int NV = Cell::NV, D = Cell::D;
    for (int a = 0; a < NV; ++a) out[a] = 0.0;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D];
    for (int q = 0; q < geo.nq; ++q) {
        double w = geo.eval(q, N, G);
        double ph = interp<NV>(N, phi), vq = interp<NV>(N, Vc);
        double gq = g_fun(dfun, ph);
        for (int a = 0; a < NV; ++a) out[a] += w * gq * vq * N[a];
    }
}

// ============================================================================= scalar reductions
// s[0] += ∫|a−b|², s[1] += ∫|a|²  for nodal fields with NC components (consistent mass)
template <class Cell, int NC>
PFX_HD void elem_l2(int n1d, const double* X, const double* A, const double* B, double* s) {
    constexpr int NV = Cell::NV, D = Cell::D;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D];
    for (int q = 0; q < geo.nq; ++q) {
        double w = geo.eval(q, N, G);
        for (int c = 0; c < NC; ++c) {
            double aq = 0.0, bq = 0.0;
            for (int a = 0; a < NV; ++a) { aq += N[a] * A[a * NC + c]; bq += N[a] * B[a * NC + c]; }
            s[0] += w * (aq - bq) * (aq - bq);
            s[1] += w * aq * aq;
        }
    }
}

// s[0..7] += ∫gψ_a, ∫ψ_a, ∫ψ_b, ∫φ², ∫|∇φ|², ∫Fφ², ∫F|∇φ|², ∫ᾱ
template <class Cell>
PFX_HD void elem_energies(const Material& m, int n1d, const double* X, const double* U, const double* phi,
                          const double* Fat, const double* abar, double* s) {
    constexpr int NV = Cell::NV, D = Cell::D, NS = SymN<D>::N;
    Geo<Cell> geo; geo.init(X, n1d);
    double N[NV], G[NV][D], e[NS], pa = 0.0, pb = 0.0;
    if constexpr (Cell::simplex) {                      // strain is cell-wise constant: split once
        geo.eval(0, N, G);
        sym_grad<Cell>(G, U, e);
        energy_split<D>(m, m.emodel, e, pa, pb);
    }
    for (int q = 0; q < geo.nq; ++q) {
        double w = geo.eval(q, N, G);
        if constexpr (!Cell::simplex) { sym_grad<Cell>(G, U, e); energy_split<D>(m, m.emodel, e, pa, pb); }
        double ph = interp<NV>(N, phi), g2 = 0.0;
        for (int i = 0; i < D; ++i) { double gi = 0.0; for (int a = 0; a < NV; ++a) gi += G[a][i] * phi[a]; g2 += gi * gi; }
        double gq = g_fun(m.dfun, ph);
        s[0] += w * gq * pa; s[1] += w * pa; s[2] += w * pb;
        s[3] += w * ph * ph; s[4] += w * g2;
        if (m.fatigue) {
            double fq = interp<NV>(N, Fat);
            s[5] += w * fq * ph * ph; s[6] += w * fq * g2; s[7] += w * interp<NV>(N, abar);
        }
    }
}

}  // namespace pfx
