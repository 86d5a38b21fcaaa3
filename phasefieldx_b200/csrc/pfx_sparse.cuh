// pfx_sparse.cuh — assembled operators for P1 tetrahedra in sliced-ELL layout (sm_100a, fp64).
//
// In 3D the per-cell data a matrix-free apply has to stream (the cached 6x6 element tangent of the nonlinear
// split, 168 B per tetrahedron ~ 1 kB per node) is as large as the assembled 3x3-block matrix itself
// (~15 blocks of 72 B per node), while contracting it costs ~10x the instructions of a block SpMV (geometry,
// strain, 21-entry contraction, shared-memory scatter and node-wise gather per cell).  So for tetrahedra the
// tangent is assembled ONCE per Newton linearisation — the role of dolfinx assemble_matrix on J_u / J_phi
// (reference solver.py:186-187, 223; here without any factorisation) — and every PCG iteration is one
// bandwidth-bound SpMV:
//   * rows = owned nodes in block order, a warp = one slice of 32 rows, entry k of the 32 rows side by side:
//     column indices and values are read fully coalesced, the input vector is gathered through L1/L2;
//   * assembly is row-wise too (one thread per row walks the cells incident to its node in a fixed order and
//     adds row a of their element matrices): atomic-free and bit-reproducible like everything else here;
//   * the diagonal blocks fall out of the assembly, so the Jacobi / block-Jacobi data needs no extra pass.
// The phase-field operator J_phi (coefficients 2H + F Gc/l, F Gc l) and the mass matrix of the L2 projection
// share the same graph with scalar entries.
#pragma once
#include "pfx_kernels.cuh"

namespace pfx {

struct SparseView {
    int64_t n_rows, n_slices;
    const int32_t* slice_k0;     // [n_slices+1]
    const int32_t* col;          // [total_k * 32]
    const int4* cell_nodes;      // [n_own_cells]
    const int32_t* nc_ptr;       // [n_rows+1]
    const int32_t* nc_code;      // (cell << 2 | vertex)
    const uint32_t* nc_kpos;     // entry positions of the cell's four nodes in the row
};

// ---- assembly of J_u from the cached element tangents; also the inverse diagonal and the inverse nodal diagonal
// blocks (block-Jacobi smoother), Dirichlet dofs decoupled as in OpDiagU / OpDiagBlkU
__global__ void __launch_bounds__(128) assemble_u_tet_kernel(SparseView sv, const double* __restrict__ X, const double* __restrict__ tan,
                                                             const uint8_t* __restrict__ mask, double* __restrict__ K,
                                                             double* __restrict__ dinv, double* __restrict__ dblk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.n_rows) return;
    const int lane = (int)(i & 31);
    const int64_t k0 = sv.slice_k0[i >> 5];
    const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
    for (int k = 0; k < len; ++k)
#pragma unroll
        for (int e = 0; e < 9; ++e) K[((k0 + k) * 9 + e) * 32 + lane] = 0.0;
    double Kd[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Kd[e] = 0.0;
    int kdiag = 0;
    for (int32_t q = sv.nc_ptr[i]; q < sv.nc_ptr[i + 1]; ++q) {
        const int32_t code = sv.nc_code[q];
        const uint32_t kp = sv.nc_kpos[q];
        const int64_t c = code >> 2;
        const int a = code & 3;
        const int4 nd = sv.cell_nodes[c];
        const int n4[4] = {nd.x, nd.y, nd.z, nd.w};
        double Xc[12], T[21], Kr[36];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int d = 0; d < 3; ++d) Xc[v * 3 + d] = X[(size_t)n4[v] * 3 + d];
#pragma unroll
        for (int t = 0; t < 21; ++t) T[t] = tan[(size_t)c * 21 + t];
        Geo<Tet> geo; geo.init(Xc, 1);
        elem_matrix_row_tan<Tet>(geo.G, T, a, Kr);
        kdiag = (int)((kp >> (8 * a)) & 255u);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == a) {
#pragma unroll
                for (int e = 0; e < 9; ++e) Kd[e] += Kr[b * 9 + e];
            } else {
                const int64_t k = k0 + (int)((kp >> (8 * b)) & 255u);
#pragma unroll
                for (int e = 0; e < 9; ++e) K[(k * 9 + e) * 32 + lane] += Kr[b * 9 + e];
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) K[((k0 + kdiag) * 9 + e) * 32 + lane] = Kd[e];
    bool fixed[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) fixed[d] = mask && mask[(size_t)i * 3 + d];
#pragma unroll
    for (int d = 0; d < 3; ++d) dinv[(size_t)i * 3 + d] = fixed[d] ? 1.0 : 1.0 / Kd[d * 4];
    if (dblk) {
        const double blk[6] = {Kd[0], Kd[4], Kd[8], 0.5 * (Kd[1] + Kd[3]), 0.5 * (Kd[2] + Kd[6]), 0.5 * (Kd[5] + Kd[7])};
        double inv[6];
        inv_sym_block<3>(blk, fixed, inv);
#pragma unroll
        for (int e = 0; e < 6; ++e) dblk[(size_t)i * 6 + e] = inv[e];
    }
}

// ---- assembly of K = sum_cells int (cm N_a N_b + cg grad N_a . grad N_b) with P1 nodal coefficients cm, cg
// (J_phi: cm = 2H + F Gc/l, cg = F Gc l; mass matrix: cm = 1, cg = 0) and its inverse diagonal (1 at masked dofs)
__global__ void __launch_bounds__(128) assemble_s_tet_kernel(SparseView sv, const double* __restrict__ X, const double* __restrict__ cm,
                                                             const double* __restrict__ cg, const uint8_t* __restrict__ mask,
                                                             double* __restrict__ K, double* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.n_rows) return;
    const int lane = (int)(i & 31);
    const int64_t k0 = sv.slice_k0[i >> 5];
    const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
    for (int k = 0; k < len; ++k) K[(k0 + k) * 32 + lane] = 0.0;
    double Kd = 0.0;
    int kdiag = 0;
    for (int32_t q = sv.nc_ptr[i]; q < sv.nc_ptr[i + 1]; ++q) {
        const int32_t code = sv.nc_code[q];
        const uint32_t kp = sv.nc_kpos[q];
        const int64_t c = code >> 2;
        const int a = code & 3;
        const int4 nd = sv.cell_nodes[c];
        const int n4[4] = {nd.x, nd.y, nd.z, nd.w};
        double Xc[12], m[4], C = 0.0, G = 0.0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
#pragma unroll
            for (int d = 0; d < 3; ++d) Xc[v * 3 + d] = X[(size_t)n4[v] * 3 + d];
            m[v] = cm[n4[v]]; C += m[v]; G += cg[n4[v]];
        }
        Geo<Tet> geo; geo.init(Xc, 1);
        const double vb = geo.vol * (1.0 / 120.0), cgm = geo.vol * 0.25 * G;
        kdiag = (int)((kp >> (8 * a)) & 255u);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double gg = geo.G[a][0] * geo.G[b][0] + geo.G[a][1] * geo.G[b][1] + geo.G[a][2] * geo.G[b][2];
            if (b == a) Kd += vb * (2.0 * C + 4.0 * m[a]) + cgm * gg;
            else K[(k0 + (int)((kp >> (8 * b)) & 255u)) * 32 + lane] += vb * (C + m[a] + m[b]) + cgm * gg;
        }
    }
    K[(k0 + kdiag) * 32 + lane] = Kd;
    if (dinv) dinv[i] = (mask && mask[i]) ? 1.0 : 1.0 / Kd;
}

// ---- y = K v (3x3 blocks); rows of Dirichlet dofs are written as zero (the input is zero there: PCG contract);
// the grid sum of v.y goes to `result` (or to the peers' mailboxes on the fused multi-GPU path)
__global__ void __launch_bounds__(kThreads) spmv_u_tet_kernel(SparseView sv, const double* __restrict__ K, const double* __restrict__ v,
                                                              double* __restrict__ y, const uint8_t* __restrict__ mask,
                                                              double* partial, double* result, unsigned int* counter,
                                                              const double* stop, const PeerView* pv) {
    __shared__ double s_red[kMaxRed * 32];
    if (stop && stop[0] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double red[1] = {0.0};
    if ((i >> 5) < sv.n_slices) {
        const int64_t k0 = sv.slice_k0[i >> 5];
        const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
        const int32_t* cp = sv.col + k0 * 32 + lane;
        const double* kp = K + k0 * 9 * 32 + lane;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        int cj = len > 0 ? cp[0] : 0;
        for (int k = 0; k < len; ++k) {
            const int cn = (k + 1 < len) ? cp[(k + 1) * 32] : 0;
            const double* vj = v + (size_t)cj * 3;
            const double v0 = vj[0], v1 = vj[1], v2 = vj[2];
            const double* e = kp + (size_t)k * 9 * 32;
            a0 += e[0] * v0 + e[32] * v1 + e[64] * v2;
            a1 += e[96] * v0 + e[128] * v1 + e[160] * v2;
            a2 += e[192] * v0 + e[224] * v1 + e[256] * v2;
            cj = cn;
        }
        if (i < sv.n_rows) {
            if (mask) {
                if (mask[(size_t)i * 3]) a0 = 0.0;
                if (mask[(size_t)i * 3 + 1]) a1 = 0.0;
                if (mask[(size_t)i * 3 + 2]) a2 = 0.0;
            }
            y[(size_t)i * 3] = a0; y[(size_t)i * 3 + 1] = a1; y[(size_t)i * 3 + 2] = a2;
            red[0] = v[(size_t)i * 3] * a0 + v[(size_t)i * 3 + 1] * a1 + v[(size_t)i * 3 + 2] * a2;
        }
    }
    grid_reduce<1>(red, partial, result, counter, s_red, pv);
}

// ---- y = K v (scalar entries)
__global__ void __launch_bounds__(kThreads) spmv_s_tet_kernel(SparseView sv, const double* __restrict__ K, const double* __restrict__ v,
                                                              double* __restrict__ y, const uint8_t* __restrict__ mask,
                                                              double* partial, double* result, unsigned int* counter,
                                                              const double* stop, const PeerView* pv) {
    __shared__ double s_red[kMaxRed * 32];
    if (stop && stop[0] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double red[1] = {0.0};
    if ((i >> 5) < sv.n_slices) {
        const int64_t k0 = sv.slice_k0[i >> 5];
        const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
        const int32_t* cp = sv.col + k0 * 32 + lane;
        const double* kp = K + k0 * 32 + lane;
        double acc0 = 0.0, acc1 = 0.0;
        int k = 0;
        for (; k + 1 < len; k += 2) {
            const int c0 = cp[k * 32], c1 = cp[(k + 1) * 32];
            acc0 += kp[(size_t)k * 32] * v[c0];
            acc1 += kp[(size_t)(k + 1) * 32] * v[c1];
        }
        if (k < len) acc0 += kp[(size_t)k * 32] * v[cp[k * 32]];
        double acc = acc0 + acc1;
        if (i < sv.n_rows) {
            if (mask && mask[i]) acc = 0.0;
            y[i] = acc;
            red[0] = v[i] * acc;
        }
    }
    grid_reduce<1>(red, partial, result, counter, s_red, pv);
}

// ============================================================================= Q1 quadrilaterals
// The matrix-free Q1 forms integrate 3 x 3 Gauss points per cell in EVERY PCG iteration; assembled once per Newton
// linearisation the same operator is a 2 x 2-block SpMV over ~9 entries per row.  Same graph / layout as above
// (four nodes per cell); the element rows come from quadrature instead of a cached constant tangent.

// row a of the element matrix of J_u at (u, phi): K[(b*2 + i)*2 + j] = int eps(N_a e_i) : dsigma(eps(N_b e_j))
PFX_HD void elem_matrix_row_u_quad(const Material& m, int n1d, const double* X, const double* phi, const double* U, int a, double* K) {
    for (int t = 0; t < 16; ++t) K[t] = 0.0;
    Geo<Quad> geo; geo.init(X, n1d);
    double N[4], G[4][2], e[3] = {0.0, 0.0, 0.0}, de[3], dsa[3], dsb[3];
    for (int q = 0; q < geo.nq; ++q) {
        const double w = geo.eval(q, N, G);
        const double gq = g_fun(m.dfun, interp<4>(N, phi)) + m.k;
        if (m.smodel != M_ISO) sym_grad<Quad>(G, U, e);
        for (int b = 0; b < 4; ++b)
            for (int j = 0; j < 2; ++j) {
                unit_strain<2>(G[b], j, de);
                dstress_split<2>(m, m.smodel, e, de, dsa, dsb);
                const double S0 = gq * dsa[0] + dsb[0], S1 = gq * dsa[1] + dsb[1], S2 = gq * dsa[2] + dsb[2];
                K[(b * 2 + 0) * 2 + j] += w * (S0 * G[a][0] + S2 * G[a][1]);
                K[(b * 2 + 1) * 2 + j] += w * (S2 * G[a][0] + S1 * G[a][1]);
            }
    }
}

__global__ void __launch_bounds__(128) assemble_u_quad_kernel(SparseView sv, Material m, int n1d, const double* __restrict__ X,
                                                              const double* __restrict__ u, const double* __restrict__ phi,
                                                              const uint8_t* __restrict__ mask, double* __restrict__ K,
                                                              double* __restrict__ dinv, double* __restrict__ dblk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.n_rows) return;
    const int lane = (int)(i & 31);
    const int64_t k0 = sv.slice_k0[i >> 5];
    const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
    for (int k = 0; k < len; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) K[((k0 + k) * 4 + e) * 32 + lane] = 0.0;
    double Kd[4] = {0.0, 0.0, 0.0, 0.0};
    int kdiag = 0;
    for (int32_t q = sv.nc_ptr[i]; q < sv.nc_ptr[i + 1]; ++q) {
        const int32_t code = sv.nc_code[q];
        const uint32_t kp = sv.nc_kpos[q];
        const int64_t c = code >> 2;
        const int a = code & 3;
        const int4 nd = sv.cell_nodes[c];
        const int n4[4] = {nd.x, nd.y, nd.z, nd.w};
        double Xc[8], Uc[8], ph[4], Kr[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            Xc[2 * v] = X[(size_t)n4[v] * 2]; Xc[2 * v + 1] = X[(size_t)n4[v] * 2 + 1];
            Uc[2 * v] = u[(size_t)n4[v] * 2]; Uc[2 * v + 1] = u[(size_t)n4[v] * 2 + 1];
            ph[v] = phi[n4[v]];
        }
        elem_matrix_row_u_quad(m, n1d, Xc, ph, Uc, a, Kr);
        kdiag = (int)((kp >> (8 * a)) & 255u);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == a) {
#pragma unroll
                for (int e = 0; e < 4; ++e) Kd[e] += Kr[b * 4 + e];
            } else {
                const int64_t k = k0 + (int)((kp >> (8 * b)) & 255u);
#pragma unroll
                for (int e = 0; e < 4; ++e) K[(k * 4 + e) * 32 + lane] += Kr[b * 4 + e];
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) K[((k0 + kdiag) * 4 + e) * 32 + lane] = Kd[e];
    bool fixed[2];
    fixed[0] = mask && mask[(size_t)i * 2]; fixed[1] = mask && mask[(size_t)i * 2 + 1];
    dinv[(size_t)i * 2] = fixed[0] ? 1.0 : 1.0 / Kd[0];
    dinv[(size_t)i * 2 + 1] = fixed[1] ? 1.0 : 1.0 / Kd[3];
    if (dblk) {
        const double blk[3] = {Kd[0], Kd[3], 0.5 * (Kd[1] + Kd[2])};
        double inv[3];
        inv_sym_block<2>(blk, fixed, inv);
        dblk[(size_t)i * 3] = inv[0]; dblk[(size_t)i * 3 + 1] = inv[1]; dblk[(size_t)i * 3 + 2] = inv[2];
    }
}

// scalar operator with P1/Q1 nodal coefficients on quadrilaterals: K = int (cm N_a N_b + cg grad N_a . grad N_b)
__global__ void __launch_bounds__(128) assemble_s_quad_kernel(SparseView sv, int n1d, const double* __restrict__ X, const double* __restrict__ cm,
                                                              const double* __restrict__ cg, const uint8_t* __restrict__ mask,
                                                              double* __restrict__ K, double* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.n_rows) return;
    const int lane = (int)(i & 31);
    const int64_t k0 = sv.slice_k0[i >> 5];
    const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
    for (int k = 0; k < len; ++k) K[(k0 + k) * 32 + lane] = 0.0;
    double Kd = 0.0;
    int kdiag = 0;
    for (int32_t q = sv.nc_ptr[i]; q < sv.nc_ptr[i + 1]; ++q) {
        const int32_t code = sv.nc_code[q];
        const uint32_t kp = sv.nc_kpos[q];
        const int64_t c = code >> 2;
        const int a = code & 3;
        const int4 nd = sv.cell_nodes[c];
        const int n4[4] = {nd.x, nd.y, nd.z, nd.w};
        double Xc[8], m4[4], g4[4], Kr[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            Xc[2 * v] = X[(size_t)n4[v] * 2]; Xc[2 * v + 1] = X[(size_t)n4[v] * 2 + 1];
            m4[v] = cm[n4[v]]; g4[v] = cg[n4[v]];
        }
        Geo<Quad> geo; geo.init(Xc, n1d);
        double N[4], G[4][2];
        for (int qp = 0; qp < geo.nq; ++qp) {
            const double w = geo.eval(qp, N, G);
            const double cmq = interp<4>(N, m4), cgq = interp<4>(N, g4);
#pragma unroll
            for (int b = 0; b < 4; ++b) Kr[b] += w * (cmq * N[a] * N[b] + cgq * (G[a][0] * G[b][0] + G[a][1] * G[b][1]));
        }
        kdiag = (int)((kp >> (8 * a)) & 255u);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == a) Kd += Kr[b];
            else K[(k0 + (int)((kp >> (8 * b)) & 255u)) * 32 + lane] += Kr[b];
        }
    }
    K[(k0 + kdiag) * 32 + lane] = Kd;
    if (dinv) dinv[i] = (mask && mask[i]) ? 1.0 : 1.0 / Kd;
}

// y = K v (2 x 2 blocks)
__global__ void __launch_bounds__(kThreads) spmv_u_quad_kernel(SparseView sv, const double* __restrict__ K, const double* __restrict__ v,
                                                               double* __restrict__ y, const uint8_t* __restrict__ mask,
                                                               double* partial, double* result, unsigned int* counter,
                                                               const double* stop, const PeerView* pv) {
    __shared__ double s_red[kMaxRed * 32];
    if (stop && stop[0] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double red[1] = {0.0};
    if ((i >> 5) < sv.n_slices) {
        const int64_t k0 = sv.slice_k0[i >> 5];
        const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
        const int32_t* cp = sv.col + k0 * 32 + lane;
        const double* kp = K + k0 * 4 * 32 + lane;
        const double2* v2 = reinterpret_cast<const double2*>(v);
        double a0 = 0.0, a1 = 0.0;
        for (int k = 0; k < len; ++k) {
            const double2 vj = v2[cp[k * 32]];
            const double* e = kp + (size_t)k * 4 * 32;
            a0 += e[0] * vj.x + e[32] * vj.y;
            a1 += e[64] * vj.x + e[96] * vj.y;
        }
        if (i < sv.n_rows) {
            if (mask) {
                if (mask[(size_t)i * 2]) a0 = 0.0;
                if (mask[(size_t)i * 2 + 1]) a1 = 0.0;
            }
            reinterpret_cast<double2*>(y)[i] = make_double2(a0, a1);
            const double2 vi = v2[i];
            red[0] = vi.x * a0 + vi.y * a1;
        }
    }
    grid_reduce<1>(red, partial, result, counter, s_red, pv);
}

// ============================================================================= P1 triangles
// Opt-in (PFX_TRI_SPARSE=1) and kept as a second implementation for the parity tests: measured on the 2 M-triangle bench
// mesh the 2 x 2-block SpMV (285 MB per apply) takes 88 us = 3.2 TB/s against 57 us for the matrix-free pipeline kernel
// of pfx_kernels.cuh, which therefore stays the default on triangles (scalar J_phi / mass: 64 us against 43 us).  With
// 7 entries of 32 B per row the sector-granular gathers of the input vector (7 x 32 sectors per warp and row slice)
// move as many bytes through L2 as the matrix stream itself — batching four entries per pass changed nothing (87.9 us)
// — whereas the matrix-free kernel stages the vector once, coalesced, in shared memory.
__global__ void __launch_bounds__(128) assemble_u_tri_kernel(SparseView sv, Material m, const double* __restrict__ X,
                                                             const double* __restrict__ u, const double* __restrict__ phi,
                                                             const uint8_t* __restrict__ mask, double* __restrict__ K,
                                                             double* __restrict__ dinv, double* __restrict__ dblk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.n_rows) return;
    const int lane = (int)(i & 31);
    const int64_t k0 = sv.slice_k0[i >> 5];
    const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
    for (int k = 0; k < len; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) K[((k0 + k) * 4 + e) * 32 + lane] = 0.0;
    double Kd[4] = {0.0, 0.0, 0.0, 0.0};
    int kdiag = 0;
    for (int32_t q = sv.nc_ptr[i]; q < sv.nc_ptr[i + 1]; ++q) {
        const int32_t code = sv.nc_code[q];
        const uint32_t kp = sv.nc_kpos[q];
        const int64_t c = code >> 2;
        const int a = code & 3;
        const int4 nd = sv.cell_nodes[c];
        const int n3[3] = {nd.x, nd.y, nd.z};
        double Xc[6], Uc[6], ph[3], T[6], Kr[12];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            Xc[2 * v] = X[(size_t)n3[v] * 2]; Xc[2 * v + 1] = X[(size_t)n3[v] * 2 + 1];
            Uc[2 * v] = u[(size_t)n3[v] * 2]; Uc[2 * v + 1] = u[(size_t)n3[v] * 2 + 1];
            ph[v] = phi[n3[v]];
        }
        elem_tangent_u<Tri>(m, Xc, ph, Uc, T);
        Geo<Tri> geo; geo.init(Xc, 1);
        elem_matrix_row_tan<Tri>(geo.G, T, a, Kr);
        kdiag = (int)((kp >> (8 * a)) & 255u);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b == a) {
#pragma unroll
                for (int e = 0; e < 4; ++e) Kd[e] += Kr[b * 4 + e];
            } else {
                const int64_t k = k0 + (int)((kp >> (8 * b)) & 255u);
#pragma unroll
                for (int e = 0; e < 4; ++e) K[(k * 4 + e) * 32 + lane] += Kr[b * 4 + e];
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) K[((k0 + kdiag) * 4 + e) * 32 + lane] = Kd[e];
    bool fixed[2];
    fixed[0] = mask && mask[(size_t)i * 2]; fixed[1] = mask && mask[(size_t)i * 2 + 1];
    dinv[(size_t)i * 2] = fixed[0] ? 1.0 : 1.0 / Kd[0];
    dinv[(size_t)i * 2 + 1] = fixed[1] ? 1.0 : 1.0 / Kd[3];
    if (dblk) {
        const double blk[3] = {Kd[0], Kd[3], 0.5 * (Kd[1] + Kd[2])};
        double inv[3];
        inv_sym_block<2>(blk, fixed, inv);
        dblk[(size_t)i * 3] = inv[0]; dblk[(size_t)i * 3 + 1] = inv[1]; dblk[(size_t)i * 3 + 2] = inv[2];
    }
}

// scalar operator with P1 nodal coefficients on triangles (closed form: int N_a N_b N_c = area/60 (1 + d_ab + d_ac + d_bc + 2 d_abc))
__global__ void __launch_bounds__(128) assemble_s_tri_kernel(SparseView sv, const double* __restrict__ X, const double* __restrict__ cm,
                                                             const double* __restrict__ cg, const uint8_t* __restrict__ mask,
                                                             double* __restrict__ K, double* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.n_rows) return;
    const int lane = (int)(i & 31);
    const int64_t k0 = sv.slice_k0[i >> 5];
    const int len = sv.slice_k0[(i >> 5) + 1] - (int)k0;
    for (int k = 0; k < len; ++k) K[(k0 + k) * 32 + lane] = 0.0;
    double Kd = 0.0;
    int kdiag = 0;
    for (int32_t q = sv.nc_ptr[i]; q < sv.nc_ptr[i + 1]; ++q) {
        const int32_t code = sv.nc_code[q];
        const uint32_t kp = sv.nc_kpos[q];
        const int64_t c = code >> 2;
        const int a = code & 3;
        const int4 nd = sv.cell_nodes[c];
        const int n3[3] = {nd.x, nd.y, nd.z};
        double Xc[6], m3v[3], C = 0.0, G = 0.0;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            Xc[2 * v] = X[(size_t)n3[v] * 2]; Xc[2 * v + 1] = X[(size_t)n3[v] * 2 + 1];
            m3v[v] = cm[n3[v]]; C += m3v[v]; G += cg[n3[v]];
        }
        Geo<Tri> geo; geo.init(Xc, 1);
        const double vb = geo.vol * (1.0 / 60.0), cgm = geo.vol * (1.0 / 3.0) * G;
        kdiag = (int)((kp >> (8 * a)) & 255u);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double gg = geo.G[a][0] * geo.G[b][0] + geo.G[a][1] * geo.G[b][1];
            if (b == a) Kd += vb * (2.0 * C + 4.0 * m3v[a]) + cgm * gg;
            else K[(k0 + (int)((kp >> (8 * b)) & 255u)) * 32 + lane] += vb * (C + m3v[a] + m3v[b]) + cgm * gg;
        }
    }
    K[(k0 + kdiag) * 32 + lane] = Kd;
    if (dinv) dinv[i] = (mask && mask[i]) ? 1.0 : 1.0 / Kd;
}

}  // namespace pfx
