// pfx_monolithic.cuh — element forms and vector kernels of the energy-controlled monolithic solvers (sm_100a, fp64).
//
// Replaces the blocked forms of reference solver_ener_variational.py:167-203 / solver_ener_non_variational.py:166-201
// (F_u, F_phi with psi_a(u) directly — no history field —, J = d(F_u, F_phi)/d(u, phi)) on the block-scatter kernel of
// pfx_kernels.cuh, and the PETSc Vec kernels of the Krylov solver standing in for the MUMPS solve of the bordered
// system (flexible GMRES, pfx_engine.cu::ec_*).  Quadrature as in the other general forms: degree-5 rule on simplices,
// n1d x n1d Gauss on Q1.
#pragma once
#include "pfx_kernels.cuh"

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

template <class Cell, int MODE>
struct OpResidualPhiE {    // staged: X u phi
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<2 * D + 1>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) { s[i] = a.X[(size_t)g * D + i]; s[D + i] = a.u[(size_t)g * D + i]; }
        s[2 * D] = a.phi[g];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], U[NV * D], ph[NV];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + i]; }
            ph[k] = nd[k][2 * D];
        }
        elem_residual_phi_e<Cell>(MODE, a.m, a.n1d, X, U, ph, a.coef, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) { a.y[g] = acc[0]; }
};

template <class Cell>
struct OpApplyUP {         // staged: X u phi v_u v_phi ; rows of Dirichlet u-dofs are written as zero (input zero there)
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = D + 1, NRED = 0;
    static constexpr int NF = odd<3 * D + 2>::value;
    static constexpr int MINB = 1;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) {
            s[i] = a.X[(size_t)g * D + i];
            s[D + i] = a.u[(size_t)g * D + i];
            double vv = a.v[(size_t)g * D + i];
            if (a.mask && a.mask[(size_t)g * D + i]) vv = 0.0;
            s[2 * D + 1 + i] = vv;
        }
        s[2 * D] = a.phi[g];
        s[3 * D + 1] = a.v2[g];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], U[NV * D], ph[NV], Vu[NV * D], Vp[NV];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + i]; Vu[k * D + i] = nd[k][2 * D + 1 + i]; }
            ph[k] = nd[k][2 * D]; Vp[k] = nd[k][3 * D + 1];
        }
        elem_apply_up<Cell>(a.m, a.n1d, X, ph, U, Vu, Vp, a.coef, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) {
        for (int i = 0; i < D; ++i) a.y[(size_t)g * D + i] = (a.mask && a.mask[(size_t)g * D + i]) ? 0.0 : acc[i];
        a.y2[g] = acc[D];
    }
};

// ----------------------------------------------------------------------------- Krylov vector kernels
// out[j] = sum_i V[j][i] * w[i] for j < k (k <= kMaxBasis), fixed summation order: one pass over w, every basis
// vector streamed once.  Grid-stride partials in `partial[block][k]`, last block reduces.
constexpr int kMaxBasis = 64;

__global__ void __launch_bounds__(kThreads) multi_dot_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ld,
                                                             const double* __restrict__ w, double* partial, double* result,
                                                             unsigned int* counter) {
    __shared__ double s_red[32];
    __shared__ unsigned int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j0 = 0; j0 < k; j0 += 8) {
        const int nj = min(8, k - j0);
        double acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const double wi = w[i];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nj) acc[j] += V[(int64_t)(j0 + j) * ld + i] * wi;
        }
        for (int j = 0; j < nj; ++j) {
            double v = warp_sum(acc[j]);
            __syncthreads();
            if (lane == 0) s_red[warp] = v;
            __syncthreads();
            if (warp == 0) {
                double t = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : 0.0;
                t = warp_sum(t);
                if (lane == 0) partial[(size_t)blockIdx.x * kMaxBasis + j0 + j] = t;
            }
        }
    }
    __threadfence();
    if (threadIdx.x == 0) s_last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int j = warp; j < k; j += (int)(blockDim.x >> 5)) {
        double t = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) t += __ldcg(&partial[(size_t)b * kMaxBasis + j]);
        t = warp_sum(t);
        if (lane == 0) result[j] = t;
    }
}

// w[i] -= sum_j h[j] V[j][i]   (h on the device: the result of multi_dot_kernel)
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ld,
                                                              const double* __restrict__ h, double sign, double* __restrict__ w) {
    __shared__ double s_h[kMaxBasis];
    if ((int)threadIdx.x < k) s_h[threadIdx.x] = sign * h[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = w[i];
        for (int j = 0; j < k; ++j) acc += s_h[j] * V[(int64_t)j * ld + i];
        w[i] = acc;
    }
}

__global__ void scale_copy_kernel(int64_t n, double* out, const double* in, double s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = s * in[i];
}

// out[i] = a[i] + s * b[i]
__global__ void add_scaled_kernel(int64_t n, double* out, const double* a, double s, const double* b) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i] + s * b[i];
}

// masked copy: out[i] = mask[i] ? 0 : in[i]
__global__ void masked_copy_kernel(int64_t n, const uint8_t* mask, const double* in, double* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (mask && mask[i]) ? 0.0 : in[i];
}

// lumped nodal value of a load vector: out[i] = max(b[i] * dinv_mass_lumped..., 0): H_i = b_i / m_i with m_i = int N_i
__global__ void lumped_divide_kernel(int64_t n, const double* b, const double* mlump, double* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fmax(b[i] / mlump[i], 0.0);
}

__global__ void fill_kernel(int64_t n, double* out, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}

}  // namespace pfx
