// pfx_monolithic.cuh — element forms and vector kernels of the energy-controlled monolithic solvers (sm_100a, fp64).
//
// Replaces the blocked forms of reference solver_ener_variational.py:167-203 / solver_ener_non_variational.py:166-201
// (F_u, F_phi with psi_a(u) directly — no history field —, J = d(F_u, F_phi)/d(u, phi)) on the block-scatter kernel of
// pfx_kernels.cuh, and the PETSc Vec kernels of the Krylov solver standing in for the MUMPS solve of the bordered
// system (flexible GMRES, pfx_engine.cu::ec_*).  Quadrature as in the other general forms: degree-5 rule on simplices,
// n1d x n1d Gauss on Q1.
#pragma once
#include "pfx_kernels.cuh"
#include "pfx_elements_coupled.cuh"

namespace pfx {

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
