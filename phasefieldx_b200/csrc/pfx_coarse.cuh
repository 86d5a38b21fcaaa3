// pfx_coarse.cuh — aggregate coarse space for the PCG: M^-1 = D^-1 + Z (Z^T A Z)^-1 Z^T.
//
// The reference factorises J_u and J_phi with MUMPS (solver.py:198-205, 237-244); the matrix-free
// PCG that replaces it needs O(1/h) iterations with a Jacobi preconditioner, which is what a
// 10^6-node load step spends its time on.  Aggregates are runs of consecutive CTA blocks of the
// RCB tree (spatially compact, contiguous node ranges); every aggregate carries its rigid-body
// modes (u: translations + rotations, phi: the constant).  The coarse matrix is small enough
// (<= kCoarseMax rows per region) that its inverse is kept explicitly, in fp32 (~100-150 MB), and
// the coarse solve is one dense mat-vec spread over the SMs.  Additive combination keeps M^-1
// symmetric positive definite for ANY symmetric positive definite coarse matrix, so Z^T A Z may
// lag behind the current tangent (it is refreshed once per load step / when iterations grow).
//
// Per PCG iteration the Jacobi path's update / direction kernels are replaced by
// (D = diagonal of A for phi, the nodal d x d diagonal blocks of A for u: block-Jacobi smoother)
//   pcg_update_coarse_kernel    x += a p ; r -= a Ap ; sums r.D^-1 r, r.r ; w = Z^T r per chunk
//   coarse_solve_kernel         y = Ainv w ; r.z = r.D^-1 r + w.y ; publishes the PCG scalars
//   pcg_direction_coarse_kernel p = D^-1 r + Z y + b p
// All sums have a fixed order (no atomics on data): results are run-to-run reproducible.
#pragma once
#include "pfx_kernels.cuh"

namespace pfx {

constexpr int kCoarseMax = 6144;     // most coarse dofs of one region (one dense inverse)
constexpr int kCoarseRegion = 1792;  // 2D default region size (coarse u-dofs)
constexpr int kCoarseTotal = 16384;  // 2D default cap of the whole coarse space (coarse u-dofs); 3D: one region of <= kCoarseMax
constexpr int kCoarseRowsPerWarp = 4;                              // coarse solve: consecutive rows per warp
constexpr int kCoarseRowsPerCta = kCoarseRowsPerWarp * (kThreads / 32);
constexpr int kCoarseChunkNodes = 256;                             // update / direction kernels: one warp per chunk of <= this many nodes
constexpr int kCoarseChunksPerCta = kThreads / 32;

// Regions: contiguous runs of aggregates.  The coarse matrix keeps only the couplings inside a
// region (block-diagonal Z^T A Z: still symmetric positive definite), so the inverse is one dense
// block per region and the coarse space can be several times larger than one dense inverse allows.
struct CoarseView {
    int n_agg, n_chunks, S, ld;       // S chunks (<= kCoarseChunkNodes nodes, one warp each) per aggregate; ld: largest row stride of a region
    const int32_t* chunk_node0;       // [n_chunks+1] owned-node range of every chunk
    const float* rel;                 // [n_owned*GD] (x - centroid(aggregate)) / size(aggregate)
    double* wpart;                    // [n_chunks*M] per-chunk pieces of Z^T r
    double* y;                        // [n_agg*M]
    const float* Ainv;                // per region r: [n_r*M][ld_r] at float offset reg_tab[r].z
    const int32_t* agg_reg;           // [n_agg] region of every aggregate
    const int4* reg_tab;              // [n_regions] {first aggregate, aggregates, float offset of the inverse, ld_r}
    const int2* cta_tab;              // coarse solve: [CTAs] {region, first row of the region handled by the CTA}
    const long long* ac_off;          // [n_regions] offset (doubles) of the region's block in the fp64 work matrix
};

template <int NC, int GD> struct Modes { static constexpr int M = NC == 1 ? 1 : (GD == 2 ? 3 : 6); };

// Third, rank-level coarse space of multi-GPU runs: the rank-local coarse matrices are block-diagonal over the ranks,
// so the modes that span several ranks (the bending of a beam cut into slabs) are lost and the iteration count grows
// with the number of GPUs (25 M-dof beam: 1.8 x at 8 GPUs).  One more additive term restores them:
//   M^-1 = B^-1 + Z A_loc^-1 Z^T + Z T E^-1 T^T Z^T,   E = T^T (Z^T A Z) T  (with the couplings ACROSS ranks),
// T = the rigid-body modes of a whole rank expressed in the modes of its aggregates ([n_agg*M][M] per rank), E of
// dimension n_ranks * M (<= 48), inverted on the host and replicated.  Per iteration every rank pushes its M numbers
// T^T w through the mailboxes (coarse3_reduce_kernel) and adds T y_g to its coarse solution (coarse_solve_kernel).
struct Coarse3View {
    const PeerView* pv = nullptr;     // null: third level off
    const double* T = nullptr;        // [n_agg*M][M] row-major
    const double* Einv = nullptr;     // [n_ranks*M][n_ranks*M]
};

// w += Z_node^T v   (modes: translations, then rotations about x, y, z / about z in 2D)
template <int NC, int GD>
__device__ __forceinline__ void restrict_add(const float* rl, const double* v, double* w) {
    if constexpr (NC == 1) {
        w[0] += v[0];
    } else if constexpr (GD == 2) {
        w[0] += v[0]; w[1] += v[1];
        w[2] += (double)rl[0] * v[1] - (double)rl[1] * v[0];
    } else {
        w[0] += v[0]; w[1] += v[1]; w[2] += v[2];
        w[3] += (double)rl[1] * v[2] - (double)rl[2] * v[1];
        w[4] += (double)rl[2] * v[0] - (double)rl[0] * v[2];
        w[5] += (double)rl[0] * v[1] - (double)rl[1] * v[0];
    }
}

// z = Z_node y
template <int NC, int GD>
__device__ __forceinline__ void prolong(const float* rl, const double* y, double* z) {
    if constexpr (NC == 1) {
        z[0] = y[0];
    } else if constexpr (GD == 2) {
        z[0] = y[0] - (double)rl[1] * y[2];
        z[1] = y[1] + (double)rl[0] * y[2];
    } else {
        z[0] = y[0] + (double)rl[2] * y[4] - (double)rl[1] * y[5];
        z[1] = y[1] - (double)rl[2] * y[3] + (double)rl[0] * y[5];
        z[2] = y[2] + (double)rl[1] * y[3] - (double)rl[0] * y[4];
    }
}

template <int NC, int GD>
__device__ __forceinline__ void load_rel(const CoarseView& cv, int nd, float* rl) {
    if constexpr (NC > 1) {
        if constexpr (GD == 2) {
            const float2 t = reinterpret_cast<const float2*>(cv.rel)[nd];
            rl[0] = t.x; rl[1] = t.y;
        } else {
            for (int k = 0; k < GD; ++k) rl[k] = cv.rel[(size_t)nd * GD + k];
        }
    }
}

template <int NC>
__device__ __forceinline__ void load_vec(const double* v, int nd, double* out) {
    if constexpr (NC == 2) {
        const double2 t = reinterpret_cast<const double2*>(v)[nd];
        out[0] = t.x; out[1] = t.y;
    } else {
        for (int k = 0; k < NC; ++k) out[k] = v[(size_t)nd * NC + k];
    }
}
template <int NC>
__device__ __forceinline__ void store_vec(double* v, int nd, const double* in) {
    if constexpr (NC == 2) {
        reinterpret_cast<double2*>(v)[nd] = make_double2(in[0], in[1]);
    } else {
        for (int k = 0; k < NC; ++k) v[(size_t)nd * NC + k] = in[k];
    }
}

// z = D^-1 r: diagonal (BLK = false, dinv[n*NC]) or nodal-block (BLK = true, inverse blocks
// [n*NC(NC+1)/2] in the order (00, 11[, 22], 01[, 02, 12])) Jacobi smoother
template <int NC, bool BLK>
__device__ __forceinline__ void smooth(const double* __restrict__ dinv, int nd, const double* rv, double* z) {
    if constexpr (!BLK) {
        double dv[NC];
        load_vec<NC>(dinv, nd, dv);
#pragma unroll
        for (int k = 0; k < NC; ++k) z[k] = dv[k] * rv[k];
    } else if constexpr (NC == 2) {
        const double* b = dinv + (size_t)nd * 3;
        const double b00 = b[0], b11 = b[1], b01 = b[2];
        z[0] = b00 * rv[0] + b01 * rv[1];
        z[1] = b01 * rv[0] + b11 * rv[1];
    } else {
        const double* b = dinv + (size_t)nd * 6;
        const double b00 = b[0], b11 = b[1], b22 = b[2], b01 = b[3], b02 = b[4], b12 = b[5];
        z[0] = b00 * rv[0] + b01 * rv[1] + b02 * rv[2];
        z[1] = b01 * rv[0] + b11 * rv[1] + b12 * rv[2];
        z[2] = b02 * rv[0] + b12 * rv[1] + b22 * rv[2];
    }
}

// x += alpha p ; r -= alpha Ap ; per-chunk Z^T r ; grid sums r.D^-1 r, r.r -> S[S_STAGE..+1]
// init != 0: r is only read (start of a solve; also ignores a stale DONE flag)
template <int NC, int GD, bool BLK>
__global__ void __launch_bounds__(kThreads) pcg_update_coarse_kernel(CoarseView cv, double* __restrict__ x, double* __restrict__ r,
                                                                     const double* __restrict__ p, const double* __restrict__ Ap,
                                                                     const double* __restrict__ dinv, double* S, int par,
                                                                     int init, double* partial, unsigned int* counter,
                                                                     const PeerView* pv = nullptr) {
    constexpr int M = Modes<NC, GD>::M;
    __shared__ double s_red[kMaxRed * 32];
    __shared__ bool is_last;
    __shared__ double s_pap;
    if (!init && S[S_DONE] != 0.0) return;
    double pap = S[S_PAP];
    if (pv && !init) {                          // fused all-reduce: p.Ap of all ranks from the mailboxes
        if (threadIdx.x < 32) {
            double sm[1];
            mbox_sum(pv, sm, 1);
            if (threadIdx.x == 0) s_pap = sm[0];
        }
        __syncthreads();
        pap = s_pap;
    }
    const double alpha = init ? 0.0 : S[S_RZ0 + par] / pap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sums[2] = {0.0, 0.0};
    // one warp per chunk (<= ~256 nodes): the restriction sums need only shuffles
    const int c1 = min(cv.n_chunks, (int)(blockIdx.x + 1) * kCoarseChunksPerCta);
    for (int ch = blockIdx.x * kCoarseChunksPerCta + warp; ch < c1; ch += kThreads / 32) {
        const int n0 = cv.chunk_node0[ch], n1 = cv.chunk_node0[ch + 1];
        double w[M];
#pragma unroll
        for (int k = 0; k < M; ++k) w[k] = 0.0;
        int nbeg = n0;
        if constexpr (NC == 1) {
            // scalar field: two nodes per lane through 128-bit accesses (chunks start on even nodes)
            if ((n0 & 1) == 0) {
                const int np2 = (n1 - n0) >> 1;
                double2* r2 = reinterpret_cast<double2*>(r + n0);
                double2* x2 = reinterpret_cast<double2*>(x + n0);
                const double2 *p2 = reinterpret_cast<const double2*>(p + n0), *A2 = reinterpret_cast<const double2*>(Ap + n0),
                              *d2 = reinterpret_cast<const double2*>(dinv + n0);
                for (int i = lane; i < np2; i += 32) {
                    double2 rv = r2[i];
                    const double2 dv = d2[i];
                    if (!init) {
                        const double2 pk = p2[i], av = A2[i];
                        double2 xv = x2[i];
                        rv.x -= alpha * av.x; rv.y -= alpha * av.y;
                        xv.x += alpha * pk.x; xv.y += alpha * pk.y;
                        r2[i] = rv; x2[i] = xv;
                    }
                    sums[0] += rv.x * dv.x * rv.x + rv.y * dv.y * rv.y;
                    sums[1] += rv.x * rv.x + rv.y * rv.y;
                    w[0] += rv.x + rv.y;
                }
                nbeg = n0 + 2 * np2;
            }
        }
#pragma unroll 2
        for (int nd = nbeg + lane; nd < n1; nd += 32) {
            double rv[NC], zv[NC];
            float rl[GD];
            load_vec<NC>(r, nd, rv);
            load_rel<NC, GD>(cv, nd, rl);
            if (!init) {
                double pk[NC], av[NC], xv[NC];
                load_vec<NC>(p, nd, pk);
                load_vec<NC>(Ap, nd, av);
                load_vec<NC>(x, nd, xv);
#pragma unroll
                for (int k = 0; k < NC; ++k) { rv[k] -= alpha * av[k]; xv[k] += alpha * pk[k]; }
                store_vec<NC>(r, nd, rv);
                store_vec<NC>(x, nd, xv);
            }
            smooth<NC, BLK>(dinv, nd, rv, zv);
#pragma unroll
            for (int k = 0; k < NC; ++k) { sums[0] += rv[k] * zv[k]; sums[1] += rv[k] * rv[k]; }
            restrict_add<NC, GD>(rl, rv, w);
        }
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const double t = warp_sum(w[k]);
            if (lane == 0) cv.wpart[(size_t)ch * M + k] = t;
        }
    }
    block_sum<2>(sums, s_red);
    if (threadIdx.x == 0) {
        partial[(size_t)blockIdx.x * 2 + 0] = sums[0];
        partial[(size_t)blockIdx.x * 2 + 1] = sums[1];
        __threadfence();
        unsigned int t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[2] = {0.0, 0.0};
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            acc[0] += __ldcg(&partial[(size_t)b * 2 + 0]);
            acc[1] += __ldcg(&partial[(size_t)b * 2 + 1]);
        }
        block_sum<2>(acc, s_red);
        if (threadIdx.x == 0) { S[S_STAGE] = acc[0]; S[S_STAGE + 1] = acc[1]; }
    }
}

// 128-bit read of the coarse inverse with an L2 evict-last policy: the matrix is re-read every
// iteration while ~300 MB of vectors stream past it
__device__ __forceinline__ float4 ld_keep(const float4* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

// y = Ainv w (fp32 matrix, fp64 accumulation), r.z = S[S_STAGE] + w.y.  A warp handles
// kCoarseRowsPerWarp consecutive rows of one region's inverse at once (each w entry read from shared
// memory once per kCoarseRowsPerWarp matrix entries), 2 x kCoarseRowsPerWarp 128-bit loads in flight per lane.
// mode bit 0: start of a solve (ignore a stale DONE flag); bit 1: only add w.y to S[S_STAGE] (an
// all-reduce / pcg_setup_kernel follows); mode 0: publish the PCG scalars, or with pv (fused
// multi-GPU path) push (r.z, r.r) of this rank to every rank's mailbox.
template <int M>
__global__ void __launch_bounds__(kThreads, 3) coarse_solve_kernel(CoarseView cv, double* S, int par, int mode, double* partial,
                                                                   unsigned int* counter, const PeerView* pv = nullptr,
                                                                   Coarse3View c3 = Coarse3View()) {
    constexpr int R = kCoarseRowsPerWarp;
    extern __shared__ __align__(16) double s_w[];          // [ld]
    __shared__ double s_red[kMaxRed * 32];
    __shared__ bool is_last;
    __shared__ double s_yg[8];
    if (!(mode & 1) && S[S_DONE] != 0.0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (c3.pv && threadIdx.x < 32) {
        // third level (pfx_coarse.cuh, "rank level"): every rank's restricted residual g_r = T_r^T w_r from the
        // mailboxes, y_g = (E^-1 g)[rows of this rank]; added to y of every aggregate below as T_a y_g
        double gall[kMaxRanks * 8];
        mbox_gather(c3.pv, gall, M);
        const int ng = c3.pv->n_ranks * M;
        if (lane < M) {
            const double* row = c3.Einv + (size_t)(c3.pv->rank * M + lane) * ng;
            double t = 0.0;
            for (int j = 0; j < ng; ++j) t += row[j] * gall[j];
            s_yg[lane] = t;
        }
    }
    const int2 ct = cv.cta_tab[blockIdx.x];
    const int4 rt = cv.reg_tab[ct.x];             // {first aggregate, aggregates, inverse offset, ld}
    const int nc = rt.y * M, ld4 = rt.w >> 2;
    const int row0 = ct.y + warp * R;             // this warp's R consecutive rows of the region
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const float4* A4[R];
#pragma unroll
    for (int r = 0; r < R; ++r) A4[r] = reinterpret_cast<const float4*>(cv.Ainv + (size_t)rt.z + (size_t)min(row0 + r, nc - 1) * rt.w);
    // two column blocks (2 x R loads of 128 bits) in flight per lane; the first pair is requested
    // before w is staged in shared memory
    float4 a0[R], a1[R];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        a0[r] = (lane < ld4) ? ld_keep(A4[r] + lane, pol) : zero4;
        a1[r] = (lane + 32 < ld4) ? ld_keep(A4[r] + lane + 32, pol) : zero4;
    }
    for (int i = threadIdx.x; i < rt.w; i += blockDim.x) {
        double t = 0.0;
        if (i < nc) {
            const int ag = rt.x + i / M, k = i % M;
            for (int s = 0; s < cv.S; ++s) t += __ldcg(&cv.wpart[((size_t)ag * cv.S + s) * M + k]);
        }
        s_w[i] = t;
    }
    __syncthreads();
    double wy[1] = {0.0};
    if (row0 < nc) {
        const double2* w2 = reinterpret_cast<const double2*>(s_w);
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
#pragma unroll 1
        for (int j4 = lane; j4 < ld4; j4 += 64) {
            float4 n0[R], n1[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                n0[r] = (j4 + 64 < ld4) ? ld_keep(A4[r] + j4 + 64, pol) : zero4;
                n1[r] = (j4 + 96 < ld4) ? ld_keep(A4[r] + j4 + 96, pol) : zero4;
            }
            {
                const double2 u0 = w2[2 * j4], u1 = w2[2 * j4 + 1];
#pragma unroll
                for (int r = 0; r < R; ++r)
                    acc[r] += (double)a0[r].x * u0.x + (double)a0[r].y * u0.y + (double)a0[r].z * u1.x + (double)a0[r].w * u1.y;
            }
            if (j4 + 32 < ld4) {
                const double2 u0 = w2[2 * (j4 + 32)], u1 = w2[2 * (j4 + 32) + 1];
#pragma unroll
                for (int r = 0; r < R; ++r)
                    acc[r] += (double)a1[r].x * u0.x + (double)a1[r].y * u0.y + (double)a1[r].z * u1.x + (double)a1[r].w * u1.y;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) { a0[r] = n0[r]; a1[r] = n1[r]; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double v = warp_sum(acc[r]);
            if (lane == 0 && row0 + r < nc) {
                if (c3.pv) {
                    const double* Ta = c3.T + ((size_t)rt.x * M + row0 + r) * M;       // row (aggregate, component) of T
#pragma unroll
                    for (int j = 0; j < M; ++j) v += Ta[j] * s_yg[j];
                }
                cv.y[(size_t)rt.x * M + row0 + r] = v; wy[0] += v * s_w[row0 + r];
            }
        }
    }
    block_sum<1>(wy, s_red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = wy[0];
        __threadfence();
        unsigned int t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[1] = {0.0};
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc[0] += __ldcg(&partial[b]);
        block_sum<1>(acc, s_red);
        if (pv && mode == 0) {
            if (threadIdx.x < 32) {
                double v[2];
                v[0] = S[S_STAGE] + __shfl_sync(0xffffffffu, acc[0], 0);
                v[1] = S[S_STAGE + 1];
                mbox_push(pv, v, 2);
            }
        } else if (threadIdx.x == 0) {
            const double rz = S[S_STAGE] + acc[0];
            if (mode & 2) S[S_STAGE] = rz;
            else pcg_finalize(S, par, rz, S[S_STAGE + 1]);
        }
    }
}

// p = D^-1 r + Z y + beta p   (Dirichlet dofs stay zero)
template <int NC, int GD, bool BLK>
__global__ void __launch_bounds__(kThreads) pcg_direction_coarse_kernel(CoarseView cv, double* __restrict__ p,
                                                                        const double* __restrict__ r,
                                                                        const double* __restrict__ dinv,
                                                                        const uint8_t* __restrict__ mask, double* S, int par,
                                                                        int init, const PeerView* pv = nullptr) {
    constexpr int M = Modes<NC, GD>::M;
    __shared__ double s_sum[2];
    if (!init && S[S_DONE] != 0.0) return;
    double rz_new = S[S_RZ0 + (par ^ 1)];
    if (pv && !init) {                          // fused all-reduce of (r.z, r.r); block 0 publishes them at its end
        if (threadIdx.x < 32) {
            double sm[2];
            mbox_sum(pv, sm, 2);
            if (threadIdx.x == 0) { s_sum[0] = sm[0]; s_sum[1] = sm[1]; }
        }
        __syncthreads();
        rz_new = s_sum[0];
    }
    const double beta = init ? 0.0 : rz_new / S[S_RZ0 + par];
    {                                            // one CTA per chunk (<= kCoarseChunkNodes nodes: one node per thread)
        const int ch = blockIdx.x, ag = ch / cv.S;
        const int n0 = cv.chunk_node0[ch], n1 = cv.chunk_node0[ch + 1];
        double y[M];
#pragma unroll
        for (int k = 0; k < M; ++k) y[k] = cv.y[(size_t)ag * M + k];
        int nbeg = n0;
        if constexpr (NC == 1) {
            if ((n0 & 1) == 0) {
                const int np2 = (n1 - n0) >> 1;
                double2* p2 = reinterpret_cast<double2*>(p + n0);
                const double2 *r2 = reinterpret_cast<const double2*>(r + n0), *d2 = reinterpret_cast<const double2*>(dinv + n0);
                const uchar2* m2 = reinterpret_cast<const uchar2*>(mask ? mask + n0 : nullptr);
                for (int i = threadIdx.x; i < np2; i += kThreads) {
                    const double2 rv = r2[i], dv = d2[i];
                    double2 pk = init ? make_double2(0.0, 0.0) : p2[i];
                    uchar2 fx = make_uchar2(0, 0);
                    if (mask) fx = m2[i];
                    pk.x = fx.x ? 0.0 : dv.x * rv.x + y[0] + beta * pk.x;
                    pk.y = fx.y ? 0.0 : dv.y * rv.y + y[0] + beta * pk.y;
                    p2[i] = pk;
                }
                nbeg = n0 + 2 * np2;
            }
        }
        for (int nd = nbeg + (int)threadIdx.x; nd < n1; nd += kThreads) {
            double rv[NC], zv[NC], z[NC], pk[NC];
            float rl[GD];
            load_vec<NC>(r, nd, rv);
            smooth<NC, BLK>(dinv, nd, rv, zv);
            load_rel<NC, GD>(cv, nd, rl);
            prolong<NC, GD>(rl, y, z);
            if (init) {
#pragma unroll
                for (int k = 0; k < NC; ++k) pk[k] = 0.0;
            } else {
                load_vec<NC>(p, nd, pk);
            }
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const bool fixed = mask && mask[(size_t)nd * NC + k];
                pk[k] = fixed ? 0.0 : zv[k] + z[k] + beta * pk[k];
            }
            store_vec<NC>(p, nd, pk);
        }
    }
    if (pv && !init && blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x == 0) pcg_finalize(S, par, s_sum[0], s_sum[1]);
    }
}

// ---- set-up of Z^T A Z by probing: one apply per (colour, mode) ---------------------------------
// v = sum over the aggregates of one colour of their mode `mode` (zero on Dirichlet dofs)
template <int NC, int GD>
__global__ void __launch_bounds__(kThreads) coarse_probe_kernel(CoarseView cv, const int32_t* agg_colour, int colour, int mode,
                                                                const uint8_t* mask, double* v) {
    constexpr int M = Modes<NC, GD>::M;
    const int ch = blockIdx.x, ag = ch / cv.S;
    const int n0 = cv.chunk_node0[ch], n1 = cv.chunk_node0[ch + 1];
    double y[M];
#pragma unroll
    for (int k = 0; k < M; ++k) y[k] = (k == mode && agg_colour[ag] == colour) ? 1.0 : 0.0;
    for (int nd = n0 + (int)threadIdx.x; nd < n1; nd += kThreads) {
        double z[NC];
        float rl[GD];
        load_rel<NC, GD>(cv, nd, rl);
        prolong<NC, GD>(rl, y, z);
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (mask && mask[(size_t)nd * NC + k]) z[k] = 0.0;
        store_vec<NC>(v, nd, z);
    }
}

template <int NC, int GD>
__global__ void __launch_bounds__(kThreads) coarse_restrict_kernel(CoarseView cv, const double* __restrict__ v) {
    constexpr int M = Modes<NC, GD>::M;
    __shared__ double s_red[kMaxRed * 32];
    const int ch = blockIdx.x;
    const int n0 = cv.chunk_node0[ch], n1 = cv.chunk_node0[ch + 1];
    double vals[M];
#pragma unroll
    for (int k = 0; k < M; ++k) vals[k] = 0.0;
    for (int nd = n0 + (int)threadIdx.x; nd < n1; nd += kThreads) {
        double vv[NC];
        float rl[GD];
        load_vec<NC>(v, nd, vv);
        load_rel<NC, GD>(cv, nd, rl);
        restrict_add<NC, GD>(rl, vv, vals);
    }
    block_sum<M>(vals, s_red);
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < M; ++k) cv.wpart[(size_t)ch * M + k] = vals[k];
}

// column (j, mode) of Z^T A Z from the restricted probe: rows of aggregate i belong to the one
// aggregate j of this colour within distance one of i (distance-2 colouring => unique); couplings
// between different regions are dropped
template <int M>
__global__ void coarse_fill_kernel(CoarseView cv, const int32_t* nbrcol, int n_col, int colour, int mode, double* Ac) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cv.n_agg * M) return;
    const int a = i / M, k = i - a * M;
    const int j = nbrcol[(size_t)a * n_col + colour];
    if (j < 0) return;
    const int rg = cv.agg_reg[a];
    if (cv.agg_reg[j] != rg) return;
    const int4 rt = cv.reg_tab[rg];
    double t = 0.0;
    for (int s = 0; s < cv.S; ++s) t += cv.wpart[((size_t)a * cv.S + s) * M + k];
    const size_t ncr = (size_t)rt.y * M;
    Ac[cv.ac_off[rg] + ((size_t)(j - rt.x) * M + mode) * ncr + (size_t)(a - rt.x) * M + k] = t;
}

// aggregates without a free dof in some mode leave an empty row/column: decouple them
__global__ void coarse_fixdiag_kernel(int nc, double* Ac, int ldc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nc && !(Ac[(size_t)i * ldc + i] > 0.0)) Ac[(size_t)i * ldc + i] = 1.0;
}

// fp32 full symmetric copy of the inverse held in the lower triangle (column-major) of Ac
__global__ void coarse_convert_kernel(int nc, const double* Ac, int ldc, float* Ainv, int ld) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= ld) return;
    float v = 0.f;
    if (col < nc) v = (float)(row >= col ? Ac[(size_t)col * ldc + row] : Ac[(size_t)row * ldc + col]);
    Ainv[(size_t)row * ld + col] = v;
}

// ---- third level ------------------------------------------------------------------------------------------
// v = Z T e_mode on the active rank (zero elsewhere, zero on Dirichlet dofs): probing vector of one column of E
template <int NC, int GD>
__global__ void __launch_bounds__(kThreads) coarse3_probe_kernel(CoarseView cv, const double* T, int active, int mode,
                                                                 const uint8_t* mask, double* v) {
    constexpr int M = Modes<NC, GD>::M;
    const int ch = blockIdx.x, ag = ch / cv.S;
    const int n0 = cv.chunk_node0[ch], n1 = cv.chunk_node0[ch + 1];
    double y[M];
#pragma unroll
    for (int k = 0; k < M; ++k) y[k] = active ? T[((size_t)ag * M + k) * M + mode] : 0.0;
    for (int nd = n0 + (int)threadIdx.x; nd < n1; nd += kThreads) {
        double z[NC];
        float rl[GD];
        load_rel<NC, GD>(cv, nd, rl);
        prolong<NC, GD>(rl, y, z);
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (mask && mask[(size_t)nd * NC + k]) z[k] = 0.0;
        store_vec<NC>(v, nd, z);
    }
}

// g = T^T w (w = Z^T r summed from the per-chunk pieces), pushed to every rank's mailbox.  One CTA.
// check_done != 0: skipped once the running PCG has converged (the launch sits inside the captured iteration).
template <int M>
__global__ void __launch_bounds__(kThreads) coarse3_reduce_kernel(CoarseView cv, const double* T, const double* S, int check_done,
                                                                  const PeerView* pv) {
    __shared__ double s_red[kMaxRed * 32];
    if (check_done && S[S_DONE] != 0.0) return;
    double g[M];
#pragma unroll
    for (int j = 0; j < M; ++j) g[j] = 0.0;
    for (int i = threadIdx.x; i < cv.n_agg * M; i += blockDim.x) {
        const int a = i / M, k = i - a * M;
        double w = 0.0;
        for (int s = 0; s < cv.S; ++s) w += __ldcg(&cv.wpart[((size_t)a * cv.S + s) * M + k]);
        const double* Ta = T + (size_t)i * M;
#pragma unroll
        for (int j = 0; j < M; ++j) g[j] += Ta[j] * w;
    }
    block_sum<M>(g, s_red);
    if (threadIdx.x < 32) {
#pragma unroll
        for (int j = 0; j < M; ++j) g[j] = __shfl_sync(0xffffffffu, g[j], 0);
        mbox_push(pv, g, M);
    }
}

// column `col` of E from the gathered g of all ranks (set-up; one warp)
template <int M>
__global__ void coarse3_collect_kernel(const PeerView* pv, int col, double* E) {
    double gall[kMaxRanks * 8];
    mbox_gather(pv, gall, M);
    const int ng = pv->n_ranks * M;
    if (threadIdx.x == 0)
        for (int i = 0; i < ng; ++i) E[(size_t)i * ng + col] = gall[i];
}

}  // namespace pfx
