// pfx_kernels.cuh — sm_100a kernels of the staggered load step (fp64, HBM-bound, no tensor cores).
//
// Block-scatter kernels: one CTA per mesh block (pfx_blocks.h).  Phase A stages the nodal
// fields of the block's owned + halo nodes in shared memory (owned range: coalesced; halo:
// gathered, normally L2 hits because neighbouring blocks are scheduled together).  Phase B
// (one thread per element, chunked) evaluates the element vector in registers and parks it in
// a shared element buffer.  Phase C (one thread per owned node) sums the node's incident
// entries in a fixed order — atomic-free, bit-reproducible — and writes its rows coalesced.
// Optional epilogue: per-block partial sums (dot products, norms) + last-block final reduce.
//
// Replaces (reference call sites): dolfinx assemble_vector/assemble_matrix on F_u/J_u/F_phi/
// J_phi (solver.py:173-245), the mass assembly of Math/projection.py:35-44, assemble_scalar of
// energy.py / errors_functions.py, and PETSc Vec/KSP kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pfx_elements.cuh"
#include "pfx_blocks.h"

namespace pfx {

constexpr int kThreads = 256;
constexpr int kMaxRed = 8;

struct BlockView {
    int nb;
    const int32_t* node0;
    const int32_t* halo_off;
    const int32_t* halo;
    const int32_t* elem_off;
    const int32_t* nprim;
    const ushort4* lconn;
    const int32_t* inc_ptr;
    const uint16_t* inc;
    int max_loc;      // smem sizing
    int max_inc;      // largest per-block incidence list
    int max_own;      // largest owned-node count of a block
    const int4* meta; // [2*nb] packed {n0, n_own, h0, n_halo | e0, n_elems, inc0, n_inc}
    const uint4* inc8; // triangles: 8 incidence codes per owned node
    int chunk;        // elements per phase-B pass
    // own cells (pfx_blocks.h): every local cell leads the list of exactly one block; compact index own_c0[b] + j
    const int32_t* nlow;       // [nb] own cells
    const int32_t* own_c0;     // [nb+1] compact own-cell index
};

struct PeerView;

struct KArgs {
    Material m;
    int n1d;
    const double* X;        // [N][D] compact coordinates, block order
    const double* u;        // [N*D]
    const double* phi;      // [N]
    const double* v;        // input vector (u- or phi-space)
    const double* H;
    const double* Fat;
    const double* f1;       // op-specific extra nodal fields
    const double* f2;
    const uint8_t* mask;    // Dirichlet flags per dof (may be null)
    double* y;              // output vector
    double* partial;        // [nb][kMaxRed]
    double* result;         // [kMaxRed] final sums (written by the last block)
    unsigned int* counter;
    const double* stop;     // optional sticky DONE flag of the running PCG: skip the launch when set
    int input_masked;       // caller guarantees v == 0 on Dirichlet dofs (PCG directions)
    int wait_halo;          // kernel waits for the halo itself (no separate halo_wait_kernel)
    int aux;                // op-specific selector (OpRhsPsi: 0 = psi_a, 1 = psi_b)
    const PeerView* pv;     // multi-GPU fused path: wait for the halo in the prologue, push the partial
                            // dot product to the peers' mailboxes in the epilogue (null otherwise)
    const double* cm;       // precomputed phi-operator coefficients 2H + F·Gc/l and F·Gc·l (per node)
    const double* cg;
    double* tan;            // cached element tangents [n_own_cells][21] (tets), compact own-cell index
    // monolithic (u, phi) operators of the energy-controlled solvers (pfx_monolithic.cuh)
    const double* v2;       // phi-part of the input vector
    double* y2;             // phi-part of the output
    double coef;            // Gc + lambda c1 (variational scheme) or Gc
};

constexpr int kMaxRanks = 8;
constexpr int kMboxStride = 16;              // doubles per (slot, rank): 8 payload + sequence word
constexpr long long kSpinLimit = 20000000LL;   // ~10 s, then S[S_ERR] is raised instead of hanging

struct PeerView {
    int rank, n_ranks, n_nbrs;
    int nbr_rank[kMaxRanks];
    double* mbox_local;                      // [4][n_ranks][kMboxStride]
    double* mbox_peer[kMaxRanks];            // mailbox base of every rank
    unsigned long long* hflag_local;         // [n_ranks]: last halo sequence pushed by rank r
    unsigned long long* hflag_peer[kMaxRanks];
    unsigned long long* seq;                 // local counters: [0] all-reduce, [1] halo
    double* err;                             // S[S_ERR]: set on spin timeout
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// S-array slots of the PCG (device-resident scalars)
enum { S_PAP = 0, S_RZ0 = 1, S_RZ1 = 2, S_RR = 3, S_THRESH = 4, S_ITERS = 5, S_DONE = 6, S_STAGE = 8, S_BREF = 10 /* ..12: largest ||b||^2 of this load step per operator */, S_ERR = 15, S_COUNT = 16 };

// warp 0, all 32 lanes: publish n <= 8 partial sums of this rank to every rank's mailbox
__device__ __forceinline__ void mbox_push(const PeerView* pv, const double* vals, int n) {
    const int lane = threadIdx.x & 31;
    unsigned long long sq = 0;
    if (lane == 0) { sq = pv->seq[0] + 1ull; pv->seq[0] = sq; }
    sq = __shfl_sync(0xffffffffu, sq, 0);
    if (lane < pv->n_ranks) {
        double* dst = pv->mbox_peer[lane] + ((size_t)(sq & 3ull) * pv->n_ranks + pv->rank) * kMboxStride;
        for (int k = 0; k < n; ++k) dst[k] = vals[k];          // every lane holds the warp-reduced sums
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(dst + 8), sq);
    }
}

// warp 0, all 32 lanes: wait for the current all-reduce sequence and return the rank-ordered sums
// (identical on every rank and in every CTA); valid in every lane of the warp
__device__ __forceinline__ void mbox_sum(const PeerView* pv, double* sums, int n) {
    const int lane = threadIdx.x & 31;
    const unsigned long long sq = pv->seq[0];
    double v[2] = {0.0, 0.0};
    if (lane < pv->n_ranks) {
        const double* src = pv->mbox_local + ((size_t)(sq & 3ull) * pv->n_ranks + lane) * kMboxStride;
        long long spins = 0;
        while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(src + 8)) != sq) {
            if (++spins > kSpinLimit) { *pv->err = 2.0; break; }
        }
        for (int k = 0; k < n; ++k) v[k] = __ldcv(src + k);
    }
    for (int k = 0; k < n; ++k) {
        double t = 0.0;
        for (int r = 0; r < pv->n_ranks; ++r) t += __shfl_sync(0xffffffffu, v[k], r);
        sums[k] = t;
    }
}

// warp 0, all 32 lanes: wait for the current all-reduce sequence and return every rank's n values unsummed
// (out[r * n + k], identical on every rank and in every CTA); valid in every lane of the warp
__device__ __forceinline__ void mbox_gather(const PeerView* pv, double* out, int n) {
    const int lane = threadIdx.x & 31;
    const unsigned long long sq = pv->seq[0];
    double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (lane < pv->n_ranks) {
        const double* src = pv->mbox_local + ((size_t)(sq & 3ull) * pv->n_ranks + lane) * kMboxStride;
        long long spins = 0;
        while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(src + 8)) != sq) {
            if (++spins > kSpinLimit) { *pv->err = 2.0; break; }
        }
        for (int k = 0; k < n; ++k) v[k] = __ldcv(src + k);
    }
    for (int r = 0; r < pv->n_ranks; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < n) out[r * n + k] = __shfl_sync(0xffffffffu, v[k], r);
}

// every thread of the CTA: wait until all neighbours pushed the current halo sequence
__device__ __forceinline__ void halo_wait_cta(const PeerView* pv) {
    if ((int)threadIdx.x < pv->n_nbrs) {
        const unsigned long long want = pv->seq[1];
        const unsigned long long* f = pv->hflag_local + pv->nbr_rank[threadIdx.x];
        long long spins = 0;
        while (ld_acquire_sys(f) < want) {
            if (++spins > kSpinLimit) { *pv->err = 1.0; break; }
        }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum of up to NR values per thread; result valid in thread 0.
template <int NR>
__device__ __forceinline__ void block_sum(double* vals, double* s_red /*[NR*32]*/) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        double v = warp_sum(vals[k]);
        if (lane == 0) s_red[k * 32 + w] = v;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double v = (lane < nw) ? s_red[k * 32 + lane] : 0.0;
            vals[k] = warp_sum(v);
        }
    }
    __syncthreads();
}

// Per-block partials -> global result by the last block to finish (fixed summation order).
// After the block sum only warp 0 stays: lane 0 publishes the partial and takes a ticket; the
// last block's warp 0 adds the partials in a fixed order (deterministic result).
template <int NR>
__device__ __forceinline__ void grid_reduce(double* vals, double* partial, double* result, unsigned int* counter,
                                            double* s_red, const PeerView* pv = nullptr) {
    block_sum<NR>(vals, s_red);
    if (threadIdx.x >= 32) return;
    unsigned int last = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) partial[(size_t)blockIdx.x * kMaxRed + k] = vals[k];
        __threadfence();
        unsigned int t = atomicInc(counter, gridDim.x - 1);
        last = (t == gridDim.x - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    double acc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32)
#pragma unroll
        for (int k = 0; k < NR; ++k) acc[k] += __ldcg(&partial[(size_t)b * kMaxRed + k]);
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = warp_sum(acc[k]);
    if (pv) { mbox_push(pv, acc, NR); return; }            // fused all-reduce: straight to the peers' mailboxes
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < NR; ++k) result[k] = acc[k];
}

// ============================================================================= scatter ops
// Each Op: NF staged doubles per local node (odd => conflict-free AoS), NOUT rows per node,
// NRED partial sums; load(), elem(), store().
enum { OP_APPLY_U, OP_RESIDUAL_U, OP_DIAG_U, OP_APPLY_PHI, OP_RESIDUAL_PHI, OP_DIAG_PHI, OP_APPLY_MASS,
       OP_DIAG_MASS, OP_RHS_PSI, OP_RHS_FATIGUE, OP_TANGENT_U, OP_DIAGBLK_U };

template <int V> struct odd { static constexpr int value = (V & 1) ? V : V + 1; };
template <class Op, class = void> struct min_blocks { static constexpr int value = (Op::NV == 3) ? 4 : (Op::D == 2 ? 3 : (Op::NOUT == 1 ? 2 : 1)); };
template <class Op> struct min_blocks<Op, decltype((void)Op::MINB)> { static constexpr int value = Op::MINB; };
template <class Op, class = void> struct own_only { static constexpr bool value = false; };
template <class Op> struct own_only<Op, decltype((void)Op::OWN)> { static constexpr bool value = Op::OWN; };

template <class Cell, int SM>
struct OpApplyU {          // staged: X[D] phi v[D] (u[D]); SM = compile-time stress split
    static constexpr bool NL = (SM != M_ISO);
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = D, NRED = 1;
    static constexpr int MINB = (NV == 3) ? 4 : (D == 2 ? 2 : 1);
    static constexpr int NF = odd<D + 1 + D + (NL ? D : 0)>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        s[D] = a.phi[g];
        for (int i = 0; i < D; ++i) {
            double vv = a.v[(size_t)g * D + i];
            if (a.mask && a.mask[(size_t)g * D + i]) vv = 0.0;
            s[D + 1 + i] = vv;
        }
        if (NL) for (int i = 0; i < D; ++i) s[2 * D + 1 + i] = a.u[(size_t)g * D + i];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], ph[NV], V[NV * D], U[NV * D];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; V[k * D + i] = nd[k][D + 1 + i]; U[k * D + i] = NL ? nd[k][2 * D + 1 + i] : 0.0; }
            ph[k] = nd[k][D];
        }
        elem_apply_u<Cell, SM>(a.m, a.n1d, X, ph, U, V, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double* red) {
        for (int i = 0; i < D; ++i) {
            double vin = a.v[(size_t)g * D + i];
            double yv = (a.mask && a.mask[(size_t)g * D + i]) ? vin : acc[i];
            a.y[(size_t)g * D + i] = yv;
            red[0] += vin * yv;
        }
    }
};

template <class Cell>
struct OpResidualU {       // staged: X phi u
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = D, NRED = 0;
    static constexpr int NF = odd<D + 1 + D>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) { s[i] = a.X[(size_t)g * D + i]; s[D + 1 + i] = a.u[(size_t)g * D + i]; }
        s[D] = a.phi[g];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], ph[NV], U[NV * D];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + 1 + i]; }
            ph[k] = nd[k][D];
        }
        elem_residual_u<Cell>(a.m, a.n1d, X, ph, U, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) {
        for (int i = 0; i < D; ++i) a.y[(size_t)g * D + i] = acc[i];
    }
};

template <class Cell>
struct OpDiagU {           // writes 1/diag (1 at Dirichlet dofs)
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = D, NRED = 0;
    static constexpr int NF = odd<D + 1 + D>::value;
    __device__ static void load(const KArgs& a, int g, double* s) { OpResidualU<Cell>::load(a, g, s); }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], ph[NV], U[NV * D];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + 1 + i]; }
            ph[k] = nd[k][D];
        }
        elem_diag_u<Cell>(a.m, a.n1d, X, ph, U, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) {
        for (int i = 0; i < D; ++i) {
            bool bc = a.mask && a.mask[(size_t)g * D + i];
            a.y[(size_t)g * D + i] = bc ? 1.0 : 1.0 / acc[i];
        }
    }
};

// inverse of a symmetric D x D block stored (00, 11[, 22], 01[, 02, 12]); Dirichlet dofs are decoupled first
template <int D>
__device__ __forceinline__ void inv_sym_block(const double* b, const bool* fixed, double* out) {
    if constexpr (D == 2) {
        double a00 = fixed[0] ? 1.0 : b[0], a11 = fixed[1] ? 1.0 : b[1], a01 = (fixed[0] || fixed[1]) ? 0.0 : b[2];
        const double id = 1.0 / (a00 * a11 - a01 * a01);
        out[0] = a11 * id; out[1] = a00 * id; out[2] = -a01 * id;
    } else {
        double a00 = fixed[0] ? 1.0 : b[0], a11 = fixed[1] ? 1.0 : b[1], a22 = fixed[2] ? 1.0 : b[2];
        double a01 = (fixed[0] || fixed[1]) ? 0.0 : b[3], a02 = (fixed[0] || fixed[2]) ? 0.0 : b[4],
               a12 = (fixed[1] || fixed[2]) ? 0.0 : b[5];
        const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
        const double id = 1.0 / (a00 * c00 + a01 * c01 + a02 * c02);
        out[0] = c00 * id; out[1] = (a00 * a22 - a02 * a02) * id; out[2] = (a00 * a11 - a01 * a01) * id;
        out[3] = c01 * id; out[4] = c02 * id; out[5] = (a01 * a02 - a00 * a12) * id;
    }
}

template <class Cell>
struct OpDiagBlkU {        // writes the inverse nodal diagonal blocks (block-Jacobi smoother of the two-level PCG)
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = D * (D + 1) / 2, NRED = 0;
    static constexpr int NF = odd<D + 1 + D>::value;
    static constexpr int MINB = (NV == 3) ? 3 : (D == 2 ? 2 : 1);
    __device__ static void load(const KArgs& a, int g, double* s) { OpResidualU<Cell>::load(a, g, s); }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], ph[NV], U[NV * D];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + 1 + i]; }
            ph[k] = nd[k][D];
        }
        elem_diagblk_u<Cell>(a.m, a.n1d, X, ph, U, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) {
        bool fixed[D];
        double inv[NOUT];
        for (int i = 0; i < D; ++i) fixed[i] = a.mask && a.mask[(size_t)g * D + i];
        inv_sym_block<D>(acc, fixed, inv);
        for (int i = 0; i < NOUT; ++i) a.y[(size_t)g * NOUT + i] = inv[i];
    }
};

// φ-space coefficient fields: cm = 2H + F·Gc/l, cg = F·Gc·l (F ≡ 1 without fatigue)
__device__ __forceinline__ void phi_coeffs(const KArgs& a, int g, double& cm, double& cg) {
    double F = a.m.fatigue ? a.Fat[g] : 1.0;
    cm = 2.0 * a.H[g] + F * a.m.Gc / a.m.l;
    cg = F * a.m.Gc * a.m.l;
}

template <class Cell>
struct OpApplyPhi {        // staged: X cm cg p
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 1;
    static constexpr int NF = odd<D + 3>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        phi_coeffs(a, g, s[D], s[D + 1]);
        double vv = a.v[g];
        if (a.mask && a.mask[g]) vv = 0.0;
        s[D + 2] = vv;
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], cm[NV], cg[NV], p[NV];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i];
            cm[k] = nd[k][D]; cg[k] = nd[k][D + 1]; p[k] = nd[k][D + 2];
        }
        elem_apply_phi<Cell>(a.n1d, X, cm, cg, p, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double* red) {
        double vin = a.v[g];
        double yv = (a.mask && a.mask[g]) ? vin : acc[0];
        a.y[g] = yv;
        red[0] += vin * yv;
    }
};

template <class Cell>
struct OpResidualPhi {     // staged: X cm cg H phi
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<D + 4>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        phi_coeffs(a, g, s[D], s[D + 1]);
        s[D + 2] = a.H[g];
        s[D + 3] = a.phi[g];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], cm[NV], cg[NV], H[NV], p[NV];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i];
            cm[k] = nd[k][D]; cg[k] = nd[k][D + 1]; H[k] = nd[k][D + 2]; p[k] = nd[k][D + 3];
        }
        elem_residual_phi<Cell>(a.n1d, X, cm, cg, H, p, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) { a.y[g] = acc[0]; }
};

template <class Cell>
struct OpDiagPhi {
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<D + 2>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        phi_coeffs(a, g, s[D], s[D + 1]);
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], cm[NV], cg[NV];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i];
            cm[k] = nd[k][D]; cg[k] = nd[k][D + 1];
        }
        elem_diag_phi<Cell>(a.n1d, X, cm, cg, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) {
        a.y[g] = (a.mask && a.mask[g]) ? 1.0 : 1.0 / acc[0];
    }
};

// Phase-field forms for a general degradation function (elem_phi_general; MODE 0 residual, 1 apply, 2 diagonal).
// Staged: X phi H F v — used instead of the three Ops above when the degradation function is not quadratic.
template <class Cell, int MODE>
struct OpPhiG {
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = (MODE == 1) ? 1 : 0;
    static constexpr int NF = odd<D + 4>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        s[D] = a.phi[g];
        s[D + 1] = a.H[g];
        s[D + 2] = a.m.fatigue ? a.Fat[g] : 1.0;
        double vv = 0.0;
        if (MODE == 1) { vv = a.v[g]; if (a.mask && a.mask[g]) vv = 0.0; }
        s[D + 3] = vv;
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], ph[NV], H[NV], F[NV], p[NV];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i];
            ph[k] = nd[k][D]; H[k] = nd[k][D + 1]; F[k] = nd[k][D + 2]; p[k] = nd[k][D + 3];
        }
        elem_phi_general<Cell>(MODE, a.m, a.n1d, X, ph, H, F, p, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double* red) {
        if (MODE == 0) {
            a.y[g] = acc[0];
        } else if (MODE == 1) {
            double vin = a.v[g];
            double yv = (a.mask && a.mask[g]) ? vin : acc[0];
            a.y[g] = yv;
            red[0] += vin * yv;
        } else {
            a.y[g] = (a.mask && a.mask[g]) ? 1.0 : 1.0 / acc[0];
        }
    }
};

template <class Cell>
struct OpApplyMass {       // staged: X p
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 1;
    static constexpr int NF = odd<D + 1>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        s[D] = a.v[g];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], p[NV];
        for (int k = 0; k < NV; ++k) { for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i]; p[k] = nd[k][D]; }
        elem_apply_mass<Cell>(a.n1d, X, p, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double* red) {
        a.y[g] = acc[0];
        red[0] += a.v[g] * acc[0];
    }
};

template <class Cell>
struct OpDiagMass {
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<D>::value;
    __device__ static void load(const KArgs& a, int g, double* s) { for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i]; }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D];
        for (int k = 0; k < NV; ++k) for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i];
        elem_diag_mass<Cell>(a.n1d, X, out);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) { a.y[g] = 1.0 / acc[0]; }
};

template <class Cell>
struct OpRhsPsi {          // staged: X u ; y = ∫ψ_a(u) N
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<2 * D>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) { s[i] = a.X[(size_t)g * D + i]; s[D + i] = a.u[(size_t)g * D + i]; }
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], U[NV * D];
        for (int k = 0; k < NV; ++k) for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + i]; }
        elem_rhs_psi<Cell>(a.m, a.n1d, X, U, out, a.aux);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) { a.y[g] = acc[0]; }
};

template <class Cell>
struct OpRhsFatigue {      // staged: X f1(=phi_old) f2(=V_c)
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<D + 2>::value;
    __device__ static void load(const KArgs& a, int g, double* s) {
        for (int i = 0; i < D; ++i) s[i] = a.X[(size_t)g * D + i];
        s[D] = a.f1[g]; s[D + 1] = a.f2[g];
    }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t) {
        double X[NV * D], p[NV], v[NV];
        for (int k = 0; k < NV; ++k) { for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i]; p[k] = nd[k][D]; v[k] = nd[k][D + 1]; }
        elem_rhs_fatigue<Cell>(a.n1d, X, p, v, out, a.m.dfun);
    }
    __device__ static void store(const KArgs& a, int g, const double* acc, double*) { a.y[g] = acc[0]; }
};

template <class Cell>
struct OpTangentU {        // staged: X phi u ; writes the cached tangent of the block's OWN cells, no nodal output
    static constexpr int D = Cell::D, NV = Cell::NV, NOUT = 1, NRED = 0;
    static constexpr int NF = odd<D + 1 + D>::value;
    static constexpr int MINB = 1;
    static constexpr bool OWN = true;       // scatter_kernel: only the own cells, `gid` = compact own-cell index
    __device__ static void load(const KArgs& a, int g, double* s) { OpResidualU<Cell>::load(a, g, s); }
    __device__ static void elem(const KArgs& a, const double* const* nd, double* out, int64_t gid) {
        double X[NV * D], ph[NV], U[NV * D], T[21];
        for (int k = 0; k < NV; ++k) {
            for (int i = 0; i < D; ++i) { X[k * D + i] = nd[k][i]; U[k * D + i] = nd[k][D + 1 + i]; }
            ph[k] = nd[k][D];
        }
        elem_tangent_u<Cell>(a.m, X, ph, U, T);
        constexpr int NT = SymN<D>::N * (SymN<D>::N + 1) / 2;
        for (int t = 0; t < NT; ++t) a.tan[(size_t)gid * NT + t] = T[t];
        for (int k = 0; k < NV; ++k) out[k] = 0.0;
    }
    __device__ static void store(const KArgs&, int, const double*, double*) {}
};

template <class Op>
__global__ void __launch_bounds__(kThreads, min_blocks<Op>::value) scatter_kernel(BlockView bv, KArgs a) {
    constexpr int NV = Op::NV, NF = Op::NF, NOUT = Op::NOUT;
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    if (a.stop && a.stop[0] != 0.0) return;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n0 = bv.node0[b], n_own = bv.node0[b + 1] - n0;
    const int h0 = bv.halo_off[b], n_loc = n_own + bv.halo_off[b + 1] - h0;
    const int e0 = bv.elem_off[b], ne = own_only<Op>::value ? bv.nlow[b] : bv.elem_off[b + 1] - e0;
    const int64_t gid0 = own_only<Op>::value ? (int64_t)bv.own_c0[b] : (int64_t)e0;
    const int chunk = bv.chunk;
    double* s_nodes = smem;
    double* s_ebuf = smem + (size_t)bv.max_loc * NF;
    uint16_t* s_inc = reinterpret_cast<uint16_t*>(smem + ((((size_t)bv.max_loc * NF + (size_t)NV * NOUT * chunk) + 1) & ~(size_t)1));   // 16-byte aligned
    constexpr int LW = NV > 4 ? 2 : 1;          // 64-bit connectivity words per cell (hexahedra: 8 x u16)
    constexpr int VB = NV > 4 ? 3 : 2;          // bits of the vertex number in an incidence code
    // prefetch this thread's first two element connectivities (independent of phase A)
    ushort4 lc0 = make_ushort4(0, 0, 0, 0), lc1 = lc0;
    if (LW == 1) {
        if (tid < ne) lc0 = bv.lconn[e0 + tid];
        if (tid + kThreads < ne) lc1 = bv.lconn[e0 + tid + kThreads];
    }
    // ---- phase A: stage nodal fields and the block's incidence list
    const int i0 = bv.inc_ptr[n0], ninc = bv.inc_ptr[n0 + n_own] - i0;
    for (int i = tid; i < n_loc; i += kThreads) {
        int g = (i < n_own) ? n0 + i : bv.halo[h0 + i - n_own];
        Op::load(a, g, s_nodes + (size_t)i * NF);
    }
    for (int i = tid; i < (ninc >> 3); i += kThreads)        // lists are padded to 8 entries per block
        reinterpret_cast<uint4*>(s_inc)[i] = reinterpret_cast<const uint4*>(bv.inc + i0)[i];
    int cur = 0, end = 0;
    if (tid < n_own) { cur = bv.inc_ptr[n0 + tid] - i0; end = bv.inc_ptr[n0 + tid + 1] - i0; }
    __syncthreads();
    double acc[NOUT];
#pragma unroll
    for (int c = 0; c < NOUT; ++c) acc[c] = 0.0;
    for (int c0 = 0; c0 < ne; c0 += chunk) {
        const int nc = min(chunk, ne - c0);
        // ---- phase B: element vectors
        for (int j = tid; j < nc; j += kThreads) {
            const double* nd[NV > 4 ? 8 : 4];
            if (LW == 1) {
                ushort4 lc = (c0 + j == tid) ? lc0 : ((c0 + j == tid + kThreads) ? lc1 : bv.lconn[e0 + c0 + j]);
                nd[0] = s_nodes + (size_t)lc.x * NF; nd[1] = s_nodes + (size_t)lc.y * NF;
                nd[2] = s_nodes + (size_t)lc.z * NF; nd[3] = s_nodes + (size_t)lc.w * NF;
            } else {
                const ushort4 la = bv.lconn[2 * (size_t)(e0 + c0 + j)], lb = bv.lconn[2 * (size_t)(e0 + c0 + j) + 1];
                nd[0] = s_nodes + (size_t)la.x * NF; nd[1] = s_nodes + (size_t)la.y * NF;
                nd[2] = s_nodes + (size_t)la.z * NF; nd[3] = s_nodes + (size_t)la.w * NF;
                nd[4 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.x * NF; nd[5 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.y * NF;
                nd[6 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.z * NF; nd[7 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.w * NF;
            }
            double out[NV * NOUT];
            Op::elem(a, nd, out, gid0 + c0 + j);
#pragma unroll
            for (int k = 0; k < NV * NOUT; ++k) s_ebuf[k * chunk + j] = out[k];
        }
        __syncthreads();
        // ---- phase C: node-wise gather of incident element rows (fixed order)
        if (tid < n_own) {
            while (cur < end) {
                int code = s_inc[cur];
                int j = (code >> VB) - c0;
                if (j >= nc) break;
                int v = code & ((1 << VB) - 1);
#pragma unroll
                for (int c = 0; c < NOUT; ++c) acc[c] += s_ebuf[(v * NOUT + c) * chunk + j];
                ++cur;
            }
        }
        if (c0 + chunk < ne) __syncthreads();
    }
    double red[Op::NRED > 0 ? Op::NRED : 1];
#pragma unroll
    for (int k = 0; k < (Op::NRED > 0 ? Op::NRED : 1); ++k) red[k] = 0.0;
    if (tid < n_own) Op::store(a, n0 + tid, acc, red);
    if (Op::NRED > 0) grid_reduce<(Op::NRED > 0 ? Op::NRED : 1)>(red, a.partial, a.result, a.counter, s_red, a.pv);
}

// ----------------------------------------------------------------------------- async staging helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a transaction count that never completes (a bug) must not hang the GPU — after ~2 s the wait
// gives up (the result is then garbage and the parity checks fail loudly)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((spins & 0x3FFu) == 0x3FFu) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) return;
        }
    }
}
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ============================================================================= hot kernel
// J_u·v on P1 triangles (config 1-3 meshes): same three phases as scatter_kernel, specialised:
// SoA double2 staging (128-bit shared loads/stores), element rows parked as double2 per vertex,
// two register-prefetched elements per thread (blocks are sized to <= 2·kThreads elements),
// reciprocal / rsqrt intrinsics, model fixed at compile time.
template <int SM>
__device__ __forceinline__ void tri_apply_elem(const Material& m, double2 x0, double2 x1, double2 x2, double p0, double p1,
                                               double p2, double2 v0, double2 v1, double2 v2, double2 u0, double2 u1,
                                               double2 u2, double2& f0, double2& f1, double2& f2) {
    const double ax = x1.x - x0.x, ay = x1.y - x0.y, bx = x2.x - x0.x, by = x2.y - x0.y;
    const double det = ax * by - ay * bx, inv = __drcp_rn(det);
    // gradients G1 = (by, -bx)/det, G2 = (-ay, ax)/det, G0 = -(G1+G2)
    const double g1x = by * inv, g1y = -bx * inv, g2x = -ay * inv, g2y = ax * inv;
    const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
    const double t0 = 1.0 - p0, t1 = 1.0 - p1, t2 = 1.0 - p2, ts = t0 + t1 + t2;
    const double gb = (t0 * t0 + t1 * t1 + t2 * t2 + ts * ts) * (1.0 / 12.0) + m.k;
    double de[3] = {g0x * v0.x + g1x * v1.x + g2x * v2.x, g0y * v0.y + g1y * v1.y + g2y * v2.y,
                    0.5 * (g0y * v0.x + g1y * v1.x + g2y * v2.x + g0x * v0.y + g1x * v1.y + g2x * v2.y)};
    double S[3];
    if (SM == M_ISO) {
        const double lt = m.lam * (de[0] + de[1]), mu2 = 2.0 * m.mu;
        S[0] = gb * (lt + mu2 * de[0]); S[1] = gb * (lt + mu2 * de[1]); S[2] = gb * (mu2 * de[2]);
    } else {
        double e[3] = {g0x * u0.x + g1x * u1.x + g2x * u2.x, g0y * u0.y + g1y * u1.y + g2y * u2.y,
                       0.5 * (g0y * u0.x + g1y * u1.x + g2y * u2.x + g0x * u0.y + g1x * u1.y + g2x * u2.y)};
        double dsa[3], dsb[3];
        if (SM == M_SPECTRAL) {
            const double a = e[0], c = e[1], b = e[2], dd = de[0] - de[1], dtr = de[0] + de[1];
            const double amc = a - c, tr = a + c;
            const double Dl = amc * amc + 4.0 * b * b + kEps * kEps;
            const double is = rsqrt(Dl), s = Dl * is;
            const double l0 = 0.5 * (tr - s), l1 = 0.5 * (tr + s);
            const double m0 = amc * is, m1 = 2.0 * b * is;
            const double ds = m0 * dd + 2.0 * m1 * de[2];
            // ½·H(λ) with H = ½(1 + sign λ); copysign form (differs from sign() only at λ == 0 exactly,
            // which the eps-regularised eigenvalues never hit; tangent only, the residual is unaffected)
            const double w0 = (0.25 + copysign(0.25, l0)) * (dtr - ds);
            const double w1 = (0.25 + copysign(0.25, l1)) * (dtr + ds);
            const double pd = 0.5 * (fmax(l1, 0.0) - fmax(l0, 0.0)) * is;
            const double dm0 = dd - m0 * ds, dm1 = 2.0 * de[2] - m1 * ds;
            const double e1x = 0.5 * (1.0 + m0), e1y = 0.5 * (1.0 - m0), e1s = 0.5 * m1;
            const double lt = m.lam * (tr > 0.0 ? 1.0 : (tr < 0.0 ? 0.0 : 0.5)) * dtr, mu2 = 2.0 * m.mu;   // tr == 0 at u = 0
            dsa[0] = lt + mu2 * (w0 * (1.0 - e1x) + w1 * e1x + pd * dm0);
            dsa[1] = lt + mu2 * (w0 * (1.0 - e1y) + w1 * e1y - pd * dm0);
            dsa[2] = mu2 * ((w1 - w0) * e1s + pd * dm1);
            const double lf = m.lam * dtr;
            dsb[0] = lf + mu2 * de[0] - dsa[0]; dsb[1] = lf + mu2 * de[1] - dsa[1]; dsb[2] = mu2 * de[2] - dsa[2];
        } else {
            dstress_split<2>(m, SM, e, de, dsa, dsb);
        }
        S[0] = gb * dsa[0] + dsb[0]; S[1] = gb * dsa[1] + dsb[1]; S[2] = gb * dsa[2] + dsb[2];
    }
    const double vol = 0.5 * fabs(det);
    S[0] *= vol; S[1] *= vol; S[2] *= vol;
    f1 = make_double2(S[0] * g1x + S[2] * g1y, S[2] * g1x + S[1] * g1y);
    f2 = make_double2(S[0] * g2x + S[2] * g2y, S[2] * g2x + S[1] * g2y);
    f0 = make_double2(-(f1.x + f2.x), -(f1.y + f2.y));
}

template <int SM>
__global__ void __launch_bounds__(kThreads, 4) apply_u_tri_kernel(BlockView bv, KArgs a) {
    constexpr bool NL = (SM != M_ISO);
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    if (a.stop && a.stop[0] != 0.0) return;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n0 = bv.node0[b], n_own = bv.node0[b + 1] - n0;
    const int h0 = bv.halo_off[b], n_loc = n_own + bv.halo_off[b + 1] - h0;
    const int e0 = bv.elem_off[b], ne = bv.elem_off[b + 1] - e0;
    const int ML = bv.max_loc, chunk = bv.chunk;
    double2* sX = reinterpret_cast<double2*>(smem);
    double2* sV = sX + ML;
    double2* sU = sV + ML;                          // only used when NL
    double* sP = reinterpret_cast<double*>(sU + (NL ? ML : 0));
    double2* sE = reinterpret_cast<double2*>(sP + ((ML + 1) & ~1));     // [3][chunk]
    uint16_t* s_inc = reinterpret_cast<uint16_t*>(sE + 3 * (size_t)chunk);
    ushort4 lc0 = make_ushort4(0, 0, 0, 0), lc1 = lc0;
    if (tid < ne) lc0 = bv.lconn[e0 + tid];
    if (tid + kThreads < ne) lc1 = bv.lconn[e0 + tid + kThreads];
    const int i0 = bv.inc_ptr[n0], ninc = bv.inc_ptr[n0 + n_own] - i0;
    int cur = 0, end = 0;
    if (tid < n_own) { cur = bv.inc_ptr[n0 + tid] - i0; end = bv.inc_ptr[n0 + tid + 1] - i0; }
    const double2* X2 = reinterpret_cast<const double2*>(a.X);
    const double2* V2 = reinterpret_cast<const double2*>(a.v);
    const double2* U2 = reinterpret_cast<const double2*>(a.u);
    const uchar2* M2 = reinterpret_cast<const uchar2*>(a.mask);
    // ---- phase A
    for (int i = tid; i < n_loc; i += kThreads) {
        int g = (i < n_own) ? n0 + i : bv.halo[h0 + i - n_own];
        double2 vv = V2[g];
        if (M2) { uchar2 mk = M2[g]; if (mk.x) vv.x = 0.0; if (mk.y) vv.y = 0.0; }
        sX[i] = X2[g]; sV[i] = vv; sP[i] = a.phi[g];
        if (NL) sU[i] = U2[g];
    }
    for (int i = tid; i < (ninc >> 3); i += kThreads)
        reinterpret_cast<uint4*>(s_inc)[i] = reinterpret_cast<const uint4*>(bv.inc + i0)[i];
    __syncthreads();
    double2 acc = make_double2(0.0, 0.0);
    for (int c0 = 0; c0 < ne; c0 += chunk) {
        const int nc = min(chunk, ne - c0);
        // ---- phase B
        for (int j = tid; j < nc; j += kThreads) {
            const int eid = c0 + j;
            ushort4 lc = (eid == tid) ? lc0 : ((eid == tid + kThreads) ? lc1 : bv.lconn[e0 + eid]);
            double2 f0, f1, f2;
            const double2 z = make_double2(0.0, 0.0);
            tri_apply_elem<SM>(a.m, sX[lc.x], sX[lc.y], sX[lc.z], sP[lc.x], sP[lc.y], sP[lc.z], sV[lc.x], sV[lc.y], sV[lc.z],
                               NL ? sU[lc.x] : z, NL ? sU[lc.y] : z, NL ? sU[lc.z] : z, f0, f1, f2);
            sE[j] = f0; sE[chunk + j] = f1; sE[2 * chunk + j] = f2;
        }
        __syncthreads();
        // ---- phase C
        if (tid < n_own) {
            while (cur < end) {
                int code = s_inc[cur];
                int j = (code >> 2) - c0;
                if (j >= nc) break;
                double2 f = sE[(code & 3) * chunk + j];
                acc.x += f.x; acc.y += f.y;
                ++cur;
            }
        }
        if (c0 + chunk < ne) __syncthreads();
    }
    double red[1] = {0.0};
    if (tid < n_own) {
        const int g = n0 + tid;
        double2 vin = V2[g];
        if (M2) { uchar2 mk = M2[g]; if (mk.x) acc.x = vin.x; if (mk.y) acc.y = vin.y; }
        reinterpret_cast<double2*>(a.y)[g] = acc;
        red[0] = vin.x * acc.x + vin.y * acc.y;
    }
    grid_reduce<1>(red, a.partial, a.result, a.counter, s_red, a.pv);
}

// ============================================================================= hot kernel, v3
// Same arithmetic as apply_u_tri_kernel, restructured for Blackwell: a persistent grid (a few CTAs per SM,
// each walking blocks b, b+grid, ...) with DOUBLE-BUFFERED asynchronous staging — the contiguous owned
// ranges of X, v, u, phi and the fixed-width incidence table arrive by 1-D bulk copies (TMA engine,
// mbarrier complete_tx), the halo nodes by per-thread cp.async — so the loads of block i+1 are in flight
// while block i is computed.  Packed block metadata and every per-thread index of the NEXT block are
// prefetched in registers one iteration ahead (no dependent global loads on the critical path), two
// unrolled cells per thread, and phase C reads a fixed-width (8) incidence table — eight independent
// 128-bit shared loads per node instead of a pointer-chasing loop.  Contract: when a.mask != null the
// INPUT is already zero on Dirichlet dofs (true for every PCG direction) and the output rows of those
// dofs are written as zero; the public pfx_apply_u wrapper restores y = v there.  Requires every block
// to list <= 2*kThreads cells (the block builder sizes triangle blocks that way; otherwise the engine
// falls back to apply_u_tri_kernel).
template <int SM>
__global__ void __launch_bounds__(kThreads, 3) apply_u_tri_v3_kernel(BlockView bv, KArgs a) {
    constexpr bool NL = (SM != M_ISO);
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    __shared__ uint64_t s_bar[2];
    if (a.stop && a.stop[0] != 0.0) return;
    const int tid = threadIdx.x;
    const int ML = bv.max_loc, chunk = bv.chunk, MO = (bv.max_own + 1) & ~1;
    const size_t oV = (size_t)ML * 16, oU = 2 * oV, oP = oV * (NL ? 3 : 2), oI = oP + (size_t)((ML + 3) & ~1) * 8;
    const size_t stage_bytes = oI + (size_t)MO * 16;
    char* const sbase = reinterpret_cast<char*>(smem);
    double2* sE = reinterpret_cast<double2*>(sbase + 2 * stage_bytes);      // [3*chunk + 1], last = zero slot
    const int zslot = 3 * chunk;
    const double2* X2 = reinterpret_cast<const double2*>(a.X);
    const double2* V2 = reinterpret_cast<const double2*>(a.v);
    const double2* U2 = reinterpret_cast<const double2*>(a.u);
    const unsigned short* M2 = reinterpret_cast<const unsigned short*>(a.mask);   // 2 flags per node, kept raw
    const uint2* LC = reinterpret_cast<const uint2*>(bv.lconn);                  // 4 x u16 per cell, kept raw
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sE[zslot] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (a.pv && a.wait_halo) halo_wait_cta(a.pv);          // neighbours' pushes of this direction have landed

    auto issue = [&](int4 ma, int4 mb, int s_, int hidx) {
        char* p = sbase + (size_t)s_ * stage_bytes;
        double2 *tX = reinterpret_cast<double2*>(p), *tV = reinterpret_cast<double2*>(p + oV), *tU = reinterpret_cast<double2*>(p + oU);
        double* tP = reinterpret_cast<double*>(p + oP);
        const int n0 = ma.x, n_own = ma.y, h0 = ma.z, n_halo = ma.w;
        if (tid == 0) {
            fence_proxy_async();
            const uint32_t nev = (uint32_t)(n_own & ~1);
            uint32_t bytes = (uint32_t)n_own * 16u * (NL ? 3u : 2u) + nev * 8u + (uint32_t)n_own * 16u;
            mbar_expect_tx(&s_bar[s_], bytes);
            bulk_g2s(tX, X2 + n0, (uint32_t)n_own * 16u, &s_bar[s_]);
            bulk_g2s(tV, V2 + n0, (uint32_t)n_own * 16u, &s_bar[s_]);
            if (NL) bulk_g2s(tU, U2 + n0, (uint32_t)n_own * 16u, &s_bar[s_]);
            if (nev) bulk_g2s(tP, a.phi + n0, nev * 8u, &s_bar[s_]);
            bulk_g2s(p + oI, bv.inc8 + n0, (uint32_t)n_own * 16u, &s_bar[s_]);
            if (n_own & 1) cp_async8(tP + n_own - 1, a.phi + n0 + n_own - 1);
        }
        for (int i = tid; i < n_halo; i += kThreads) {
            int g = (i == tid) ? hidx : bv.halo[h0 + i];
            cp_async16(tX + n_own + i, X2 + g);
            cp_async16(tV + n_own + i, V2 + g);
            if (NL) cp_async16(tU + n_own + i, U2 + g);
            cp_async8(tP + n_own + i, a.phi + g);
        }
        cp_async_commit();
        (void)mb;
    };

    const int stride = gridDim.x, nb = bv.nb;
    int blk = blockIdx.x;                       // grid <= nb: every CTA owns at least one block
    int4 ma = bv.meta[2 * blk], mb = bv.meta[2 * blk + 1];
    int4 na = ma, nbm = mb;
    bool has_next = blk + stride < nb;
    if (has_next) { na = bv.meta[2 * (blk + stride)]; nbm = bv.meta[2 * (blk + stride) + 1]; }
    issue(ma, mb, 0, tid < ma.w ? bv.halo[ma.z + tid] : 0);
    uint2 lc0 = make_uint2(0u, 0u), lc1 = lc0;
    if (tid < mb.y) lc0 = LC[mb.x + tid];
    if (tid + kThreads < mb.y) lc1 = LC[mb.x + tid + kThreads];
    unsigned int mk = 0u;
    if (M2 && tid < ma.y) mk = M2[ma.x + tid];
    int hnext = (has_next && tid < na.w) ? bv.halo[na.z + tid] : 0;
    double red = 0.0;
    uint32_t phases = 0u;
    for (int it = 0;; ++it) {
        const int s = it & 1;
        const int n0 = ma.x, n_own = ma.y, ne = mb.y;
        // prefetch next block into the other stage; start the metadata of the one after
        int4 fa = na, fb = nbm;
        const bool has_nn = blk + 2 * stride < nb;
        if (has_next) {
            issue(na, nbm, s ^ 1, hnext);
            if (has_nn) { fa = bv.meta[2 * (blk + 2 * stride)]; fb = bv.meta[2 * (blk + 2 * stride) + 1]; }
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        mbar_wait(&s_bar[s], (phases >> s) & 1u);
        phases ^= (1u << s);
        __syncthreads();
        const char* p = sbase + (size_t)s * stage_bytes;
        const double2 *sX = reinterpret_cast<const double2*>(p), *sV = reinterpret_cast<const double2*>(p + oV),
                      *sU = reinterpret_cast<const double2*>(p + oU);
        const double* sP = reinterpret_cast<const double*>(p + oP);
        const uint4* s_inc8 = reinterpret_cast<const uint4*>(p + oI);
        // ---- phase B: two cells per thread
        const double2 z = make_double2(0.0, 0.0);
        if (tid < ne) {
            double2 f0, f1, f2;
            const int vx = lc0.x & 0xFFFF, vy = lc0.x >> 16, vz = lc0.y & 0xFFFF;
            tri_apply_elem<SM>(a.m, sX[vx], sX[vy], sX[vz], sP[vx], sP[vy], sP[vz], sV[vx], sV[vy],
                                   sV[vz], NL ? sU[vx] : z, NL ? sU[vy] : z, NL ? sU[vz] : z, f0, f1, f2);
            sE[tid] = f0; sE[chunk + tid] = f1; sE[2 * chunk + tid] = f2;
        }
        if (tid + kThreads < ne) {
            double2 f0, f1, f2;
            const int vx = lc1.x & 0xFFFF, vy = lc1.x >> 16, vz = lc1.y & 0xFFFF;
            tri_apply_elem<SM>(a.m, sX[vx], sX[vy], sX[vz], sP[vx], sP[vy], sP[vz], sV[vx], sV[vy],
                                   sV[vz], NL ? sU[vx] : z, NL ? sU[vy] : z, NL ? sU[vz] : z, f0, f1, f2);
            sE[tid + kThreads] = f0; sE[chunk + tid + kThreads] = f1; sE[2 * chunk + tid + kThreads] = f2;
        }
        __syncthreads();
        // ---- per-thread indices of the NEXT block (consumed one iteration later)
        uint2 lc0n = make_uint2(0u, 0u), lc1n = lc0n;
        unsigned int mkn = 0u;
        int hnn = 0;
        if (has_next) {
            if (tid < nbm.y) lc0n = LC[nbm.x + tid];
            if (tid + kThreads < nbm.y) lc1n = LC[nbm.x + tid + kThreads];
            if (M2 && tid < na.y) mkn = M2[na.x + tid];
            if (has_nn && tid < fa.w) hnn = bv.halo[fa.z + tid];
        }
        // ---- phase C: eight independent incident rows per owned node, fixed order
        if (tid < n_own) {
            const uint4 c8 = s_inc8[tid];
            const unsigned cw[4] = {c8.x, c8.y, c8.z, c8.w};
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int code = (int)((cw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
                const int idx = min((code & 3) * chunk + (code >> 2), zslot);
                const double2 f = sE[idx];
                acc.x += f.x; acc.y += f.y;
            }
            if ((c8.w >> 16) == 0xFFFEu) {                    // rare: valence > 8, rest from the CSR list
                int k0 = bv.inc_ptr[n0 + tid] + 7, k1 = bv.inc_ptr[n0 + tid + 1];
                for (int k = k0; k < k1; ++k) {
                    const int code = bv.inc[k];
                    if (code >= 0xFFFE) break;
                    const double2 f = sE[(code & 3) * chunk + (code >> 2)];
                    acc.x += f.x; acc.y += f.y;
                }
            }
            if (mk & 0xFFu) acc.x = 0.0;
            if (mk & 0xFF00u) acc.y = 0.0;
            const double2 vin = sV[tid];
            reinterpret_cast<double2*>(a.y)[n0 + tid] = acc;
            red += vin.x * acc.x + vin.y * acc.y;
        }
        if (!has_next) break;
        __syncthreads();                         // stage s and sE are free again
        blk += stride;
        ma = na; mb = nbm; na = fa; nbm = fb;
        lc0 = lc0n; lc1 = lc1n; mk = mkn; hnext = hnn;
        has_next = has_nn;
    }
    double r1[1] = {red};
    grid_reduce<1>(r1, a.partial, a.result, a.counter, s_red, a.pv);
}

// ============================================================================= phi-apply, v3 pipeline
// J_phi·p on P1 triangles with the same persistent / TMA double-buffered / prefetched structure as
// apply_u_tri_v3_kernel; staged fields: X (double2), cm, cg, p (doubles; cm = 2H + F·Gc/l and
// cg = F·Gc·l are refreshed once per phi-solve by phi_coeff_kernel).  Same input/output contract.
__global__ void phi_coeff_kernel(int64_t n, Material m, const double* H, const double* Fat, double* cm, double* cg) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double F = m.fatigue ? Fat[i] : 1.0;
        cm[i] = 2.0 * H[i] + F * m.Gc / m.l;
        cg[i] = F * m.Gc * m.l;
    }
}

__device__ __forceinline__ void tri_apply_phi_elem(double2 x0, double2 x1, double2 x2, double m0, double m1, double m2,
                                                   double c0, double c1, double c2, double p0, double p1, double p2,
                                                   double& f0, double& f1, double& f2) {
    const double ax = x1.x - x0.x, ay = x1.y - x0.y, bx = x2.x - x0.x, by = x2.y - x0.y;
    const double det = ax * by - ay * bx, inv = __drcp_rn(det);
    const double g1x = by * inv, g1y = -bx * inv, g2x = -ay * inv, g2y = ax * inv;
    const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
    const double vol = 0.5 * fabs(det);
    const double C = m0 + m1 + m2, P = p0 + p1 + p2, Q = m0 * p0 + m1 * p1 + m2 * p2;
    const double cgm = vol * (c0 + c1 + c2) * (1.0 / 3.0);
    const double gpx = g0x * p0 + g1x * p1 + g2x * p2, gpy = g0y * p0 + g1y * p1 + g2y * p2;
    const double vb = vol * (1.0 / 60.0), cpq = C * P + Q;
    f0 = vb * (cpq + C * p0 + m0 * P + 2.0 * m0 * p0) + cgm * (g0x * gpx + g0y * gpy);
    f1 = vb * (cpq + C * p1 + m1 * P + 2.0 * m1 * p1) + cgm * (g1x * gpx + g1y * gpy);
    f2 = vb * (cpq + C * p2 + m2 * P + 2.0 * m2 * p2) + cgm * (g2x * gpx + g2y * gpy);
}

__global__ void __launch_bounds__(kThreads, 4) apply_phi_tri_v3_kernel(BlockView bv, KArgs a) {
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    __shared__ uint64_t s_bar[2];
    if (a.stop && a.stop[0] != 0.0) return;
    const int tid = threadIdx.x;
    const int ML = bv.max_loc, chunk = bv.chunk, MO = (bv.max_own + 1) & ~1;
    const size_t sc = (size_t)((ML + 3) & ~1) * 8;                       // bytes of one scalar array
    const size_t oM = (size_t)ML * 16, oG = oM + sc, oP = oG + sc, oI = oP + sc;
    const size_t stage_bytes = oI + (size_t)MO * 16;
    char* const sbase = reinterpret_cast<char*>(smem);
    double* sE = reinterpret_cast<double*>(sbase + 2 * stage_bytes);     // [3*chunk + 2], slot 3*chunk = zero
    const int zslot = 3 * chunk;
    const double2* X2 = reinterpret_cast<const double2*>(a.X);
    const uint2* LC = reinterpret_cast<const uint2*>(bv.lconn);
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sE[zslot] = 0.0;
    }
    __syncthreads();
    if (a.pv && a.wait_halo) halo_wait_cta(a.pv);

    auto issue = [&](int4 ma, int s_, int hidx) {
        char* p = sbase + (size_t)s_ * stage_bytes;
        double2* tX = reinterpret_cast<double2*>(p);
        double *tM = reinterpret_cast<double*>(p + oM), *tG = reinterpret_cast<double*>(p + oG), *tP = reinterpret_cast<double*>(p + oP);
        const int n0 = ma.x, n_own = ma.y, h0 = ma.z, n_halo = ma.w;
        if (tid == 0) {
            fence_proxy_async();
            const uint32_t nev = (uint32_t)(n_own & ~1);
            mbar_expect_tx(&s_bar[s_], (uint32_t)n_own * 32u + 3u * nev * 8u);
            bulk_g2s(tX, X2 + n0, (uint32_t)n_own * 16u, &s_bar[s_]);
            if (nev) {
                bulk_g2s(tM, a.cm + n0, nev * 8u, &s_bar[s_]);
                bulk_g2s(tG, a.cg + n0, nev * 8u, &s_bar[s_]);
                bulk_g2s(tP, a.v + n0, nev * 8u, &s_bar[s_]);
            }
            bulk_g2s(p + oI, bv.inc8 + n0, (uint32_t)n_own * 16u, &s_bar[s_]);
            if (n_own & 1) {
                const int t = n_own - 1;
                cp_async8(tM + t, a.cm + n0 + t); cp_async8(tG + t, a.cg + n0 + t); cp_async8(tP + t, a.v + n0 + t);
            }
        }
        for (int i = tid; i < n_halo; i += kThreads) {
            int g = (i == tid) ? hidx : bv.halo[h0 + i];
            cp_async16(tX + n_own + i, X2 + g);
            cp_async8(tM + n_own + i, a.cm + g);
            cp_async8(tG + n_own + i, a.cg + g);
            cp_async8(tP + n_own + i, a.v + g);
        }
        cp_async_commit();
    };

    const int stride = gridDim.x, nb = bv.nb;
    int blk = blockIdx.x;
    int4 ma = bv.meta[2 * blk], mb = bv.meta[2 * blk + 1];
    int4 na = ma, nbm = mb;
    bool has_next = blk + stride < nb;
    if (has_next) { na = bv.meta[2 * (blk + stride)]; nbm = bv.meta[2 * (blk + stride) + 1]; }
    issue(ma, 0, tid < ma.w ? bv.halo[ma.z + tid] : 0);
    uint2 lc0 = make_uint2(0u, 0u), lc1 = lc0;
    if (tid < mb.y) lc0 = LC[mb.x + tid];
    if (tid + kThreads < mb.y) lc1 = LC[mb.x + tid + kThreads];
    unsigned int mk = 0u;
    if (a.mask && tid < ma.y) mk = a.mask[ma.x + tid];
    int hnext = (has_next && tid < na.w) ? bv.halo[na.z + tid] : 0;
    double red = 0.0;
    uint32_t phases = 0u;
    for (int it = 0;; ++it) {
        const int s = it & 1;
        const int n0 = ma.x, n_own = ma.y, ne = mb.y;
        int4 fa = na, fb = nbm;
        const bool has_nn = blk + 2 * stride < nb;
        if (has_next) {
            issue(na, s ^ 1, hnext);
            if (has_nn) { fa = bv.meta[2 * (blk + 2 * stride)]; fb = bv.meta[2 * (blk + 2 * stride) + 1]; }
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        mbar_wait(&s_bar[s], (phases >> s) & 1u);
        phases ^= (1u << s);
        __syncthreads();
        const char* p = sbase + (size_t)s * stage_bytes;
        const double2* sX = reinterpret_cast<const double2*>(p);
        const double *sM = reinterpret_cast<const double*>(p + oM), *sG = reinterpret_cast<const double*>(p + oG),
                     *sP = reinterpret_cast<const double*>(p + oP);
        const uint4* s_inc8 = reinterpret_cast<const uint4*>(p + oI);
        if (tid < ne) {
            const int vx = lc0.x & 0xFFFF, vy = lc0.x >> 16, vz = lc0.y & 0xFFFF;
            double f0, f1, f2;
            tri_apply_phi_elem(sX[vx], sX[vy], sX[vz], sM[vx], sM[vy], sM[vz], sG[vx], sG[vy], sG[vz], sP[vx], sP[vy], sP[vz], f0, f1, f2);
            sE[tid] = f0; sE[chunk + tid] = f1; sE[2 * chunk + tid] = f2;
        }
        if (tid + kThreads < ne) {
            const int vx = lc1.x & 0xFFFF, vy = lc1.x >> 16, vz = lc1.y & 0xFFFF;
            double f0, f1, f2;
            tri_apply_phi_elem(sX[vx], sX[vy], sX[vz], sM[vx], sM[vy], sM[vz], sG[vx], sG[vy], sG[vz], sP[vx], sP[vy], sP[vz], f0, f1, f2);
            sE[tid + kThreads] = f0; sE[chunk + tid + kThreads] = f1; sE[2 * chunk + tid + kThreads] = f2;
        }
        __syncthreads();
        uint2 lc0n = make_uint2(0u, 0u), lc1n = lc0n;
        unsigned int mkn = 0u;
        int hnn = 0;
        if (has_next) {
            if (tid < nbm.y) lc0n = LC[nbm.x + tid];
            if (tid + kThreads < nbm.y) lc1n = LC[nbm.x + tid + kThreads];
            if (a.mask && tid < na.y) mkn = a.mask[na.x + tid];
            if (has_nn && tid < fa.w) hnn = bv.halo[fa.z + tid];
        }
        if (tid < n_own) {
            const uint4 c8 = s_inc8[tid];
            const unsigned cw[4] = {c8.x, c8.y, c8.z, c8.w};
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int code = (int)((cw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
                acc += sE[min((code & 3) * chunk + (code >> 2), zslot)];
            }
            if ((c8.w >> 16) == 0xFFFEu) {
                int k0 = bv.inc_ptr[n0 + tid] + 7, k1 = bv.inc_ptr[n0 + tid + 1];
                for (int k = k0; k < k1; ++k) {
                    const int code = bv.inc[k];
                    if (code >= 0xFFFE) break;
                    acc += sE[(code & 3) * chunk + (code >> 2)];
                }
            }
            if (mk) acc = 0.0;
            a.y[n0 + tid] = acc;
            red += sP[tid] * acc;
        }
        if (!has_next) break;
        __syncthreads();
        blk += stride;
        ma = na; mb = nbm; na = fa; nbm = fb;
        lc0 = lc0n; lc1 = lc1n; mk = mkn; hnext = hnn;
        has_next = has_nn;
    }
    double r1[1] = {red};
    grid_reduce<1>(r1, a.partial, a.result, a.counter, s_red, a.pv);
}

// ============================================================================= small meshes: whole PCG in one kernel
// Configs 1 and 3 (5 089 nodes, 80 blocks) are launch-latency bound: three launches per CG iteration
// cost ~13 µs while the arithmetic is ~1 µs.  Here the ENTIRE Jacobi-PCG solve of J_u δ = b runs in one
// cooperative launch, one CTA per mesh block: coordinates, phi, u, the incidence table and the
// connectivities are staged ONCE and stay in shared memory / registers for the whole solve; per
// iteration a CTA only (i) recomputes p on its halo from the neighbours' r (written to global before
// the reduction barrier), (ii) runs phases B/C, and (iii) takes part in two grid barriers that also
// carry the two dot products (per-block partials, summed by every CTA in the same fixed order, so
// all CTAs take identical decisions and the result is bit-reproducible).
struct CoopArgs {
    const double* F;          // right-hand side (Dirichlet entries ignored)
    const double* dinv;       // Jacobi data
    double* x;                // solution (w_x)
    double* r;                // residual, also the halo exchange buffer (w_r)
    double* p;                // direction, saved on exit / reloaded on resume (w_p)
    double* partial;          // [2][nb][2]
    unsigned long long* bar;  // grid barrier counter (monotonic, zeroed by the host before the launch)
    double* S;                // PCG scalars (S_RR, S_ITERS, S_DONE, S_THRESH, S_RZ0)
    double rtol2;
    int bref_slot;            // >= 0: per-load-step absolute floor (see pcg_setup_kernel)
    int max_iters;            // iterations in this launch
    int resume;               // 0: start from x = 0; 1: continue from the saved x, r, p
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void coop_barrier(unsigned long long* bar, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ull) : "memory");
        long long spins = 0;
        while (ld_acquire_gpu(bar) < target) {
            if (++spins > kSpinLimit) break;
        }
    }
    __syncthreads();
}

template <int SM>
__global__ void __launch_bounds__(kThreads, 1) pcg_coop_u_tri_kernel(BlockView bv, KArgs a, CoopArgs ca) {
    constexpr bool NL = (SM != M_ISO);
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    __shared__ double s_tot[2];
    const int tid = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
    const int4 ma = bv.meta[2 * b], mb = bv.meta[2 * b + 1];
    const int n0 = ma.x, n_own = ma.y, h0 = ma.z, n_halo = ma.w, n_loc = n_own + n_halo, e0 = mb.x, ne = mb.y;
    const int ML = bv.max_loc, chunk = 2 * kThreads;
    double2* sX = reinterpret_cast<double2*>(smem);
    double2* sU = sX + ML;
    double2* sV = sU + ML;                   // p on every local node
    double2* sD = sV + ML;                   // Jacobi data on every local node
    double2* sE = sD + ML;                   // [3*chunk + 1]
    double* sP = reinterpret_cast<double*>(sE + 3 * chunk + 1);
    int* sH = reinterpret_cast<int*>(sP + ((ML + 1) & ~1));
    uint4* sI = reinterpret_cast<uint4*>(sH + ((ML + 3) & ~3));
    const int zslot = 3 * chunk;
    const double2* X2 = reinterpret_cast<const double2*>(a.X);
    const double2* U2 = reinterpret_cast<const double2*>(a.u);
    const double2* D2 = reinterpret_cast<const double2*>(ca.dinv);
    const unsigned short* M2 = reinterpret_cast<const unsigned short*>(a.mask);
    const uint2* LC = reinterpret_cast<const uint2*>(bv.lconn);
    // ---- one-time staging
    for (int i = tid; i < n_loc; i += kThreads) {
        const int g = (i < n_own) ? n0 + i : bv.halo[h0 + i - n_own];
        sX[i] = X2[g]; sP[i] = a.phi[g]; sD[i] = D2[g]; sH[i] = g;
        if (NL) sU[i] = U2[g];
    }
    if (tid < n_own) sI[tid] = bv.inc8[n0 + tid];
    if (tid == 0) sE[zslot] = make_double2(0.0, 0.0);
    uint2 lc0 = make_uint2(0u, 0u), lc1 = lc0;
    if (tid < ne) lc0 = LC[e0 + tid];
    if (tid + kThreads < ne) lc1 = LC[e0 + tid + kThreads];
    unsigned int mk = 0u;
    double2 x = make_double2(0.0, 0.0), r = x, dv = x;
    const bool own = tid < n_own;
    double2* Xg = reinterpret_cast<double2*>(ca.x);
    double2* Rg = reinterpret_cast<double2*>(ca.r);
    double2* Pg = reinterpret_cast<double2*>(ca.p);
    const double2* Fg = reinterpret_cast<const double2*>(ca.F);
    if (own) {
        if (M2) mk = M2[n0 + tid];
        dv = D2[n0 + tid];
        if (ca.resume) { x = Xg[n0 + tid]; r = Rg[n0 + tid]; }
        else {
            r = Fg[n0 + tid];
            if (mk & 0xFFu) r.x = 0.0;
            if (mk & 0xFF00u) r.y = 0.0;
            Rg[n0 + tid] = r;
        }
    }
    unsigned long long gen = 0;
    int red_par = 0;
    // grid-wide sums of two per-thread values; every CTA obtains bit-identical totals
    auto grid_sum2 = [&](double v0, double v1, double& t0, double& t1) {
        double vals[2] = {v0, v1};
        block_sum<2>(vals, s_red);
        double* part = ca.partial + (size_t)red_par * nb * 2;
        if (tid == 0) { part[2 * b] = vals[0]; part[2 * b + 1] = vals[1]; }
        ++gen;
        coop_barrier(ca.bar, gen * (unsigned long long)nb);
        if (tid < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int k = tid; k < nb; k += 32) { s0 += __ldcg(part + 2 * k); s1 += __ldcg(part + 2 * k + 1); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (tid == 0) { s_tot[0] = s0; s_tot[1] = s1; }
        }
        __syncthreads();
        t0 = s_tot[0]; t1 = s_tot[1];
        red_par ^= 1;
        __syncthreads();
    };
    // ---- start-up: z = D^-1 r, p = z (or the saved p), rz, rr
    double rz, rr;
    {
        double2 z = make_double2(dv.x * r.x, dv.y * r.y);
        grid_sum2(own ? r.x * z.x + r.y * z.y : 0.0, own ? r.x * r.x + r.y * r.y : 0.0, rz, rr);
        // (the barrier inside grid_sum2 also published every block's r for the halo reads below)
        for (int i = tid; i < n_loc; i += kThreads) {
            const int g = sH[i];
            if (ca.resume) sV[i] = Pg[g];
            else { const double2 rg = __ldcg(Rg + g); const double2 d = sD[i]; sV[i] = make_double2(d.x * rg.x, d.y * rg.y); }
        }
        __syncthreads();
    }
    double thresh = ca.S[S_THRESH];
    if (!ca.resume) {
        double ref = rr;
        if (ca.bref_slot >= 0) ref = fmax(ca.S[ca.bref_slot], rr);
        thresh = ca.rtol2 * ref;
        if (b == 0 && tid == 0) { if (ca.bref_slot >= 0) ca.S[ca.bref_slot] = ref; ca.S[S_THRESH] = thresh; ca.S[S_ITERS] = 0.0; }
    }
    int it = 0;
    const double2 zz = make_double2(0.0, 0.0);
    while (it < ca.max_iters && rr > thresh && rr > 0.0) {
        // ---- phase B / C: Ap on the owned nodes
        if (tid < ne) {
            const int vx = lc0.x & 0xFFFF, vy = lc0.x >> 16, vz = lc0.y & 0xFFFF;
            double2 f0, f1, f2;
            tri_apply_elem<SM>(a.m, sX[vx], sX[vy], sX[vz], sP[vx], sP[vy], sP[vz], sV[vx], sV[vy], sV[vz],
                               NL ? sU[vx] : zz, NL ? sU[vy] : zz, NL ? sU[vz] : zz, f0, f1, f2);
            sE[tid] = f0; sE[chunk + tid] = f1; sE[2 * chunk + tid] = f2;
        }
        if (tid + kThreads < ne) {
            const int vx = lc1.x & 0xFFFF, vy = lc1.x >> 16, vz = lc1.y & 0xFFFF;
            double2 f0, f1, f2;
            tri_apply_elem<SM>(a.m, sX[vx], sX[vy], sX[vz], sP[vx], sP[vy], sP[vz], sV[vx], sV[vy], sV[vz],
                               NL ? sU[vx] : zz, NL ? sU[vy] : zz, NL ? sU[vz] : zz, f0, f1, f2);
            sE[tid + kThreads] = f0; sE[chunk + tid + kThreads] = f1; sE[2 * chunk + tid + kThreads] = f2;
        }
        __syncthreads();
        double2 Ap = zz, pv = zz;
        if (own) {
            const uint4 c8 = sI[tid];
            const unsigned cw[4] = {c8.x, c8.y, c8.z, c8.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int code = (int)((cw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
                const double2 f = sE[min((code & 3) * chunk + (code >> 2), zslot)];
                Ap.x += f.x; Ap.y += f.y;
            }
            if ((c8.w >> 16) == 0xFFFEu) {
                int k0 = bv.inc_ptr[n0 + tid] + 7, k1 = bv.inc_ptr[n0 + tid + 1];
                for (int k = k0; k < k1; ++k) {
                    const int code = bv.inc[k];
                    if (code >= 0xFFFE) break;
                    const double2 f = sE[(code & 3) * chunk + (code >> 2)];
                    Ap.x += f.x; Ap.y += f.y;
                }
            }
            if (mk & 0xFFu) Ap.x = 0.0;
            if (mk & 0xFF00u) Ap.y = 0.0;
            pv = sV[tid];
        }
        double pAp, dummy;
        grid_sum2(own ? pv.x * Ap.x + pv.y * Ap.y : 0.0, 0.0, pAp, dummy);
        const double alpha = rz / pAp;
        double lz = 0.0, lr = 0.0;
        if (own) {
            x.x += alpha * pv.x; x.y += alpha * pv.y;
            r.x -= alpha * Ap.x; r.y -= alpha * Ap.y;
            Rg[n0 + tid] = r;                                   // neighbours read their halo r after the barrier
            lz = r.x * dv.x * r.x + r.y * dv.y * r.y;
            lr = r.x * r.x + r.y * r.y;
        }
        double rz_new, rr_new;
        grid_sum2(lz, lr, rz_new, rr_new);
        const double beta = rz_new / rz;
        rz = rz_new; rr = rr_new;
        ++it;
        // ---- p = D^-1 r + beta p on every local node (halo r from the neighbours)
        for (int i = tid; i < n_loc; i += kThreads) {
            const double2 rg = (i < n_own) ? ((i == tid) ? r : __ldcg(Rg + n0 + i)) : __ldcg(Rg + sH[i]);
            const double2 d = sD[i];
            double2 pp = sV[i];
            pp.x = d.x * rg.x + beta * pp.x; pp.y = d.y * rg.y + beta * pp.y;
            sV[i] = pp;
        }
        __syncthreads();
        // the next write of Rg happens after the next grid barrier (p.Ap), so these reads are safe
    }
    // ---- save state and publish the PCG bookkeeping
    if (own) { Xg[n0 + tid] = x; Pg[n0 + tid] = sV[tid]; }
    if (b == 0 && tid == 0) {
        ca.S[S_RR] = rr; ca.S[S_RZ0] = rz;
        ca.S[S_ITERS] += (double)it;
        ca.S[S_DONE] = (!(rr > thresh) || !(rr > 0.0)) ? 1.0 : 0.0;
    }
}

// same structure for the phase-field operator J_phi (scalar unknown; coefficients cm, cg per node)
__global__ void __launch_bounds__(kThreads, 1) pcg_coop_phi_tri_kernel(BlockView bv, KArgs a, CoopArgs ca) {
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    __shared__ double s_tot[2];
    const int tid = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
    const int4 ma = bv.meta[2 * b], mb = bv.meta[2 * b + 1];
    const int n0 = ma.x, n_own = ma.y, h0 = ma.z, n_halo = ma.w, n_loc = n_own + n_halo, e0 = mb.x, ne = mb.y;
    const int ML = bv.max_loc, chunk = 2 * kThreads, MLe = (ML + 1) & ~1;
    double2* sX = reinterpret_cast<double2*>(smem);
    double* sM = reinterpret_cast<double*>(sX + ML);     // cm
    double* sG = sM + MLe;                               // cg
    double* sV = sG + MLe;                               // p on every local node
    double* sD = sV + MLe;                               // Jacobi data
    double* sE = sD + MLe;                               // [3*chunk + 2]
    int* sH = reinterpret_cast<int*>(sE + 3 * chunk + 2);
    uint4* sI = reinterpret_cast<uint4*>(sH + ((ML + 3) & ~3));
    const int zslot = 3 * chunk;
    const double2* X2 = reinterpret_cast<const double2*>(a.X);
    const uint2* LC = reinterpret_cast<const uint2*>(bv.lconn);
    for (int i = tid; i < n_loc; i += kThreads) {
        const int g = (i < n_own) ? n0 + i : bv.halo[h0 + i - n_own];
        sX[i] = X2[g]; sM[i] = a.cm[g]; sG[i] = a.cg[g]; sD[i] = ca.dinv[g]; sH[i] = g;
    }
    if (tid < n_own) sI[tid] = bv.inc8[n0 + tid];
    if (tid == 0) sE[zslot] = 0.0;
    uint2 lc0 = make_uint2(0u, 0u), lc1 = lc0;
    if (tid < ne) lc0 = LC[e0 + tid];
    if (tid + kThreads < ne) lc1 = LC[e0 + tid + kThreads];
    unsigned int mk = 0u;
    double x = 0.0, r = 0.0, dv = 0.0;
    const bool own = tid < n_own;
    if (own) {
        if (a.mask) mk = a.mask[n0 + tid];
        dv = ca.dinv[n0 + tid];
        if (ca.resume) { x = ca.x[n0 + tid]; r = ca.r[n0 + tid]; }
        else { r = mk ? 0.0 : ca.F[n0 + tid]; ca.r[n0 + tid] = r; }
    }
    unsigned long long gen = 0;
    int red_par = 0;
    auto grid_sum2 = [&](double v0, double v1, double& t0, double& t1) {
        double vals[2] = {v0, v1};
        block_sum<2>(vals, s_red);
        double* part = ca.partial + (size_t)red_par * nb * 2;
        if (tid == 0) { part[2 * b] = vals[0]; part[2 * b + 1] = vals[1]; }
        ++gen;
        coop_barrier(ca.bar, gen * (unsigned long long)nb);
        if (tid < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int k = tid; k < nb; k += 32) { s0 += __ldcg(part + 2 * k); s1 += __ldcg(part + 2 * k + 1); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (tid == 0) { s_tot[0] = s0; s_tot[1] = s1; }
        }
        __syncthreads();
        t0 = s_tot[0]; t1 = s_tot[1];
        red_par ^= 1;
        __syncthreads();
    };
    double rz, rr;
    grid_sum2(own ? r * dv * r : 0.0, own ? r * r : 0.0, rz, rr);
    for (int i = tid; i < n_loc; i += kThreads) {
        const int g = sH[i];
        sV[i] = ca.resume ? ca.p[g] : sD[i] * __ldcg(ca.r + g);
    }
    __syncthreads();
    double thresh = ca.S[S_THRESH];
    if (!ca.resume) {
        double ref = rr;
        if (ca.bref_slot >= 0) ref = fmax(ca.S[ca.bref_slot], rr);
        thresh = ca.rtol2 * ref;
        if (b == 0 && tid == 0) { if (ca.bref_slot >= 0) ca.S[ca.bref_slot] = ref; ca.S[S_THRESH] = thresh; ca.S[S_ITERS] = 0.0; }
    }
    int it = 0;
    while (it < ca.max_iters && rr > thresh && rr > 0.0) {
        if (tid < ne) {
            const int vx = lc0.x & 0xFFFF, vy = lc0.x >> 16, vz = lc0.y & 0xFFFF;
            double f0, f1, f2;
            tri_apply_phi_elem(sX[vx], sX[vy], sX[vz], sM[vx], sM[vy], sM[vz], sG[vx], sG[vy], sG[vz], sV[vx], sV[vy], sV[vz], f0, f1, f2);
            sE[tid] = f0; sE[chunk + tid] = f1; sE[2 * chunk + tid] = f2;
        }
        if (tid + kThreads < ne) {
            const int vx = lc1.x & 0xFFFF, vy = lc1.x >> 16, vz = lc1.y & 0xFFFF;
            double f0, f1, f2;
            tri_apply_phi_elem(sX[vx], sX[vy], sX[vz], sM[vx], sM[vy], sM[vz], sG[vx], sG[vy], sG[vz], sV[vx], sV[vy], sV[vz], f0, f1, f2);
            sE[tid + kThreads] = f0; sE[chunk + tid + kThreads] = f1; sE[2 * chunk + tid + kThreads] = f2;
        }
        __syncthreads();
        double Ap = 0.0, pv = 0.0;
        if (own) {
            const uint4 c8 = sI[tid];
            const unsigned cw[4] = {c8.x, c8.y, c8.z, c8.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int code = (int)((cw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
                Ap += sE[min((code & 3) * chunk + (code >> 2), zslot)];
            }
            if ((c8.w >> 16) == 0xFFFEu) {
                int k0 = bv.inc_ptr[n0 + tid] + 7, k1 = bv.inc_ptr[n0 + tid + 1];
                for (int k = k0; k < k1; ++k) {
                    const int code = bv.inc[k];
                    if (code >= 0xFFFE) break;
                    Ap += sE[(code & 3) * chunk + (code >> 2)];
                }
            }
            if (mk) Ap = 0.0;
            pv = sV[tid];
        }
        double pAp, dummy;
        grid_sum2(own ? pv * Ap : 0.0, 0.0, pAp, dummy);
        const double alpha = rz / pAp;
        double lz = 0.0, lr = 0.0;
        if (own) {
            x += alpha * pv;
            r -= alpha * Ap;
            ca.r[n0 + tid] = r;
            lz = r * dv * r; lr = r * r;
        }
        double rz_new, rr_new;
        grid_sum2(lz, lr, rz_new, rr_new);
        const double beta = rz_new / rz;
        rz = rz_new; rr = rr_new;
        ++it;
        for (int i = tid; i < n_loc; i += kThreads) {
            const double rg = (i < n_own) ? ((i == tid) ? r : __ldcg(ca.r + n0 + i)) : __ldcg(ca.r + sH[i]);
            sV[i] = sD[i] * rg + beta * sV[i];
        }
        __syncthreads();
    }
    if (own) { ca.x[n0 + tid] = x; ca.p[n0 + tid] = sV[tid]; }
    if (b == 0 && tid == 0) {
        ca.S[S_RR] = rr; ca.S[S_RZ0] = rz;
        ca.S[S_ITERS] += (double)it;
        ca.S[S_DONE] = (!(rr > thresh) || !(rr > 0.0)) ? 1.0 : 0.0;
    }
}

// ============================================================================= element reductions
// One thread per PRIMARY element of the block; sums NR scalars over the whole mesh.
enum { RED_L2_U, RED_L2_PHI, RED_ENERGY };

template <class Cell, int KIND>
__global__ void __launch_bounds__(kThreads) reduce_kernel(BlockView bv, KArgs a) {
    constexpr int NV = Cell::NV, D = Cell::D;
    constexpr int NR = (KIND == RED_ENERGY) ? 8 : 2;
    constexpr int NC = (KIND == RED_L2_U) ? D : 1;
    constexpr int NF = odd<(KIND == RED_ENERGY) ? (2 * D + 3) : (D + 2 * NC)>::value;
    extern __shared__ double smem[];
    __shared__ double s_red[kMaxRed * 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n0 = bv.node0[b], n_own = bv.node0[b + 1] - n0;
    const int h0 = bv.halo_off[b], n_loc = n_own + bv.halo_off[b + 1] - h0;
    double* s_nodes = smem;
    for (int i = tid; i < n_loc; i += kThreads) {
        int g = (i < n_own) ? n0 + i : bv.halo[h0 + i - n_own];
        double* s = s_nodes + (size_t)i * NF;
        for (int k = 0; k < D; ++k) s[k] = a.X[(size_t)g * D + k];
        if (KIND == RED_ENERGY) {
            for (int k = 0; k < D; ++k) s[D + k] = a.u[(size_t)g * D + k];
            s[2 * D] = a.phi[g];
            s[2 * D + 1] = a.m.fatigue ? a.Fat[g] : 1.0;
            s[2 * D + 2] = a.m.fatigue ? a.f1[g] : 0.0;
        } else {
            for (int k = 0; k < NC; ++k) { s[D + k] = a.f1[(size_t)g * NC + k]; s[D + NC + k] = a.f2[(size_t)g * NC + k]; }
        }
    }
    __syncthreads();
    double vals[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) vals[k] = 0.0;
    const int e0 = bv.elem_off[b], ne = bv.nprim[b];
    for (int j = tid; j < ne; j += kThreads) {
        const double* nd[NV > 4 ? 8 : 4];
        if (NV <= 4) {
            ushort4 lc = bv.lconn[e0 + j];
            nd[0] = s_nodes + (size_t)lc.x * NF; nd[1] = s_nodes + (size_t)lc.y * NF;
            nd[2] = s_nodes + (size_t)lc.z * NF; nd[3] = s_nodes + (size_t)lc.w * NF;
        } else {
            const ushort4 la = bv.lconn[2 * (size_t)(e0 + j)], lb = bv.lconn[2 * (size_t)(e0 + j) + 1];
            nd[0] = s_nodes + (size_t)la.x * NF; nd[1] = s_nodes + (size_t)la.y * NF;
            nd[2] = s_nodes + (size_t)la.z * NF; nd[3] = s_nodes + (size_t)la.w * NF;
            nd[4 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.x * NF; nd[5 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.y * NF;
            nd[6 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.z * NF; nd[7 % (NV > 4 ? 8 : 4)] = s_nodes + (size_t)lb.w * NF;
        }
        double X[NV * D];
        for (int k = 0; k < NV; ++k) for (int i = 0; i < D; ++i) X[k * D + i] = nd[k][i];
        if (KIND == RED_ENERGY) {
            double U[NV * D], ph[NV], F[NV], ab[NV];
            for (int k = 0; k < NV; ++k) {
                for (int i = 0; i < D; ++i) U[k * D + i] = nd[k][D + i];
                ph[k] = nd[k][2 * D]; F[k] = nd[k][2 * D + 1]; ab[k] = nd[k][2 * D + 2];
            }
            elem_energies<Cell>(a.m, a.n1d, X, U, ph, F, ab, vals);
        } else {
            double A[NV * NC], B[NV * NC];
            for (int k = 0; k < NV; ++k) for (int i = 0; i < NC; ++i) { A[k * NC + i] = nd[k][D + i]; B[k * NC + i] = nd[k][D + NC + i]; }
            elem_l2<Cell, NC>(a.n1d, X, A, B, vals);
        }
    }
    grid_reduce<NR>(vals, a.partial, a.result, a.counter, s_red);
}

// ============================================================================= vector kernels
// PCG state scalars (device resident).  Indices into the scalar array S.


// Publishes the two sums of an update step: rz_next, rr; bumps the iteration count and sets
// the sticky DONE flag (all later launches of the captured graph then return immediately).
__device__ __forceinline__ void pcg_finalize(double* S, int par, double rz_new, double rr_new) {
    S[S_RZ0 + (par ^ 1)] = rz_new;
    S[S_RR] = rr_new;
    S[S_ITERS] += 1.0;
    if (!(rr_new > S[S_THRESH])) S[S_DONE] = 1.0;     // also stops on NaN
}

// x += α p ; r −= α Ap ; sums Σ r·Dinv·r, Σ r·r     (α = rz[par]/pAp)
// finalize != 0 (one GPU): the last block publishes directly; else the sums go to S[S_STAGE..]
// for the all-reduce + pcg_publish_kernel.
__global__ void __launch_bounds__(kThreads) pcg_update_kernel(int64_t n, double* __restrict__ x, double* __restrict__ r,
                                                              const double* __restrict__ p, const double* __restrict__ Ap,
                                                              const double* __restrict__ dinv, double* S, int par,
                                                              int finalize, double* partial, unsigned int* counter,
                                                              const PeerView* pv = nullptr) {
    __shared__ double s_red[kMaxRed * 32];
    __shared__ bool is_last;
    __shared__ double s_pap;
    if (S[S_DONE] != 0.0) return;
    double pap = S[S_PAP];
    if (pv) {                                   // fused all-reduce: p.Ap of all ranks from the mailboxes
        if (threadIdx.x < 32) {
            double sm[1];
            mbox_sum(pv, sm, 1);
            if (threadIdx.x == 0) s_pap = sm[0];
        }
        __syncthreads();
        pap = s_pap;
    }
    const double alpha = S[S_RZ0 + par] / pap;
    double vals[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, gs = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2 *p2 = reinterpret_cast<const double2*>(p), *A2 = reinterpret_cast<const double2*>(Ap),
                  *d2 = reinterpret_cast<const double2*>(dinv);
    for (int64_t i = i0; i < n2; i += gs) {                  // 128-bit streams
        const double2 pv = p2[i], av = A2[i], dv = d2[i];
        double2 rv = r2[i], xv = x2[i];
        rv.x -= alpha * av.x; rv.y -= alpha * av.y;
        xv.x += alpha * pv.x; xv.y += alpha * pv.y;
        x2[i] = xv; r2[i] = rv;
        vals[0] += rv.x * dv.x * rv.x + rv.y * dv.y * rv.y;
        vals[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && i0 == 0) {                                // odd tail
        const int64_t i = n - 1;
        double ri = r[i] - alpha * Ap[i];
        x[i] += alpha * p[i];
        r[i] = ri;
        vals[0] += ri * dinv[i] * ri;
        vals[1] += ri * ri;
    }
    block_sum<2>(vals, s_red);
    if (threadIdx.x == 0) {
        partial[(size_t)blockIdx.x * kMaxRed + 0] = vals[0];
        partial[(size_t)blockIdx.x * kMaxRed + 1] = vals[1];
        __threadfence();
        unsigned int t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[2] = {0.0, 0.0};
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            acc[0] += __ldcg(&partial[(size_t)b * kMaxRed + 0]);
            acc[1] += __ldcg(&partial[(size_t)b * kMaxRed + 1]);
        }
        block_sum<2>(acc, s_red);
        if (pv) {
            if (threadIdx.x < 32) {
                acc[0] = __shfl_sync(0xffffffffu, acc[0], 0);
                acc[1] = __shfl_sync(0xffffffffu, acc[1], 0);
                mbox_push(pv, acc, 2);
            }
        } else if (threadIdx.x == 0) {
            if (finalize) pcg_finalize(S, par, acc[0], acc[1]);
            else { S[S_STAGE] = acc[0]; S[S_STAGE + 1] = acc[1]; }
        }
    }
}

// multi-GPU: after the all-reduce of S[S_STAGE..S_STAGE+1]
__global__ void pcg_publish_kernel(double* S, int par) {
    if (S[S_DONE] != 0.0) return;
    pcg_finalize(S, par, S[S_STAGE], S[S_STAGE + 1]);
}

// p = Dinv r + β p   (β = rz[par^1]/rz[par])
__global__ void __launch_bounds__(kThreads) pcg_direction_kernel(int64_t n, double* __restrict__ p, const double* __restrict__ r,
                                                                 const double* __restrict__ dinv, double* S, int par,
                                                                 const PeerView* pv = nullptr) {
    __shared__ double s_sum[2];
    if (S[S_DONE] != 0.0) return;
    double rz_new = S[S_RZ0 + (par ^ 1)];
    if (pv) {                                   // fused all-reduce of (r.z, r.r); block 0 publishes them at its end
        if (threadIdx.x < 32) {
            double sm[2];
            mbox_sum(pv, sm, 2);
            if (threadIdx.x == 0) { s_sum[0] = sm[0]; s_sum[1] = sm[1]; }
        }
        __syncthreads();
        rz_new = s_sum[0];
    }
    const double beta = rz_new / S[S_RZ0 + par];
    const int64_t n2 = n >> 1, gs = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2 *r2 = reinterpret_cast<const double2*>(r), *d2 = reinterpret_cast<const double2*>(dinv);
    for (int64_t i = i0; i < n2; i += gs) {
        const double2 rv = r2[i], dv = d2[i];
        double2 pv = p2[i];
        pv.x = dv.x * rv.x + beta * pv.x; pv.y = dv.y * rv.y + beta * pv.y;
        p2[i] = pv;
    }
    if ((n & 1) && i0 == 0) p[n - 1] = dinv[n - 1] * r[n - 1] + beta * p[n - 1];
    if (pv && blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x == 0) pcg_finalize(S, par, s_sum[0], s_sum[1]);
    }
}

// generic fused vector helpers -------------------------------------------------------------
// out = a + s*b  (b may be null => copy/scale)
__global__ void axpby_kernel(int64_t n, double* out, const double* a, double sa, const double* b, double sb) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (a ? sa * a[i] : 0.0) + (b ? sb * b[i] : 0.0);
}

// identity rows of the Dirichlet dofs: y = v there (dolfinx assemble_matrix(bcs) puts 1 on the diagonal)
__global__ void mask_restore_kernel(int64_t n, const uint8_t* mask, const double* v, double* y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (mask[i]) y[i] = v[i];
}

__global__ void mask_zero_kernel(int64_t n, const uint8_t* mask, double* v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (mask[i]) v[i] = 0.0;
}

// SNES residual assembly (dolfinx NonlinearProblem idiom):
//   F_free = f_int + lift (− f_ext),  F_bc = x − g ;  rhs b = F on free dofs, 0 on Dirichlet dofs;
//   also delta_bc = F_bc.  Partial sums: ||F||², ||x||².
__global__ void __launch_bounds__(kThreads) snes_residual_kernel(int64_t n, const double* fint, const double* lift,
                                                                 const double* fext, const double* x, const double* g,
                                                                 const uint8_t* mask, double* F, double* partial,
                                                                 double* result, unsigned int* counter) {
    __shared__ double s_red[kMaxRed * 32];
    double vals[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double f;
        if (mask[i]) f = x[i] - g[i];
        else f = fint[i] + (lift ? lift[i] : 0.0) - (fext ? fext[i] : 0.0);
        F[i] = f;
        vals[0] += f * f;
        vals[1] += x[i] * x[i];
    }
    grid_reduce<2>(vals, partial, result, counter, s_red);
}

// dbc = mask ? g − x : 0 ; result[0] = Σ dbc²
__global__ void __launch_bounds__(kThreads) bc_gap_kernel(int64_t n, const double* x, const double* g, const uint8_t* mask,
                                                          double* dbc, double* partial, double* result,
                                                          unsigned int* counter) {
    __shared__ double s_red[kMaxRed * 32];
    double vals[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double d = mask[i] ? g[i] - x[i] : 0.0;
        dbc[i] = d;
        vals[0] += d * d;
    }
    grid_reduce<1>(vals, partial, result, counter, s_red);
}

// PCG start: r = mask ? 0 : F ; x(=δ) = 0 ; p = Dinv r ; sums rz, rr
__global__ void __launch_bounds__(kThreads) pcg_init_kernel(int64_t n, const double* F, const uint8_t* mask, const double* dinv,
                                                            const double* x0, const double* Ax0, double* x, double* r,
                                                            double* p, double* partial, double* result,
                                                            unsigned int* counter) {
    __shared__ double s_red[kMaxRed * 32];
    double vals[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double ri = (mask && mask[i]) ? 0.0 : F[i];
        if (x0) { ri -= Ax0[i]; x[i] = x0[i]; } else x[i] = 0.0;
        if (mask && mask[i]) ri = 0.0;
        r[i] = ri;
        double zi = dinv[i] * ri;
        p[i] = zi;
        vals[0] += ri * zi;
        vals[1] += ri * ri;
    }
    grid_reduce<2>(vals, partial, result, counter, s_red);
}

// Newton update x −= δ on free dofs, x = g on Dirichlet dofs; sums ||δ_total||² for the stol test
__global__ void __launch_bounds__(kThreads) newton_update_kernel(int64_t n, double* x, const double* delta, const double* F,
                                                                 const uint8_t* mask, double* partial, double* result,
                                                                 unsigned int* counter) {
    __shared__ double s_red[kMaxRed * 32];
    double vals[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double d = mask[i] ? F[i] : delta[i];
        x[i] -= d;
        vals[0] += d * d;
    }
    grid_reduce<1>(vals, partial, result, counter, s_red);
}

// nodal history / fatigue updates (solver.py:357-375, 401-411)
__global__ void nodal_max_kernel(int64_t n, double* out, const double* a, const double* b) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fmax(a[i], b[i]);
}

__global__ void fatigue_update_kernel(int64_t n, const double* alpha_c, const double* alpha_n, const double* abar_n,
                                      double* abar_c, double* Fat, double dt, double alpha_T) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double da = alpha_c[i] - alpha_n[i];
        double hv = (da / dt < 0.0) ? 0.0 : 1.0;            // np.heaviside(x, 1)
        double ab = abar_n[i] + fabs(da) * hv;
        abar_c[i] = ab;
        double t = 2.0 * alpha_T / (ab + alpha_T);
        Fat[i] = (ab >= alpha_T) ? t * t : 1.0;
    }
}

// gather-sum of a vector over an index list (reactions): result[c] = Σ_{k, comp(idx)=c} f[idx[k]]
__global__ void __launch_bounds__(kThreads) gather_sum_kernel(int64_t n, const int32_t* idx, const double* f, int d,
                                                              double* partial, double* result, unsigned int* counter) {
    __shared__ double s_red[kMaxRed * 32];
    double vals[3] = {0.0, 0.0, 0.0};
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int i = idx[k];
        int c = i % d;
        double v = f[i];
        vals[0] += (c == 0) ? v : 0.0; vals[1] += (c == 1) ? v : 0.0; vals[2] += (c == 2) ? v : 0.0;
    }
    grid_reduce<3>(vals, partial, result, counter, s_red);
}

// scatter values into a vector at an index list (Dirichlet values, tractions)
__global__ void scatter_set_kernel(int64_t n, const int32_t* idx, const double* vals, double* out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[idx[k]] = vals[k];
}

// halo exchange pack: buf[k] = v[idx[k]*nc + c]
__global__ void pack_kernel(int64_t n, const int32_t* idx, int nc, const double* v, double* buf) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n * nc; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = k / nc; int c = (int)(k - j * nc);
        buf[k] = v[(int64_t)idx[j] * nc + c];
    }
}

// ============================================================================= NVLink peer path
// One process per GPU; peers' buffers are mapped with CUDA IPC (pfx_comm_p2p_import).  Halo values
// are PUSHED straight into the neighbours' ghost ranges with plain stores over NVLink and the CG
// scalars are all-reduced through per-rank mailboxes — no NCCL call inside the PCG iteration.
// Every rank runs the same launch sequence, so sequence counters advance in lock step; waits are
// bounded (a timeout raises a flag instead of hanging the GPU).
// values of the send list -> the neighbours' ghost ranges; the last block raises the halo flags
__global__ void __launch_bounds__(kThreads) halo_push_kernel(int64_t n_send, const int32_t* idx, const uint8_t* nbr,
                                                             const int64_t* dst, int nc, const double* v,
                                                             double* const* peer_vec /*[n_nbrs]*/, PeerView pv,
                                                             unsigned int* counter, const double* stop) {
    if (stop && stop[0] != 0.0) return;
    __shared__ bool is_last;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_send * nc; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = k / nc;
        int c = (int)(k - j * nc);
        peer_vec[nbr[j]][dst[j] * nc + c] = v[(int64_t)idx[j] * nc + c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x < pv.n_nbrs) {
        __threadfence_system();
        unsigned long long sq = pv.seq[1] + 1ull;
        st_release_sys(pv.hflag_peer[pv.nbr_rank[threadIdx.x]] + pv.rank, sq);
    }
    if (is_last) {
        __syncthreads();
        if (threadIdx.x == 0) pv.seq[1] = pv.seq[1] + 1ull;
    }
}

// blocks the stream until every neighbour's push of the current halo sequence has landed
__global__ void halo_wait_kernel(PeerView pv, const double* stop) {
    if (stop && stop[0] != 0.0) return;
    if ((int)threadIdx.x < pv.n_nbrs) {
        const unsigned long long want = pv.seq[1];
        const unsigned long long* f = pv.hflag_local + pv.nbr_rank[threadIdx.x];
        long long spins = 0;
        while (ld_acquire_sys(f) < want) {
            if (++spins > kSpinLimit) { *pv.err = 1.0; break; }
        }
    }
}

// all-reduce (sum) of n <= 8 staged doubles through the mailboxes, then the PCG bookkeeping:
// mode 0: S[S_PAP] = sum0 ; mode 1: pcg_finalize(par, sum0, sum1) ; mode 2: stage[] = sums
__global__ void p2p_allreduce_kernel(PeerView pv, double* S, double* stage, int n, int mode, int par) {
    if (mode != 2 && S[S_DONE] != 0.0) return;
    const int lane = threadIdx.x;
    unsigned long long sq = 0;
    if (lane == 0) { sq = pv.seq[0] + 1ull; pv.seq[0] = sq; }
    sq = __shfl_sync(0xffffffffu, sq, 0);
    const int slot = (int)(sq & 3ull);
    if (lane < pv.n_ranks) {
        double* dst = pv.mbox_peer[lane] + ((size_t)slot * pv.n_ranks + pv.rank) * kMboxStride;
        for (int k = 0; k < n; ++k) dst[k] = stage[k];
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(dst + 8), sq);
    }
    double vals[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) vals[k] = 0.0;
    if (lane < pv.n_ranks) {
        const double* src = pv.mbox_local + ((size_t)slot * pv.n_ranks + lane) * kMboxStride;
        long long spins = 0;
        while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(src + 8)) != sq) {
            if (++spins > kSpinLimit) { *pv.err = 2.0; break; }
        }
        for (int k = 0; k < n; ++k) vals[k] = __ldcv(src + k);
    }
    // fixed-order sum over ranks (identical on every rank)
    double sums[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double t = 0.0;
        for (int r = 0; r < pv.n_ranks; ++r) t += __shfl_sync(0xffffffffu, vals[k], r);
        sums[k] = t;
    }
    if (lane == 0) {
        if (mode == 0) S[S_PAP] = sums[0];
        else if (mode == 1) pcg_finalize(S, par, sums[0], sums[1]);
        else for (int k = 0; k < n; ++k) stage[k] = sums[k];
    }
}

// L2 flush helper for timing
__global__ void flush_kernel(int64_t n, double* buf) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] += 1.0;
}

}  // namespace pfx
