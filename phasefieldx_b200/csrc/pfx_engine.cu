// pfx_engine.cu — context, device-resident staggered load step, and the C-ABI of include/pfx_b200.h.
//
// Host control flow restates pfx/Element/Phase_Field_Fracture/solver/solver.py:319-411 (stagger
// loop) with dolfinx NonlinearProblem/PETSc SNES semantics and a Jacobi-PCG in
// place of the MUMPS factorisation.  All arithmetic on fields runs in the kernels of
// pfx_kernels.cuh; there is no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/pfx_b200.h"
#include "pfx_blocks.h"
#include "pfx_comm.h"
#include "pfx_kernels.cuh"
#include "pfx_coarse.cuh"
#include "pfx_dense.cuh"
#include "pfx_sparse.cuh"
#include "pfx_monolithic.cuh"

using namespace pfx;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return PFX_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

namespace {

enum LinKind { LIN_U = 0, LIN_PHI = 1, LIN_MASS = 2 };

struct BCSet {
    std::vector<int32_t> dofs;    // internal (block-order) dof ids
    int32_t* d_dofs = nullptr;
    int32_t* d_owned = nullptr;   // subset on owned nodes (reactions)
    int64_t n_owned = 0;
    double* d_vals = nullptr;
    int64_t n = 0;
};

struct Traction {
    int32_t* d_nodes = nullptr;
    double* d_w = nullptr;
    int64_t n = 0;
    double T[3] = {0, 0, 0};
};

// aggregate coarse space of the PCG (pfx_coarse.cuh); op 0 = J_u, op 1 = J_phi
struct CoarseOp {
    float* Ainv = nullptr;          // one dense fp32 inverse per region
    int m = 0, ld = 0;              // modes per aggregate; largest row stride of a region
    int n_ctas = 0;                 // coarse solve: one row per warp
    int4* reg_tab = nullptr;
    int2* cta_tab = nullptr;
    long long* ac_off = nullptr;
    std::vector<int4> h_reg;        // {first aggregate, aggregates, float offset of the inverse, ld}
    std::vector<long long> h_ac_off;
    int64_t ainv_floats = 0;
    bool valid = false, stale = true;
    int64_t ref_its = -1;           // most PCG iterations of a solve in the load step(s) right after the last refresh
    bool ref_open = false;
    int open_steps = 0;
    cudaGraphExec_t graph = nullptr;
};

// assembled operators of tetrahedral meshes (pfx_sparse.cuh)
struct Sparse {
    bool enabled = false;
    SparseView sv;
    double *Ku = nullptr, *Kphi = nullptr, *Km = nullptr;
    int64_t total_k = 0;
    bool phi_fresh = false;          // Kphi belongs to the current H / fatigue field
};

struct Coarse {
    bool enabled = false;
    int n_agg = 0, S = 1, n_chunks = 0, n_col = 0, n_reg = 1;
    int32_t *chunk_node0 = nullptr, *colour = nullptr, *nbrcol = nullptr, *agg_reg = nullptr;
    float* rel = nullptr;
    double *wpart = nullptr, *y = nullptr, *Ac = nullptr, *partial = nullptr, *work = nullptr;
    double* dblk = nullptr;          // inverse nodal diagonal blocks of J_u (block-Jacobi smoother)
    double *workT = nullptr, *workL = nullptr;   // dense inverse (pfx_dense.cuh): panel and diagonal-block inverses; `work` = Y
    int* info = nullptr;
    CoarseOp op[2];
    bool refresh_every_step = false;
    bool block_jacobi = true;        // PFX_BLOCK_JACOBI=0: plain diagonal smoother for J_u
    int64_t setups = 0;
    // third, rank-level coarse space of J_u on the fused multi-GPU path (pfx_coarse.cuh: Coarse3View)
    std::vector<double> h_cen, h_half;   // aggregate centroids [n_agg*3] and half extents [n_agg]
    double *T3 = nullptr, *E3 = nullptr, *Einv3 = nullptr;
    bool lvl3_ready = false, lvl3_on = false;
    bool lvl3_wanted = true;             // PFX_NO_LEVEL3=1 switches it off
};

}  // namespace

struct pfx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int cell_type = 0, nv = 0, dim = 0;
    pfx_params prm;
    Material mat;
    BlockMesh bm;
    BlockView bv;
    int64_t n_nodes = 0, n_owned = 0;
    // fields (block order, length n_nodes or n_nodes*dim)
    double *X = nullptr, *u = nullptr, *u_old = nullptr, *phi = nullptr, *phi_old = nullptr, *H = nullptr,
           *Vc = nullptr, *Vn = nullptr, *Fat = nullptr, *alpha_c = nullptr, *alpha_n = nullptr, *abar_c = nullptr,
           *abar_n = nullptr, *dinv_mass = nullptr;
    // work vectors (u-space sized)
    double *w_r = nullptr, *w_p = nullptr, *w_Ap = nullptr, *w_x = nullptr, *w_F = nullptr, *w_dinv = nullptr,
           *w_t0 = nullptr, *w_t1 = nullptr, *g_u = nullptr, *g_phi = nullptr, *fext = nullptr;
    uint8_t *mask_u = nullptr, *mask_phi = nullptr;
    bool has_fext = false;
    std::vector<BCSet> bcs_u, bcs_phi;
    std::vector<Traction> tractions;
    // reductions
    double* partial = nullptr;
    double* S = nullptr;          // PCG scalars
    double* red = nullptr;        // generic result buffer [kMaxRed]
    unsigned int* counter = nullptr;
    double* h_pinned = nullptr;   // [64]
    int vec_grid = 0;
    // graphs
    cudaGraphExec_t graph[3] = {nullptr, nullptr, nullptr};
    int graph_iters = 0;
    bool use_graph = true;
    // comm
    nccl_comm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    std::vector<int32_t> nbrs;
    std::vector<int64_t> send_off, recv_off;
    int32_t* d_send_idx = nullptr;
    double* d_send_buf = nullptr;
    int64_t n_send = 0;
    // NVLink peer path
    bool p2p = false;
    PeerView pv;
    double* mbox = nullptr;           // mailboxes + halo flags + counters (IPC-exported)
    uint8_t* d_send_nbr = nullptr;
    PeerView* d_pv = nullptr;          // device copy of pv (fused waits / mailbox pushes inside the big kernels)
    bool p2p_fused = true;             // PFX_P2P_FUSED=0: separate wait / all-reduce kernels
    int64_t* d_send_dst = nullptr;
    double** d_peer_vec[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<void*> ipc_opened;
    // misc
    std::vector<void*> allocs;
    std::string err;
    int64_t launches = 0;
    int64_t floor_solves = 0;     // PCG solves of the current load step that met the per-step absolute floor before iterating
    // per-phase device timers of pfx_load_step: events recorded at phase boundaries, read after the final sync
    std::vector<cudaEvent_t> ph_events;
    std::vector<int> ph_ids;
    size_t ph_used = 0;
    bool ph_on = false;
    double* flush_buf = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t flush_n = 0;
    bool configuring = false;
    bool verbose = false;         // PFX_VERBOSE=1: per-Newton-iteration trace on stderr
    bool cg_adaptive = true;      // PFX_CG_ADAPTIVE=0: every PCG solve to cg_rtol of its own right-hand side
    double* tan = nullptr;        // cached element tangents [n_own_cells][21] (tets)
    bool tan_fresh = false;       // the cached tangents (and the matrices assembled from them) belong to the current u, phi
    double *cm = nullptr, *cg = nullptr;   // phi-operator nodal coefficients (triangles, v3 path)
    double *ones = nullptr, *zeros = nullptr;   // coefficients that turn the phi operator into the mass matrix
    int phi_v3_ctas = 0;
    Sparse sp;
    int coop_capacity = 0;        // co-resident CTAs of the whole-PCG cooperative kernel (0: unavailable)
    unsigned long long* coop_bar = nullptr;
    bool phi_coeff_fresh = false;
    bool generic_apply = false;   // PFX_GENERIC_APPLY=1: use the generic scatter kernel for J_u·v on triangles
    int apply_variant = 3;        // PFX_APPLY_VARIANT: 3 persistent TMA-staged kernel (default), 1 one CTA per block
    int v3_ctas = 0, num_sms = 148;
    Coarse coarse;
    void* ec = nullptr;           // EnergyCtl of the energy-controlled solvers (pfx_ec.inc), created on first use
};

namespace {

template <class T>
int dev_alloc(pfx_ctx* c, T** p, size_t n) {
    CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    CK(cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), c->stream));
    c->allocs.push_back((void*)*p);
    return PFX_OK;
}

template <class T>
int dev_upload(pfx_ctx* c, T** p, const std::vector<T>& h) {
    int st = dev_alloc(c, p, h.size());
    if (st) return st;
    if (!h.empty()) CK(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PFX_OK;
}

// phase ids of pfx_step_report.phase_ms
enum { PH_U_SETUP = 0, PH_U_COARSE = 1, PH_U_PCG = 2, PH_PROJECTION = 3, PH_FATIGUE = 4, PH_PHI_SETUP = 5, PH_PHI_COARSE = 6,
       PH_PHI_PCG = 7, PH_NORMS = 8, PH_OTHER = 9, PH_COUNT = 10 };

// the phase `id` starts now on the context stream (no-op outside pfx_load_step)
inline void mark(pfx_ctx* c, int id) {
    if (!c->ph_on) return;
    if (c->ph_used == c->ph_events.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { c->ph_on = false; return; }
        c->ph_events.push_back(e);
        c->ph_ids.push_back(0);
    }
    cudaEventRecord(c->ph_events[c->ph_used], c->stream);
    c->ph_ids[c->ph_used++] = id;
}

inline int vgrid(const pfx_ctx* c, int64_t n) {
    int64_t g = (n + kThreads - 1) / kThreads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, c->vec_grid));
}

// --------------------------------------------------------------------------- kernel dispatch
template <class Op>
int launch_scatter_op(pfx_ctx* c, const KArgs& a) {
    BlockView bv = c->bv;
    int per_elem = Op::NV * Op::NOUT;
    // element buffer of at most `kb` KB; whole multiples of the CTA width keep phase B balanced (every thread
    // the same number of cells per chunk).  48 KB: 512 tets per chunk for the vector operators (measured on the
    // 3D beam: 312 us per PCG iteration with 320-cell chunks, 266 us with 512; 768 and 1024 drop to one CTA per SM)
    static const int kb = getenv("PFX_CHUNK_KB") ? atoi(getenv("PFX_CHUNK_KB")) : 48;
    int cap = std::max(32, (kb * 1024 / (per_elem * 8)) / 32 * 32);
    if (cap >= kThreads) cap = cap / kThreads * kThreads;
    bv.chunk = std::min((c->bm.max_elems + 31) / 32 * 32, cap);
    size_t smem = ((((size_t)c->bm.max_loc * Op::NF + (size_t)per_elem * bv.chunk) + 1) & ~(size_t)1) * sizeof(double) +
                  ((size_t)c->bv.max_inc * sizeof(uint16_t) + 15) / 16 * 16;
    if (c->configuring) {           // pfx_create: opt in to > 48 KB dynamic shared memory, no launch
        if (smem + 4096 > 48 * 1024)      // static (reduction scratch) + dynamic must stay below 48 KB unless opted in
            CK(cudaFuncSetAttribute(scatter_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return PFX_OK;
    }
    scatter_kernel<Op><<<c->bm.nb, kThreads, smem, c->stream>>>(bv, a);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

template <int SM>
int launch_apply_u_tri(pfx_ctx* c, const KArgs& a) {
    BlockView bv = c->bv;
    bv.chunk = std::min((c->bm.max_elems + 31) / 32 * 32, 2 * kThreads + 64);
    size_t ML = (size_t)c->bm.max_loc;
    size_t smem = (ML * (SM != M_ISO ? 3 : 2)) * sizeof(double2) + ((ML + 1) & ~(size_t)1) * sizeof(double) +
                  3 * (size_t)bv.chunk * sizeof(double2) + ((size_t)c->bv.max_inc * sizeof(uint16_t) + 15) / 16 * 16;
    if (c->configuring) {
        CK(cudaFuncSetAttribute(apply_u_tri_kernel<SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return PFX_OK;
    }
    apply_u_tri_kernel<SM><<<c->bm.nb, kThreads, smem, c->stream>>>(bv, a);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

template <int SM>
int launch_apply_u_tri_v3(pfx_ctx* c, const KArgs& a) {
    BlockView bv = c->bv;
    bv.chunk = 2 * kThreads;
    size_t ML = (size_t)c->bm.max_loc, MO = ((size_t)c->bm.max_own + 1) & ~(size_t)1;
    size_t stage = ML * 16 * (SM != M_ISO ? 3 : 2) + ((ML + 3) & ~(size_t)1) * 8 + MO * 16;
    size_t smem = 2 * stage + (3 * (size_t)bv.chunk + 1) * sizeof(double2);
    if (c->configuring) {
        CK(cudaFuncSetAttribute(apply_u_tri_v3_kernel<SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, apply_u_tri_v3_kernel<SM>, kThreads, smem));
        c->v3_ctas = std::max(1, nb) * c->num_sms;
        return PFX_OK;
    }
    int grid = std::min(c->bm.nb, c->v3_ctas);
    apply_u_tri_v3_kernel<SM><<<grid, kThreads, smem, c->stream>>>(bv, a);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

int launch_apply_phi_tri_v3(pfx_ctx* c, const KArgs& a) {
    BlockView bv = c->bv;
    bv.chunk = 2 * kThreads;
    size_t ML = (size_t)c->bm.max_loc, MO = ((size_t)c->bm.max_own + 1) & ~(size_t)1;
    size_t stage = ML * 16 + 3 * (((ML + 3) & ~(size_t)1) * 8) + MO * 16;
    size_t smem = 2 * stage + (3 * (size_t)bv.chunk + 2) * sizeof(double);
    if (c->configuring) {
        CK(cudaFuncSetAttribute(apply_phi_tri_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, apply_phi_tri_v3_kernel, kThreads, smem));
        c->phi_v3_ctas = std::max(1, nb) * c->num_sms;
        return PFX_OK;
    }
    int grid = std::min(c->bm.nb, c->phi_v3_ctas);
    apply_phi_tri_v3_kernel<<<grid, kThreads, smem, c->stream>>>(bv, a);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

// whole-PCG cooperative kernel (small triangle meshes, one GPU)
template <int SM>
int coop_u_tri(pfx_ctx* c, const KArgs& a, CoopArgs ca, bool configure) {
    size_t ML = (size_t)c->bm.max_loc, MO = ((size_t)c->bm.max_own + 1) & ~(size_t)1;
    size_t smem = 4 * ML * 16 + (3 * (size_t)(2 * kThreads) + 1) * 16 + ((ML + 1) & ~(size_t)1) * 8 +
                  ((ML + 3) & ~(size_t)3) * 4 + MO * 16;
    if (configure) {
        CK(cudaFuncSetAttribute(pcg_coop_u_tri_kernel<SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, coop = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_coop_u_tri_kernel<SM>, kThreads, smem));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
        c->coop_capacity = coop ? per_sm * c->num_sms : 0;
        return PFX_OK;
    }
    BlockView bv = c->bv;
    KArgs aa = a;
    void* args[] = {(void*)&bv, (void*)&aa, (void*)&ca};
    CK(cudaLaunchCooperativeKernel((void*)pcg_coop_u_tri_kernel<SM>, dim3(c->bm.nb), dim3(kThreads), args, smem, c->stream));
    c->launches++;
    return PFX_OK;
}

int coop_phi(pfx_ctx* c, const KArgs& a, CoopArgs ca, bool configure) {
    size_t ML = (size_t)c->bm.max_loc, MO = ((size_t)c->bm.max_own + 1) & ~(size_t)1, MLe = (ML + 1) & ~(size_t)1;
    size_t smem = ML * 16 + 4 * MLe * 8 + (3 * (size_t)(2 * kThreads) + 2) * 8 + ((ML + 3) & ~(size_t)3) * 4 + MO * 16 + 16;
    if (configure) {
        CK(cudaFuncSetAttribute(pcg_coop_phi_tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_coop_phi_tri_kernel, kThreads, smem));
        c->coop_capacity = std::min(c->coop_capacity, per_sm * c->num_sms);
        return PFX_OK;
    }
    BlockView bv = c->bv;
    KArgs aa = a;
    void* args[] = {(void*)&bv, (void*)&aa, (void*)&ca};
    CK(cudaLaunchCooperativeKernel((void*)pcg_coop_phi_tri_kernel, dim3(c->bm.nb), dim3(kThreads), args, smem, c->stream));
    c->launches++;
    return PFX_OK;
}

int coop_u(pfx_ctx* c, const KArgs& a, const CoopArgs& ca, bool configure) {
    if (c->mat.smodel == M_ISO) return coop_u_tri<M_ISO>(c, a, ca, configure);
    if (c->mat.smodel == M_SPECTRAL) return coop_u_tri<M_SPECTRAL>(c, a, ca, configure);
    return coop_u_tri<M_VOLDEV>(c, a, ca, configure);
}

template <class Cell>
int launch_scatter_cell(pfx_ctx* c, int op, const KArgs& a) {
    if (a.m.dfun != 0 || c->configuring) {
        // general degradation function: quadrature-based phase-field forms; the u-forms go through the generic
        // kernels (the hand-tuned triangle kernels have (1-phi)^2 built in)
        if (op == OP_APPLY_PHI) { int st = launch_scatter_op<OpPhiG<Cell, 1>>(c, a); if (st || !c->configuring) return st; }
        if (op == OP_RESIDUAL_PHI) { int st = launch_scatter_op<OpPhiG<Cell, 0>>(c, a); if (st || !c->configuring) return st; }
        if (op == OP_DIAG_PHI) { int st = launch_scatter_op<OpPhiG<Cell, 2>>(c, a); if (st || !c->configuring) return st; }
        if (op == OP_APPLY_U && !c->configuring) {
            if (a.m.smodel == M_ISO) return launch_scatter_op<OpApplyU<Cell, M_ISO>>(c, a);
            if (a.m.smodel == M_SPECTRAL) return launch_scatter_op<OpApplyU<Cell, M_SPECTRAL>>(c, a);
            return launch_scatter_op<OpApplyU<Cell, M_VOLDEV>>(c, a);
        }
    }
    if (op == OP_APPLY_MASS && std::is_same<Cell, Tri>::value && c->ones && !c->configuring && c->apply_variant >= 3 &&
        c->bm.max_elems <= 2 * kThreads && a.mask == nullptr) {
        KArgs m = a;                      // M = the phi operator with coefficients (1, 0)
        m.cm = c->ones; m.cg = c->zeros;
        return launch_apply_phi_tri_v3(c, m);
    }
    if (op == OP_APPLY_PHI && std::is_same<Cell, Tri>::value && c->cm && (a.m.dfun == 0 || c->configuring)) {
        if (c->configuring) { int st = launch_apply_phi_tri_v3(c, a); if (st) return st; }
        else if (c->apply_variant >= 3 && c->bm.max_elems <= 2 * kThreads && a.input_masked && c->phi_coeff_fresh)
            return launch_apply_phi_tri_v3(c, a);
    }
    switch (op) {
        case OP_APPLY_U:
            if (std::is_same<Cell, Tri>::value && c->configuring) {      // opt the persistent variant in
                int st = (a.m.smodel == M_ISO) ? launch_apply_u_tri_v3<M_ISO>(c, a)
                         : (a.m.smodel == M_SPECTRAL ? launch_apply_u_tri_v3<M_SPECTRAL>(c, a)
                                                     : launch_apply_u_tri_v3<M_VOLDEV>(c, a));
                if (st) return st;
            } else if (std::is_same<Cell, Tri>::value && c->apply_variant >= 3 && c->bm.max_elems <= 2 * kThreads &&
                       (a.mask == nullptr || a.input_masked)) {
                if (a.m.smodel == M_ISO) return launch_apply_u_tri_v3<M_ISO>(c, a);
                if (a.m.smodel == M_SPECTRAL) return launch_apply_u_tri_v3<M_SPECTRAL>(c, a);
                return launch_apply_u_tri_v3<M_VOLDEV>(c, a);
            }
            if (std::is_same<Cell, Tri>::value && !c->generic_apply) {
                if (a.m.smodel == M_ISO) return launch_apply_u_tri<M_ISO>(c, a);
                if (a.m.smodel == M_SPECTRAL) return launch_apply_u_tri<M_SPECTRAL>(c, a);
                return launch_apply_u_tri<M_VOLDEV>(c, a);
            }
            if (a.m.smodel == M_ISO) return launch_scatter_op<OpApplyU<Cell, M_ISO>>(c, a);
            if (a.m.smodel == M_SPECTRAL) return launch_scatter_op<OpApplyU<Cell, M_SPECTRAL>>(c, a);
            return launch_scatter_op<OpApplyU<Cell, M_VOLDEV>>(c, a);
        case OP_RESIDUAL_U: return launch_scatter_op<OpResidualU<Cell>>(c, a);
        case OP_DIAG_U: return launch_scatter_op<OpDiagU<Cell>>(c, a);
        case OP_APPLY_PHI: return launch_scatter_op<OpApplyPhi<Cell>>(c, a);
        case OP_RESIDUAL_PHI: return launch_scatter_op<OpResidualPhi<Cell>>(c, a);
        case OP_DIAG_PHI: return launch_scatter_op<OpDiagPhi<Cell>>(c, a);
        case OP_APPLY_MASS: return launch_scatter_op<OpApplyMass<Cell>>(c, a);
        case OP_DIAG_MASS: return launch_scatter_op<OpDiagMass<Cell>>(c, a);
        case OP_RHS_PSI: return launch_scatter_op<OpRhsPsi<Cell>>(c, a);
        case OP_RHS_FATIGUE: return launch_scatter_op<OpRhsFatigue<Cell>>(c, a);
        case OP_TANGENT_U:
            if constexpr (Cell::simplex) return launch_scatter_op<OpTangentU<Cell>>(c, a);
            break;
        case OP_DIAGBLK_U: return launch_scatter_op<OpDiagBlkU<Cell>>(c, a);
    }
    if (c->configuring) return PFX_OK;
    c->err = "unknown op";
    return PFX_ERR_BAD_ARG;
}

int launch_scatter(pfx_ctx* c, int op, const KArgs& a) {
    if (c->cell_type == PFX_CELL_TRIANGLE) return launch_scatter_cell<Tri>(c, op, a);
    if (c->cell_type == PFX_CELL_QUADRILATERAL) return launch_scatter_cell<Quad>(c, op, a);
    if (c->cell_type == PFX_CELL_HEXAHEDRON) return launch_scatter_cell<Hex>(c, op, a);
    return launch_scatter_cell<Tet>(c, op, a);
}

template <class Cell, int KIND>
int launch_reduce_k(pfx_ctx* c, const KArgs& a) {
    constexpr int D = Cell::D;
    constexpr int NC = (KIND == RED_L2_U) ? D : 1;
    constexpr int NF = odd<(KIND == RED_ENERGY) ? (2 * D + 3) : (D + 2 * NC)>::value;
    size_t smem = (size_t)c->bm.max_loc * NF * sizeof(double);
    if (c->configuring) {
        if (smem + 4096 > 48 * 1024)
            CK(cudaFuncSetAttribute(reduce_kernel<Cell, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return PFX_OK;
    }
    reduce_kernel<Cell, KIND><<<c->bm.nb, kThreads, smem, c->stream>>>(c->bv, a);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

template <int KIND>
int launch_reduce(pfx_ctx* c, const KArgs& a) {
    if (c->cell_type == PFX_CELL_TRIANGLE) return launch_reduce_k<Tri, KIND>(c, a);
    if (c->cell_type == PFX_CELL_QUADRILATERAL) return launch_reduce_k<Quad, KIND>(c, a);
    if (c->cell_type == PFX_CELL_HEXAHEDRON) return launch_reduce_k<Hex, KIND>(c, a);
    return launch_reduce_k<Tet, KIND>(c, a);
}

KArgs base_args(pfx_ctx* c) {
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.m = c->mat;
    a.n1d = c->prm.quad_points_1d;
    a.X = c->X; a.u = c->u; a.phi = c->phi; a.H = c->H; a.Fat = c->Fat;
    a.partial = c->partial; a.result = c->red; a.counter = c->counter;
    a.tan = c->tan;
    a.cm = c->cm; a.cg = c->cg;
    return a;
}

// --------------------------------------------------------------------------- communication
int allreduce(pfx_ctx* c, double* buf, int n) {
    if (c->n_ranks <= 1) return PFX_OK;
    if (c->p2p) {
        p2p_allreduce_kernel<<<1, 32, 0, c->stream>>>(c->pv, c->S, buf, n, 2, 0);
        c->launches++;
        CK(cudaGetLastError());
        return PFX_OK;
    }
    int r = nccl().AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c->comm, c->stream);
    if (r != 0) { c->err = std::string("ncclAllReduce: ") + nccl().GetErrorString(r); return PFX_ERR_COMM; }
    return PFX_OK;
}

// forward halo: owners send the values of their boundary nodes, ghosts receive in place
int exchange(pfx_ctx* c, double* v, int nc, const double* stop = nullptr, bool consumer_waits = false) {
    if (c->n_ranks <= 1 || c->nbrs.empty()) return PFX_OK;
    if (c->p2p) {
        int vid = v == c->w_p ? 0 : (v == c->u ? 1 : (v == c->phi ? 2 : (v == c->Vc ? 3 : (v == c->alpha_c ? 4 : -1))));
        if (vid < 0) { c->err = "exchange: vector is not peer-mapped"; return PFX_ERR_COMM; }
        halo_push_kernel<<<vgrid(c, std::max<int64_t>(c->n_send * nc, 1)), kThreads, 0, c->stream>>>(
            c->n_send, c->d_send_idx, c->d_send_nbr, c->d_send_dst, nc, v, c->d_peer_vec[vid], c->pv, c->counter + 1, stop);
        if (!consumer_waits) halo_wait_kernel<<<1, 32, 0, c->stream>>>(c->pv, stop);
        c->launches += 2;
        CK(cudaGetLastError());
        return PFX_OK;
    }
    if (c->n_send > 0) {
        pack_kernel<<<vgrid(c, c->n_send * nc), kThreads, 0, c->stream>>>(c->n_send, c->d_send_idx, nc, v, c->d_send_buf);
        c->launches++;
    }
    int r = nccl().GroupStart();
    for (size_t k = 0; k < c->nbrs.size() && r == 0; ++k) {
        int64_t ns = c->send_off[k + 1] - c->send_off[k], nr = c->recv_off[k + 1] - c->recv_off[k];
        if (ns > 0) r = nccl().Send(c->d_send_buf + c->send_off[k] * nc, (size_t)(ns * nc), kNcclFloat64, c->nbrs[k], c->comm, c->stream);
        if (nr > 0 && r == 0)
            r = nccl().Recv(v + (c->n_owned + c->recv_off[k]) * nc, (size_t)(nr * nc), kNcclFloat64, c->nbrs[k], c->comm, c->stream);
    }
    int r2 = nccl().GroupEnd();
    if (r != 0 || r2 != 0) { c->err = std::string("halo exchange: ") + nccl().GetErrorString(r ? r : r2); return PFX_ERR_COMM; }
    return PFX_OK;
}

// device result buffer -> host (after optional all-reduce); synchronises the stream
int fetch(pfx_ctx* c, double* dev, int n, double* out) {
    int st = allreduce(c, dev, n);
    if (st) return st;
    CK(cudaMemcpyAsync(c->h_pinned, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) out[i] = c->h_pinned[i];
    return PFX_OK;
}

// --------------------------------------------------------------------------- Dirichlet data
int rebuild_bc(pfx_ctx* c, bool is_u) {
    int64_t n = c->n_nodes * (is_u ? c->dim : 1);
    uint8_t* mask = is_u ? c->mask_u : c->mask_phi;
    std::vector<BCSet>& sets = is_u ? c->bcs_u : c->bcs_phi;
    std::vector<uint8_t> hm(n, 0);
    for (auto& s : sets)
        for (int32_t d : s.dofs) hm[d] = 1;
    CK(cudaMemcpyAsync(mask, hm.data(), n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->coarse.op[is_u ? 0 : 1].stale = true;      // the masked modes changed
    return PFX_OK;
}

// values of all BC sets -> g (later ids win: launched in id order on one stream)
int push_bc_values(pfx_ctx* c, bool is_u) {
    double* g = is_u ? c->g_u : c->g_phi;
    std::vector<BCSet>& sets = is_u ? c->bcs_u : c->bcs_phi;
    for (auto& s : sets)
        if (s.n > 0) {
            scatter_set_kernel<<<vgrid(c, s.n), kThreads, 0, c->stream>>>(s.n, s.d_dofs, s.d_vals, g);
            c->launches++;
        }
    CK(cudaGetLastError());
    return PFX_OK;
}

__global__ void traction_add_kernel(int64_t n, const int32_t* nodes, const double* w, double t0, double t1, double t2,
                                    int d, double* fext) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        double* f = fext + (int64_t)nodes[k] * d;
        f[0] += w[k] * t0;
        f[1] += w[k] * t1;
        if (d == 3) f[2] += w[k] * t2;
    }
}

int rebuild_fext(pfx_ctx* c) {
    if (!c->has_fext) return PFX_OK;
    int64_t n = c->n_nodes * c->dim;
    CK(cudaMemsetAsync(c->fext, 0, n * sizeof(double), c->stream));
    for (auto& t : c->tractions)
        if (t.n > 0) {
            traction_add_kernel<<<vgrid(c, t.n), kThreads, 0, c->stream>>>(t.n, t.d_nodes, t.d_w, t.T[0], t.T[1], t.T[2], c->dim, c->fext);
            c->launches++;
        }
    CK(cudaGetLastError());
    return PFX_OK;
}

// --------------------------------------------------------------------------- PCG
// bref_slot >= 0: the stop threshold is floored at rtol^2 * (largest ||b||^2 this operator has seen in
// the current load step), so later Newton / stagger iterations, whose right-hand sides are orders of
// magnitude smaller, are solved to the same ABSOLUTE accuracy as the first solve instead of to 11
// digits of their own tiny residual (the reference's MUMPS solves are exact; what parity needs is the
// absolute error of u, phi).
__global__ void pcg_setup_kernel(double* S, double rtol2, int bref_slot) {
    double rz = S[S_STAGE], rr = S[S_STAGE + 1];
    S[S_RZ0] = rz; S[S_RZ1] = 0.0; S[S_RR] = rr; S[S_PAP] = 1.0;
    double ref = rr;
    if (bref_slot >= 0) {
        ref = fmax(S[bref_slot], rr);
        S[bref_slot] = ref;
    }
    S[S_THRESH] = rtol2 * ref;
    S[S_ITERS] = 0.0;
    S[S_DONE] = (rr > 0.0 && rr > S[S_THRESH]) ? 0.0 : 1.0;
}

// does the apply of this operator run on a kernel that waits for the halo in its own prologue?
bool apply_waits_for_halo(const pfx_ctx* c, int kind) {
    if (c->sp.enabled) return false;            // assembled operators: the SpMV does not wait itself
    if (c->cell_type != PFX_CELL_TRIANGLE || c->apply_variant < 3 || c->bm.max_elems > 2 * kThreads || c->mat.dfun != 0) return false;
    return kind == LIN_U || (kind == LIN_PHI && c->cm && c->phi_coeff_fresh);
}

int apply_lin(pfx_ctx* c, int kind, const double* v, double* y, const uint8_t* mask, double* result, const double* stop,
              int input_masked = 0, const PeerView* pv = nullptr, int wait_halo = 0) {
    if (c->sp.enabled && (mask == nullptr || input_masked)) {
        // tetrahedra: one SpMV with the assembled operator
        const Sparse& sp = c->sp;
        const int grid = (int)((sp.sv.n_slices * 32 + kThreads - 1) / kThreads);
        const double* K = nullptr;
        if (kind == LIN_U && c->tan_fresh) K = sp.Ku;
        else if (kind == LIN_PHI && sp.phi_fresh) K = sp.Kphi;
        else if (kind == LIN_MASS) K = sp.Km;
        if (K) {
            if (kind == LIN_U && c->dim == 3) spmv_u_tet_kernel<<<grid, kThreads, 0, c->stream>>>(sp.sv, K, v, y, mask, c->partial, result, c->counter, stop, pv);
            else if (kind == LIN_U) spmv_u_quad_kernel<<<grid, kThreads, 0, c->stream>>>(sp.sv, K, v, y, mask, c->partial, result, c->counter, stop, pv);
            else spmv_s_tet_kernel<<<grid, kThreads, 0, c->stream>>>(sp.sv, K, v, y, mask, c->partial, result, c->counter, stop, pv);
            c->launches++;
            CK(cudaGetLastError());
            return PFX_OK;
        }
    }
    KArgs a = base_args(c);
    a.v = v; a.y = y; a.mask = mask; a.result = result; a.stop = stop; a.input_masked = input_masked;
    a.pv = pv; a.wait_halo = wait_halo;
    int op = kind == LIN_U ? OP_APPLY_U : (kind == LIN_PHI ? OP_APPLY_PHI : OP_APPLY_MASS);
    return launch_scatter(c, op, a);
}

// tetrahedra: element tangents of the current (u, phi) and the matrix assembled from them, together with the
// Jacobi / block-Jacobi data of the PCG (the diagonal blocks fall out of the assembly)
int refresh_tangent(pfx_ctx* c) {
    if (!c->sp.enabled) return PFX_OK;
    const bool blk = c->coarse.enabled && c->coarse.block_jacobi;
    const int grid = (int)((c->sp.sv.n_rows + 127) / 128);
    if (c->cell_type == PFX_CELL_QUADRILATERAL || c->cell_type == PFX_CELL_TRIANGLE) {
        if (c->cell_type == PFX_CELL_TRIANGLE)
            assemble_u_tri_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->mat, c->X, c->u, c->phi, c->mask_u, c->sp.Ku, c->w_dinv,
                                                               blk ? c->coarse.dblk : nullptr);
        else
            assemble_u_quad_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->mat, c->prm.quad_points_1d, c->X, c->u, c->phi, c->mask_u, c->sp.Ku,
                                                                c->w_dinv, blk ? c->coarse.dblk : nullptr);
        c->launches++;
        CK(cudaGetLastError());
        c->tan_fresh = true;
        return PFX_OK;
    }
    KArgs t = base_args(c);
    int st = launch_scatter(c, OP_TANGENT_U, t);
    if (st) return st;
    assemble_u_tet_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->X, c->tan, c->mask_u, c->sp.Ku, c->w_dinv, blk ? c->coarse.dblk : nullptr);
    c->launches++;
    CK(cudaGetLastError());
    c->tan_fresh = true;
    return PFX_OK;
}

// phase-field coefficient fields of the current H / fatigue field; tetrahedra: J_phi assembled from them (and 1/diag)
int refresh_phi_operator(pfx_ctx* c) {
    c->phi_coeff_fresh = false;
    c->sp.phi_fresh = false;
    if (!c->cm || c->mat.dfun != 0) return PFX_OK;
    const int64_t nall = c->n_nodes;
    phi_coeff_kernel<<<vgrid(c, nall), kThreads, 0, c->stream>>>(nall, c->mat, c->H, c->Fat, c->cm, c->cg);
    c->launches++;
    c->phi_coeff_fresh = true;
    if (c->sp.enabled) {
        const int grid = (int)((c->sp.sv.n_rows + 127) / 128);
        if (c->cell_type == PFX_CELL_QUADRILATERAL)
            assemble_s_quad_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->prm.quad_points_1d, c->X, c->cm, c->cg, c->mask_phi, c->sp.Kphi, c->w_dinv);
        else if (c->cell_type == PFX_CELL_TRIANGLE)
            assemble_s_tri_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->X, c->cm, c->cg, c->mask_phi, c->sp.Kphi, c->w_dinv);
        else
            assemble_s_tet_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->X, c->cm, c->cg, c->mask_phi, c->sp.Kphi, c->w_dinv);
        c->launches++;
        c->sp.phi_fresh = true;
    }
    CK(cudaGetLastError());
    return PFX_OK;
}

// --------------------------------------------------------------------------- coarse space
// Aggregates = runs of consecutive blocks; distance-2 colouring of the aggregate graph for the
// probing set-up.  Host work, once per mesh.
int build_coarse(pfx_ctx* c, const pfx_mesh* mesh) {
    const BlockMesh& bm = c->bm;
    Coarse& cs = c->coarse;
    const char* nc_env = getenv("PFX_NO_COARSE");
    if ((nc_env && nc_env[0] == '1') || bm.nb < 32) return PFX_OK;
    const char* rf = getenv("PFX_COARSE_REFRESH");
    if (rf) cs.refresh_every_step = (rf[0] == 's');               // "step": every load step; default: when iterations grow
    const char* bj = getenv("PFX_BLOCK_JACOBI");
    if (bj && bj[0] == '0') cs.block_jacobi = false;
    const int D = c->dim, m_u = D == 2 ? 3 : 6;
    // aggregates of g consecutive blocks; regions of at most `reg_cap` coarse u-dofs (dense inverse per
    // region).  Defaults: at most kCoarseTotal coarse dofs in all, regions of <= kCoarseRegion.
    const char* eg = getenv("PFX_COARSE_AGG_BLOCKS");
    const char* er = getenv("PFX_COARSE_REGION");
    // 2D: many small aggregates (one per block up to kCoarseTotal dofs) in regions of kCoarseRegion — refining the
    // aggregates pays (iterations ~ sqrt(H/h)) and compact regions lose little.  3D: block aggregates are only
    // ~6 cells across, and cutting the coarse matrix into regions costs far more than coarser aggregates do
    // (beam, 8 M tets: 1 499 iterations per step with 7 regions, 576 with one region of 5 730 dofs), so the
    // whole coarse space is one region of at most kCoarseMax dofs.
    const char* et = getenv("PFX_COARSE_TOTAL");            // cap of the whole (per-rank) coarse space, coarse u-dofs
    const int total_cap = (et && atoi(et) >= m_u * 8) ? atoi(et) : (D == 3 ? kCoarseMax : kCoarseTotal);
    const int reg_want = er ? atoi(er) : (c->prm.coarse_region > 0 ? c->prm.coarse_region : (D == 3 ? kCoarseMax : kCoarseRegion));
    const int reg_cap = std::min(kCoarseMax, std::max(m_u * 8, reg_want));
    int g = (bm.nb * m_u + total_cap - 1) / total_cap;
    if (reg_cap >= kCoarseMax) g = std::max(g, (bm.nb * m_u + kCoarseMax - 1) / kCoarseMax);      // one region holds them all
    if (c->prm.coarse_agg_blocks > 0) g = c->prm.coarse_agg_blocks;
    if (eg && atoi(eg) > 0) g = atoi(eg);
    g = std::max(g, 1);
    const int reg_aggs = std::max(1, reg_cap / m_u);
    AggregateSet as;
    std::string er_ = build_aggregates(bm, mesh->n_cells, mesh->coords, mesh->cells, g, reg_aggs, kCoarseChunkNodes, as);
    if (!er_.empty()) { c->err = er_; return PFX_ERR_BAD_ARG; }
    const int n_agg = as.n_agg, n_reg = as.n_reg;
    cs.n_agg = n_agg; cs.n_reg = n_reg; cs.S = as.S; cs.n_chunks = as.n_chunks; cs.n_col = as.n_col;
    const std::vector<int32_t>&chunk0 = as.chunk_node0, &colour = as.colour, &nbrcol = as.nbrcol, &agg_reg = as.agg_reg,
                              &reg_agg0 = as.reg_agg0;
    const std::vector<float>& rel = as.rel;
    cs.h_cen = as.cen; cs.h_half = as.half;
    { const char* n3 = getenv("PFX_NO_LEVEL3"); cs.lvl3_wanted = !(n3 && n3[0] == '1'); }
    int st;
    if ((st = dev_upload(c, &cs.chunk_node0, chunk0))) return st;
    if ((st = dev_upload(c, &cs.colour, colour))) return st;
    if ((st = dev_upload(c, &cs.nbrcol, nbrcol))) return st;
    if ((st = dev_upload(c, &cs.rel, rel))) return st;
    if ((st = dev_upload(c, &cs.agg_reg, agg_reg))) return st;
    int64_t ac_doubles = 0;
    int nc_max = 0;
    for (int k = 0; k < 2; ++k) {
        CoarseOp& op = cs.op[k];
        op.m = k == 0 ? m_u : 1;
        std::vector<int2> ctas;
        int64_t off = 0, aoff = 0;
        for (int r = 0; r < n_reg; ++r) {
            const int na = reg_agg0[r + 1] - reg_agg0[r], ncr = na * op.m, ldr = (ncr + 3) / 4 * 4;
            op.h_reg.push_back(make_int4(reg_agg0[r], na, (int)off, ldr));
            op.h_ac_off.push_back(aoff);
            off += (int64_t)ncr * ldr;
            aoff += (int64_t)ncr * ncr;
            op.ld = std::max(op.ld, ldr);
            nc_max = std::max(nc_max, ncr);
            for (int row = 0; row < ncr; row += kCoarseRowsPerCta) ctas.push_back(make_int2(r, row));
        }
        if (off > (int64_t)0x7fffffff) { c->err = "coarse space: inverse too large"; return PFX_ERR_BAD_ARG; }
        op.ainv_floats = off;
        op.n_ctas = (int)ctas.size();
        ac_doubles = std::max(ac_doubles, aoff);
        if ((st = dev_upload(c, &op.reg_tab, op.h_reg))) return st;
        if ((st = dev_upload(c, &op.cta_tab, ctas))) return st;
        if ((st = dev_upload(c, &op.ac_off, op.h_ac_off))) return st;
        if ((st = dev_alloc(c, &op.Ainv, (size_t)off))) return st;
    }
    if ((st = dev_alloc(c, &cs.wpart, (size_t)cs.n_chunks * m_u))) return st;
    if ((st = dev_alloc(c, &cs.y, (size_t)n_agg * m_u))) return st;
    if ((st = dev_alloc(c, &cs.Ac, (size_t)ac_doubles))) return st;
    if ((st = dev_alloc(c, &cs.partial, (size_t)std::max(2 * cs.n_chunks, cs.op[0].n_ctas)))) return st;
    if ((st = dev_alloc(c, &cs.info, (size_t)1))) return st;
    if ((st = dev_alloc(c, &cs.dblk, (size_t)c->n_nodes * (D * (D + 1) / 2)))) return st;
    // work space of the dense inverses (pfx_dense.cuh): Y (largest region squared), a 64-wide panel, the diagonal blocks
    if ((st = dev_alloc(c, &cs.work, (size_t)nc_max * nc_max))) return st;
    if ((st = dev_alloc(c, &cs.workT, (size_t)nc_max * kDenseNB))) return st;
    if ((st = dev_alloc(c, &cs.workL, (size_t)((nc_max + kDenseNB - 1) / kDenseNB) * kDenseNB * kDenseNB))) return st;
    {   // the coarse solve stages w of one region (up to kCoarseMax doubles) in dynamic shared memory
        const int smem = kCoarseMax * (int)sizeof(double);
        CK(cudaFuncSetAttribute(coarse_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(coarse_solve_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(coarse_solve_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    cs.enabled = true;
    return PFX_OK;
}

CoarseView coarse_view(const pfx_ctx* c, int kind) {
    const Coarse& cs = c->coarse;
    CoarseView v;
    v.n_agg = cs.n_agg; v.n_chunks = cs.n_chunks; v.S = cs.S; v.ld = cs.op[kind].ld;
    v.chunk_node0 = cs.chunk_node0; v.rel = cs.rel; v.wpart = cs.wpart; v.y = cs.y; v.Ainv = cs.op[kind].Ainv;
    v.agg_reg = cs.agg_reg; v.reg_tab = cs.op[kind].reg_tab; v.cta_tab = cs.op[kind].cta_tab; v.ac_off = cs.op[kind].ac_off;
    return v;
}

bool coarse_active(const pfx_ctx* c, int kind) {
    return c->coarse.enabled && kind != LIN_MASS;
}

template <int NC, int GD>
int coarse_setup_t(pfx_ctx* c, int kind, const uint8_t* mask) {
    constexpr int M = Modes<NC, GD>::M;
    Coarse& cs = c->coarse;
    CoarseOp& op = cs.op[kind];
    CoarseView cv = coarse_view(c, kind);
    const int nc = cs.n_agg * M;
    auto now = [&]() { cudaStreamSynchronize(c->stream); return std::chrono::steady_clock::now(); };
    std::chrono::steady_clock::time_point t0, t1, t2;
    if (c->verbose) t0 = now();
    const long long ac_total = op.h_ac_off.back() + (long long)op.h_reg.back().y * M * op.h_reg.back().y * M;
    CK(cudaMemsetAsync(cs.Ac, 0, (size_t)ac_total * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->w_p, 0, (size_t)c->n_nodes * NC * sizeof(double), c->stream));
    for (int col = 0; col < cs.n_col; ++col)
        for (int mode = 0; mode < M; ++mode) {
            coarse_probe_kernel<NC, GD><<<cs.n_chunks, kThreads, 0, c->stream>>>(cv, cs.colour, col, mode, mask, c->w_p);
            int st = apply_lin(c, kind, c->w_p, c->w_Ap, mask, c->red, nullptr, 1);
            if (st) return st;
            coarse_restrict_kernel<NC, GD><<<cs.n_chunks, kThreads, 0, c->stream>>>(cv, c->w_Ap);
            coarse_fill_kernel<M><<<(nc + 127) / 128, 128, 0, c->stream>>>(cv, cs.nbrcol, cs.n_col, col, mode, cs.Ac);
            c->launches += 3;
        }
    CK(cudaGetLastError());
    if (c->verbose) t1 = now();
    op.valid = false;
    // factorise and invert region by region; the status words are collected and read once
    std::vector<int> h_info(2 * cs.n_reg, 0);
    bool ok = true;
    for (int r = 0; r < cs.n_reg && ok; ++r) {
        const int ncr = op.h_reg[r].y * M;
        double* A = cs.Ac + op.h_ac_off[r];
        coarse_fixdiag_kernel<<<(ncr + 127) / 128, 128, 0, c->stream>>>(ncr, A, ncr);
        CK(cudaMemsetAsync(cs.info, 0, sizeof(int), c->stream));
        c->launches += dense_spd_inverse(c->stream, ncr, A, cs.work, cs.workT, cs.workL, cs.info);
        CK(cudaMemcpyAsync(&h_info[2 * r], cs.info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        dim3 grid((op.h_reg[r].w + 127) / 128, ncr);
        coarse_convert_kernel<<<grid, 128, 0, c->stream>>>(ncr, A, ncr, op.Ainv + op.h_reg[r].z, op.h_reg[r].w);
        c->launches += 2;
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    for (int v : h_info) ok = ok && v == 0;
    if (c->n_ranks > 1) {          // every rank must take the same PCG path: all or none
        double bad = ok ? 0.0 : 1.0, tot = 0.0;
        CK(cudaMemcpyAsync(c->S + S_STAGE, &bad, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        int st = fetch(c, c->S + S_STAGE, 1, &tot);
        if (st) return st;
        ok = tot == 0.0;
    }
    op.valid = ok;
    if (!ok && c->verbose) fprintf(stderr, "[pfx] coarse matrix of operator %d is not positive definite: Jacobi only\n", kind);
    if (c->verbose) {
        t2 = now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[pfx] coarse set-up op %d: %d aggregates x %d modes in %d regions, %d colours: probing %.2f ms, inverses %.2f ms\n",
                kind, cs.n_agg, M, cs.n_reg, cs.n_col, ms(t0, t1), ms(t1, t2));
    }
    op.stale = false;
    op.ref_its = -1;
    op.ref_open = true;
    op.open_steps = 0;
    cs.setups++;
    return PFX_OK;
}

// Third, rank-level coarse space of J_u (pfx_coarse.cuh: Coarse3View): T of this rank's aggregates, E by probing with
// halo exchange (n_ranks * M applies), inverse on the host.  Collective: every rank runs the same sequence.
template <int NC, int GD>
int coarse3_setup_t(pfx_ctx* c, const uint8_t* mask) {
    constexpr int M = Modes<NC, GD>::M;
    Coarse& cs = c->coarse;
    const int nR = c->n_ranks, ng = nR * M, n_agg = cs.n_agg;
    int st;
    if (!cs.lvl3_ready) {
        // rank centroid / half extent from the aggregates; T_a = [[I, C(q_a)], [0, (s_a / s_R) I]], q_a = (c_a - c_R) / s_R
        double cR[3] = {0, 0, 0}, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int a = 0; a < n_agg; ++a)
            for (int k = 0; k < GD; ++k) {
                const double x = cs.h_cen[(size_t)a * 3 + k];
                cR[k] += x / n_agg; lo[k] = std::min(lo[k], x - cs.h_half[a]); hi[k] = std::max(hi[k], x + cs.h_half[a]);
            }
        double sR = 0.0;
        for (int k = 0; k < GD; ++k) sR = std::max(sR, 0.5 * (hi[k] - lo[k]));
        if (!(sR > 0.0)) sR = 1.0;
        std::vector<double> T((size_t)n_agg * M * M, 0.0);
        for (int a = 0; a < n_agg; ++a) {
            double q[3] = {0, 0, 0};
            for (int k = 0; k < GD; ++k) q[k] = (cs.h_cen[(size_t)a * 3 + k] - cR[k]) / sR;
            const double sc = cs.h_half[a] / sR;
            double* Ta = T.data() + (size_t)a * M * M;
            auto at = [&](int i, int j) -> double& { return Ta[i * M + j]; };
            if (GD == 2) {
                at(0, 0) = 1.0; at(1, 1) = 1.0; at(0, 2) = -q[1]; at(1, 2) = q[0]; at(2, 2) = sc;
            } else {
                for (int k = 0; k < 3; ++k) { at(k, k) = 1.0; at(3 + k, 3 + k) = sc; }
                // translation part of omega x q: (w_y q_z - w_z q_y, w_z q_x - w_x q_z, w_x q_y - w_y q_x)
                at(0, 4) = q[2];  at(0, 5) = -q[1];
                at(1, 3) = -q[2]; at(1, 5) = q[0];
                at(2, 3) = q[1];  at(2, 4) = -q[0];
            }
        }
        if ((st = dev_upload(c, &cs.T3, T))) return st;
        if ((st = dev_alloc(c, &cs.E3, (size_t)ng * ng))) return st;
        if ((st = dev_alloc(c, &cs.Einv3, (size_t)ng * ng))) return st;
        cs.lvl3_ready = true;
    }
    CoarseView cv = coarse_view(c, LIN_U);
    CK(cudaMemsetAsync(cs.E3, 0, (size_t)ng * ng * sizeof(double), c->stream));
    for (int r = 0; r < nR; ++r)
        for (int m = 0; m < M; ++m) {
            CK(cudaMemsetAsync(c->w_p, 0, (size_t)c->n_nodes * NC * sizeof(double), c->stream));
            coarse3_probe_kernel<NC, GD><<<cs.n_chunks, kThreads, 0, c->stream>>>(cv, cs.T3, r == c->rank ? 1 : 0, m, mask, c->w_p);
            c->launches++;
            if ((st = exchange(c, c->w_p, NC))) return st;
            if ((st = apply_lin(c, LIN_U, c->w_p, c->w_Ap, mask, c->red, nullptr, 1))) return st;
            coarse_restrict_kernel<NC, GD><<<cs.n_chunks, kThreads, 0, c->stream>>>(cv, c->w_Ap);
            coarse3_reduce_kernel<M><<<1, kThreads, 0, c->stream>>>(cv, cs.T3, c->S, 0, c->d_pv);
            coarse3_collect_kernel<M><<<1, 32, 0, c->stream>>>(c->d_pv, r * M + m, cs.E3);
            c->launches += 3;
        }
    CK(cudaGetLastError());
    std::vector<double> E((size_t)ng * ng), Ei((size_t)ng * ng, 0.0);
    CK(cudaMemcpyAsync(E.data(), cs.E3, E.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // symmetrise, decouple empty rows, Cholesky (E must be positive definite), inverse by two triangular solves per column
    bool ok = true;
    for (int i = 0; i < ng; ++i)
        for (int j = 0; j < i; ++j) { const double v = 0.5 * (E[(size_t)i * ng + j] + E[(size_t)j * ng + i]); E[(size_t)i * ng + j] = E[(size_t)j * ng + i] = v; }
    for (int i = 0; i < ng; ++i)
        if (!(E[(size_t)i * ng + i] > 0.0)) {
            for (int j = 0; j < ng; ++j) E[(size_t)i * ng + j] = E[(size_t)j * ng + i] = 0.0;
            E[(size_t)i * ng + i] = 1.0;
        }
    std::vector<double> Lc(E);
    for (int j = 0; j < ng && ok; ++j) {
        double d = Lc[(size_t)j * ng + j];
        for (int k = 0; k < j; ++k) d -= Lc[(size_t)j * ng + k] * Lc[(size_t)j * ng + k];
        if (!(d > 0.0) || !std::isfinite(d)) { ok = false; break; }
        d = std::sqrt(d);
        Lc[(size_t)j * ng + j] = d;
        for (int i = j + 1; i < ng; ++i) {
            double v = Lc[(size_t)i * ng + j];
            for (int k = 0; k < j; ++k) v -= Lc[(size_t)i * ng + k] * Lc[(size_t)j * ng + k];
            Lc[(size_t)i * ng + j] = v / d;
        }
    }
    if (ok) {
        std::vector<double> y(ng);
        for (int col = 0; col < ng; ++col) {
            for (int i = 0; i < ng; ++i) {                       // L y = e_col
                double v = (i == col) ? 1.0 : 0.0;
                for (int k = 0; k < i; ++k) v -= Lc[(size_t)i * ng + k] * y[k];
                y[i] = v / Lc[(size_t)i * ng + i];
            }
            for (int i = ng - 1; i >= 0; --i) {                  // L^T x = y
                double v = y[i];
                for (int k = i + 1; k < ng; ++k) v -= Lc[(size_t)k * ng + i] * Ei[(size_t)k * ng + col];
                Ei[(size_t)i * ng + col] = v / Lc[(size_t)i * ng + i];
            }
        }
        CK(cudaMemcpyAsync(cs.Einv3, Ei.data(), Ei.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if (c->verbose) fprintf(stderr, "[pfx] rank-level coarse space: %d dofs, %s\n", ng, ok ? "on" : "matrix not positive definite: off");
    if (ok != cs.lvl3_on && cs.op[LIN_U].graph) {        // the captured iteration changes with the third level
        cudaGraphExecDestroy(cs.op[LIN_U].graph);
        cs.op[LIN_U].graph = nullptr;
    }
    cs.lvl3_on = ok;
    return PFX_OK;
}

int coarse_setup(pfx_ctx* c, int kind, const uint8_t* mask) {
    if (kind == LIN_PHI) return c->dim == 2 ? coarse_setup_t<1, 2>(c, kind, mask) : coarse_setup_t<1, 3>(c, kind, mask);
    int st = c->dim == 2 ? coarse_setup_t<2, 2>(c, kind, mask) : coarse_setup_t<3, 3>(c, kind, mask);
    if (st) return st;
    Coarse& cs = c->coarse;
    if (c->n_ranks > 1 && c->p2p && c->p2p_fused && cs.lvl3_wanted && cs.op[LIN_U].valid)
        return c->dim == 2 ? coarse3_setup_t<2, 2>(c, mask) : coarse3_setup_t<3, 3>(c, mask);
    return PFX_OK;
}

// tail of one PCG iteration / start of a solve (init) with the coarse space: update, coarse solve, direction
// fused != 0 (multi-GPU NVLink path): the three kernels do their all-reduces through the peers' mailboxes
template <int NC, int GD, bool BLK>
int coarse_tail_t(pfx_ctx* c, int kind, int par, int init, const uint8_t* mask, const double* dinv, double rtol2, int bref_slot,
                  int fused) {
    constexpr int M = Modes<NC, GD>::M;
    Coarse& cs = c->coarse;
    CoarseView cv = coarse_view(c, kind);
    const bool multi = c->n_ranks > 1;
    const PeerView* pv = (fused && !init) ? c->d_pv : nullptr;
    int st;
    if (BLK) dinv = cs.dblk;
    const int vgrid_c = (cs.n_chunks + kCoarseChunksPerCta - 1) / kCoarseChunksPerCta;
    pcg_update_coarse_kernel<NC, GD, BLK><<<vgrid_c, kThreads, 0, c->stream>>>(cv, c->w_x, c->w_r, c->w_p, c->w_Ap, dinv, c->S, par, init,
                                                                                  cs.partial, c->counter, pv);
    const int solve_ctas = cs.op[kind].n_ctas;
    const int mode = init ? 3 : ((multi && !pv) ? 2 : 0);
    Coarse3View c3;
    if (kind == LIN_U && cs.lvl3_on && fused) {      // rank-level coarse space: one more mailbox hop of M numbers
        c3.pv = c->d_pv; c3.T = cs.T3; c3.Einv = cs.Einv3;
        coarse3_reduce_kernel<M><<<1, kThreads, 0, c->stream>>>(cv, cs.T3, c->S, init ? 0 : 1, c->d_pv);
        c->launches++;
    }
    coarse_solve_kernel<M><<<solve_ctas, kThreads, (size_t)cv.ld * sizeof(double), c->stream>>>(cv, c->S, par, mode, cs.partial,
                                                                                              c->counter, pv, c3);
    c->launches += 2;
    if (init) {
        if ((st = allreduce(c, c->S + S_STAGE, 2))) return st;
        pcg_setup_kernel<<<1, 1, 0, c->stream>>>(c->S, rtol2, bref_slot);
        c->launches++;
    } else if (multi && !pv) {
        if (c->p2p) {
            p2p_allreduce_kernel<<<1, 32, 0, c->stream>>>(c->pv, c->S, c->S + S_STAGE, 2, 1, par);
        } else {
            if ((st = allreduce(c, c->S + S_STAGE, 2))) return st;
            pcg_publish_kernel<<<1, 1, 0, c->stream>>>(c->S, par);
        }
        c->launches++;
    }
    pcg_direction_coarse_kernel<NC, GD, BLK><<<cs.n_chunks, kThreads, 0, c->stream>>>(cv, c->w_p, c->w_r, dinv, mask, c->S, par, init, pv);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

int coarse_tail(pfx_ctx* c, int kind, int par, int init, const uint8_t* mask, const double* dinv, double rtol2 = 0.0,
                int bref_slot = -1, int fused = 0) {
    if (kind == LIN_PHI)
        return c->dim == 2 ? coarse_tail_t<1, 2, false>(c, kind, par, init, mask, dinv, rtol2, bref_slot, fused)
                           : coarse_tail_t<1, 3, false>(c, kind, par, init, mask, dinv, rtol2, bref_slot, fused);
    if (!c->coarse.block_jacobi)
        return c->dim == 2 ? coarse_tail_t<2, 2, false>(c, kind, par, init, mask, dinv, rtol2, bref_slot, fused)
                           : coarse_tail_t<3, 3, false>(c, kind, par, init, mask, dinv, rtol2, bref_slot, fused);
    return c->dim == 2 ? coarse_tail_t<2, 2, true>(c, kind, par, init, mask, dinv, rtol2, bref_slot, fused)
                       : coarse_tail_t<3, 3, true>(c, kind, par, init, mask, dinv, rtol2, bref_slot, fused);
}

// one PCG iteration with the coarse space.  Multi-GPU: every rank keeps the coarse space of its own
// aggregates (block-diagonal coarse matrix, no coarse-level communication); the scalars are reduced
// exactly as on the Jacobi path.
int pcg_iteration_coarse(pfx_ctx* c, int kind, int par, const uint8_t* mask, const double* dinv) {
    const int nc = (kind == LIN_U) ? c->dim : 1;
    const bool multi = c->n_ranks > 1;
    int st;
    if (multi && c->p2p && c->p2p_fused) {
        const bool cw = apply_waits_for_halo(c, kind);
        if ((st = exchange(c, c->w_p, nc, c->S + S_DONE, cw))) return st;
        if ((st = apply_lin(c, kind, c->w_p, c->w_Ap, mask, c->S + S_STAGE, c->S + S_DONE, 1, c->d_pv, cw ? 1 : 0))) return st;
        return coarse_tail(c, kind, par, 0, mask, dinv, 0.0, -1, 1);
    }
    if ((st = exchange(c, c->w_p, nc, c->S + S_DONE))) return st;
    if ((st = apply_lin(c, kind, c->w_p, c->w_Ap, mask, multi ? c->S + S_STAGE : c->S + S_PAP, c->S + S_DONE, 1))) return st;
    if (multi && c->p2p) {
        p2p_allreduce_kernel<<<1, 32, 0, c->stream>>>(c->pv, c->S, c->S + S_STAGE, 1, 0, par);
        c->launches++;
    } else if (multi) {
        if ((st = allreduce(c, c->S + S_STAGE, 1))) return st;
        axpby_kernel<<<1, 1, 0, c->stream>>>(1, c->S + S_PAP, c->S + S_STAGE, 1.0, nullptr, 0.0);
        c->launches++;
    }
    return coarse_tail(c, kind, par, 0, mask, dinv);
}

// one PCG iteration (parity par) enqueued on the stream
int pcg_iteration(pfx_ctx* c, int kind, int par, int64_t n, const uint8_t* mask, const double* dinv) {
    int nc = (kind == LIN_U) ? c->dim : 1;
    bool multi = c->n_ranks > 1;
    if (multi && c->p2p && c->p2p_fused) {
        // fused NVLink path: 4 launches, no collective call — halo push | apply (waits for the halo in its
        // prologue, pushes p.Ap to the peers' mailboxes in its epilogue) | update (sums p.Ap from the
        // mailboxes, pushes r.z, r.r) | direction (sums them, publishes the PCG state)
        const bool cw = apply_waits_for_halo(c, kind);
        int st = exchange(c, c->w_p, nc, c->S + S_DONE, cw);
        if (st) return st;
        st = apply_lin(c, kind, c->w_p, c->w_Ap, mask, c->S + S_STAGE, c->S + S_DONE, 1, c->d_pv, cw ? 1 : 0);
        if (st) return st;
        int g = vgrid(c, n);
        pcg_update_kernel<<<g, kThreads, 0, c->stream>>>(n, c->w_x, c->w_r, c->w_p, c->w_Ap, dinv, c->S, par, 0, c->partial,
                                                        c->counter, c->d_pv);
        pcg_direction_kernel<<<g, kThreads, 0, c->stream>>>(n, c->w_p, c->w_r, dinv, c->S, par, c->d_pv);
        c->launches += 2;
        CK(cudaGetLastError());
        return PFX_OK;
    }
    int st = exchange(c, c->w_p, nc, c->S + S_DONE);
    if (st) return st;
    st = apply_lin(c, kind, c->w_p, c->w_Ap, mask, multi ? c->S + S_STAGE : c->S + S_PAP, c->S + S_DONE, 1);
    if (st) return st;
    if (multi && c->p2p) {
        p2p_allreduce_kernel<<<1, 32, 0, c->stream>>>(c->pv, c->S, c->S + S_STAGE, 1, 0, par);
        c->launches++;
    } else if (multi) {
        // skipped launches leave the stage untouched; the guarded copy below keeps S_PAP sane
        st = allreduce(c, c->S + S_STAGE, 1);
        if (st) return st;
        axpby_kernel<<<1, 1, 0, c->stream>>>(1, c->S + S_PAP, c->S + S_STAGE, 1.0, nullptr, 0.0);
        c->launches++;
    }
    int g = vgrid(c, n);
    pcg_update_kernel<<<g, kThreads, 0, c->stream>>>(n, c->w_x, c->w_r, c->w_p, c->w_Ap, dinv, c->S, par, multi ? 0 : 1,
                                                    c->partial, c->counter);
    c->launches++;
    if (multi && c->p2p) {
        p2p_allreduce_kernel<<<1, 32, 0, c->stream>>>(c->pv, c->S, c->S + S_STAGE, 2, 1, par);
        c->launches++;
    } else if (multi) {
        st = allreduce(c, c->S + S_STAGE, 2);
        if (st) return st;
        pcg_publish_kernel<<<1, 1, 0, c->stream>>>(c->S, par);
        c->launches++;
    }
    pcg_direction_kernel<<<g, kThreads, 0, c->stream>>>(n, c->w_p, c->w_r, dinv, c->S, par);
    c->launches++;
    CK(cudaGetLastError());
    return PFX_OK;
}

int build_graph(pfx_ctx* c, int kind, int64_t n, const uint8_t* mask, const double* dinv, bool coarse_on = false) {
    int iters = c->graph_iters;
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int st = PFX_OK;
    for (int k = 0; k < iters && st == PFX_OK; ++k)
        st = coarse_on ? pcg_iteration_coarse(c, kind, k & 1, mask, dinv) : pcg_iteration(c, kind, k & 1, n, mask, dinv);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (st) return st;
    CK(e);
    CK(cudaGraphInstantiate(coarse_on ? &c->coarse.op[kind].graph : &c->graph[kind], g, 0));
    CK(cudaGraphDestroy(g));
    return PFX_OK;
}

// Solve A x = b (A = J_u | J_phi | M with Dirichlet rows/cols = identity), Jacobi preconditioned.
// b in w_F (Dirichlet entries ignored), solution in w_x.  Returns iterations through `iters`.
int pcg_solve(pfx_ctx* c, int kind, double rtol, int64_t* iters, bool* converged) {
    int nc = (kind == LIN_U) ? c->dim : 1;
    int64_t n = c->n_owned * nc;
    const uint8_t* mask = kind == LIN_U ? c->mask_u : (kind == LIN_PHI ? c->mask_phi : nullptr);
    const double* dinv = kind == LIN_MASS ? c->dinv_mass : c->w_dinv;
    *converged = false;
    *iters = 0;
    if (c->mat.dfun == 0 && (kind == LIN_U || (kind == LIN_PHI && c->cm && c->phi_coeff_fresh)) && c->n_ranks == 1 &&
        c->coop_capacity >= c->bm.nb && c->coop_bar) {
        mark(c, kind == LIN_U ? PH_U_PCG : PH_PHI_PCG);
        // small mesh: the whole solve in one cooperative launch (re-launched only if max_iters runs out)
        KArgs a = base_args(c);
        a.mask = mask;
        CoopArgs ca;
        ca.F = c->w_F; ca.dinv = dinv; ca.x = c->w_x; ca.r = c->w_r; ca.p = c->w_p; ca.partial = c->partial;
        ca.bar = c->coop_bar; ca.S = c->S; ca.rtol2 = rtol * rtol;
        ca.bref_slot = c->cg_adaptive ? S_BREF + kind : -1;
        ca.resume = 0;
        for (;;) {
            ca.max_iters = (int)std::min<int64_t>(4096, (int64_t)c->prm.cg_max_it - *iters);
            CK(cudaMemsetAsync(c->coop_bar, 0, sizeof(unsigned long long), c->stream));
            int st = (kind == LIN_U) ? coop_u(c, a, ca, false) : coop_phi(c, a, ca, false);
            if (st) return st;
            CK(cudaMemcpyAsync(c->h_pinned, c->S, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            *iters = (int64_t)c->h_pinned[S_ITERS];
            double rr = c->h_pinned[S_RR];
            if (!(rr == rr)) { c->err = "NaN in PCG"; return PFX_ERR_NAN; }
            if (c->h_pinned[S_DONE] != 0.0) { *converged = true; break; }
            if (*iters >= c->prm.cg_max_it) break;
            ca.resume = 1;
        }
        if (*converged && *iters == 0 && ca.bref_slot >= 0 && c->h_pinned[S_RR] > 0.0) c->floor_solves++;
        return PFX_OK;
    }
    int g = vgrid(c, n);
    int st;
    bool coarse_on = false;
    if (coarse_active(c, kind)) {
        if (c->coarse.op[kind].stale) {
            mark(c, kind == LIN_U ? PH_U_COARSE : PH_PHI_COARSE);
            if ((st = coarse_setup(c, kind, mask))) return st;
        }
        coarse_on = c->coarse.op[kind].valid;
    }
    if (kind != LIN_MASS) mark(c, kind == LIN_U ? PH_U_PCG : PH_PHI_PCG);
    const int bref_slot = (c->cg_adaptive && kind != LIN_MASS) ? S_BREF + kind : -1;
    pcg_init_kernel<<<g, kThreads, 0, c->stream>>>(n, c->w_F, mask, dinv, nullptr, nullptr, c->w_x, c->w_r, c->w_p,
                                                  c->partial, c->S + S_STAGE, c->counter);
    c->launches++;
    if (coarse_on) {
        // z0 = (D^-1 + Z Ainv Z^T) r0: the init pass of the coarse kernels redoes the sums and the direction
        if ((st = coarse_tail(c, kind, 0, 1, mask, dinv, rtol * rtol, bref_slot, (c->n_ranks > 1 && c->p2p && c->p2p_fused) ? 1 : 0))) return st;
    } else {
        if ((st = allreduce(c, c->S + S_STAGE, 2))) return st;
        pcg_setup_kernel<<<1, 1, 0, c->stream>>>(c->S, rtol * rtol, bref_slot);
        c->launches++;
    }
    CK(cudaGetLastError());
    int64_t max_it = c->prm.cg_max_it;
    for (;;) {
        if (c->use_graph) {
            cudaGraphExec_t& gx = coarse_on ? c->coarse.op[kind].graph : c->graph[kind];
            if (!gx) {
                st = build_graph(c, kind, n, mask, dinv, coarse_on);
                if (st) return st;
            }
            CK(cudaGraphLaunch(gx, c->stream));
            c->launches += (int64_t)c->graph_iters * (coarse_on ? 4 : 3);
        } else {
            for (int k = 0; k < c->graph_iters; ++k) {
                st = coarse_on ? pcg_iteration_coarse(c, kind, k & 1, mask, dinv) : pcg_iteration(c, kind, k & 1, n, mask, dinv);
                if (st) return st;
            }
        }
        CK(cudaMemcpyAsync(c->h_pinned, c->S, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        *iters = (int64_t)c->h_pinned[S_ITERS];
        double rr = c->h_pinned[S_RR];
        if (c->h_pinned[S_ERR] != 0.0) { c->err = "peer wait timed out (halo/all-reduce)"; return PFX_ERR_COMM; }
        if (!(rr == rr)) { c->err = "NaN in PCG"; return PFX_ERR_NAN; }
        if (c->h_pinned[S_DONE] != 0.0) { *converged = true; break; }
        if (*iters >= max_it) break;
    }
    if (*converged && *iters == 0 && bref_slot >= 0 && c->h_pinned[S_RR] > 0.0) c->floor_solves++;
    if (coarse_on) {                 // lagged coarse matrix: refresh when it stops paying
        CoarseOp& op = c->coarse.op[kind];
        if (op.ref_open) op.ref_its = std::max(op.ref_its, *iters);
        else if (*iters > op.ref_its * 3 / 2 + 10) op.stale = true;
    }
    return PFX_OK;
}

// inverse nodal diagonal blocks of the current J_u (smoother of the two-level PCG)
int block_diag(pfx_ctx* c) {
    if (!coarse_active(c, LIN_U) || !c->coarse.block_jacobi || c->sp.enabled) return PFX_OK;
    KArgs d = base_args(c);
    d.y = c->coarse.dblk; d.mask = c->mask_u;
    return launch_scatter(c, OP_DIAGBLK_U, d);
}

// --------------------------------------------------------------------------- SNES-like Newton
struct NewtonOut {
    int its = 0, reason = 0;
    double fnorm = 0;
    int64_t cg_its = 0;
};

// is_u: displacement problem (solver.py:198-205) else phase-field problem (:237-244)
int newton(pfx_ctx* c, bool is_u, NewtonOut* out) {
    const int nc = is_u ? c->dim : 1;
    const int64_t n = c->n_owned * nc;
    double* x = is_u ? c->u : c->phi;
    const double* g = is_u ? c->g_u : c->g_phi;
    const uint8_t* mask = is_u ? c->mask_u : c->mask_phi;
    const int kind = is_u ? LIN_U : LIN_PHI;
    const int gr = vgrid(c, n);
    double ttol = 0, snorm = 0, xnorm = 0, vals[4];
    int st;
    c->phi_coeff_fresh = false;
    c->sp.phi_fresh = false;
    if (!is_u && (st = refresh_phi_operator(c))) return st;    // H and the fatigue field are fixed during the phi solve
    const bool sparse_op = c->sp.enabled && (is_u || c->sp.phi_fresh);   // operator + Jacobi data come from the assembly
    for (int it = 0;; ++it) {
        mark(c, is_u ? PH_U_SETUP : PH_PHI_SETUP);
        // residual with lifting (dolfinx assemble_residual idiom)
        KArgs a = base_args(c);
        a.y = c->w_t0;
        st = launch_scatter(c, is_u ? OP_RESIDUAL_U : OP_RESIDUAL_PHI, a);
        if (st) return st;
        const double* lift = nullptr;
        if (it == 0) {
            // gap on every local dof (ghost entries feed the halo elements of the apply); no gap (every stagger
            // iteration after the first of a load step) => no lifting term and no linearisation needed yet
            int64_t nall = c->n_nodes * nc;
            bc_gap_kernel<<<vgrid(c, nall), kThreads, 0, c->stream>>>(nall, x, g, mask, c->w_p, c->partial, c->red, c->counter);
            c->launches++;
            double gap2 = 0.0;
            if ((st = fetch(c, c->red, 1, &gap2))) return st;
            if (gap2 != 0.0) {
                if (is_u && (st = refresh_tangent(c))) return st;
                st = apply_lin(c, kind, c->w_p, c->w_t1, nullptr, c->red, nullptr);
                if (st) return st;
                lift = c->w_t1;
            }
        }
        snes_residual_kernel<<<gr, kThreads, 0, c->stream>>>(n, c->w_t0, lift, (is_u && c->has_fext) ? c->fext : nullptr, x, g,
                                                            mask, c->w_F, c->partial, c->red, c->counter);
        c->launches++;
        CK(cudaGetLastError());
        st = fetch(c, c->red, 2, vals);
        if (st) return st;
        double fnorm = std::sqrt(vals[0]);
        xnorm = std::sqrt(vals[1]);
        out->fnorm = fnorm;
        out->its = it;
        if (!(fnorm == fnorm) || std::isinf(fnorm)) { out->reason = -4; return PFX_OK; }
        if (it == 0) ttol = fnorm * c->prm.snes_rtol;
        if (fnorm < c->prm.snes_atol) { out->reason = 2; return PFX_OK; }
        if (it > 0) {
            if (fnorm <= ttol) { out->reason = 3; return PFX_OK; }
            if (snorm < c->prm.snes_stol * xnorm) { out->reason = 4; return PFX_OK; }
        }
        if (it >= c->prm.snes_max_it) { out->reason = -5; return PFX_OK; }
        if (sparse_op) {
            if (is_u && !c->tan_fresh && (st = refresh_tangent(c))) return st;
        } else {
            // Jacobi data of the current tangent
            KArgs d = base_args(c);
            d.y = c->w_dinv; d.mask = mask;
            st = launch_scatter(c, is_u ? OP_DIAG_U : OP_DIAG_PHI, d);
            if (st) return st;
            if (is_u && (st = block_diag(c))) return st;
        }
        int64_t cg = 0; bool conv = false;
        std::chrono::steady_clock::time_point tp0;
        if (c->verbose) { cudaStreamSynchronize(c->stream); tp0 = std::chrono::steady_clock::now(); }
        st = pcg_solve(c, kind, c->prm.cg_rtol, &cg, &conv);
        if (st) return st;
        out->cg_its += cg;
        if (c->verbose) {
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
            fprintf(stderr, "[pfx] %s newton it %d: |F| %.3e  pcg its %lld  |r|/|b| %.2e  %.2f ms (%.1f us/it)\n", is_u ? "u  " : "phi",
                    it, fnorm, (long long)cg, std::sqrt(c->h_pinned[S_RR]) / std::max(fnorm, 1e-300), ms, 1e3 * ms / std::max<int64_t>(cg, 1));
        }
        mark(c, is_u ? PH_U_SETUP : PH_PHI_SETUP);
        if (!conv) { out->reason = -3; return PFX_OK; }      // linear solve failed (KSP analogue)
        newton_update_kernel<<<gr, kThreads, 0, c->stream>>>(n, x, c->w_x, c->w_F, mask, c->partial, c->red, c->counter);
        c->launches++;
        c->tan_fresh = false;
        CK(cudaGetLastError());
        st = fetch(c, c->red, 1, vals);
        if (st) return st;
        snorm = std::sqrt(vals[0]);
        st = exchange(c, x, nc);
        if (st) return st;
    }
}

// L2 projection: solve M h = b (b already in w_F), result copied to `dst` (Math/projection.py:19-59)
int project_solve(pfx_ctx* c, double* dst, int64_t* cg_its) {
    bool conv = false; int64_t it = 0;
    int st = pcg_solve(c, LIN_MASS, c->prm.proj_rtol, &it, &conv);
    if (st) return st;
    *cg_its += it;
    if (!conv) { c->err = "projection PCG did not converge"; return PFX_ERR_NOT_CONVERGED_PROJ; }
    CK(cudaMemcpyAsync(dst, c->w_x, c->n_owned * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return exchange(c, dst, 1);
}

int l2_error(pfx_ctx* c, bool is_u, const double* a_, const double* b_, double* err) {
    KArgs a = base_args(c);
    a.f1 = a_; a.f2 = b_;
    int st = is_u ? launch_reduce<RED_L2_U>(c, a) : launch_reduce<RED_L2_PHI>(c, a);
    if (st) return st;
    double v[2];
    st = fetch(c, c->red, 2, v);
    if (st) return st;
    double num = std::sqrt(v[0]), den = std::sqrt(v[1]);
    *err = (den == 0.0) ? 0.0 : num / den;
    return PFX_OK;
}

const double* field_ptr(pfx_ctx* c, int f, int* nc) {
    *nc = 1;
    switch (f) {
        case PFX_FIELD_U: *nc = c->dim; return c->u;
        case PFX_FIELD_U_OLD: *nc = c->dim; return c->u_old;
        case PFX_FIELD_PHI: return c->phi;
        case PFX_FIELD_PHI_OLD: return c->phi_old;
        case PFX_FIELD_H: return c->H;
        case PFX_FIELD_VC: return c->Vc;
        case PFX_FIELD_VN: return c->Vn;
        case PFX_FIELD_FATIGUE: return c->Fat;
        case PFX_FIELD_ALPHA_BAR: return c->abar_c;
        case PFX_FIELD_ALPHA_BAR_N: return c->abar_n;
        case PFX_FIELD_ALPHA_C: return c->alpha_c;
        case PFX_FIELD_ALPHA_N: return c->alpha_n;
    }
    return nullptr;
}

// device buffers of a redefined BC / traction set are released before the new ones are made
template <class T>
void dev_release(pfx_ctx* c, T** p) {
    if (!*p) return;
    auto it = std::find(c->allocs.begin(), c->allocs.end(), (void*)*p);
    if (it != c->allocs.end()) c->allocs.erase(it);
    cudaFree(*p);
    *p = nullptr;
}

int set_bc_common(pfx_ctx* c, bool is_u, int bc_id, const int32_t* dofs, int64_t n) {
    if (!c || bc_id < 0 || n < 0 || (n > 0 && !dofs)) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    std::vector<BCSet>& sets = is_u ? c->bcs_u : c->bcs_phi;
    if ((size_t)bc_id >= sets.size()) sets.resize(bc_id + 1);
    BCSet& s = sets[bc_id];
    int nc = is_u ? c->dim : 1;
    int64_t ndof = c->n_nodes * nc;
    s.dofs.resize(n);
    std::vector<int32_t> owned;
    for (int64_t k = 0; k < n; ++k) {
        if (dofs[k] < 0 || dofs[k] >= ndof) { c->err = "bc dof out of range"; return PFX_ERR_BAD_ARG; }
        int32_t node = dofs[k] / nc, comp = dofs[k] % nc;
        s.dofs[k] = c->bm.iperm[node] * nc + comp;
        if (c->bm.iperm[node] < c->n_owned) owned.push_back(s.dofs[k]);
    }
    s.n = n; s.n_owned = (int64_t)owned.size();
    CK(cudaStreamSynchronize(c->stream));
    dev_release(c, &s.d_dofs); dev_release(c, &s.d_owned); dev_release(c, &s.d_vals);
    int st = dev_upload(c, &s.d_dofs, s.dofs);
    if (st) return st;
    st = dev_upload(c, &s.d_owned, owned);
    if (st) return st;
    st = dev_alloc(c, &s.d_vals, (size_t)n);
    if (st) return st;
    return rebuild_bc(c, is_u);
}

int set_bc_values_common(pfx_ctx* c, bool is_u, int bc_id, const double* values, int64_t n) {
    if (!c) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    std::vector<BCSet>& sets = is_u ? c->bcs_u : c->bcs_phi;
    if (bc_id < 0 || (size_t)bc_id >= sets.size() || sets[bc_id].n != n || (n > 0 && !values)) {
        c->err = "bc values: unknown id or length mismatch";
        return PFX_ERR_BAD_ARG;
    }
    if (n > 0) CK(cudaMemcpyAsync(sets[bc_id].d_vals, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return push_bc_values(c, is_u);
}

#include "pfx_ec.inc"

EnergyCtl& ec_of(pfx_ctx* c) {
    if (!c->ec) c->ec = new EnergyCtl();
    return *static_cast<EnergyCtl*>(c->ec);
}

}  // namespace

// ================================================================================= C-ABI
extern "C" {

const char* pfx_version(void) { return "pfx-b200 0.1.0 (sm_100a)"; }

const char* pfx_status_string(int s) {
    switch (s) {
        case PFX_OK: return "ok";
        case PFX_ERR_BAD_ARG: return "bad argument";
        case PFX_ERR_CUDA: return "CUDA error";
        case PFX_ERR_NOT_CONVERGED_U: return "displacement solve did not converge";
        case PFX_ERR_NOT_CONVERGED_PHI: return "phase-field solve did not converge";
        case PFX_ERR_NOT_CONVERGED_PROJ: return "projection solve did not converge";
        case PFX_ERR_NAN: return "NaN encountered";
        case PFX_ERR_UNSUPPORTED: return "unsupported configuration";
        case PFX_ERR_COMM: return "communication error";
        case PFX_ERR_NO_DEVICE: return "no CUDA device";
    }
    return "unknown";
}

int pfx_device_count(int* n) {
    if (!n) return PFX_ERR_BAD_ARG;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; cudaGetLastError(); return PFX_ERR_NO_DEVICE; }
    return PFX_OK;
}

void pfx_default_params(pfx_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->lambda_ = 121.15384615384615; p->mu = 80.76923076923077; p->Gc = 0.0027; p->l = 0.015; p->k = 0.0;
    p->model = PFX_MODEL_ISOTROPIC; p->irreversibility = 0; p->fatigue = 0; p->fatigue_val = 0.05625;
    p->snes_rtol = 1e-8; p->snes_atol = 1e-9; p->snes_stol = 1e-8; p->snes_max_it = 50000;
    p->cg_rtol = 1e-11; p->cg_max_it = 200000; p->proj_rtol = 1e-13;
    p->quad_points_1d = 3; p->block_nodes = 0; p->cg_chunk = 0;
}

const char* pfx_last_error(const pfx_ctx* c) { return c ? c->err.c_str() : "null context"; }

int pfx_create(const pfx_mesh* mesh, const pfx_params* prm, int device, pfx_ctx** out) {
    if (!mesh || !prm || !out || !mesh->coords || !mesh->cells || mesh->n_nodes <= 0 || mesh->n_cells <= 0) return PFX_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return PFX_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return PFX_ERR_BAD_ARG;
    pfx_ctx* c = new pfx_ctx();
    *out = c;                       // returned even on failure so pfx_last_error() can be read
    c->device = device;
    c->prm = *prm;
    c->cell_type = mesh->cell_type;
    switch (mesh->cell_type) {
        case PFX_CELL_TRIANGLE: c->nv = 3; c->dim = 2; break;
        case PFX_CELL_QUADRILATERAL: c->nv = 4; c->dim = 2; break;
        case PFX_CELL_TETRAHEDRON: c->nv = 4; c->dim = 3; break;
        case PFX_CELL_HEXAHEDRON: c->nv = 8; c->dim = 3; break;
        default: c->err = "unknown cell type"; return PFX_ERR_BAD_ARG;
    }
    if (prm->quad_points_1d < 1 || prm->quad_points_1d > 4) { c->err = "quad_points_1d must be 1..4"; return PFX_ERR_BAD_ARG; }
    Material& m = c->mat;
    m.lam = prm->lambda_; m.mu = prm->mu; m.Gc = prm->Gc; m.l = prm->l; m.k = prm->k;
    m.fatigue = prm->fatigue; m.alpha_T = prm->fatigue_val;
    if (prm->degradation_function != PFX_DEGRADATION_QUADRATIC && prm->degradation_function != PFX_DEGRADATION_BORDEN) {
        c->err = "unknown degradation function"; return PFX_ERR_BAD_ARG;
    }
    m.dfun = prm->degradation_function;
    switch (prm->model) {
        case PFX_MODEL_ISOTROPIC: m.smodel = M_ISO; m.emodel = M_ISO; break;
        case PFX_MODEL_SPECTRAL: m.smodel = M_SPECTRAL; m.emodel = M_SPECTRAL; break;
        case PFX_MODEL_DEVIATORIC: m.smodel = M_VOLDEV; m.emodel = M_VOLDEV; break;
        case PFX_MODEL_HYBRID_SPECTRAL: m.smodel = M_ISO; m.emodel = M_SPECTRAL; break;
        case PFX_MODEL_HYBRID_DEVIATORIC: m.smodel = M_ISO; m.emodel = M_VOLDEV; break;
        default: c->err = "unknown model"; return PFX_ERR_BAD_ARG;
    }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->n_nodes = mesh->n_nodes;
    c->n_owned = mesh->n_owned > 0 ? mesh->n_owned : mesh->n_nodes;
    const char* ng = getenv("PFX_NO_GRAPH");
    c->use_graph = !(ng && ng[0] == '1');
    const char* ga = getenv("PFX_GENERIC_APPLY");
    c->generic_apply = (ga && ga[0] == '1');
    const char* vb = getenv("PFX_VERBOSE");
    c->verbose = (vb && vb[0] == '1');
    const char* ca = getenv("PFX_CG_ADAPTIVE");
    if (ca) c->cg_adaptive = (ca[0] != '0');
    const char* av = getenv("PFX_APPLY_VARIANT");
    if (av) c->apply_variant = atoi(av);

    // ---- blocks
    int target = prm->block_nodes;
    if (target <= 0) {
        // fill the machine (>= 2 CTAs per SM when the mesh allows), cap by the CTA width
        int64_t want = c->n_owned / (148 * 2);
        target = (int)std::min<int64_t>(kThreads - 16, std::max<int64_t>(64, want));
    }
    target = std::min(target, kThreads);
    for (int64_t e = 0; e < mesh->n_cells * c->nv; ++e)
        if (mesh->cells[e] < 0 || mesh->cells[e] >= mesh->n_nodes) { c->err = "cell vertex out of range"; return PFX_ERR_BAD_ARG; }
    std::string er;
    for (int attempt = 0;; ++attempt) {
        er = build_blocks(c->nv, c->dim, c->n_nodes, c->n_owned, mesh->n_cells, mesh->coords, mesh->cells,
                          mesh->cell_primary, target, c->bm);
        if (!er.empty()) { c->err = er; return PFX_ERR_BAD_ARG; }
        if (prm->block_nodes > 0 || target <= 64 || attempt >= 8) break;
        if (c->cell_type == PFX_CELL_TRIANGLE) {
            // triangles: keep every block within two element passes of the CTA (2 x kThreads)
            if (c->bm.max_elems <= 2 * kThreads) break;
            target = std::max(64, target - std::max(4, (c->bm.max_elems - 2 * kThreads) / 4));
        } else {
            break;
        }
    }
    BlockMesh& bm = c->bm;
    int32_t *d_node0, *d_hoff, *d_halo, *d_eoff, *d_nprim, *d_incp;
    uint16_t *d_lconn, *d_inc;
    int st;
    if ((st = dev_upload(c, &d_node0, bm.node0))) return st;
    if ((st = dev_upload(c, &d_hoff, bm.halo_off))) return st;
    if ((st = dev_upload(c, &d_halo, bm.halo))) return st;
    if ((st = dev_upload(c, &d_eoff, bm.elem_off))) return st;
    if ((st = dev_upload(c, &d_nprim, bm.nprim))) return st;
    if ((st = dev_upload(c, &d_incp, bm.inc_ptr))) return st;
    if ((st = dev_upload(c, &d_lconn, bm.lconn))) return st;
    if ((st = dev_upload(c, &d_inc, bm.inc))) return st;
    c->bv.nb = bm.nb; c->bv.node0 = d_node0; c->bv.halo_off = d_hoff; c->bv.halo = d_halo; c->bv.elem_off = d_eoff;
    c->bv.nprim = d_nprim; c->bv.lconn = (const ushort4*)d_lconn; c->bv.inc_ptr = d_incp; c->bv.inc = d_inc;
    c->bv.max_loc = bm.max_loc; c->bv.chunk = 256; c->bv.max_own = bm.max_own;
    {
        int32_t* d_meta; uint16_t* d_inc8;
        if ((st = dev_upload(c, &d_meta, bm.meta))) return st;
        if ((st = dev_upload(c, &d_inc8, bm.inc8))) return st;
        c->bv.meta = (const int4*)d_meta; c->bv.inc8 = (const uint4*)d_inc8;
    }
    {
        int mi = 0;
        for (int b = 0; b < bm.nb; ++b) mi = std::max(mi, bm.inc_ptr[bm.node0[b + 1]] - bm.inc_ptr[bm.node0[b]]);
        c->bv.max_inc = mi;
    }
    {   // own cells (compact per-cell index)
        int32_t *d_nlow, *d_c0;
        if ((st = dev_upload(c, &d_nlow, bm.nlow))) return st;
        if ((st = dev_upload(c, &d_c0, bm.own_c0))) return st;
        c->bv.nlow = d_nlow; c->bv.own_c0 = d_c0;
    }

    // ---- fields
    int64_t N = c->n_nodes, D = c->dim;
    std::vector<double> hx(N * D);
    for (int64_t i = 0; i < N; ++i)
        for (int k = 0; k < D; ++k) hx[i * D + k] = mesh->coords[3 * (int64_t)bm.perm[i] + k];
    if ((st = dev_upload(c, &c->X, hx))) return st;
    double** vecs[] = {&c->u, &c->u_old, &c->w_r, &c->w_p, &c->w_Ap, &c->w_x, &c->w_F, &c->w_dinv, &c->w_t0, &c->w_t1, &c->g_u};
    for (auto v : vecs)
        if ((st = dev_alloc(c, v, (size_t)(N * D)))) return st;
    double** scal[] = {&c->phi, &c->phi_old, &c->H, &c->Vc, &c->Vn, &c->Fat, &c->alpha_c, &c->alpha_n, &c->abar_c, &c->abar_n, &c->dinv_mass, &c->g_phi};
    for (auto v : scal)
        if ((st = dev_alloc(c, v, (size_t)N))) return st;
    if ((st = dev_alloc(c, &c->mask_u, (size_t)(N * D)))) return st;
    if ((st = dev_alloc(c, &c->mask_phi, (size_t)N))) return st;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->vec_grid = prop.multiProcessorCount * 8;
    c->num_sms = prop.multiProcessorCount;
    if ((st = dev_alloc(c, &c->partial, (size_t)std::max(bm.nb, c->vec_grid) * kMaxRed))) return st;
    if ((st = dev_alloc(c, &c->S, (size_t)S_COUNT))) return st;
    if ((st = dev_alloc(c, &c->red, (size_t)kMaxRed))) return st;
    if ((st = dev_alloc(c, &c->counter, (size_t)4))) return st;
    CK(cudaMallocHost((void**)&c->h_pinned, 64 * sizeof(double)));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    c->graph_iters = prm->cg_chunk > 0 ? (prm->cg_chunk + 1) / 2 * 2 : 32;

    // tetrahedra: assembled operators (pfx_sparse.cuh) from cached element tangents; PFX_NO_SPARSE=1 keeps the
    // matrix-free scatter kernels (slow in 3D; an independent cross-check of the assembly)
    {
        const char* ns = getenv("PFX_NO_SPARSE");
        // triangles stay matrix-free by default: measured on the 2 M-triangle bench mesh the 2 x 2-block SpMV takes 88 us
        // (3.2 TB/s on 285 MB; the gathers of the input vector cost as much L2 traffic as the matrix stream) against 57 us
        // for the persistent matrix-free kernel, and the scalar SpMV 64 us against 43 us.  PFX_TRI_SPARSE=1 selects it.
        const char* tsp = getenv("PFX_TRI_SPARSE");
        const bool tri_sparse = c->cell_type == PFX_CELL_TRIANGLE && tsp && tsp[0] == '1';
        if ((c->cell_type == PFX_CELL_TETRAHEDRON || c->cell_type == PFX_CELL_QUADRILATERAL || tri_sparse) && !(ns && ns[0] == '1')) {
            NodeGraph ng;
            std::string eg = build_node_graph(bm, mesh->n_cells, mesh->cells, ng);
            if (!eg.empty()) { c->err = eg; return PFX_ERR_BAD_ARG; }
            Sparse& sp = c->sp;
            int32_t *d_k0, *d_col, *d_cn, *d_ncp, *d_ncc;
            uint32_t* d_nck;
            if ((st = dev_upload(c, &d_k0, ng.slice_k0))) return st;
            if ((st = dev_upload(c, &d_col, ng.col))) return st;
            if ((st = dev_upload(c, &d_cn, ng.cell_nodes))) return st;
            if ((st = dev_upload(c, &d_ncp, ng.nc_ptr))) return st;
            if ((st = dev_upload(c, &d_ncc, ng.nc_code))) return st;
            if ((st = dev_upload(c, &d_nck, ng.nc_kpos))) return st;
            sp.sv.n_rows = ng.n_rows; sp.sv.n_slices = ng.n_slices; sp.sv.slice_k0 = d_k0; sp.sv.col = d_col;
            sp.sv.cell_nodes = (const int4*)d_cn; sp.sv.nc_ptr = d_ncp; sp.sv.nc_code = d_ncc; sp.sv.nc_kpos = d_nck;
            sp.total_k = ng.total_k;
            if ((st = dev_alloc(c, &sp.Ku, (size_t)ng.total_k * (c->dim * c->dim) * 32))) return st;
            if ((st = dev_alloc(c, &sp.Kphi, (size_t)ng.total_k * 32))) return st;
            if ((st = dev_alloc(c, &sp.Km, (size_t)ng.total_k * 32))) return st;
            if (c->dim == 3 && (st = dev_alloc(c, &c->tan, (size_t)21 * (size_t)std::max<int64_t>(bm.n_own_cells, 1)))) return st;
            sp.enabled = true;
        }
    }
    if (c->cell_type == PFX_CELL_TRIANGLE || c->sp.enabled) {
        if ((st = dev_alloc(c, &c->cm, (size_t)N + 2))) return st;
        if ((st = dev_alloc(c, &c->cg, (size_t)N + 2))) return st;
        if ((st = dev_alloc(c, &c->zeros, (size_t)N + 2))) return st;
        std::vector<double> one((size_t)N + 2, 1.0);
        if ((st = dev_upload(c, &c->ones, one))) return st;
    }
    // ---- opt every kernel instantiation of this mesh/model in to its shared-memory size
    {
        c->configuring = true;
        KArgs z = base_args(c);
        for (int op = OP_APPLY_U; op <= OP_DIAGBLK_U; ++op)
            if ((st = launch_scatter(c, op, z))) return st;
        if ((st = launch_reduce<RED_L2_U>(c, z))) return st;
        if ((st = launch_reduce<RED_L2_PHI>(c, z))) return st;
        if ((st = launch_reduce<RED_ENERGY>(c, z))) return st;
        if (c->cell_type == PFX_CELL_TRIANGLE && bm.max_elems <= 2 * kThreads && !getenv("PFX_NO_COOP")) {
            if ((st = coop_u(c, z, CoopArgs(), true))) return st;
            if ((st = coop_phi(c, z, CoopArgs(), true))) return st;
            if ((st = dev_alloc(c, &c->coop_bar, (size_t)2))) return st;
        }
        c->configuring = false;
    }
    // ---- constant data: 1/diag(M) (tetrahedra: the assembled mass matrix as well)
    if (c->sp.enabled) {
        const int grid = (int)((c->sp.sv.n_rows + 127) / 128);
        if (c->cell_type == PFX_CELL_QUADRILATERAL)
            assemble_s_quad_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->prm.quad_points_1d, c->X, c->ones, c->zeros, nullptr, c->sp.Km, c->dinv_mass);
        else if (c->cell_type == PFX_CELL_TRIANGLE)
            assemble_s_tri_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->X, c->ones, c->zeros, nullptr, c->sp.Km, c->dinv_mass);
        else
            assemble_s_tet_kernel<<<grid, 128, 0, c->stream>>>(c->sp.sv, c->X, c->ones, c->zeros, nullptr, c->sp.Km, c->dinv_mass);
        c->launches++;
        CK(cudaGetLastError());
    } else {
        KArgs a = base_args(c);
        a.y = c->dinv_mass;
        if ((st = launch_scatter(c, OP_DIAG_MASS, a))) return st;
    }
    CK(cudaStreamSynchronize(c->stream));
    return build_coarse(c, mesh);
}

int pfx_destroy(pfx_ctx* c) {
    if (!c) return PFX_OK;
    cudaSetDevice(c->device);
    for (cudaEvent_t e : c->ph_events) cudaEventDestroy(e);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 3; ++k)
        if (c->graph[k]) cudaGraphExecDestroy(c->graph[k]);
    for (int k = 0; k < 2; ++k)
        if (c->coarse.op[k].graph) cudaGraphExecDestroy(c->coarse.op[k].graph);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->comm) nccl().CommDestroy(c->comm);
    for (void* p : c->allocs) cudaFree(p);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete static_cast<EnergyCtl*>(c->ec);
    delete c;
    return PFX_OK;
}

int pfx_set_bc_u(pfx_ctx* c, int id, const int32_t* dofs, int64_t n) { return set_bc_common(c, true, id, dofs, n); }
int pfx_set_bc_phi(pfx_ctx* c, int id, const int32_t* dofs, int64_t n) { return set_bc_common(c, false, id, dofs, n); }
int pfx_set_bc_values_u(pfx_ctx* c, int id, const double* v, int64_t n) { return set_bc_values_common(c, true, id, v, n); }
int pfx_set_bc_values_phi(pfx_ctx* c, int id, const double* v, int64_t n) { return set_bc_values_common(c, false, id, v, n); }

int pfx_set_traction(pfx_ctx* c, int t_id, const int32_t* nodes, const double* w, int64_t n) {
    if (!c || t_id < 0 || n < 0 || (n > 0 && (!nodes || !w))) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    if ((size_t)t_id >= c->tractions.size()) c->tractions.resize(t_id + 1);
    Traction& t = c->tractions[t_id];
    CK(cudaStreamSynchronize(c->stream));
    dev_release(c, &t.d_nodes); dev_release(c, &t.d_w);
    std::vector<int32_t> nn(n);
    std::vector<double> ww(w, w + n);
    for (int64_t k = 0; k < n; ++k) {
        if (nodes[k] < 0 || nodes[k] >= c->n_nodes) { c->err = "traction node out of range"; return PFX_ERR_BAD_ARG; }
        nn[k] = c->bm.iperm[nodes[k]];
    }
    int st;
    if ((st = dev_upload(c, &t.d_nodes, nn))) return st;
    if ((st = dev_upload(c, &t.d_w, ww))) return st;
    t.n = n;
    if (!c->fext && (st = dev_alloc(c, &c->fext, (size_t)(c->n_nodes * c->dim)))) return st;
    c->has_fext = true;
    return rebuild_fext(c);
}

int pfx_set_traction_value(pfx_ctx* c, int t_id, const double* T) {
    if (!c || !T || t_id < 0 || (size_t)t_id >= c->tractions.size()) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    for (int k = 0; k < c->dim; ++k) c->tractions[t_id].T[k] = T[k];
    return rebuild_fext(c);
}

int pfx_get_field(pfx_ctx* c, int field, double* host, int64_t n) {
    if (!c || !host) return PFX_ERR_BAD_ARG;
    int nc;
    const double* d = field_ptr(c, field, &nc);
    if (!d || n != c->n_nodes * nc) { c->err = "get_field: unknown field or wrong length"; return PFX_ERR_BAD_ARG; }
    CK(cudaSetDevice(c->device));
    std::vector<double> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < c->n_nodes; ++i)
        for (int k = 0; k < nc; ++k) host[(int64_t)c->bm.perm[i] * nc + k] = tmp[i * nc + k];
    return PFX_OK;
}

int pfx_set_field(pfx_ctx* c, int field, const double* host, int64_t n) {
    if (!c || !host) return PFX_ERR_BAD_ARG;
    int nc;
    double* d = const_cast<double*>(field_ptr(c, field, &nc));
    if (!d || n != c->n_nodes * nc) { c->err = "set_field: unknown field or wrong length"; return PFX_ERR_BAD_ARG; }
    CK(cudaSetDevice(c->device));
    std::vector<double> tmp(n);
    for (int64_t i = 0; i < c->n_nodes; ++i)
        for (int k = 0; k < nc; ++k) tmp[i * nc + k] = host[(int64_t)c->bm.perm[i] * nc + k];
    CK(cudaMemcpyAsync(d, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->tan_fresh = false;
    c->coarse.op[0].stale = c->coarse.op[1].stale = true;
    return PFX_OK;
}

// ------------------------------------------------------------------------------- load step
int pfx_load_step(pfx_ctx* c, const pfx_step_opts* o, pfx_step_report* rep) {
    if (!c || !o || !rep) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    const int64_t N = c->n_nodes, D = c->dim;
    int cap = o->history_capacity;
    int32_t *hu = rep->hist_u_its, *hp = rep->hist_phi_its;
    double *heu = rep->hist_err_u, *hep = rep->hist_err_phi, *hfu = rep->hist_fnorm_u, *hfp = rep->hist_fnorm_phi;
    memset(rep, 0, sizeof(*rep));
    rep->hist_u_its = hu; rep->hist_phi_its = hp; rep->hist_err_u = heu; rep->hist_err_phi = hep;
    rep->hist_fnorm_u = hfu; rep->hist_fnorm_phi = hfp;
    double err_phi = 1.0, err_u = 1.0;
    int it = 0, st;
    CK(cudaEventRecord(c->ev0, c->stream));
    c->ph_on = true; c->ph_used = 0; c->floor_solves = 0;
    mark(c, PH_OTHER);
    CK(cudaMemsetAsync(c->S + S_BREF, 0, 3 * sizeof(double), c->stream));      // per-step PCG reference norms
    for (int k = 0; k < 2 && c->coarse.enabled; ++k) {
        CoarseOp& op = c->coarse.op[k];
        if (c->coarse.refresh_every_step) op.stale = true;
        // the reference iteration count is the largest of the solves in the step of the refresh and the next one
        if (op.ref_open && op.ref_its >= 0 && ++op.open_steps >= 2) op.ref_open = false;
    }
    const int gN = vgrid(c, N);
    while ((err_phi > o->stagger_tol || err_u > o->stagger_tol || it < o->min_stagger_iter) && it < o->max_stagger_iter) {
        ++it;
        // ---- displacement (solver.py:333-352)
        NewtonOut nu;
        if ((st = newton(c, true, &nu))) return st;
        rep->u_newton_its = nu.its; rep->u_reason = nu.reason; rep->fnorm_u = nu.fnorm; rep->cg_its_u += nu.cg_its;
        mark(c, PH_NORMS);
        if ((st = l2_error(c, true, c->u, c->u_old, &err_u))) return st;
        if (nu.reason <= 0) {
            rep->stagger_iters = it;
            c->err = "Solver did not converge, got " + std::to_string(nu.reason) + ".";
            return PFX_ERR_NOT_CONVERGED_U;
        }
        CK(cudaMemcpyAsync(c->u_old, c->u, N * D * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        // ---- history (solver.py:355-360)
        mark(c, PH_PROJECTION);
        KArgs a = base_args(c);
        a.y = c->w_F;
        if ((st = launch_scatter(c, OP_RHS_PSI, a))) return st;
        if ((st = project_solve(c, c->Vc, &rep->cg_its_proj))) return st;
        if (c->prm.irreversibility) {
            nodal_max_kernel<<<gN, kThreads, 0, c->stream>>>(N, c->H, c->Vc, c->Vn);
            c->launches++;
        } else {
            CK(cudaMemcpyAsync(c->H, c->Vc, N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
        // ---- fatigue (solver.py:362-375)
        if (c->prm.fatigue) {
            mark(c, PH_FATIGUE);
            KArgs f = base_args(c);
            f.f1 = c->phi_old; f.f2 = c->Vc; f.y = c->w_F;
            if ((st = launch_scatter(c, OP_RHS_FATIGUE, f))) return st;
            if ((st = project_solve(c, c->alpha_c, &rep->cg_its_proj))) return st;
            fatigue_update_kernel<<<gN, kThreads, 0, c->stream>>>(N, c->alpha_c, c->alpha_n, c->abar_n, c->abar_c, c->Fat,
                                                                 o->dt, c->prm.fatigue_val);
            c->launches++;
        }
        // ---- phase field (solver.py:380-398)
        NewtonOut np;
        if ((st = newton(c, false, &np))) return st;
        rep->phi_newton_its = np.its; rep->phi_reason = np.reason; rep->fnorm_phi = np.fnorm; rep->cg_its_phi += np.cg_its;
        mark(c, PH_NORMS);
        if ((st = l2_error(c, false, c->phi, c->phi_old, &err_phi))) return st;
        if (np.reason <= 0) {
            rep->stagger_iters = it;
            c->err = "Solver did not converge, got " + std::to_string(np.reason) + ".";
            return PFX_ERR_NOT_CONVERGED_PHI;
        }
        CK(cudaMemcpyAsync(c->phi_old, c->phi, N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        if (it <= cap) {
            int k = it - 1;
            if (hu) hu[k] = nu.its;
            if (hp) hp[k] = np.its;
            if (heu) heu[k] = err_u;
            if (hep) hep[k] = err_phi;
            if (hfu) hfu[k] = nu.fnorm;
            if (hfp) hfp[k] = np.fnorm;
            rep->history_len = it;
        }
    }
    // ---- end of step (solver.py:401-411)
    mark(c, PH_OTHER);
    if (c->prm.irreversibility) {
        nodal_max_kernel<<<gN, kThreads, 0, c->stream>>>(N, c->Vn, c->Vc, c->Vn);
        c->launches++;
    }
    if (c->prm.fatigue) {
        CK(cudaMemcpyAsync(c->abar_n, c->abar_c, N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->alpha_n, c->alpha_c, N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    rep->gpu_ms = ms;
    // per-phase split: time from each phase mark to the next one (the last phase runs to ev1)
    for (size_t k = 0; k < c->ph_used; ++k) {
        float d = 0.f;
        cudaEvent_t next = (k + 1 < c->ph_used) ? c->ph_events[k + 1] : c->ev1;
        if (cudaEventElapsedTime(&d, c->ph_events[k], next) == cudaSuccess) rep->phase_ms[c->ph_ids[k]] += d;
    }
    c->ph_on = false;
    rep->cg_floor_solves = c->floor_solves;
    rep->stagger_iters = it;
    rep->err_u = err_u; rep->err_phi = err_phi;
    return PFX_OK;
}

// ------------------------------------------------------------------------------- energy-controlled solvers
void pfx_ec_default_opts(pfx_ec_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->c1 = 1.0; o->c2 = 1.0; o->variational = 1;
    o->newton_rtol = 1e-9; o->newton_atol = 1e-10; o->newton_max_it = 50;       // dolfinx NewtonSolver defaults
    o->lin_rtol = 1e-10; o->gmres_restart = 40; o->lin_max_it = 2000; o->inner_rtol = 0.1;
}

int pfx_ec_solve(pfx_ctx* c, const pfx_ec_opts* o, double tau, double* lambda_inout, pfx_ec_report* rep) {
    if (!c || !o || !lambda_inout || !rep) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    return ec_solve(c, o, tau, lambda_inout, rep);
}

int pfx_ec_checkpoint(pfx_ctx* c, int restore) {
    if (!c) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    const size_t nu = (size_t)c->n_nodes * c->dim * sizeof(double), np = (size_t)c->n_nodes * sizeof(double);
    if (restore) {
        CK(cudaMemcpyAsync(c->u, c->u_old, nu, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->phi, c->phi_old, np, cudaMemcpyDeviceToDevice, c->stream));
        c->tan_fresh = false;
    } else {
        CK(cudaMemcpyAsync(c->u_old, c->u, nu, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->phi_old, c->phi, np, cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return PFX_OK;
}

// ------------------------------------------------------------------------------- outputs
static int solve_single(pfx_ctx* c, bool is_u, pfx_step_report* rep) {
    if (!c || !rep) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    memset(rep, 0, sizeof(*rep));
    c->ph_on = false; c->floor_solves = 0;
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(c->S + S_BREF, 0, 3 * sizeof(double), c->stream));
    NewtonOut no;
    int st = newton(c, is_u, &no);
    if (st) return st;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    rep->gpu_ms = ms;
    rep->cg_floor_solves = c->floor_solves;
    if (is_u) { rep->u_newton_its = no.its; rep->u_reason = no.reason; rep->fnorm_u = no.fnorm; rep->cg_its_u = no.cg_its; }
    else { rep->phi_newton_its = no.its; rep->phi_reason = no.reason; rep->fnorm_phi = no.fnorm; rep->cg_its_phi = no.cg_its; }
    if (no.reason <= 0) {
        c->err = "Solver did not converge, got " + std::to_string(no.reason) + ".";
        return is_u ? PFX_ERR_NOT_CONVERGED_U : PFX_ERR_NOT_CONVERGED_PHI;
    }
    return PFX_OK;
}

// the two sub-problems of the staggered step on their own (single-field solvers of the reference)
int pfx_solve_u(pfx_ctx* c, pfx_step_report* rep) { return solve_single(c, true, rep); }
int pfx_solve_phi(pfx_ctx* c, pfx_step_report* rep) { return solve_single(c, false, rep); }

int pfx_reaction(pfx_ctx* c, int bc_id, double R[3]) {
    if (!c || !R || bc_id < 0 || (size_t)bc_id >= c->bcs_u.size()) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    KArgs a = base_args(c);
    a.y = c->w_t0;
    int st = launch_scatter(c, OP_RESIDUAL_U, a);
    if (st) return st;
    BCSet& s = c->bcs_u[bc_id];
    gather_sum_kernel<<<vgrid(c, std::max<int64_t>(s.n_owned, 1)), kThreads, 0, c->stream>>>(s.n_owned, s.d_owned, c->w_t0, c->dim,
                                                                                              c->partial, c->red, c->counter);
    c->launches++;
    CK(cudaGetLastError());
    double v[3];
    if ((st = fetch(c, c->red, 3, v))) return st;
    R[0] = v[0]; R[1] = v[1]; R[2] = (c->dim == 3) ? v[2] : 0.0;
    return PFX_OK;
}

int pfx_energies(pfx_ctx* c, double out[11]) {
    if (!c || !out) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    KArgs a = base_args(c);
    a.f1 = c->abar_c;
    int st = launch_reduce<RED_ENERGY>(c, a);
    if (st) return st;
    double s[8];
    if ((st = fetch(c, c->red, 8, s))) return st;
    const Material& m = c->mat;
    double E = s[0] + s[2], PSI_a = s[1], PSI_b = s[2];
    double gamma_phi = s[3] / (2.0 * m.l), gamma_grad = s[4] * m.l / 2.0, gamma = gamma_phi + gamma_grad;
    double W_phi, W_grad, alpha_acum;
    if (m.fatigue) { W_phi = s[5] * m.Gc / (2.0 * m.l); W_grad = s[6] * m.Gc * m.l / 2.0; alpha_acum = s[7]; }
    else { W_phi = m.Gc * gamma_phi; W_grad = m.Gc * gamma_grad; alpha_acum = 0.0; }
    double W = W_phi + W_grad;
    double vals[11] = {E + W, E, W, W_phi, W_grad, gamma, gamma_phi, gamma_grad, PSI_a, PSI_b, alpha_acum};
    for (int k = 0; k < 11; ++k) out[k] = vals[k];
    return PFX_OK;
}

// ------------------------------------------------------------------------------- operator hooks
static int host_apply(pfx_ctx* c, int op, int nc, const double* v, double* y, const uint8_t* mask) {
    int64_t n = c->n_nodes * nc;
    std::vector<double> tmp(n);
    const bool is_apply = (op == OP_APPLY_U || op == OP_APPLY_PHI || op == OP_APPLY_MASS);
    if (v) {
        for (int64_t i = 0; i < c->n_nodes; ++i)
            for (int k = 0; k < nc; ++k) tmp[i * nc + k] = v[(int64_t)c->bm.perm[i] * nc + k];
        CK(cudaMemcpyAsync(c->w_p, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    int st;
    if (is_apply) {
        // the operators of the PCG path: they take an input that is zero on Dirichlet dofs and write zero rows there;
        // the identity rows of the assembled operator (y = v on those dofs) are restored afterwards
        if (mask) {
            CK(cudaMemcpyAsync(c->w_t0, c->w_p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            mask_zero_kernel<<<vgrid(c, n), kThreads, 0, c->stream>>>(n, mask, c->w_p);
        }
        if (op == OP_APPLY_U && (st = refresh_tangent(c))) return st;            // linearise at the current state
        if (op == OP_APPLY_PHI && (st = refresh_phi_operator(c))) return st;
        st = apply_lin(c, op == OP_APPLY_U ? LIN_U : (op == OP_APPLY_PHI ? LIN_PHI : LIN_MASS), c->w_p, c->w_Ap, mask, c->red, nullptr,
                       mask ? 1 : 0);
        if (op == OP_APPLY_PHI) c->phi_coeff_fresh = c->sp.phi_fresh = false;
        if (st) return st;
        if (mask) mask_restore_kernel<<<vgrid(c, n), kThreads, 0, c->stream>>>(n, mask, c->w_t0, c->w_Ap);
    } else {
        KArgs a = base_args(c);
        a.v = c->w_p; a.y = c->w_Ap; a.mask = mask;
        if ((st = launch_scatter(c, op, a))) return st;
    }
    CK(cudaMemcpyAsync(tmp.data(), c->w_Ap, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    for (int64_t i = 0; i < c->n_owned; ++i)
        for (int k = 0; k < nc; ++k) y[(int64_t)c->bm.perm[i] * nc + k] = tmp[i * nc + k];
    return PFX_OK;
}

int pfx_apply_u(pfx_ctx* c, const double* v, double* y, int use_bc) {
    if (!c || !v || !y) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    return host_apply(c, OP_APPLY_U, c->dim, v, y, use_bc ? c->mask_u : nullptr);
}
int pfx_apply_phi(pfx_ctx* c, const double* v, double* y, int use_bc) {
    if (!c || !v || !y) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    return host_apply(c, OP_APPLY_PHI, 1, v, y, use_bc ? c->mask_phi : nullptr);
}
int pfx_apply_mass(pfx_ctx* c, const double* v, double* y) {
    if (!c || !v || !y) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    return host_apply(c, OP_APPLY_MASS, 1, v, y, nullptr);
}
int pfx_residual_u(pfx_ctx* c, double* f) {
    if (!c || !f) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    return host_apply(c, OP_RESIDUAL_U, c->dim, nullptr, f, nullptr);
}
int pfx_diag_u(pfx_ctx* c, double* d) {
    if (!c || !d) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    if (c->sp.enabled) {                               // the diagonal falls out of the assembly
        int st0 = refresh_tangent(c);
        if (st0) return st0;
        int64_t n = c->n_owned * c->dim;
        std::vector<double> tmp(n);
        CK(cudaMemcpyAsync(tmp.data(), c->w_dinv, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int64_t i = 0; i < c->n_owned; ++i)
            for (int k = 0; k < c->dim; ++k) d[(int64_t)c->bm.perm[i] * c->dim + k] = 1.0 / tmp[i * c->dim + k];
        return PFX_OK;
    }
    int st = host_apply(c, OP_DIAG_U, c->dim, nullptr, d, c->mask_u);
    if (st) return st;
    for (int64_t i = 0; i < c->n_owned; ++i)                                 // kernel stores 1/diag
        for (int k = 0; k < c->dim; ++k) { double& v = d[(int64_t)c->bm.perm[i] * c->dim + k]; v = 1.0 / v; }
    return PFX_OK;
}
int pfx_rhs_psi(pfx_ctx* c, double* b) {
    if (!c || !b) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    return host_apply(c, OP_RHS_PSI, 1, nullptr, b, nullptr);
}

// L2 projection of psi_a (which = 0) or psi_b (which = 1) of the current u onto the nodal space, to the host
// (Math/projection.py:19-59 as called by solver_anisotropic.py:236-237); single GPU
int pfx_project_psi(pfx_ctx* c, int which, double* host) {
    if (!c || !host || which < 0 || which > 1) return PFX_ERR_BAD_ARG;
    if (c->n_ranks > 1) { c->err = "pfx_project_psi: single-GPU contexts only"; return PFX_ERR_BAD_ARG; }
    CK(cudaSetDevice(c->device));
    KArgs a = base_args(c);
    a.y = c->w_F; a.aux = which;
    int st = launch_scatter(c, OP_RHS_PSI, a);
    if (st) return st;
    bool conv = false; int64_t it = 0;
    if ((st = pcg_solve(c, LIN_MASS, c->prm.proj_rtol, &it, &conv))) return st;
    if (!conv) { c->err = "projection PCG did not converge"; return PFX_ERR_NOT_CONVERGED_PROJ; }
    std::vector<double> tmp(c->n_nodes);
    CK(cudaMemcpyAsync(tmp.data(), c->w_x, c->n_owned * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < c->n_owned; ++i) host[c->bm.perm[i]] = tmp[i];
    return PFX_OK;
}

int pfx_l2_error(pfx_ctx* c, int fa, int fb, double* err) {
    if (!c || !err) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    int na, nb;
    const double *a = field_ptr(c, fa, &na), *b = field_ptr(c, fb, &nb);
    if (!a || !b || na != nb) return PFX_ERR_BAD_ARG;
    return l2_error(c, na > 1, a, b, err);
}

int pfx_algorithmic_bytes(pfx_ctx* c, int kind, double* bytes) {
    if (!c || !bytes) return PFX_ERR_BAD_ARG;
    double Ne = (double)c->bm.n_cells, Nn = (double)c->n_owned, d = c->dim, nv = c->nv;
    double conn = 4.0 * nv * Ne;
    bool nl = c->mat.smodel != M_ISO;
    switch (kind) {
        case 0:                                                                                  // u-apply
            *bytes = conn + Nn * (8 * d + 8 + 8 * d + 8 * d + (nl ? 8 * d : 0));
            break;
        case 5:      // bytes the u-apply implementation streams: with cached element tangents X, v, y per node + 21 doubles per cell
            *bytes = conn + Nn * (8 * d + 8 + 8 * d + 8 * d + (nl ? 8 * d : 0));
            if (c->sp.enabled) *bytes = (double)c->sp.total_k * 32.0 * (8.0 * d * d + 4.0) + Nn * (8 * d + 8 * d);
            break;
        case 1: *bytes = conn + Nn * (8 * d + 24 + (c->mat.fatigue ? 8 : 0)); break;             // phi-apply
        case 2: *bytes = conn + Nn * (8 * d + 16); break;                                        // mass-apply
        case 3: *bytes = conn + Nn * (8 * d + 8 + 8 * d + 8 * d); break;                         // residual_u
        case 4:                                                                                  // u-PCG iteration
            *bytes = conn + Nn * (8 * d + 8 + 8 * d + 8 * d + (nl ? 8 * d : 0)) + 11.0 * 8.0 * d * Nn;
            if (coarse_active(c, LIN_U)) {
                // two-level iteration: inverse diagonal blocks instead of the diagonal (read twice), the mode
                // geometry (fp32, read twice), the Dirichlet flags, and the fp32 coarse inverse
                const CoarseOp& op = c->coarse.op[LIN_U];
                if (c->coarse.block_jacobi) *bytes += 2.0 * 8.0 * Nn * (d * (d + 1) / 2 - d);
                *bytes += Nn * (2.0 * 4.0 * d + d) + 4.0 * (double)op.ainv_floats;
            }
            break;
        default: return PFX_ERR_BAD_ARG;
    }
    return PFX_OK;
}

int pfx_time_operator(pfx_ctx* c, int kind, int reps, int flush_l2, float* ms_out) {
    if (!c || !ms_out || reps <= 0) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    if (flush_l2 && !c->flush_buf) {
        c->flush_n = (int64_t)256 * 1024 * 1024 / 8;
        int st = dev_alloc(c, &c->flush_buf, (size_t)c->flush_n);
        if (st) return st;
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    bool timed_coarse = false;
    if (kind == 0 || kind == 4) {                          // tetrahedra: linearise + assemble at the current state first
        int st = refresh_tangent(c);
        if (st) return st;
    }
    if (kind == 0) {
        int64_t nall = c->n_nodes * c->dim;
        mask_zero_kernel<<<vgrid(c, nall), kThreads, 0, c->stream>>>(nall, c->mask_u, c->w_p);
        CK(cudaGetLastError());
    }
    if (kind == 1) {
        mask_zero_kernel<<<vgrid(c, c->n_nodes), kThreads, 0, c->stream>>>(c->n_nodes, c->mask_phi, c->w_p);
        int st = refresh_phi_operator(c);
        if (st) return st;
    }
    // kind 4: a genuine PCG on J_u with right-hand side = current contents of w_p and a
    // tolerance of zero, so every timed iteration does full work
    if (kind == 4) {
        int st = PFX_OK;
        if (!c->sp.enabled) {
            KArgs d = base_args(c);
            d.y = c->w_dinv; d.mask = c->mask_u;
            if ((st = launch_scatter(c, OP_DIAG_U, d))) return st;
            if ((st = block_diag(c))) return st;
        }
        int64_t n = c->n_owned * c->dim;
        CK(cudaMemcpyAsync(c->w_F, c->w_p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        pcg_init_kernel<<<vgrid(c, n), kThreads, 0, c->stream>>>(n, c->w_F, c->mask_u, c->w_dinv, nullptr, nullptr, c->w_x, c->w_r,
                                                               c->w_p, c->partial, c->S + S_STAGE, c->counter);
        if (coarse_active(c, LIN_U)) {
            if ((st = coarse_setup(c, LIN_U, c->mask_u))) return st;
            timed_coarse = c->coarse.op[LIN_U].valid;
        }
        if (timed_coarse) {
            if ((st = coarse_tail(c, LIN_U, 0, 1, c->mask_u, c->w_dinv, 0.0, -1, (c->n_ranks > 1 && c->p2p && c->p2p_fused) ? 1 : 0))) return st;
        } else {
            if ((st = allreduce(c, c->S + S_STAGE, 2))) return st;
            pcg_setup_kernel<<<1, 1, 0, c->stream>>>(c->S, 0.0, -1);
        }
        CK(cudaStreamSynchronize(c->stream));
    }
    for (int r = 0; r < reps; ++r) {
        if (flush_l2) {
            flush_kernel<<<c->vec_grid, kThreads, 0, c->stream>>>(c->flush_n, c->flush_buf);
        }
        CK(cudaEventRecord(e0, c->stream));
        int st = PFX_OK;
        if (kind == 0) st = apply_lin(c, LIN_U, c->w_p, c->w_Ap, c->mask_u, c->red, nullptr, 1);   // the launch PCG makes
        else if (kind == 1) st = apply_lin(c, LIN_PHI, c->w_p, c->w_Ap, c->mask_phi, c->red, nullptr, 1);
        else if (kind == 2) st = apply_lin(c, LIN_MASS, c->w_p, c->w_Ap, nullptr, c->red, nullptr);
        else if (kind == 3) { KArgs a = base_args(c); a.y = c->w_t0; st = launch_scatter(c, OP_RESIDUAL_U, a); }
        else if (kind == 4) {
            st = timed_coarse ? pcg_iteration_coarse(c, LIN_U, r & 1, c->mask_u, c->w_dinv)
                              : pcg_iteration(c, LIN_U, r & 1, c->n_owned * c->dim, c->mask_u, c->w_dinv);
        } else st = PFX_ERR_BAD_ARG;
        if (st) return st;
        CK(cudaEventRecord(e1, c->stream));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_out[r], e0, e1));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kind == 1) c->phi_coeff_fresh = c->sp.phi_fresh = false;
    return PFX_OK;
}

// test hook: the in-tree dense SPD inverse used for the coarse matrices (pfx_dense.cuh)
int pfx_dense_spd_inverse(int device, int n, const double* A_host, double* Ainv_host, int* info_out) {
    if (n <= 0 || !A_host || !Ainv_host) return PFX_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PFX_ERR_NO_DEVICE;
    double *A = nullptr, *Y = nullptr, *T = nullptr, *L = nullptr;
    int* info = nullptr;
    const size_t nn = (size_t)n * n, nblk = (n + kDenseNB - 1) / kDenseNB;
    bool ok = cudaMalloc(&A, nn * 8) == cudaSuccess && cudaMalloc(&Y, nn * 8) == cudaSuccess &&
              cudaMalloc(&T, (size_t)n * kDenseNB * 8) == cudaSuccess && cudaMalloc(&L, nblk * kDenseNB * kDenseNB * 8) == cudaSuccess &&
              cudaMalloc(&info, sizeof(int)) == cudaSuccess;
    int h_info = 0;
    if (ok) {
        cudaMemcpy(A, A_host, nn * 8, cudaMemcpyHostToDevice);
        cudaMemset(info, 0, sizeof(int));
        dense_spd_inverse(nullptr, n, A, Y, T, L, info);
        ok = cudaDeviceSynchronize() == cudaSuccess;
        cudaMemcpy(Ainv_host, A, nn * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(A); cudaFree(Y); cudaFree(T); cudaFree(L); cudaFree(info);
    if (info_out) *info_out = h_info;
    return ok ? PFX_OK : PFX_ERR_CUDA;
}

int pfx_launch_count(pfx_ctx* c, int64_t* n) {
    if (!c || !n) return PFX_ERR_BAD_ARG;
    *n = c->launches;
    return PFX_OK;
}

int pfx_coarse_stats(pfx_ctx* c, int64_t out[4]) {
    if (!c || !out) return PFX_ERR_BAD_ARG;
    out[0] = c->coarse.enabled ? 1 : 0; out[1] = c->coarse.n_agg; out[2] = c->coarse.n_col; out[3] = c->coarse.setups;
    return PFX_OK;
}

int pfx_block_stats(pfx_ctx* c, int64_t out[8]) {
    if (!c || !out) return PFX_ERR_BAD_ARG;
    out[0] = c->bm.nb; out[1] = c->bm.max_own; out[2] = c->bm.max_loc; out[3] = c->bm.max_elems;
    out[4] = c->bm.total_elems; out[5] = c->bm.n_cells; out[6] = (int64_t)c->bm.halo.size(); out[7] = c->n_owned;
    return PFX_OK;
}

// ------------------------------------------------------------------------------- partition API
struct pfx_partition {
    Partition P;
};

int pfx_partition_create(int32_t cell_type, int64_t n_nodes, int64_t n_cells, const double* coords, const int32_t* cells,
                         int n_ranks, const int32_t* cell_rank_in, pfx_partition** out) {
    if (!coords || !cells || !out || n_nodes <= 0 || n_cells <= 0) return PFX_ERR_BAD_ARG;
    int nv = (cell_type == PFX_CELL_TRIANGLE) ? 3 : (cell_type == PFX_CELL_HEXAHEDRON ? 8 : 4);
    pfx_partition* p = new pfx_partition();
    std::string er = build_partition(nv, n_nodes, n_cells, coords, cells, n_ranks, cell_rank_in, p->P);
    if (!er.empty()) { delete p; return PFX_ERR_BAD_ARG; }
    *out = p;
    return PFX_OK;
}
int pfx_partition_destroy(pfx_partition* p) { delete p; return PFX_OK; }
int pfx_partition_graph_cell_rank(int32_t cell_type, int64_t n_nodes, int64_t n_cells, const int32_t* cells, int n_ranks,
                                  int32_t* cell_rank, int64_t* edge_cut) {
    if (!cells || !cell_rank || n_cells <= 0 || n_ranks < 1) return PFX_ERR_BAD_ARG;
    int nv = cell_type == PFX_CELL_TRIANGLE ? 3 : 4, dim = cell_type == PFX_CELL_TETRAHEDRON ? 3 : 2;
    if (cell_type < PFX_CELL_TRIANGLE || cell_type > PFX_CELL_TETRAHEDRON) return PFX_ERR_BAD_ARG;      // hexahedra: RCB only
    return metis_cell_rank(nv, dim, n_nodes, n_cells, cells, n_ranks, cell_rank, edge_cut).empty() ? PFX_OK : PFX_ERR_BAD_ARG;
}
int pfx_partition_cell_rank(const pfx_partition* p, int32_t* cr) {
    if (!p || !cr) return PFX_ERR_BAD_ARG;
    std::copy(p->P.cell_rank.begin(), p->P.cell_rank.end(), cr);
    return PFX_OK;
}
int pfx_partition_node_owner(const pfx_partition* p, int32_t* no) {
    if (!p || !no) return PFX_ERR_BAD_ARG;
    std::copy(p->P.node_owner.begin(), p->P.node_owner.end(), no);
    return PFX_OK;
}
int pfx_partition_rank_sizes(const pfx_partition* p, int rank, int64_t out[5]) {
    if (!p || !out || rank < 0 || rank >= p->P.n_ranks) return PFX_ERR_BAD_ARG;
    const RankPart& R = p->P.ranks[rank];
    out[0] = (int64_t)R.l2g.size(); out[1] = R.n_owned; out[2] = (int64_t)R.cell_gid.size();
    out[3] = (int64_t)R.send_idx.size(); out[4] = (int64_t)R.neighbors.size();
    return PFX_OK;
}
int pfx_partition_rank_maps(const pfx_partition* p, int rank, int64_t* l2g, int64_t* cell_gid, int32_t* ghost_owner,
                            uint8_t* cell_primary) {
    if (!p || rank < 0 || rank >= p->P.n_ranks) return PFX_ERR_BAD_ARG;
    const RankPart& R = p->P.ranks[rank];
    if (l2g) std::copy(R.l2g.begin(), R.l2g.end(), l2g);
    if (cell_gid) std::copy(R.cell_gid.begin(), R.cell_gid.end(), cell_gid);
    if (ghost_owner) std::copy(R.ghost_owner.begin(), R.ghost_owner.end(), ghost_owner);
    if (cell_primary) std::copy(R.cell_primary.begin(), R.cell_primary.end(), cell_primary);
    return PFX_OK;
}
int pfx_partition_rank_plan(const pfx_partition* p, int rank, int32_t* nbrs, int64_t* send_off, int32_t* send_idx,
                            int64_t* recv_off) {
    if (!p || rank < 0 || rank >= p->P.n_ranks) return PFX_ERR_BAD_ARG;
    const RankPart& R = p->P.ranks[rank];
    if (nbrs) std::copy(R.neighbors.begin(), R.neighbors.end(), nbrs);
    if (send_off) std::copy(R.send_off.begin(), R.send_off.end(), send_off);
    if (send_idx) std::copy(R.send_idx.begin(), R.send_idx.end(), send_idx);
    if (recv_off) std::copy(R.recv_off.begin(), R.recv_off.end(), recv_off);
    return PFX_OK;
}

int pfx_comm_unique_id(uint8_t id[128]) {
    if (!id) return PFX_ERR_BAD_ARG;
    std::string er = nccl().load();
    if (!er.empty()) return PFX_ERR_COMM;
    nccl_uid u;
    if (nccl().GetUniqueId(&u) != 0) return PFX_ERR_COMM;
    memcpy(id, u.internal, 128);
    return PFX_OK;
}

int pfx_comm_attach(pfx_ctx* c, int rank, int n_ranks, const uint8_t id[128], int n_nbrs, const int32_t* nbrs,
                    const int64_t* send_off, const int32_t* send_idx, const int64_t* recv_off) {
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    std::string er = nccl().load();
    if (!er.empty()) { c->err = er; return PFX_ERR_COMM; }
    nccl_uid u;
    memcpy(u.internal, id, 128);
    int r = nccl().CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != 0) { c->err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r); return PFX_ERR_COMM; }
    c->rank = rank; c->n_ranks = n_ranks;
    c->nbrs.assign(nbrs, nbrs + n_nbrs);
    c->send_off.assign(send_off, send_off + n_nbrs + 1);
    c->recv_off.assign(recv_off, recv_off + n_nbrs + 1);
    c->n_send = n_nbrs > 0 ? c->send_off[n_nbrs] : 0;
    std::vector<int32_t> idx(c->n_send);
    for (int64_t k = 0; k < c->n_send; ++k) idx[k] = c->bm.iperm[send_idx[k]];
    int st;
    if ((st = dev_upload(c, &c->d_send_idx, idx))) return st;
    if ((st = dev_alloc(c, &c->d_send_buf, (size_t)(c->n_send * c->dim)))) return st;
    // graphs captured before the attach (none normally) would miss the collectives
    for (int k = 0; k < 3; ++k)
        if (c->graph[k]) { cudaGraphExecDestroy(c->graph[k]); c->graph[k] = nullptr; }
    for (int k = 0; k < 2; ++k)
        if (c->coarse.op[k].graph) { cudaGraphExecDestroy(c->coarse.op[k].graph); c->coarse.op[k].graph = nullptr; }
    // the coarse space is used by all ranks or by none (a rank with a tiny part has none)
    if (n_ranks > 1) {
        double mine = c->coarse.enabled ? 1.0 : 0.0, tot = 0.0;
        CK(cudaMemcpyAsync(c->S + S_STAGE, &mine, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if ((st = fetch(c, c->S + S_STAGE, 1, &tot))) return st;
        if (tot != (double)n_ranks) c->coarse.enabled = false;
    }
    return PFX_OK;
}

// ------------------------------------------------------------------------------- NVLink peer path
int pfx_comm_p2p_export(pfx_ctx* c, uint8_t* handles) {
    if (!c || !handles) return PFX_ERR_BAD_ARG;
    CK(cudaSetDevice(c->device));
    if (!c->mbox) {
        int st = dev_alloc(c, &c->mbox, (size_t)4 * kMaxRanks * kMboxStride + 2 * kMaxRanks + 16);
        if (st) return st;
        CK(cudaStreamSynchronize(c->stream));
    }
    void* vecs[6] = {c->w_p, c->u, c->phi, c->Vc, c->alpha_c, c->mbox};
    for (int k = 0; k < 6; ++k) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, vecs[k]));
        static_assert(sizeof(h) == 64, "IPC handle size");
        memcpy(handles + 64 * k, &h, 64);
    }
    return PFX_OK;
}

int pfx_comm_p2p_import(pfx_ctx* c, const uint8_t* all, const int64_t* peer_ghost_start) {
    if (!c || !all || (c->nbrs.size() && !peer_ghost_start)) return PFX_ERR_BAD_ARG;
    if (c->n_ranks < 2 || c->n_ranks > kMaxRanks || !c->mbox) { c->err = "p2p import: attach + export first (2..8 ranks)"; return PFX_ERR_BAD_ARG; }
    CK(cudaSetDevice(c->device));
    const int R = c->n_ranks, nn = (int)c->nbrs.size();
    std::vector<std::vector<void*>> peer(R, std::vector<void*>(6, nullptr));
    void* mine[6] = {c->w_p, c->u, c->phi, c->Vc, c->alpha_c, c->mbox};
    auto need = [&](int r) { return std::find(c->nbrs.begin(), c->nbrs.end(), r) != c->nbrs.end(); };
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < 6; ++k) {
            if (r == c->rank) { peer[r][k] = mine[k]; continue; }
            if (k < 5 && !need(r)) continue;                       // vectors only of neighbours, mailbox of everyone
            cudaIpcMemHandle_t h;
            memcpy(&h, all + (size_t)r * PFX_P2P_HANDLE_BYTES + 64 * k, 64);
            CK(cudaIpcOpenMemHandle(&peer[r][k], h, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(peer[r][k]);
        }
    PeerView& pv = c->pv;
    memset(&pv, 0, sizeof(pv));
    pv.rank = c->rank; pv.n_ranks = R; pv.n_nbrs = nn;
    for (int j = 0; j < nn; ++j) pv.nbr_rank[j] = c->nbrs[j];
    const size_t flag_off = (size_t)4 * kMaxRanks * kMboxStride;   // in doubles
    pv.mbox_local = c->mbox;
    pv.hflag_local = reinterpret_cast<unsigned long long*>(c->mbox + flag_off);
    pv.seq = reinterpret_cast<unsigned long long*>(c->mbox + flag_off + kMaxRanks);
    pv.err = c->S + S_ERR;
    for (int r = 0; r < R; ++r) {
        pv.mbox_peer[r] = reinterpret_cast<double*>(peer[r][5]);
        pv.hflag_peer[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(peer[r][5]) + flag_off);
    }
    // per-entry destination of the send list
    std::vector<uint8_t> nbr(c->n_send);
    std::vector<int64_t> dst(c->n_send);
    for (int j = 0; j < nn; ++j)
        for (int64_t k = c->send_off[j]; k < c->send_off[j + 1]; ++k) { nbr[k] = (uint8_t)j; dst[k] = peer_ghost_start[j] + (k - c->send_off[j]); }
    int st;
    if ((st = dev_upload(c, &c->d_send_nbr, nbr))) return st;
    if ((st = dev_upload(c, &c->d_send_dst, dst))) return st;
    for (int v = 0; v < 5; ++v) {
        std::vector<double*> tab(std::max(nn, 1), nullptr);
        for (int j = 0; j < nn; ++j) tab[j] = reinterpret_cast<double*>(peer[c->nbrs[j]][v]);
        if ((st = dev_upload(c, &c->d_peer_vec[v], tab))) return st;
    }
    {
        std::vector<PeerView> one(1, pv);
        if ((st = dev_upload(c, &c->d_pv, one))) return st;
    }
    const char* np = getenv("PFX_NO_P2P");
    c->p2p = !(np && np[0] == '1');
    const char* pf = getenv("PFX_P2P_FUSED");
    if (pf) c->p2p_fused = (pf[0] != '0');
    for (int k = 0; k < 3; ++k)
        if (c->graph[k]) { cudaGraphExecDestroy(c->graph[k]); c->graph[k] = nullptr; }
    return PFX_OK;
}

}  // extern "C"
