/*
 * pfx_b200.h — C-ABI of libpfx_b200.so, the B200-native staggered phase-field fracture
 * load-step engine behind the PhaseFieldX driver API.
 *
 * Every entry point is `extern "C"`, takes plain pointers/sizes, returns an int status
 * (0 = PFX_OK, >0 = PFX_ERR_*), never throws and never calls exit().  The caller owns all
 * host arrays passed in (the library copies them during the call and retains no host
 * pointer); the library owns the device memory behind the opaque `pfx_ctx*`.
 * One host thread drives a context; distinct contexts are independent.
 *
 * Each function cites the reference interface it replaces (paths relative to the
 * reference repository root, `src/phasefieldx/` abbreviated as `pfx/`).
 *
 * Layouts: coordinates [N_n,3] f64 (dolfinx pads to 3), cells [N_e,n_v] i32,
 * u interleaved [N_n*d] f64, scalar nodal fields [N_n] f64, all in the CALLER's node order.
 */
#ifndef PFX_B200_H
#define PFX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---------------------------------------------------------------- */
enum {
    PFX_OK = 0,
    PFX_ERR_BAD_ARG = 1,
    PFX_ERR_CUDA = 2,
    PFX_ERR_NOT_CONVERGED_U = 3,     /* SNES-like reason <= 0 on the displacement solve   */
    PFX_ERR_NOT_CONVERGED_PHI = 4,   /* ... on the phase-field solve                      */
    PFX_ERR_NOT_CONVERGED_PROJ = 5,  /* mass-matrix PCG of the L2 projection              */
    PFX_ERR_NAN = 6,
    PFX_ERR_UNSUPPORTED = 7,
    PFX_ERR_COMM = 8,
    PFX_ERR_NO_DEVICE = 9
};

/* ---- enumerations ---------------------------------------------------------------- */
enum { PFX_CELL_TRIANGLE = 0, PFX_CELL_QUADRILATERAL = 1, PFX_CELL_TETRAHEDRON = 2,
       PFX_CELL_HEXAHEDRON = 3 /* Q1, tensor vertex order x + 2y + 4z (dolfinx) */ };
/* pfx/Element/Phase_Field_Fracture/split_energy_stress_tangent_functions.py:234-354 */
enum { PFX_MODEL_ISOTROPIC = 0, PFX_MODEL_SPECTRAL = 1, PFX_MODEL_DEVIATORIC = 2,
       PFX_MODEL_HYBRID_SPECTRAL = 3, PFX_MODEL_HYBRID_DEVIATORIC = 4 };
/* degradation function g(phi): quadratic (1-phi)^2; borden 3(1-phi)^2 - 2(1-phi)^3 as coded in the reference (s = 0).
 * alessi / sargado are not built. */
enum { PFX_DEGRADATION_QUADRATIC = 0, PFX_DEGRADATION_BORDEN = 1 };
enum { PFX_FIELD_U = 0, PFX_FIELD_PHI = 1, PFX_FIELD_H = 2, PFX_FIELD_VC = 3, PFX_FIELD_VN = 4,
       PFX_FIELD_FATIGUE = 5, PFX_FIELD_ALPHA_BAR = 6, PFX_FIELD_ALPHA_C = 7, PFX_FIELD_U_OLD = 8,
       PFX_FIELD_PHI_OLD = 9, PFX_FIELD_ALPHA_N = 10, PFX_FIELD_ALPHA_BAR_N = 11 };

/* ---- plain-data structs ---------------------------------------------------------- */
typedef struct pfx_ctx pfx_ctx;

typedef struct {
    int32_t cell_type;        /* PFX_CELL_*                                                */
    int64_t n_nodes;          /* nodes held by this context (owned first, then ghosts)     */
    int64_t n_cells;
    const double* coords;     /* [n_nodes,3]  (msh.geometry.x)                             */
    const int32_t* cells;     /* [n_cells,n_v] (V.dofmap.list; quads in tensor order)      */
    /* distributed runs only (NULL / 0 for one GPU): see pfx_partition_create                   */
    int64_t n_owned;          /* 0 => all nodes owned                                      */
    const uint8_t* cell_primary; /* [n_cells] 1 if this rank integrates the cell in scalar
                                    reductions (energies, norms); NULL => all              */
} pfx_mesh;

/* pfx/Element/Phase_Field_Fracture/Input.py:66-102 */
typedef struct {
    double lambda_, mu, Gc, l, k;
    int32_t model;            /* PFX_MODEL_*                                               */
    int32_t irreversibility;  /* 1 = "miehe": H = max(V_c, V_n); 0: H = V_c (solver.py:357) */
    int32_t fatigue;          /* 0/1 (solver.py:362-375)                                   */
    double fatigue_val;       /* alpha_T of the asymptotic f (fatigue_degradation_functions.py:19-38) */
    /* solver controls (PETSc option dicts are hard-coded in solver.py:188-196, 227-235)   */
    double snes_rtol, snes_atol, snes_stol;  /* 1e-8, 1e-9, 1e-8                            */
    int32_t snes_max_it;                     /* 50000                                      */
    /* PCG stop test standing in for the MUMPS solve: ||r|| <= cg_rtol * max(||b||, largest ||b|| this operator has
     * seen in the current load step) — i.e. every solve of a step reaches the same ABSOLUTE accuracy as its first
     * one (default 1e-11; PFX_CG_ADAPTIVE=0: purely relative test).  A solve whose right-hand side is already
     * below that floor returns delta = 0 after 0 iterations and is counted in pfx_step_report.cg_floor_solves;
     * Newton then ends through the SNES stol test (reason 4). */
    double cg_rtol;
    int32_t cg_max_it;        /* default 200000                                            */
    double proj_rtol;         /* mass-PCG tolerance of project() (default 1e-13) */
    int32_t quad_points_1d;   /* Gauss points per direction for Q1 cells (default 3)       */
    int32_t block_nodes;      /* owned nodes per CTA tile, 0 = auto                         */
    int32_t cg_chunk;         /* CG iterations per CUDA-graph launch, 0 = auto             */
    int32_t degradation_function; /* PFX_DEGRADATION_*: g of g_degradation_functions.py:12-64 (Input.degradation_function) */
    /* aggregate coarse space of the two-level PCG, 0 = automatic: coarse u-dofs per dense region (2D default 1792: many
     * small aggregates in compact regions; one region of <= 6144 suits slender, weakly supported bodies such as the
     * three-point bending beam, whose global bending mode a region-wise coarse matrix cannot represent) and CTA blocks
     * per aggregate.  The environment variables PFX_COARSE_REGION / PFX_COARSE_AGG_BLOCKS override both. */
    int32_t coarse_region, coarse_agg_blocks;
} pfx_params;

typedef struct {
    double dt;                /* solver.py:55, used by the fatigue heaviside argument      */
    int32_t min_stagger_iter, max_stagger_iter;
    double stagger_tol;       /* solver.py:58-60                                            */
    int32_t history_capacity; /* length of the per-stagger-iteration arrays below (may be 0) */
} pfx_step_opts;

/* per-stagger-iteration log lines of solver.py:326-350, 378-396 */
typedef struct {
    int32_t stagger_iters;
    int32_t u_newton_its, phi_newton_its;        /* of the LAST stagger iteration (phasefieldx.conv) */
    int32_t u_reason, phi_reason;                /* SNES-like converged reason (2,3,4 / <=0)  */
    double err_u, err_phi;                       /* eval_error_L2_normalized, errors_functions.py:190-228 */
    double fnorm_u, fnorm_phi;
    int64_t cg_its_u, cg_its_phi, cg_its_proj;   /* totals over the load step                */
    double gpu_ms;            /* CUDA-event time of the whole step on the context stream   */
    int32_t history_len;
    int32_t* hist_u_its;      /* caller-allocated [history_capacity] or NULL                */
    int32_t* hist_phi_its;
    double* hist_err_u;
    double* hist_err_phi;
    double* hist_fnorm_u;
    double* hist_fnorm_phi;
    /* device time per phase of the step (CUDA events on the context stream), ms, summed over the stagger
     * iterations: [0] u residual / tangent / diagonal set-up, [1] u coarse-space set-up, [2] u PCG,
     * [3] projection + history (solver.py:355-360), [4] fatigue block, [5] phi set-up, [6] phi coarse-space
     * set-up, [7] phi PCG, [8] L2 error norms, [9] everything else */
    double phase_ms[10];
    int64_t cg_floor_solves;  /* PCG solves that met the per-step absolute floor before the first iteration */
} pfx_step_report;

/* ---- library-level ---------------------------------------------------------------- */
const char* pfx_version(void);
int pfx_device_count(int* n);
const char* pfx_status_string(int status);

/* ---- context ---------------------------------------------------------------------- */
/* Replaces the form/solver construction of solver.py:148-245 (Functions, F_u/J_u, F_phi/J_phi,
 * NonlinearProblem objects).  `device` = CUDA ordinal. */
int pfx_create(const pfx_mesh* mesh, const pfx_params* params, int device, pfx_ctx** out);
int pfx_destroy(pfx_ctx* ctx);
const char* pfx_last_error(const pfx_ctx* ctx);
void pfx_default_params(pfx_params* p);

/* ---- Dirichlet data: dolfinx.fem.dirichletbc objects built by pfx/Boundary/boundary_conditions.py:23-234.
 * `dofs` are unrolled (bc.dof_indices()[0]); BCs are applied in id order, later ids win on
 * shared dofs (dolfinx set_bc order).  Values are re-sent every step after the user callback
 * (solver.py:306-309): one value per dof. */
int pfx_set_bc_u(pfx_ctx* ctx, int bc_id, const int32_t* dofs, int64_t n);
int pfx_set_bc_values_u(pfx_ctx* ctx, int bc_id, const double* values, int64_t n);
int pfx_set_bc_phi(pfx_ctx* ctx, int bc_id, const int32_t* dofs, int64_t n);
int pfx_set_bc_values_phi(pfx_ctx* ctx, int bc_id, const double* values, int64_t n);
/* Neumann data of solver.py:179-184 (T_list_u): pre-integrated nodal weights w_i = ∫ N_i ds of
 * the tagged boundary (host side computes them once); load vector = w_i * T_c, T re-sent per step. */
int pfx_set_traction(pfx_ctx* ctx, int t_id, const int32_t* nodes, const double* weights, int64_t n);
int pfx_set_traction_value(pfx_ctx* ctx, int t_id, const double* T /* [d] */);

/* ---- the hot path: one load step = the whole stagger loop of solver.py:319-411 on device */
int pfx_load_step(pfx_ctx* ctx, const pfx_step_opts* opts, pfx_step_report* out);

/* ---- the two sub-problems on their own: the single-field solvers of the reference are
 * the staggered step with one field frozen.
 * pfx_solve_u: one SNES solve of F_u = 0 at the current phi (phi = 0, k = 0, isotropic model =
 *   Element/Elasticity/solver/solver.py:110-141); fills u_newton_its, u_reason, fnorm_u, cg_its_u, gpu_ms.
 * pfx_solve_phi: one SNES solve of the phase-field problem at the current H, fatigue field
 *   (H = 0, Gc = 1 = Element/Phase_Field/solver/solver.py:104-125 with case AT2); fills the phi members.
 * Non-convergence returns PFX_ERR_NOT_CONVERGED_U / _PHI like pfx_load_step. */
int pfx_solve_u(pfx_ctx* ctx, pfx_step_report* out);
int pfx_solve_phi(pfx_ctx* ctx, pfx_step_report* out);

/* ---- energy-controlled monolithic solvers: one call = one Newton solve of the blocked system
 * (u, phi, lambda) at the control value tau —
 *   pfx/Element/Phase_Field_Fracture/solver/solver_ener_variational.py:167-259, 360 (`solver_snap.solve()`) and
 *   solver_ener_non_variational.py:166-270 (variational = 0: no lambda term in F_phi, J_phi,lambda = None).
 *   F_u = F_int(u, phi) - (1 - lambda c2) int T.w ds,  F_phi = int g'(phi) psi_a(u) v + (Gc + lambda c1) K_gamma phi,
 *   F_lambda = c1 gamma(phi) + c2 int T.u ds - tau.
 * The traction is the one registered with pfx_set_traction(ctx, 0, ...) / pfx_set_traction_value (T_list_u[0] of the
 * reference signature); Dirichlet data: the u-sets only, already satisfied by u (homogeneous in every reference
 * set-up).  Newton: dolfinx NewtonSolver semantics (residual criterion; defaults rtol 1e-9, atol 1e-10, max_it 50, the
 * defaults scifem's BlockedNewtonSolver inherits).  The MUMPS solve of the bordered Jacobian is a flexible GMRES,
 * block-preconditioned by the engine's PCG solvers (lin_rtol: relative residual of the linear solve; inner_rtol: of the
 * block solves inside the preconditioner).  A solve that does not converge returns PFX_OK with converged = 0 (the
 * reference catches the solver's exception and shrinks dtau, :486-514); the caller restores the last converged state
 * with pfx_ec_checkpoint(ctx, 1).  Single-GPU contexts, non-hybrid models, no fatigue. */
typedef struct {
    double c1, c2;            /* solver_ener_variational.py:57-58                                */
    int32_t variational;      /* 1: solver_ener_variational, 0: solver_ener_non_variational        */
    double newton_rtol, newton_atol;
    int32_t newton_max_it;
    double lin_rtol;          /* default 1e-10                                                     */
    int32_t gmres_restart;    /* default 40 (<= 62)                                                */
    int32_t lin_max_it;       /* default 2000                                                      */
    double inner_rtol;        /* default 0.1                                                       */
} pfx_ec_opts;

typedef struct {
    int32_t newton_its, converged;
    int32_t reason;           /* 1 converged; -5 max_it; -3 linear solve failed; -4 NaN */
    double residual, residual0;
    int64_t gmres_its, cg_its_u, cg_its_phi;
    double lambda_, gamma, external;   /* lambda, gamma(phi) and int T.u ds after the solve          */
    double gpu_ms;
} pfx_ec_report;

void pfx_ec_default_opts(pfx_ec_opts* o);
int pfx_ec_solve(pfx_ctx* ctx, const pfx_ec_opts* opts, double tau, double* lambda_inout, pfx_ec_report* out);
/* restore = 0: remember (u, phi) as the last converged state (u_c, phi_c of the reference, :367-369);
 * restore = 1: roll back to it (:489-491). */
int pfx_ec_checkpoint(pfx_ctx* ctx, int restore);

/* ---- per-step outputs --------------------------------------------------------------- */
/* pfx/Reactions/reactions_forces.py:12-65, called per BC at solver.py:428-431 */
int pfx_reaction(pfx_ctx* ctx, int bc_id, double R[3]);
/* total.energy columns of solver.py:455: EplusW,E,W,W_phi,W_gradphi,gamma,gamma_phi,gamma_gradphi,PSI_a,PSI_b,alpha_acum */
int pfx_energies(pfx_ctx* ctx, double out[11]);
/* fields in caller node order; `host` has n_nodes*(d or 1) doubles */
int pfx_get_field(pfx_ctx* ctx, int field, double* host, int64_t n);
int pfx_set_field(pfx_ctx* ctx, int field, const double* host, int64_t n);

/* ---- operator-level hooks (bench + kernel parity tests) ----------------------------- */
/* y = J_u(u,phi)·v (solver.py:186-187), J_phi(H,Fatigue)·v (:223), mass·v (Math/projection.py:35-40),
 * with Dirichlet rows/cols = identity when use_bc != 0 (dolfinx assemble_matrix(bcs)). */
int pfx_apply_u(pfx_ctx* ctx, const double* v, double* y, int use_bc);
int pfx_apply_phi(pfx_ctx* ctx, const double* v, double* y, int use_bc);
int pfx_apply_mass(pfx_ctx* ctx, const double* v, double* y);
int pfx_residual_u(pfx_ctx* ctx, double* f);                 /* F_u_form, solver.py:176-177   */
int pfx_diag_u(pfx_ctx* ctx, double* diag);                  /* Jacobi preconditioner data    */
int pfx_rhs_psi(pfx_ctx* ctx, double* b);                    /* ∫ psi_a(u) N_i, projection RHS */
/* nodal L2 projection of psi_a (which = 0) / psi_b (which = 1) of the current u, host array [n_nodes]
 * (Math/projection.py:19-59 as used for output by solver_anisotropic.py:236-237); single-GPU contexts */
int pfx_project_psi(pfx_ctx* ctx, int which, double* host);
int pfx_l2_error(pfx_ctx* ctx, int field_a, int field_b, double* err); /* errors_functions.py:190-228 */
/* Device-resident timing of `reps` back-to-back launches of one operator (kind: 0 apply_u,
 * 1 apply_phi, 2 apply_mass, 3 residual_u, 4 one full u-PCG iteration); CUDA events on the
 * context stream, L2 flushed between launches when flush_l2 != 0.  ms_out[reps]. */
int pfx_time_operator(pfx_ctx* ctx, int kind, int reps, int flush_l2, float* ms_out);
/* algorithmic bytes of one launch of `kind` for this mesh (SURVEY §8d formulas); kind 5: the bytes the u-apply
 * implementation actually streams (with cached element tangents: 168 B per tetrahedron + X, v, y per node) */
int pfx_algorithmic_bytes(pfx_ctx* ctx, int kind, double* bytes);
/* test hook: inverse of a dense symmetric positive definite matrix (n x n, column-major host arrays; the lower
 * triangle of the result is valid) by the in-tree blocked Cholesky that inverts the coarse matrices of the
 * two-level PCG — the MUMPS stand-in has no library kernel on its path.  info_out: 1 if a pivot was not positive. */
int pfx_dense_spd_inverse(int device, int n, const double* A_host, double* Ainv_host, int* info_out);
/* kernels launched by this context since creation (bench `gpu_launches`) */
int pfx_launch_count(pfx_ctx* ctx, int64_t* n);
int pfx_block_stats(pfx_ctx* ctx, int64_t out[8]);
/* aggregate coarse space of the PCG (the stand-in for the reference's MUMPS factorisation,
 * solver.py:198-205): out = {active (0/1), aggregates, colours of the probing set-up,
 * set-ups done so far}.  Inactive on meshes of fewer than 32 blocks or with PFX_NO_COARSE=1. */
int pfx_coarse_stats(pfx_ctx* ctx, int64_t out[4]);

/* ---- output formatting (host only; replaces the ASCII conversion inside dolfinx
 * VTKFile.write_function as called by reference solver.py:476-477) -------------------------
 * Write n_rows rows of ncomp_in doubles as shortest round-trip decimal text, each number
 * followed by one blank, rows padded with "0 " to ncomp_out components (dolfinx pads vectors
 * to 3).  Returns the bytes written, or -1 when cap < n_rows*ncomp_out*25 (i64: n*21). */
int64_t pfx_format_f64(const double* v, int64_t n_rows, int32_t ncomp_in, int32_t ncomp_out, char* out,
                       int64_t cap);
int64_t pfx_format_i64(const int64_t* v, int64_t n, char* out, int64_t cap);

/* ---- multi-GPU: partition + ghost maps (dolfinx IndexMap/Scatterer + graph partitioner) ---
 * Host-only, deterministic.  cell_rank_in may supply an external partition (replay of a
 * dolfinx partition) or be NULL (recursive coordinate bisection of cell centroids).
 * Two-call protocol: pfx_partition_create builds everything, the getters copy out. */
typedef struct pfx_partition pfx_partition;
int pfx_partition_create(int32_t cell_type, int64_t n_nodes, int64_t n_cells, const double* coords,
                         const int32_t* cells, int n_ranks, const int32_t* cell_rank_in,
                         pfx_partition** out);
int pfx_partition_destroy(pfx_partition* p);
/* Graph partition: cell_rank[n_cells] by METIS k-way (the toolkit's libmetis_static.a, default options, deterministic)
 * on the cell dual graph (cells adjacent when they share a facet) — what dolfinx's graph partitioner does at mesh
 * creation.  Pass the result to pfx_partition_create as cell_rank_in.  edge_cut (may be NULL): facets cut. */
int pfx_partition_graph_cell_rank(int32_t cell_type, int64_t n_nodes, int64_t n_cells, const int32_t* cells, int n_ranks,
                                  int32_t* cell_rank /* [n_cells] */, int64_t* edge_cut);
int pfx_partition_cell_rank(const pfx_partition* p, int32_t* cell_rank /* [n_cells] */);
int pfx_partition_node_owner(const pfx_partition* p, int32_t* node_owner /* [n_nodes] */);
/* sizes for rank r: out = {n_local_nodes, n_owned, n_local_cells, n_send_total, n_neighbors} */
int pfx_partition_rank_sizes(const pfx_partition* p, int rank, int64_t out[5]);
/* local->global node map (owned first, ghosts grouped by owner rank ascending, each group by
 * global id), local cells as global cell ids, ghost owners, primary flags */
int pfx_partition_rank_maps(const pfx_partition* p, int rank, int64_t* local_to_global_node,
                            int64_t* local_cell_global, int32_t* ghost_owner, uint8_t* cell_primary);
/* exchange plan of rank r: neighbors[n_neighbors], send_offsets[n_neighbors+1] into
 * send_local_idx[n_send_total] (local owned indices to pack), recv_offsets[n_neighbors+1]
 * (ghost ranges, relative to n_owned) */
int pfx_partition_rank_plan(const pfx_partition* p, int rank, int32_t* neighbors, int64_t* send_offsets,
                            int32_t* send_local_idx, int64_t* recv_offsets);

/* Attach the exchange plan + NCCL communicator to a context (one process per GPU).
 * nccl_unique_id: 128 bytes from pfx_comm_unique_id() on rank 0, broadcast by the host
 * (torch.distributed).  After this call halo exchange and scalar all-reduces run inside
 * pfx_load_step / pfx_apply_* on the context stream. */
int pfx_comm_unique_id(uint8_t id[128]);
int pfx_comm_attach(pfx_ctx* ctx, int rank, int n_ranks, const uint8_t id[128], int n_neighbors,
                    const int32_t* neighbors, const int64_t* send_offsets, const int32_t* send_local_idx,
                    const int64_t* recv_offsets);

/* NVLink peer path (optional, after pfx_comm_attach): map the peers' exchange buffers with CUDA
 * IPC so that halo values are pushed with direct stores into the neighbours' ghost ranges and the
 * CG scalars are all-reduced through mailboxes — no NCCL call inside the PCG iteration.
 * pfx_comm_p2p_export fills PFX_P2P_HANDLE_BYTES bytes; the host all-gathers them (rank-major) and
 * passes the table to pfx_comm_p2p_import together with peer_ghost_start[j] = position (in nodes,
 * in neighbour j's local numbering) where this rank's values start in neighbour j's ghost range. */
#define PFX_P2P_HANDLE_BYTES 384
int pfx_comm_p2p_export(pfx_ctx* ctx, uint8_t* handles /* [PFX_P2P_HANDLE_BYTES] */);
int pfx_comm_p2p_import(pfx_ctx* ctx, const uint8_t* all_handles /* [n_ranks][PFX_P2P_HANDLE_BYTES] */,
                        const int64_t* peer_ghost_start /* [n_neighbors] */);

#ifdef __cplusplus
}
#endif
#endif /* PFX_B200_H */
