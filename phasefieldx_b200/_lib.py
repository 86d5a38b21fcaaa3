"""
ctypes binding of libpfx_b200.so (include/pfx_b200.h) and a thin numpy-facing `Engine`.

There is no CPU fallback: if the shared object is missing or no CUDA device is visible the
calls raise.  The oracle under /oracle is test infrastructure and is never imported here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfx_b200.so")

CELL = {"triangle": 0, "quadrilateral": 1, "tetrahedron": 2, "hexahedron": 3}
FIELD = {"u": 0, "phi": 1, "H": 2, "V_c": 3, "V_n": 4, "fatigue": 5, "alpha_cum_bar_c": 6, "alpha_c": 7,
         "u_old": 8, "phi_old": 9, "alpha_n": 10, "alpha_cum_bar_n": 11}
ENERGY_COLUMNS = ("EplusW", "E", "W", "W_phi", "W_gradphi", "gamma", "gamma_phi", "gamma_gradphi", "PSI_a", "PSI_b",
                  "alpha_acum")

PFX_ERR_NOT_CONVERGED_U, PFX_ERR_NOT_CONVERGED_PHI = 3, 4


class PfxMesh(C.Structure):
    _fields_ = [("cell_type", C.c_int32), ("n_nodes", C.c_int64), ("n_cells", C.c_int64),
                ("coords", C.POINTER(C.c_double)), ("cells", C.POINTER(C.c_int32)), ("n_owned", C.c_int64),
                ("cell_primary", C.POINTER(C.c_uint8))]


class PfxParams(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("mu", C.c_double), ("Gc", C.c_double), ("l", C.c_double), ("k", C.c_double),
                ("model", C.c_int32), ("irreversibility", C.c_int32), ("fatigue", C.c_int32),
                ("fatigue_val", C.c_double), ("snes_rtol", C.c_double), ("snes_atol", C.c_double),
                ("snes_stol", C.c_double), ("snes_max_it", C.c_int32), ("cg_rtol", C.c_double),
                ("cg_max_it", C.c_int32), ("proj_rtol", C.c_double), ("quad_points_1d", C.c_int32),
                ("block_nodes", C.c_int32), ("cg_chunk", C.c_int32), ("degradation_function", C.c_int32),
                ("coarse_region", C.c_int32), ("coarse_agg_blocks", C.c_int32)]


class PfxStepOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("min_stagger_iter", C.c_int32), ("max_stagger_iter", C.c_int32),
                ("stagger_tol", C.c_double), ("history_capacity", C.c_int32)]


class PfxStepReport(C.Structure):
    _fields_ = [("stagger_iters", C.c_int32), ("u_newton_its", C.c_int32), ("phi_newton_its", C.c_int32),
                ("u_reason", C.c_int32), ("phi_reason", C.c_int32), ("err_u", C.c_double), ("err_phi", C.c_double),
                ("fnorm_u", C.c_double), ("fnorm_phi", C.c_double), ("cg_its_u", C.c_int64), ("cg_its_phi", C.c_int64),
                ("cg_its_proj", C.c_int64), ("gpu_ms", C.c_double), ("history_len", C.c_int32), ("hist_u_its", C.POINTER(C.c_int32)),
                ("hist_phi_its", C.POINTER(C.c_int32)), ("hist_err_u", C.POINTER(C.c_double)),
                ("hist_err_phi", C.POINTER(C.c_double)), ("hist_fnorm_u", C.POINTER(C.c_double)),
                ("hist_fnorm_phi", C.POINTER(C.c_double)), ("phase_ms", C.c_double * 10), ("cg_floor_solves", C.c_int64)]


class PfxEcOpts(C.Structure):
    _fields_ = [("c1", C.c_double), ("c2", C.c_double), ("variational", C.c_int32), ("newton_rtol", C.c_double),
                ("newton_atol", C.c_double), ("newton_max_it", C.c_int32), ("lin_rtol", C.c_double), ("gmres_restart", C.c_int32),
                ("lin_max_it", C.c_int32), ("inner_rtol", C.c_double)]


class PfxEcReport(C.Structure):
    _fields_ = [("newton_its", C.c_int32), ("converged", C.c_int32), ("reason", C.c_int32), ("residual", C.c_double),
                ("residual0", C.c_double), ("gmres_its", C.c_int64), ("cg_its_u", C.c_int64), ("cg_its_phi", C.c_int64),
                ("lambda_", C.c_double), ("gamma", C.c_double), ("external", C.c_double), ("gpu_ms", C.c_double)]


PHASES = ("u_setup", "u_coarse_setup", "u_pcg", "projection", "fatigue", "phi_setup", "phi_coarse_setup", "phi_pcg", "norms",
          "other")


# every symbol include/pfx_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "pfx_version", "pfx_device_count", "pfx_status_string", "pfx_create", "pfx_destroy", "pfx_last_error",
    "pfx_default_params", "pfx_set_bc_u", "pfx_set_bc_values_u", "pfx_set_bc_phi", "pfx_set_bc_values_phi",
    "pfx_set_traction", "pfx_set_traction_value", "pfx_load_step", "pfx_solve_u", "pfx_solve_phi", "pfx_ec_default_opts", "pfx_ec_solve",
    "pfx_ec_checkpoint", "pfx_reaction", "pfx_energies", "pfx_get_field",
    "pfx_set_field", "pfx_apply_u", "pfx_apply_phi", "pfx_apply_mass", "pfx_residual_u", "pfx_diag_u", "pfx_rhs_psi", "pfx_project_psi",
    "pfx_l2_error", "pfx_dense_spd_inverse", "pfx_time_operator", "pfx_algorithmic_bytes", "pfx_launch_count", "pfx_block_stats", "pfx_coarse_stats",
    "pfx_partition_create", "pfx_partition_destroy", "pfx_partition_graph_cell_rank", "pfx_partition_cell_rank", "pfx_partition_node_owner",
    "pfx_partition_rank_sizes", "pfx_partition_rank_maps", "pfx_partition_rank_plan", "pfx_comm_unique_id",
    "pfx_comm_attach", "pfx_comm_p2p_export", "pfx_comm_p2p_import", "pfx_format_f64", "pfx_format_i64",
]

_lib = None


def load():
    """Load libpfx_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m phasefieldx_b200.build` "
                           "(phasefieldx_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.pfx_version.restype = C.c_char_p
    lib.pfx_status_string.restype = C.c_char_p
    lib.pfx_status_string.argtypes = [C.c_int]
    lib.pfx_last_error.restype = C.c_char_p
    lib.pfx_last_error.argtypes = [C.c_void_p]
    lib.pfx_create.argtypes = [C.POINTER(PfxMesh), C.POINTER(PfxParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.pfx_destroy.argtypes = [C.c_void_p]
    lib.pfx_default_params.argtypes = [C.POINTER(PfxParams)]
    lib.pfx_default_params.restype = None
    pd, pi, p8 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    pl = C.POINTER(C.c_int64)
    for name in ("pfx_set_bc_u", "pfx_set_bc_phi"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int, pi, C.c_int64]
    for name in ("pfx_set_bc_values_u", "pfx_set_bc_values_phi"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int, pd, C.c_int64]
    lib.pfx_set_traction.argtypes = [C.c_void_p, C.c_int, pi, pd, C.c_int64]
    lib.pfx_set_traction_value.argtypes = [C.c_void_p, C.c_int, pd]
    lib.pfx_load_step.argtypes = [C.c_void_p, C.POINTER(PfxStepOpts), C.POINTER(PfxStepReport)]
    lib.pfx_project_psi.argtypes = [C.c_void_p, C.c_int, pd]
    lib.pfx_solve_u.argtypes = [C.c_void_p, C.POINTER(PfxStepReport)]
    lib.pfx_solve_phi.argtypes = [C.c_void_p, C.POINTER(PfxStepReport)]
    lib.pfx_ec_default_opts.argtypes = [C.POINTER(PfxEcOpts)]
    lib.pfx_ec_default_opts.restype = None
    lib.pfx_ec_solve.argtypes = [C.c_void_p, C.POINTER(PfxEcOpts), C.c_double, pd, C.POINTER(PfxEcReport)]
    lib.pfx_ec_checkpoint.argtypes = [C.c_void_p, C.c_int]
    lib.pfx_reaction.argtypes = [C.c_void_p, C.c_int, pd]
    lib.pfx_energies.argtypes = [C.c_void_p, pd]
    lib.pfx_get_field.argtypes = [C.c_void_p, C.c_int, pd, C.c_int64]
    lib.pfx_set_field.argtypes = [C.c_void_p, C.c_int, pd, C.c_int64]
    lib.pfx_apply_u.argtypes = [C.c_void_p, pd, pd, C.c_int]
    lib.pfx_apply_phi.argtypes = [C.c_void_p, pd, pd, C.c_int]
    lib.pfx_apply_mass.argtypes = [C.c_void_p, pd, pd]
    lib.pfx_residual_u.argtypes = [C.c_void_p, pd]
    lib.pfx_diag_u.argtypes = [C.c_void_p, pd]
    lib.pfx_rhs_psi.argtypes = [C.c_void_p, pd]
    lib.pfx_l2_error.argtypes = [C.c_void_p, C.c_int, C.c_int, pd]
    lib.pfx_time_operator.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.pfx_algorithmic_bytes.argtypes = [C.c_void_p, C.c_int, pd]
    lib.pfx_dense_spd_inverse.argtypes = [C.c_int, C.c_int, pd, pd, C.POINTER(C.c_int)]
    lib.pfx_launch_count.argtypes = [C.c_void_p, pl]
    lib.pfx_block_stats.argtypes = [C.c_void_p, pl]
    lib.pfx_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.pfx_partition_create.argtypes = [C.c_int32, C.c_int64, C.c_int64, pd, pi, C.c_int, pi, C.POINTER(C.c_void_p)]
    lib.pfx_partition_destroy.argtypes = [C.c_void_p]
    lib.pfx_partition_graph_cell_rank.argtypes = [C.c_int32, C.c_int64, C.c_int64, pi, C.c_int, pi, pl]
    lib.pfx_partition_cell_rank.argtypes = [C.c_void_p, pi]
    lib.pfx_partition_node_owner.argtypes = [C.c_void_p, pi]
    lib.pfx_partition_rank_sizes.argtypes = [C.c_void_p, C.c_int, pl]
    lib.pfx_partition_rank_maps.argtypes = [C.c_void_p, C.c_int, pl, pl, pi, p8]
    lib.pfx_partition_rank_plan.argtypes = [C.c_void_p, C.c_int, pi, pl, pi, pl]
    lib.pfx_comm_unique_id.argtypes = [p8]
    lib.pfx_comm_attach.argtypes = [C.c_void_p, C.c_int, C.c_int, p8, C.c_int, pi, pl, pi, pl]
    lib.pfx_comm_p2p_export.argtypes = [C.c_void_p, p8]
    lib.pfx_comm_p2p_import.argtypes = [C.c_void_p, p8, pl]
    lib.pfx_coarse_stats.argtypes = [C.c_void_p, pl]
    lib.pfx_format_f64.restype = C.c_int64
    lib.pfx_format_f64.argtypes = [pd, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
    lib.pfx_format_i64.restype = C.c_int64
    lib.pfx_format_i64.argtypes = [pl, C.c_int64, C.c_void_p, C.c_int64]
    _lib = lib
    return lib


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _pl(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


MODEL = {("isotropic", None): 0, ("anisotropic", "spectral"): 1, ("anisotropic", "deviatoric"): 2,
         ("hybrid", "spectral"): 3, ("hybrid", "deviatoric"): 4}


DEGRADATION = {"quadratic": 0, "borden": 1}


def check_degradation_function(name):
    """What the reference does with each `Input.degradation_function` (g_degradation_functions.py:127-174):
    quadratic, borden -> on the GPU path; sargado -> its derivative evaluates exp(100 * 100) while the forms are
    built, so the reference raises OverflowError before the first step (:97-124) and so does
    this; alessi -> as coded its derivative has the wrong sign and a pole at phi = 0.01 (:67-94) — not restated;
    anything else -> the reference's ValueError."""
    if name in DEGRADATION:
        return
    if name == "sargado":
        raise OverflowError("math range error")
    if name == "alessi":
        raise NotImplementedError("degradation_function 'alessi' is not on the GPU path: as coded in the reference its derivative "
                                  "has the wrong sign and a pole at phi = 0.01 (g_degradation_functions.py:67-94)")
    raise ValueError("Invalid degradation_type. Please choose from 'quadratic', 'borden', 'alessi', or 'sargado'.")


def model_code(degradation, split_energy):
    """Dispatch of reference split_energy_stress_tangent_functions.py:234-354."""
    if degradation == "isotropic":
        return 0
    try:
        return MODEL[(degradation, split_energy)]
    except KeyError:
        raise ValueError("Invalid")


class PfxError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def graph_cell_rank(cell_type, n_nodes, cells, n_ranks):
    """(cell_rank, edge_cut) of the METIS k-way partition of the cell dual graph (pfx_partition_graph_cell_rank)."""
    lib = load()
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    out = np.zeros(len(cells), dtype=np.int32)
    cut = np.zeros(1, dtype=np.int64)
    st = lib.pfx_partition_graph_cell_rank(CELL[cell_type], int(n_nodes), len(cells), _pi(cells), int(n_ranks), _pi(out), _pl(cut))
    if st:
        raise PfxError(st, "graph partition failed")
    return out, int(cut[0])


class Partition:
    """Host-only rank partition + ghost maps (pfx_partition_* of the C-ABI)."""

    def __init__(self, cell_type, coords, cells, n_ranks, cell_rank=None):
        self.lib = load()
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.n_ranks = n_ranks
        self.n_nodes, self.n_cells = len(coords), len(cells)
        h = C.c_void_p()
        cr = None if cell_rank is None else np.ascontiguousarray(cell_rank, dtype=np.int32)
        st = self.lib.pfx_partition_create(CELL[cell_type], self.n_nodes, self.n_cells, _pd(coords), _pi(cells), n_ranks,
                                           None if cr is None else _pi(cr), C.byref(h))
        if st:
            raise PfxError(st, "pfx_partition_create failed")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.pfx_partition_destroy(self.h)
            self.h = None

    def cell_rank(self):
        out = np.empty(self.n_cells, dtype=np.int32)
        self.lib.pfx_partition_cell_rank(self.h, _pi(out))
        return out

    def node_owner(self):
        out = np.empty(self.n_nodes, dtype=np.int32)
        self.lib.pfx_partition_node_owner(self.h, _pi(out))
        return out

    def rank(self, r):
        sz = np.zeros(5, dtype=np.int64)
        self.lib.pfx_partition_rank_sizes(self.h, r, _pl(sz))
        n_local, n_owned, n_cells, n_send, n_nb = (int(v) for v in sz)
        l2g = np.empty(n_local, dtype=np.int64)
        cell_gid = np.empty(n_cells, dtype=np.int64)
        ghost_owner = np.empty(n_local - n_owned, dtype=np.int32)
        primary = np.empty(n_cells, dtype=np.uint8)
        self.lib.pfx_partition_rank_maps(self.h, r, _pl(l2g), _pl(cell_gid), _pi(ghost_owner), _p8(primary))
        nbrs = np.empty(n_nb, dtype=np.int32)
        send_off = np.zeros(n_nb + 1, dtype=np.int64)
        recv_off = np.zeros(n_nb + 1, dtype=np.int64)
        send_idx = np.empty(n_send, dtype=np.int32)
        self.lib.pfx_partition_rank_plan(self.h, r, _pi(nbrs), _pl(send_off), _pi(send_idx), _pl(recv_off))
        return dict(n_owned=n_owned, l2g=l2g, cell_gid=cell_gid, ghost_owner=ghost_owner, cell_primary=primary,
                    neighbors=nbrs, send_off=send_off, send_idx=send_idx, recv_off=recv_off)


class Engine:
    """One pfx_ctx: device-resident state of a staggered phase-field fracture problem."""

    def __init__(self, coords, cells, cell_type, lambda_, mu, Gc, l, k=0.0, model=0, irreversibility=False,
                 fatigue=False, fatigue_val=0.05625, device=0, n_owned=0, cell_primary=None, degradation_function="quadratic",
                 **solver_opts):
        self.lib = load()
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        if self.coords.ndim != 2 or self.coords.shape[1] != 3:
            raise ValueError("coords must be [N,3]")
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.cell_type = cell_type
        self.dim = 3 if cell_type in ("tetrahedron", "hexahedron") else 2
        self.n_nodes = len(self.coords)
        p = PfxParams()
        self.lib.pfx_default_params(C.byref(p))
        p.lambda_, p.mu, p.Gc, p.l, p.k = lambda_, mu, Gc, l, k
        p.model, p.irreversibility, p.fatigue, p.fatigue_val = model, int(irreversibility), int(fatigue), fatigue_val
        try:
            p.degradation_function = DEGRADATION[degradation_function]
        except KeyError:
            raise NotImplementedError(f"degradation_function {degradation_function!r}: only 'quadratic' and 'borden' are on "
                                      "the GPU path (SURVEY §8f N2)")
        for key, val in solver_opts.items():
            if not hasattr(p, key):
                raise TypeError(f"unknown solver option {key!r}")
            setattr(p, key, val)
        self.params = p
        m = PfxMesh()
        m.cell_type = CELL[cell_type]
        m.n_nodes, m.n_cells = self.n_nodes, len(self.cells)
        m.coords, m.cells = _pd(self.coords), _pi(self.cells)
        m.n_owned = n_owned
        self._prim = None if cell_primary is None else np.ascontiguousarray(cell_primary, dtype=np.uint8)
        m.cell_primary = None if self._prim is None else _p8(self._prim)
        h = C.c_void_p()
        st = self.lib.pfx_create(C.byref(m), C.byref(p), device, C.byref(h))
        self.h = h if h.value else None
        if st:
            msg = self._err() if self.h else self.lib.pfx_status_string(st).decode()
            self.close()
            raise PfxError(st, f"pfx_create failed: {msg}")
        self._bc_len = {}

    # ---------------------------------------------------------------- plumbing
    def _err(self):
        return self.lib.pfx_last_error(self.h).decode()

    def _ck(self, st):
        if st:
            raise PfxError(st, f"{self.lib.pfx_status_string(st).decode()}: {self._err()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pfx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- boundary data
    def set_bc_u(self, bc_id, dofs):
        d = np.ascontiguousarray(dofs, dtype=np.int32)
        self._ck(self.lib.pfx_set_bc_u(self.h, bc_id, _pi(d), len(d)))

    def set_bc_values_u(self, bc_id, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self.lib.pfx_set_bc_values_u(self.h, bc_id, _pd(v), len(v)))

    def set_bc_phi(self, bc_id, dofs):
        d = np.ascontiguousarray(dofs, dtype=np.int32)
        self._ck(self.lib.pfx_set_bc_phi(self.h, bc_id, _pi(d), len(d)))

    def set_bc_values_phi(self, bc_id, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self.lib.pfx_set_bc_values_phi(self.h, bc_id, _pd(v), len(v)))

    def set_traction(self, t_id, nodes, weights):
        n = np.ascontiguousarray(nodes, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self._ck(self.lib.pfx_set_traction(self.h, t_id, _pi(n), _pd(w), len(n)))

    def set_traction_value(self, t_id, T):
        t = np.zeros(3)
        t[: len(T)] = T
        self._ck(self.lib.pfx_set_traction_value(self.h, t_id, _pd(t)))

    # ---------------------------------------------------------------- hot path
    def load_step(self, dt=1.0, min_stagger_iter=2, max_stagger_iter=10000, stagger_tol=1e-8, history=64):
        o = PfxStepOpts(dt, min_stagger_iter, max_stagger_iter, stagger_tol, history)
        r = PfxStepReport()
        hist = {"u_its": np.zeros(history, np.int32), "phi_its": np.zeros(history, np.int32),
                "err_u": np.zeros(history), "err_phi": np.zeros(history), "fnorm_u": np.zeros(history),
                "fnorm_phi": np.zeros(history)}
        r.hist_u_its, r.hist_phi_its = _pi(hist["u_its"]), _pi(hist["phi_its"])
        r.hist_err_u, r.hist_err_phi = _pd(hist["err_u"]), _pd(hist["err_phi"])
        r.hist_fnorm_u, r.hist_fnorm_phi = _pd(hist["fnorm_u"]), _pd(hist["fnorm_phi"])
        st = self.lib.pfx_load_step(self.h, C.byref(o), C.byref(r))
        rep = dict(stagger=r.stagger_iters, u_its=r.u_newton_its, phi_its=r.phi_newton_its, u_reason=r.u_reason,
                   phi_reason=r.phi_reason, err_u=r.err_u, err_phi=r.err_phi, fnorm_u=r.fnorm_u, fnorm_phi=r.fnorm_phi,
                   cg_its_u=r.cg_its_u, cg_its_phi=r.cg_its_phi, cg_its_proj=r.cg_its_proj, gpu_ms=r.gpu_ms,
                   phase_ms=dict(zip(PHASES, (float(v) for v in r.phase_ms))), cg_floor_solves=r.cg_floor_solves,
                   history={k: v[: r.history_len].copy() for k, v in hist.items()}, status=st)
        if st in (PFX_ERR_NOT_CONVERGED_U, PFX_ERR_NOT_CONVERGED_PHI):
            # reference solver.py:343-344 / :389-390
            reason = r.u_reason if st == PFX_ERR_NOT_CONVERGED_U else r.phi_reason
            raise RuntimeError(f"Solver did not converge, got {reason}.")
        self._ck(st)
        return rep

    def _solve_single(self, fn, is_u):
        r = PfxStepReport()
        st = fn(self.h, C.byref(r))
        if st in (PFX_ERR_NOT_CONVERGED_U, PFX_ERR_NOT_CONVERGED_PHI):
            raise RuntimeError(f"Solver did not converge, got {r.u_reason if is_u else r.phi_reason}.")
        self._ck(st)
        if is_u:
            return dict(its=r.u_newton_its, reason=r.u_reason, fnorm=r.fnorm_u, cg_its=r.cg_its_u, gpu_ms=r.gpu_ms,
                        cg_floor_solves=r.cg_floor_solves)
        return dict(its=r.phi_newton_its, reason=r.phi_reason, fnorm=r.fnorm_phi, cg_its=r.cg_its_phi, gpu_ms=r.gpu_ms,
                    cg_floor_solves=r.cg_floor_solves)

    def project_psi(self, which):
        """Nodal L2 projection of psi_a ("a") or psi_b ("b") of the current displacement field."""
        out = np.zeros(self.n_nodes)
        self._ck(self.lib.pfx_project_psi(self.h, {"a": 0, "b": 1}[which], _pd(out)))
        return out

    def solve_u(self):
        """One Newton (SNES) solve of the displacement problem at the current phi."""
        return self._solve_single(self.lib.pfx_solve_u, True)

    def solve_phi(self):
        """One Newton (SNES) solve of the phase-field problem at the current H / fatigue field."""
        return self._solve_single(self.lib.pfx_solve_phi, False)

    def ec_solve(self, tau, lam, c1=1.0, c2=1.0, variational=True, **opts):
        """One Newton solve of the energy-controlled blocked system (u, phi, lambda) at `tau`
        (reference solver_ener_[non_]variational.py, `solver_snap.solve()`); returns the report with the new lambda.
        opts: newton_rtol, newton_atol, newton_max_it, lin_rtol, gmres_restart, lin_max_it, inner_rtol."""
        o = PfxEcOpts()
        self.lib.pfx_ec_default_opts(C.byref(o))
        o.c1, o.c2, o.variational = float(c1), float(c2), int(bool(variational))
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown energy-controlled solver option {k!r}")
            setattr(o, k, v)
        lam_io = np.array([float(lam)])
        r = PfxEcReport()
        self._ck(self.lib.pfx_ec_solve(self.h, C.byref(o), float(tau), _pd(lam_io), C.byref(r)))
        return dict(its=r.newton_its, converged=bool(r.converged), reason=r.reason, residual=r.residual, residual0=r.residual0,
                    gmres_its=r.gmres_its, cg_its_u=r.cg_its_u, cg_its_phi=r.cg_its_phi, lam=r.lambda_, gamma=r.gamma,
                    external=r.external, gpu_ms=r.gpu_ms)

    def ec_checkpoint(self, restore=False):
        """Remember (restore=False) or roll back to (restore=True) the last converged (u, phi)."""
        self._ck(self.lib.pfx_ec_checkpoint(self.h, 1 if restore else 0))

    # ---------------------------------------------------------------- outputs
    def reaction(self, bc_id):
        R = np.zeros(3)
        self._ck(self.lib.pfx_reaction(self.h, bc_id, _pd(R)))
        return R

    def energies(self):
        e = np.zeros(11)
        self._ck(self.lib.pfx_energies(self.h, _pd(e)))
        return dict(zip(ENERGY_COLUMNS, e.tolist()))

    def get_field(self, name):
        nc = self.dim if name in ("u", "u_old") else 1
        out = np.zeros(self.n_nodes * nc)
        self._ck(self.lib.pfx_get_field(self.h, FIELD[name], _pd(out), len(out)))
        return out

    def set_field(self, name, values):
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        self._ck(self.lib.pfx_set_field(self.h, FIELD[name], _pd(v), len(v)))

    # ---------------------------------------------------------------- operator hooks
    def _op(self, fn, v, n, *extra):
        y = np.zeros(n)
        if v is None:
            self._ck(fn(self.h, _pd(y), *extra))
        else:
            v = np.ascontiguousarray(v, dtype=np.float64)
            self._ck(fn(self.h, _pd(v), _pd(y), *extra))
        return y

    def apply_u(self, v, use_bc=False):
        return self._op(self.lib.pfx_apply_u, v, self.n_nodes * self.dim, int(use_bc))

    def apply_phi(self, v, use_bc=False):
        return self._op(self.lib.pfx_apply_phi, v, self.n_nodes, int(use_bc))

    def apply_mass(self, v):
        return self._op(self.lib.pfx_apply_mass, v, self.n_nodes)

    def residual_u(self):
        return self._op(self.lib.pfx_residual_u, None, self.n_nodes * self.dim)

    def diag_u(self):
        return self._op(self.lib.pfx_diag_u, None, self.n_nodes * self.dim)

    def rhs_psi(self):
        return self._op(self.lib.pfx_rhs_psi, None, self.n_nodes)

    def l2_error(self, a, b):
        e = C.c_double()
        self._ck(self.lib.pfx_l2_error(self.h, FIELD[a], FIELD[b], C.byref(e)))
        return e.value

    def time_operator(self, kind, reps=10, flush_l2=True):
        ms = (C.c_float * reps)()
        self._ck(self.lib.pfx_time_operator(self.h, kind, reps, int(flush_l2), ms))
        return np.array(ms[:], dtype=np.float64)

    def algorithmic_bytes(self, kind):
        b = C.c_double()
        self._ck(self.lib.pfx_algorithmic_bytes(self.h, kind, C.byref(b)))
        return b.value

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.lib.pfx_launch_count(self.h, C.byref(n)))
        return n.value

    def block_stats(self):
        s = np.zeros(8, dtype=np.int64)
        self._ck(self.lib.pfx_block_stats(self.h, _pl(s)))
        return dict(zip(("n_blocks", "max_owned", "max_local", "max_elems", "elems_with_overlap", "n_cells",
                         "halo_entries", "n_owned"), (int(v) for v in s)))

    def coarse_stats(self):
        s = np.zeros(4, dtype=np.int64)
        self._ck(self.lib.pfx_coarse_stats(self.h, _pl(s)))
        return dict(zip(("active", "aggregates", "colours", "setups"), (int(v) for v in s)))

    # ---------------------------------------------------------------- multi-GPU
    def comm_attach(self, rank, n_ranks, unique_id, plan):
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        nb = np.ascontiguousarray(plan["neighbors"], dtype=np.int32)
        so = np.ascontiguousarray(plan["send_off"], dtype=np.int64)
        si = np.ascontiguousarray(plan["send_idx"], dtype=np.int32)
        ro = np.ascontiguousarray(plan["recv_off"], dtype=np.int64)
        self._ck(self.lib.pfx_comm_attach(self.h, rank, n_ranks, _p8(uid), len(nb), _pi(nb), _pl(so), _pi(si), _pl(ro)))

    def p2p_export(self):
        h = np.zeros(P2P_HANDLE_BYTES, dtype=np.uint8)
        self._ck(self.lib.pfx_comm_p2p_export(self.h, _p8(h)))
        return h.tobytes()

    def p2p_import(self, all_handles, ghost_start):
        a = np.frombuffer(b"".join(all_handles), dtype=np.uint8).copy()
        g = np.ascontiguousarray(ghost_start, dtype=np.int64)
        self._ck(self.lib.pfx_comm_p2p_import(self.h, _p8(a), _pl(g)))


P2P_HANDLE_BYTES = 384


def peer_ghost_start(partition, rank):
    """For every neighbour q of `rank` (ascending): position, in q's local node numbering, where
    the values `rank` owns start inside q's ghost range (host-side, from the deterministic partition)."""
    R = partition.rank(rank)
    out = []
    for q in R["neighbors"]:
        Q = partition.rank(int(q))
        j = int(np.nonzero(Q["neighbors"] == rank)[0][0])
        out.append(Q["n_owned"] + int(Q["recv_off"][j]))
    return np.array(out, dtype=np.int64)


def comm_unique_id():
    lib = load()
    uid = np.zeros(128, dtype=np.uint8)
    st = lib.pfx_comm_unique_id(_p8(uid))
    if st:
        raise PfxError(st, "pfx_comm_unique_id failed (NCCL not loadable?)")
    return uid.tobytes()


def format_f64(values, ncomp_out=None):
    """Shortest round-trip ASCII of a [N] or [N, ncomp] float64 array (blank separated, rows
    padded with zeros to ncomp_out) as bytes — the VTU writer's number formatting."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    n = v.shape[0]
    nc = 1 if v.ndim == 1 else v.shape[1]
    nco = nc if ncomp_out is None else int(ncomp_out)
    buf = np.empty(max(1, n * nco * 25), dtype=np.uint8)
    m = load().pfx_format_f64(_pd(v), n, nc, nco, buf.ctypes.data, buf.size)
    if m < 0:
        raise PfxError(2, "pfx_format_f64: bad arguments")
    return buf[:max(0, m - 1)].tobytes()


def format_i64(values):
    v = np.ascontiguousarray(values, dtype=np.int64).ravel()
    buf = np.empty(max(1, v.size * 21), dtype=np.uint8)
    m = load().pfx_format_i64(_pl(v), v.size, buf.ctypes.data, buf.size)
    if m < 0:
        raise PfxError(2, "pfx_format_i64: bad arguments")
    return buf[:max(0, m - 1)].tobytes()


def dense_spd_inverse(A, device=0):
    """(A^-1 as a full symmetric matrix, info) by the library's own blocked Cholesky (pfx_dense_spd_inverse)."""
    lib = load()
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    out = np.zeros((n, n), order="F")
    info = C.c_int(0)
    st = lib.pfx_dense_spd_inverse(device, n, A.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(info))
    if st:
        raise PfxError(st, "dense inverse failed")
    low = np.tril(out)
    return low + np.tril(out, -1).T, info.value


def device_count():
    n = C.c_int(0)
    load().pfx_device_count(C.byref(n))
    return n.value
