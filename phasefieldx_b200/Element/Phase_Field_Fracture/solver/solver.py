"""
Staggered phase-field fracture / fatigue solver — drop-in for reference
`src/phasefieldx/Element/Phase_Field_Fracture/solver/solver.py:43-480`.

Same signature, same result files (`phasefieldx.conv`, `top.dof`, `<bc>.reaction`,
`total.energy`, `parameters.input`, `simulation.log`, `paraview-solutions_vtu/*.vtu|pvd`), same
`RuntimeError` on non-convergence.  The time loop, user callbacks, logging and file writing stay
on the host (as in the reference); each load step is ONE call into the CUDA library
(`pfx_load_step`, include/pfx_b200.h), which runs the whole stagger loop of solver.py:319-411 on
the GPU.  There is no CPU path: without libpfx_b200.so or without a CUDA device this raises.
"""
import os
import time

import numpy as np

from phasefieldx_b200 import _lib, fem
from phasefieldx_b200.files import append_results_to_file, prepare_simulation
from phasefieldx_b200.io_vtk import VTKFile
from phasefieldx_b200.io_xdmf import XDMFFile
from phasefieldx_b200.Logger.library_versions import (log_end_analysis, log_library_versions, log_model_information,
                                                      log_system_info, set_logger)


class _NullLogger:
    def info(self, *a, **k):
        pass

    error = info


def _traction_sets(T_list_u):
    sets = []
    for T, ds in (T_list_u or []):
        if not isinstance(ds, fem.Measure):
            raise NotImplementedError("tractions need a measure from get_ds_bound_from_marker")
        nodes, w = ds.nodal_weights()
        sets.append((T, nodes, w))
    return sets


def solve(Data, msh, final_time, V_u, V_Φ, bc_list_u=[], bc_list_phi=[], update_boundary_conditions=None,
          f_list_u=None, T_list_u=None, update_loading=None, ds_bound=None, dt=1.0, path=None,
          bcs_list_u_names=None, min_stagger_iter=2, max_stagger_iter=10000, stagger_error_tol=1e-8,
          device=0, solver_options=None, vtu_every=1, vtu_format="ascii", vtu_parallel=False, vtu_keep=None,
          step_callback=None):
    """Solve the phase-field fracture and fatigue problem (reference signature; `f_list_u` and
    `ds_bound` are accepted and unused exactly as in the reference).
    Extra keyword-only knobs (not in the reference): `device`, `solver_options` (pfx_params
    overrides such as cg_rtol), `vtu_every` (VTU cadence; default every step as the reference),
    `vtu_format` ("ascii" = the reference's layout, "binary" = raw appended data), `vtu_parallel` (multi-rank
    runs: every rank writes its own piece + a .pvtu index, dolfinx's parallel layout, instead of gathering the
    fields on rank 0), `vtu_keep` (keep only the k most recent step files on disk; default: all, as the reference),
    `step_callback(step, report, reactions, energies)` called on every rank after the outputs of a load step."""
    if path is None:
        path = os.getcwd()
    result_folder_name = Data.results_folder_name
    # one process per GPU under torchrun (the reference's `mpirun`): rank 0 owns the files
    dist, rank, world = None, 0, 1
    try:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized() and _dist.get_world_size() > 1:
            dist, rank, world = _dist, _dist.get_rank(), _dist.get_world_size()
            device = int(os.environ.get("LOCAL_RANK", device))
    except ImportError:
        pass
    if rank == 0:
        prepare_simulation(path, result_folder_name)
    if dist is not None:
        dist.barrier()
    logger = set_logger(result_folder_name) if rank == 0 else _NullLogger()
    log_system_info(logger)
    log_library_versions(logger)
    Data.save_log_info(logger)
    if rank == 0:
        Data.save_parameters_to_csv(os.path.join(result_folder_name, "parameters.input"))
    log_model_information(msh, logger)
    logger.info("========== Stagger settings ===========")
    logger.info(f"  minimum stagger iterations: {min_stagger_iter}")
    logger.info(f"  maximum stagger iterations: {max_stagger_iter}")
    logger.info(f"  stagger error tolerance: {stagger_error_tol}")
    if bcs_list_u_names is None:
        bcs_list_u_names = [f"bc_u_{i}" for i in range(len(bc_list_u))]
    _lib.check_degradation_function(Data.degradation_function)
    if Data.fatigue and Data.fatigue_degradation_function != "asymptotic":
        raise NotImplementedError("fatigue_degradation_function must be 'asymptotic' (reference returns an "
                                  "un-raised ValueError for 'logarithmic', fatigue_degradation_functions.py:64-65)")

    # ---- device context (replaces the form / NonlinearProblem construction, solver.py:148-245)
    points, cells, ctype = fem.extract_mesh(msh)
    split = None if Data.split_energy == "no" else Data.split_energy
    gdim = 3 if ctype in ("tetrahedron", "hexahedron") else 2
    rmesh, keep_u, keep_phi = None, [], []
    engine_kw = dict(k=float(Data.k), model=_lib.model_code(Data.degradation, split),
                     irreversibility=(Data.irreversibility == "miehe"), fatigue=bool(Data.fatigue),
                     fatigue_val=float(Data.fatigue_val), device=device, degradation_function=Data.degradation_function,
                     **(solver_options or {}))
    if dist is None:
        eng = _lib.Engine(points, cells, ctype, Data.lambda_, Data.mu, Data.Gc, Data.l, **engine_kw)
    else:
        from phasefieldx_b200.distributed import RankMesh, attach, gather_field
        rmesh = RankMesh(points, cells, ctype, world, rank)
        eng = _lib.Engine(rmesh.coords, rmesh.cells, ctype, Data.lambda_, Data.mu, Data.Gc, Data.l,
                          n_owned=rmesh.n_owned, cell_primary=rmesh.cell_primary, **engine_kw)
    for i, bc in enumerate(bc_list_u):
        dofs = fem.extract_bc(bc)[0]
        if rmesh is not None:
            keep, dofs = rmesh.local_dofs(dofs, gdim)
            keep_u.append(keep)
        eng.set_bc_u(i, dofs)
    for i, bc in enumerate(bc_list_phi):
        dofs = fem.extract_bc(bc)[0]
        if rmesh is not None:
            keep, dofs = rmesh.local_dofs(dofs, 1)
            keep_phi.append(keep)
        eng.set_bc_phi(i, dofs)
    tractions = _traction_sets(T_list_u)
    for i, (_, nodes, w) in enumerate(tractions):
        if rmesh is not None:
            keep, nodes = rmesh.local_nodes(nodes)
            w = np.asarray(w)[keep]
        eng.set_traction(i, nodes, w)
    if rmesh is not None:
        attach(eng, rmesh, dist)
    for key, val in (("ksp_type", "two-level pcg (block-Jacobi + aggregate coarse space), matrix-free on GPU"), ("snes_linesearch_type", "none"),
                     ("snes_max_it", eng.params.snes_max_it), ("snes_rtol", eng.params.snes_rtol),
                     ("snes_atol", eng.params.snes_atol), ("cg_rtol", eng.params.cg_rtol)):
        logger.info(f"   {key}: {val}")

    start = time.perf_counter()
    logger.info(f" start time: {start}")
    vtk_sol = None
    pieces = bool(vtu_parallel) and dist is not None
    if Data.save_solution_vtu and (rank == 0 or pieces):
        vtk_sol = VTKFile(None, os.path.join(result_folder_name, "paraview-solutions_vtu", "phasefieldx.pvd"), "w", fmt=vtu_format,
                          rank=rank if pieces else 0, world=world if pieces else 1, keep_last=vtu_keep)
    xdmf_phi = xdmf_u = None
    if getattr(Data, "save_solution_xdmf", False) and rank == 0:
        # reference solver.py:259-278: phi.xdmf and u.xdmf in paraview-solutions_xdmf/ (heavy data in raw side files)
        xdir = os.path.join(result_folder_name, "paraview-solutions_xdmf")
        xdmf_phi, xdmf_u = XDMFFile(None, os.path.join(xdir, "phi.xdmf"), "w"), XDMFFile(None, os.path.join(xdir, "u.xdmf"), "w")
        xdmf_phi.write_mesh(msh)
        xdmf_u.write_mesh(msh)
    logger.info(" S t a r t i n g    A n a l y s i s ")
    logger.info(" ---------------------------------- ")
    logger.info(" ---------------------------------- ")

    d = eng.dim
    t, step = 0, 0
    try:
        while t < final_time:
            logger.info(f"\n\nSolution at (pseudo) time = {t}, dt = {dt}, Step = {step} ")
            logger.info("===========================================================================")
            if update_boundary_conditions is not None:
                bc_ux, bc_uy, bc_uz = update_boundary_conditions(bc_list_u, t)
            else:
                bc_ux, bc_uy, bc_uz = 0, 0, 0
            if update_loading is not None:
                update_loading(T_list_u, t)
            # re-read every mutable value after the callbacks (solver.py:306-312)
            for i, bc in enumerate(bc_list_u):
                vals = fem.extract_bc(bc)[1]
                eng.set_bc_values_u(i, vals if rmesh is None else vals[keep_u[i]])
            for i, bc in enumerate(bc_list_phi):
                vals = fem.extract_bc(bc)[1]
                eng.set_bc_values_phi(i, vals if rmesh is None else vals[keep_phi[i]])
            for i, (T, _, _) in enumerate(tractions):
                eng.set_traction_value(i, np.asarray(T.value, dtype=np.float64).reshape(-1))

            try:
                rep = eng.load_step(dt=dt, min_stagger_iter=min_stagger_iter, max_stagger_iter=max_stagger_iter,
                                    stagger_tol=stagger_error_tol, history=256)
            except RuntimeError as exc:
                logger.error(str(exc))
                raise
            h = rep["history"]
            for k in range(len(h["u_its"])):
                logger.info(f" Stagger Iteration: {k + 1}")
                logger.info(" ---------------------- ")
                logger.info(">>> Solving phase for dofs: u ")
                logger.info(f" Newton iterations: {h['u_its'][k]}")
                logger.info(f" Residual norm u: {h['fnorm_u'][k]}")
                logger.info(f" L2 error in u   direction:  {h['err_u'][k]}")
                logger.info(">>> Solving phase for dofs: Φ ")
                logger.info(f" Newton iterations: {h['phi_its'][k]}")
                logger.info(f" Residual norm Φ: {h['fnorm_phi'][k]}")
                logger.info(f" L2 error in Φ direction:  {h['err_phi'][k]}")
            logger.info(f" PCG iterations (u, Φ, projection): {rep['cg_its_u']}, {rep['cg_its_phi']}, {rep['cg_its_proj']}")

            logger.info("\n\n Saving results: ")
            if rank == 0:
                append_results_to_file(os.path.join(result_folder_name, "phasefieldx.conv"),
                                       '#step\tstagger\titerPhi\titerU', step, rep["stagger"], rep["phi_its"], rep["u_its"])
                append_results_to_file(os.path.join(result_folder_name, "top.dof"),
                                       '#step\tUx\tUy\tUz\tphi', step, bc_ux, bc_uy, bc_uz, 0.0)
            # reactions and energies are collective (global values on every rank); unlike the reference
            # (solver.py:427) the reaction files are also written in multi-rank runs
            reactions = []
            for i in range(len(bc_list_u)):
                R = eng.reaction(i)
                reactions.append(R)
                if rank == 0:
                    append_results_to_file(os.path.join(result_folder_name, bcs_list_u_names[i] + ".reaction"),
                                           '#step\tRx\tRy\tRz', step, float(R[0]), float(R[1]), float(R[2]))
            en = eng.energies()
            header = '#step\tEplusW\tE\tW\tW_phi\tW_gradphi\tgamma\tgamma_phi\tgamma_gradphi\tPSI_a\tPSI_b\talpha_acum'
            if rank == 0:
                append_results_to_file(os.path.join(result_folder_name, "total.energy"), header, step,
                                       *[en[c] for c in _lib.ENERGY_COLUMNS])
            if getattr(Data, "save_solution_xdmf", False):
                if rmesh is None:
                    phi_x, u_x = eng.get_field("phi"), eng.get_field("u").reshape(-1, d)
                else:
                    dev = f"cuda:{device}"
                    phi_x, u_x = gather_field(eng, rmesh, "phi", dist, dev), gather_field(eng, rmesh, "u", dist, dev)
                if xdmf_phi is not None:
                    xdmf_phi.write_function("phi", phi_x, step)
                    xdmf_u.write_function("u", u_x, step)
            if Data.save_solution_vtu and (step % max(1, int(vtu_every)) == 0):
                if rmesh is None or pieces:
                    phi_out, u_out = eng.get_field("phi"), eng.get_field("u").reshape(-1, d)
                else:
                    dev = f"cuda:{device}"
                    phi_out, u_out = gather_field(eng, rmesh, "phi", dist, dev), gather_field(eng, rmesh, "u", dist, dev)
                if vtk_sol is not None:
                    vtk_sol.write_function(rmesh.piece(ctype) if pieces else msh, {"phi": phi_out, "u": u_out}, step, background=True)
            if step_callback is not None:
                step_callback(step, rep, reactions, en)
            t += dt
            step += 1
    finally:
        if vtk_sol is not None:
            vtk_sol.close()
        for xf in (xdmf_phi, xdmf_u):
            if xf is not None:
                xf.close()
        if dist is not None:
            dist.barrier()          # peers keep each other's buffers mapped: close together
        eng.close()
    log_end_analysis(logger, time.perf_counter() - start)
