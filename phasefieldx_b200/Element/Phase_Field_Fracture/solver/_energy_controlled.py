"""
Shared implementation of the energy-controlled drop-in solvers (`solver_ener_variational.solve`,
`solver_ener_non_variational.solve`): the adaptive d-tau loop, logging and result files of reference
`solver_ener_variational.py:339-527` / `solver_ener_non_variational.py:344-532` on the host, one `pfx_ec_solve` call
(include/pfx_b200.h: the monolithic Newton solve of (u, phi, lambda) on the GPU) per attempted step.
"""
import os
import time

import numpy as np

from phasefieldx_b200 import _lib, fem
from phasefieldx_b200.Element._single_field import Run
from phasefieldx_b200.files import append_results_to_file

GOLDEN_RATIO = 2 / (1 + 5 ** 0.5)


def solve(variational, Data, msh, final_gamma, V_u, V_Φ, bc_list_u=[], bc_list_phi=[], f_list_u=None, T_list_u=None, ds_bound=None,
          dtau=0.0005, dtau_min=1e-12, dtau_max=0.1, path=None, bcs_list_u_names=None, c1=1.0, c2=1.0, threshold_gamma_save=None,
          device=0, solver_options=None, vtu_format="ascii", max_steps=None):
    if not T_list_u:
        raise ValueError("T_list_u[0] = (traction, measure) is required (reference solver_ener_variational.py:177)")
    T0, ds0 = T_list_u[0][0], T_list_u[0][1]
    if not isinstance(ds0, fem.Measure):
        raise NotImplementedError("tractions need a measure from get_ds_bound_from_marker")
    _lib.check_degradation_function(Data.degradation_function)
    if Data.degradation == "hybrid":
        raise NotImplementedError("the energy-controlled GPU solver needs sigma_a = d psi_a / d eps (no 'hybrid' degradation)")
    n_react = len(bc_list_u) - 1 if bc_list_phi != [] else len(bc_list_u)          # reference :141-146
    if bcs_list_u_names is None:
        bcs_list_u_names = [f"bc_u_{i}" for i in range(n_react)]
    ec_opts = dict(solver_options or {})
    engine_keys = ("cg_rtol", "cg_max_it", "block_nodes", "cg_chunk", "quad_points_1d", "proj_rtol", "coarse_region", "coarse_agg_blocks")
    engine_kw = {k: ec_opts.pop(k) for k in engine_keys if k in ec_opts}
    # traction-loaded, weakly supported bodies (the bending beams of the reference examples): one dense coarse region
    engine_kw.setdefault("coarse_region", 6144)
    options = {"newton": "dolfinx NewtonSolver semantics (residual criterion)", "rtol": ec_opts.get("newton_rtol", 1e-9),
               "atol": ec_opts.get("newton_atol", 1e-10), "max_it": ec_opts.get("newton_max_it", 50),
               "linear solver": "flexible GMRES on the bordered system, block-preconditioned by PCG solves, on GPU"}
    run = Run(Data, msh, path, options, vtu_format)
    logger, folder = run.logger, run.folder
    points, cells, ctype = fem.extract_mesh(msh)
    split = None if Data.split_energy == "no" else Data.split_energy
    eng = _lib.Engine(points, cells, ctype, Data.lambda_, Data.mu, Data.Gc, Data.l, k=float(Data.k),
                      model=_lib.model_code(Data.degradation, split), irreversibility=False, fatigue=False, device=device,
                      degradation_function=Data.degradation_function, **engine_kw)
    d = eng.dim
    header = '#step\tEplusW\tE\tW\tW_phi\tW_gradphi\tgamma\tgamma_phi\tgamma_gradphi\tPSI_a\tPSI_b\texternal'
    try:
        for i, bc in enumerate(bc_list_u):
            dofs, vals = fem.extract_bc(bc)
            if np.any(np.asarray(vals) != 0.0):
                raise NotImplementedError("the energy-controlled GPU solver takes homogeneous displacement Dirichlet data")
            eng.set_bc_u(i, dofs)
            eng.set_bc_values_u(i, vals)
        nodes, w = ds0.nodal_weights()
        if len(nodes) == 0:
            raise ValueError("the traction measure T_list_u[0][1] holds no facets: the constraint row of lambda would be empty")
        eng.set_traction(0, nodes, w)
        eng.set_traction_value(0, np.asarray(T0.value, dtype=np.float64).reshape(-1))
        # ---- initial crack from phase-field Dirichlet data (reference :276-288)
        gamma0 = 0.0
        if bc_list_phi != []:
            logger.info("\n\nBoundary condition for phase-field ")
            for i, bc in enumerate(bc_list_phi):
                dofs, vals = fem.extract_bc(bc)
                eng.set_bc_phi(i, dofs)
                eng.set_bc_values_phi(i, vals)
            logger.info("\n\nSolving boundary value problem for phase-field ")
            eng.solve_phi()
            gamma0 = eng.energies()["gamma"]
            for i in range(len(bc_list_phi)):          # the monolithic solve carries no phase-field Dirichlet data
                eng.set_bc_phi(i, np.zeros(0, np.int32))
        logger.info(f"\n\nGamma0 = {gamma0}")
        eng.ec_checkpoint()

        tau = gamma0 + dtau
        last_saved_gamma = gamma0
        if threshold_gamma_save is None:
            threshold_gamma_save = 0.0
        step, gamma, lam, lam_c = 0, 0.0, 0.0, 0.0
        while gamma < final_gamma and (max_steps is None or step < max_steps):
            tau += dtau
            logger.info(f"\n\nSolution at tau = {tau}, dtau = {dtau}, Step = {step} ")
            logger.info("===========================================================================")
            logger.info(">>> Solving phase for dofs: u, phi, lambda ")
            logger.info(f">>> tau(t) = {tau}")
            rep = eng.ec_solve(tau, lam, c1=c1, c2=c2, variational=variational, **ec_opts)
            if not rep["converged"]:
                why = ("Newton solver did not converge because maximum number of iterations reached" if rep["reason"] == -5
                       else f"Newton solver did not converge (reason {rep['reason']})")
                logger.info(f"Solver failed at tau={tau} with error: {why}. Reducing dt.")
                eng.ec_checkpoint(restore=True)
                lam = lam_c
                tau -= dtau
                dtau = max(dtau * GOLDEN_RATIO, dtau_min)
                if dtau <= dtau_min:
                    logger.info("Minimum dtau reached. Stopping.")
                    if run.vtk is not None:
                        run.vtk.write_function(msh, {"phi": eng.get_field("phi"), "u": eng.get_field("u").reshape(-1, d)}, step)
                    break
                continue
            lam = rep["lam"]
            n_iterations = rep["its"]
            logger.info(f" Newton iterations: {n_iterations}")
            logger.info(f" Residual norm: {rep['residual']}")
            logger.info(f" GMRES iterations: {rep['gmres_its']}, inner PCG iterations (u, Φ): {rep['cg_its_u']}, {rep['cg_its_phi']}")
            eng.ec_checkpoint()
            lam_c = lam
            if n_iterations <= 2:
                dtau = min(dtau * (1 + GOLDEN_RATIO), dtau_max)
                logger.info(f"Converged in ({n_iterations} iterations), increasing dt to {dtau}")
            else:
                logger.info(f"Slow convergence ({n_iterations} iterations), reducing dt to {dtau}")
            step += 1

            logger.info("\n\n Saving results: ")
            # the reference logs the KSP residual norm of its direct solve here; this column holds the Newton residual
            append_results_to_file(os.path.join(folder, "phasefieldx.conv"), '#step\titerations\tResidualNorm', step, n_iterations,
                                   rep["residual"])
            append_results_to_file(os.path.join(folder, "lambda.dof"), '#step\tlambda', step, lam)
            append_results_to_file(os.path.join(folder, "tau.input"), '#step\ttau', step, tau)
            u_now = eng.get_field("u").reshape(-1, d)
            umax = [float(u_now[:, k].max()) for k in range(d)] + [0.0] * (3 - d)
            append_results_to_file(os.path.join(folder, "max_u.dof"), '#step\tUx\tUy\tUz\tphi', step, umax[0], umax[1], umax[2], 0.0)
            for i in range(n_react):
                R = eng.reaction(i)
                append_results_to_file(os.path.join(folder, bcs_list_u_names[i] + ".reaction"), '#step\tRx\tRy\tRz', step,
                                       float(R[0]), float(R[1]), float(R[2]))
            en = eng.energies()
            gamma = en["gamma"]
            external_energy = rep["external"]
            check = c1 * gamma + c2 * external_energy - tau
            logger.info(f"Check constraint: c1*gamma + c2*external_energy - tau = {c1}*{gamma:.6e} + {c2}*{external_energy:.6e} - {tau:.6e} = {check:.6e}")
            append_results_to_file(os.path.join(folder, "total.energy"), header, step, en["EplusW"], en["E"], en["W"], en["W_phi"],
                                   en["W_gradphi"], gamma, en["gamma_phi"], en["gamma_gradphi"], en["PSI_a"], en["PSI_b"], external_energy)
            logger.info(f"gamma = {gamma:.6e}, gamma_phi = {en['gamma_phi']:.6e}, gamma_gradphi = {en['gamma_gradphi']:.6e}")
            if step == 1 or threshold_gamma_save == 0.0 or abs(gamma - last_saved_gamma) >= threshold_gamma_save:
                if run.vtk is not None:
                    logger.info(f"Saving VTU Paraview files at step={step}, gamma={gamma:.6f}, tau={tau:.6f}")
                    run.vtk.write_function(msh, {"phi": eng.get_field("phi"), "u": u_now}, step, background=True)
                run.write_xdmf({"phi": eng.get_field("phi"), "u": u_now} if run._xdmf_on else {}, step)
                last_saved_gamma = gamma
    finally:
        run.finish()
        eng.close()
