"""
Elasticity solver — drop-in for reference `src/phasefieldx/Element/Elasticity/solver/solver.py:27-297`.

Same signature and result files (`phasefieldx.conv` = step, iterations; `top.dof`; `<bc>.reaction`;
`total.energy` = step, E; VTU with `u`).  The displacement problem is the u-half of the staggered
phase-field step with phi = 0 (g = 1, no residual stiffness, unsplit law), so every load step is one
call of `pfx_solve_u` (include/pfx_b200.h) on the GPU; there is no CPU path.
"""
import os

import numpy as np

from phasefieldx_b200 import _lib, fem
from phasefieldx_b200.Element._single_field import Run
from phasefieldx_b200.Element.Phase_Field_Fracture.solver.solver import _traction_sets
from phasefieldx_b200.files import append_results_to_file


def solve(Data, msh, final_time, V_u, bc_list_u=[], update_boundary_conditions=None, f_list_u=None, T_list_u=None,
          update_loading=None, ds_bound=None, dt=1.0, path=None, quadrature_degree=2, bcs_list_u_names=None,
          device=0, solver_options=None, vtu_format="ascii"):
    """Reference signature; `f_list_u`, `ds_bound` and `quadrature_degree` are accepted and unused (P1/Q1
    elements with a constant tangent are integrated exactly by the kernels' fixed rules)."""
    options = {"ksp_type": "two-level pcg, matrix-free on GPU", "snes_linesearch_type": "none", "snes_max_it": 50000,
               "snes_rtol": 1e-8, "snes_atol": 1e-9}
    run = Run(Data, msh, path, options, vtu_format)
    logger, folder = run.logger, run.folder
    if bcs_list_u_names is None:
        bcs_list_u_names = [f"bc_u_{i}" for i in range(len(bc_list_u))]
    points, cells, ctype = fem.extract_mesh(msh)
    eng = _lib.Engine(points, cells, ctype, Data.lambda_, Data.mu, 1.0, 1.0, k=0.0, model=_lib.model_code("isotropic", None),
                      device=device, **(solver_options or {}))
    d = eng.dim
    try:
        for i, bc in enumerate(bc_list_u):
            eng.set_bc_u(i, fem.extract_bc(bc)[0])
        tractions = _traction_sets(T_list_u)
        for i, (_, nodes, w) in enumerate(tractions):
            eng.set_traction(i, nodes, w)
        t, step = 0, 0
        while t < final_time:
            run.step_header(t, dt, step)
            if bc_list_u is not None and update_boundary_conditions is not None:
                bc_ux, bc_uy, bc_uz = update_boundary_conditions(bc_list_u, t)
            else:
                bc_ux, bc_uy, bc_uz = 0, 0, 0
            if T_list_u is not None and update_loading is not None:
                update_loading(T_list_u, t)
            for i, bc in enumerate(bc_list_u):
                eng.set_bc_values_u(i, fem.extract_bc(bc)[1])
            for i, (T, _, _) in enumerate(tractions):
                eng.set_traction_value(i, np.asarray(T.value, dtype=np.float64).reshape(-1))
            logger.info(">>> Solving phase for dofs: u ")
            try:
                rep = eng.solve_u()
            except RuntimeError as exc:
                logger.error(str(exc))
                raise
            logger.info(f" Newton iterations: {rep['its']}")
            logger.info(f" Residual norm u: {rep['fnorm']}")
            logger.info(f" Converged reason {rep['reason']}.")
            logger.info(f" PCG iterations: {rep['cg_its']}")
            logger.info("\n\n Saving results: ")
            append_results_to_file(os.path.join(folder, "phasefieldx.conv"), '#step\titerations', step, rep["its"])
            append_results_to_file(os.path.join(folder, "top.dof"), '#step\tUx\tUy\tUz', step, bc_ux, bc_uy, bc_uz)
            for i in range(len(bc_list_u)):
                R = eng.reaction(i)
                append_results_to_file(os.path.join(folder, bcs_list_u_names[i] + ".reaction"), '#step\tRx\tRy\tRz', step,
                                       float(R[0]), float(R[1]), float(R[2]))
            append_results_to_file(os.path.join(folder, "total.energy"), '#step\tE', step, eng.energies()["E"])
            if run.vtk is not None:
                run.vtk.write_function(msh, {"u": eng.get_field("u").reshape(-1, d)}, step, background=True)
            run.write_xdmf({"u": eng.get_field("u").reshape(-1, d)} if run._xdmf_on else {}, step)
            t += dt
            step += 1
    finally:
        run.finish()
        eng.close()
