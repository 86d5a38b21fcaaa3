"""
Elasticity with an energy split — drop-in for reference
`src/phasefieldx/Element/Phase_Field_Fracture/solver/solver_anisotropic.py:27-310`.

F_u = (sigma_a(u) + sigma_b(u)) : eps(du): the displacement half of the staggered step with phi = 0 and no
residual stiffness, evaluated through the split stresses.  Result files as in the reference:
`phasefieldx.conv` (step, iterations), `top.dof`, `<bc>.reaction`, `total.energy` (step, E, PSI_a, PSI_b),
VTU with `u` and the nodal L2 projections `PSI_a`, `PSI_b`.  Every step is one `pfx_solve_u` call plus two
`pfx_project_psi` calls (include/pfx_b200.h) on the GPU; there is no CPU path.

E, the energy of the UNSPLIT law (reference `calculate_elastic_energy`), is evaluated as u.F_int(u)/2 — the
work of the internal forces, exact because the split stresses are homogeneous of degree one — so that
`PSI_a + PSI_b == E` stays the independent consistency check it is in the reference's
test_anisotropic_decomposition.py.
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
    """Reference signature; `f_list_u`, `ds_bound`, `quadrature_degree` are accepted and unused."""
    options = {"ksp_type": "two-level pcg, matrix-free on GPU", "snes_linesearch_type": "none", "snes_max_it": 50000,
               "snes_rtol": 1e-8, "snes_atol": 1e-9}
    run = Run(Data, msh, path, options, vtu_format)
    logger, folder = run.logger, run.folder
    if bcs_list_u_names is None:
        bcs_list_u_names = [f"bc_u_{i}" for i in range(len(bc_list_u))]
    points, cells, ctype = fem.extract_mesh(msh)
    split = None if Data.split_energy == "no" else Data.split_energy
    # reference solver_anisotropic.py:130-133, 236-237: F_u = (sigma_a + sigma_b)(u, Data) : eps(du) and the output
    # energies psi_a / psi_b(u, Data) dispatch on Data.degradation — "isotropic": psi_a = psi, psi_b = 0 whatever
    # split_energy says; "hybrid": unsplit stress, split energies; "anisotropic": both split.  sigma_a + sigma_b is
    # the full stress in every case.
    eng = _lib.Engine(points, cells, ctype, Data.lambda_, Data.mu, 1.0, 1.0, k=0.0,
                      model=_lib.model_code(Data.degradation, split), device=device, **(solver_options or {}))
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
            logger.info("\n\n Saving results: ")
            append_results_to_file(os.path.join(folder, "phasefieldx.conv"), '#step\titerations', step, rep["its"])
            append_results_to_file(os.path.join(folder, "top.dof"), '#step\tUx\tUy\tUz', step, bc_ux, bc_uy, bc_uz)
            for i in range(len(bc_list_u)):
                R = eng.reaction(i)
                append_results_to_file(os.path.join(folder, bcs_list_u_names[i] + ".reaction"), '#step\tRx\tRy\tRz', step,
                                       float(R[0]), float(R[1]), float(R[2]))
            u = eng.get_field("u")
            en = eng.energies()
            E = 0.5 * float(u @ eng.residual_u())
            append_results_to_file(os.path.join(folder, "total.energy"), '#step\tE\tPSI_a\tPSI_b', step, E, en["PSI_a"], en["PSI_b"])
            if run.vtk is not None:
                run.vtk.write_function(msh, {"u": u.reshape(-1, d), "PSI_a": eng.project_psi("a"), "PSI_b": eng.project_psi("b")},
                                       step, background=True)
            run.write_xdmf({"u": u.reshape(-1, d)} if run._xdmf_on else {}, step)
            t += dt
            step += 1
    finally:
        run.finish()
        eng.close()
