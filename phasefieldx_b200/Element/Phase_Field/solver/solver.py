"""
Phase-field solver — drop-in for reference `src/phasefieldx/Element/Phase_Field/solver/solver.py:25-287`.

Crack-surface boundary value problem (AT2): (1/l) phi v + l grad(phi).grad(v) = 0 with Dirichlet data —
the phi-half of the staggered step with H = 0 and Gc = 1 (AT1: the same operator with a rescaled length and a constant
history field, see below) — so every step is one call of
`pfx_solve_phi` (include/pfx_b200.h) on the GPU.  Result files as in the reference: `phasefieldx.conv`
(step, iterations), `total.energy` (step, gamma, gamma_phi, gamma_gradphi), VTU with `phi`.
"""
import os

import numpy as np

from phasefieldx_b200 import _lib, fem
from phasefieldx_b200.Element._single_field import Run
from phasefieldx_b200.files import append_results_to_file


def solve(Data, msh, final_time, V_Φ, bc_list_phi=[], update_boundary_conditions=None, update_loading=None,
          ds_bound=None, dt=1.0, path=None, quadrature_degree=2, case="AT2", V_gradient_Φ=None, device=0,
          solver_options=None, vtu_format="ascii"):
    """Reference signature; `ds_bound`, `quadrature_degree` are accepted and unused.  `update_loading` (a UFL source
    term -(f v / l + l grad_f . grad v) dx built from `ufl.SpatialCoordinate`, reference solver.py:124-130) changes the
    solution and has no counterpart here: passing it raises instead of silently dropping the term."""
    if update_loading is not None:
        raise NotImplementedError("update_loading (UFL source term of reference Phase_Field/solver/solver.py:124-130) is not "
                                  "available on the GPU path")
    if case not in ("AT2", "AT1"):
        raise NotImplementedError("geometric crack functions on the GPU path: 'AT2' and 'AT1'; 'WU' and 'DOUBLE' "
                                  "(Element/Phase_Field/geometric_crack.py:24-27) give indefinite Jacobians, which the "
                                  "PCG path cannot solve")
    if V_gradient_Φ is not None:
        raise NotImplementedError("gradient output (V_gradient_Φ) is not available")
    options = {"ksp_type": "two-level pcg, matrix-free on GPU", "snes_linesearch_type": "none", "snes_max_it": 50000,
               "snes_rtol": 1e-8, "snes_atol": 1e-9}
    run = Run(Data, msh, path, options, vtu_format)
    logger, folder = run.logger, run.folder
    points, cells, ctype = fem.extract_mesh(msh)
    # AT2: (1/l) phi v + l grad phi . grad v = 0 is the phase-field operator of the engine with H = 0, Gc = 1.
    # AT1 (w = phi, c0 = 8/3: (1/(c0 l)) v + (2 l/c0) grad phi . grad v = 0, geometric_crack.py:22-23, 47-48, 72-73) is the
    # same operator with l' = sqrt(2) l and the constant history field H = -1/(2 l'): the mass term 2 H + 1/l' vanishes
    # and K phi = -(1/l'^2) int N = -(1/(2 l^2)) int N.
    l_eng = float(Data.l) * (2.0 ** 0.5 if case == "AT1" else 1.0)
    eng = _lib.Engine(points, cells, ctype, 1.0, 1.0, 1.0, l_eng, k=0.0, model=_lib.model_code("isotropic", None),
                      device=device, **(solver_options or {}))
    if case == "AT1":
        eng.set_field("H", np.full(eng.n_nodes, -(1.0 / l_eng) / 2.0))
    try:
        for i, bc in enumerate(bc_list_phi):
            eng.set_bc_phi(i, fem.extract_bc(bc)[0])
        t, step = 0, 0
        while t < final_time:
            run.step_header(t, dt, step)
            if update_boundary_conditions is not None:
                update_boundary_conditions(bc_list_phi, t)
            for i, bc in enumerate(bc_list_phi):
                eng.set_bc_values_phi(i, fem.extract_bc(bc)[1])
            logger.info(">>> Solving phase for dofs: Φ ")
            try:
                rep = eng.solve_phi()
            except RuntimeError as exc:
                logger.error(str(exc))
                raise
            logger.info(f" Newton iterations: {rep['its']}")
            logger.info(f" Residual norm Φ: {rep['fnorm']}")
            logger.info(f" PCG iterations: {rep['cg_its']}")
            logger.info("\n\n Saving results: ")
            append_results_to_file(os.path.join(folder, "phasefieldx.conv"), '#step\titerations', step, rep["its"])
            en = eng.energies()
            if case == "AT1":
                # calculate_crack_surface_energy(case='AT1'): (1/(c0 l)) int phi + (l/c0) int |grad phi|^2, c0 = 8/3
                c0, l = 8.0 / 3.0, float(Data.l)
                g_phi = float(eng.apply_mass(eng.get_field("phi")).sum()) / (c0 * l)
                g_grad = (2.0 * en["gamma_gradphi"] / l_eng) * l / c0
                en = {"gamma": g_phi + g_grad, "gamma_phi": g_phi, "gamma_gradphi": g_grad}
            append_results_to_file(os.path.join(folder, "total.energy"), '#step\tgamma\tgamma_phi\tgamma_gradphi', step,
                                   en["gamma"], en["gamma_phi"], en["gamma_gradphi"])
            if run.vtk is not None:
                run.vtk.write_function(msh, {"phi": eng.get_field("phi")}, step, background=True)
            run.write_xdmf({"phi": eng.get_field("phi")} if run._xdmf_on else {}, step)
            t += dt
            step += 1
    finally:
        run.finish()
        eng.close()
