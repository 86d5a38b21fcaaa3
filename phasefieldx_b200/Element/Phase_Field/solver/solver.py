"""
Phase-field solver — drop-in for reference `src/phasefieldx/Element/Phase_Field/solver/solver.py:25-287`.

Crack-surface boundary value problem (AT2): (1/l) phi v + l grad(phi).grad(v) = 0 with Dirichlet data —
the phi-half of the staggered step with H = 0 and Gc = 1 — so every step is one call of
`pfx_solve_phi` (include/pfx_b200.h) on the GPU.  Result files as in the reference: `phasefieldx.conv`
(step, iterations), `total.energy` (step, gamma, gamma_phi, gamma_gradphi), VTU with `phi`.
"""
import os

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
    if case != "AT2":
        raise NotImplementedError("only the AT2 geometric crack function is on the GPU path (SURVEY §8f N2)")
    if V_gradient_Φ is not None:
        raise NotImplementedError("gradient output (V_gradient_Φ) is not available")
    options = {"ksp_type": "two-level pcg, matrix-free on GPU", "snes_linesearch_type": "none", "snes_max_it": 50000,
               "snes_rtol": 1e-8, "snes_atol": 1e-9}
    run = Run(Data, msh, path, options, vtu_format)
    logger, folder = run.logger, run.folder
    points, cells, ctype = fem.extract_mesh(msh)
    eng = _lib.Engine(points, cells, ctype, 1.0, 1.0, 1.0, float(Data.l), k=0.0, model=_lib.model_code("isotropic", None),
                      device=device, **(solver_options or {}))
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
