"""
Energy-controlled PFF solver, non-variational scheme — drop-in for reference
`src/phasefieldx/Element/Phase_Field_Fracture/solver/solver_ener_non_variational.py:40-532` (no lambda term in F_phi,
J_phi,lambda = None; otherwise identical to the variational module).  One `pfx_ec_solve` call per attempted step.
"""
from . import _energy_controlled


def solve(Data, msh, final_gamma, V_u, V_Φ, bc_list_u=[], bc_list_phi=[], f_list_u=None, T_list_u=None, ds_bound=None,
          dtau=0.0005, dtau_min=1e-12, dtau_max=0.1, path=None, bcs_list_u_names=None, c1=1.0, c2=1.0,
          threshold_gamma_save=None, **gpu_options):
    """Reference signature; extra keyword-only knobs: device, solver_options, vtu_format, max_steps."""
    return _energy_controlled.solve(False, Data, msh, final_gamma, V_u, V_Φ, bc_list_u, bc_list_phi, f_list_u, T_list_u, ds_bound,
                                    dtau, dtau_min, dtau_max, path, bcs_list_u_names, c1, c2, threshold_gamma_save, **gpu_options)
