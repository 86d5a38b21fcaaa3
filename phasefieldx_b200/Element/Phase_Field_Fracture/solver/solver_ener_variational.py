"""
Energy-controlled PFF solver, variational scheme — drop-in for reference
`src/phasefieldx/Element/Phase_Field_Fracture/solver/solver_ener_variational.py:41-527` (same signature and result
files: phasefieldx.conv, lambda.dof, tau.input, max_u.dof, <bc>.reaction, total.energy with the `external` column,
VTU).  Every attempted step is one `pfx_ec_solve` call on the GPU (include/pfx_b200.h); see `_energy_controlled.py`.
"""
from . import _energy_controlled


def solve(Data, msh, final_gamma, V_u, V_Φ, bc_list_u=[], bc_list_phi=[], f_list_u=None, T_list_u=None, ds_bound=None,
          dtau=0.0005, dtau_min=1e-12, dtau_max=0.1, path=None, bcs_list_u_names=None, c1=1.0, c2=1.0,
          threshold_gamma_save=None, **gpu_options):
    """Reference signature; extra keyword-only knobs: device, solver_options (pfx_ec_opts / pfx_params overrides),
    vtu_format, max_steps."""
    return _energy_controlled.solve(True, Data, msh, final_gamma, V_u, V_Φ, bc_list_u, bc_list_phi, f_list_u, T_list_u, ds_bound,
                                    dtau, dtau_min, dtau_max, path, bcs_list_u_names, c1, c2, threshold_gamma_save, **gpu_options)
