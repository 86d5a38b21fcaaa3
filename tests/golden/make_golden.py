#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from /root/reference (run in the build
container only; the GPU box has no /root/reference).

  c1_mesh.npz        config-1 mesh parsed from reference examples/PhaseFieldFracture/mesh/mesh.msh
  ref_1711.npz       first rows of the stored outputs of example 1711 (bottom.reaction,
                     total.energy, phasefieldx.conv, top.dof)
  ref_1713.npz       same for example 1713 (structured 200x100 Q1 mesh, no mesh file needed)
  ref_1712.npz       same for example 1712 (shear, spectral)
  ref_1800.npz       first 40 rows of the fatigue example 1800 (total.energy incl. alpha_acum, conv, dof)
  ref_1711_final.vtu.gz  the reference's stored final VTU of example 1711 (5 089 points / 9 990 triangles in dolfinx ordering,
                     phi, u), gzip-compressed as is: pins the VTU writer / reader to the reference's own file layout
  ref_energy_controlled.npz  first 12 rows of the stored outputs of the energy-controlled examples
                     (examples/PhaseFieldFractureEnergyControlled/results_three_point_[non_]var: total.energy,
                     bottom_left.reaction, lambda.dof; results_central_cracked_non_var: total.energy, bottom.reaction)
  ref_degradation.npz   g and dg of the reference's own g_degradation_functions.py (quadratic, borden) on a grid of
                     phi, executed unmodified (plain Python arithmetic; loaded by file path)
  ref_constitutive.npz  outputs of the reference's own constitutive layer
                     (Math/invariants.py, Math/tensor_decomposition.py, Materials/elastic_isotropic.py,
                     split_energy_stress_tangent_functions.py) executed UNMODIFIED under the torch-fp64
                     `ufl` stand-in of tests/golden/ufl_shim.py, at seeded random strains.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
EX = os.path.join(REF, "examples", "PhaseFieldFracture")


def table(path, nrows):
    a = np.loadtxt(path, skiprows=1)
    return a[:nrows]


def main():
    from phasefieldx_b200 import mesh as M
    md = M.read_from_msh(os.path.join(EX, "mesh", "mesh.msh"), None, 0, 2)
    msh = md.mesh
    np.savez_compressed(os.path.join(HERE, "c1_mesh.npz"), points=msh.geometry.x[:, :2], cells=msh.cells,
                        bottom_nodes=M.facet_nodes(msh, md.facet_tags.find(9)),
                        top_nodes=M.facet_nodes(msh, md.facet_tags.find(10)))
    for tag, folder, n in (("1711", "1711_Single_Edge_Notched_Tension_Test", 150),
                           ("1713", "1713_Symmetry_Single_Edge_Notched_Tension_Test", 60),
                           ("1712", "1712_Single_Edge_Notched_Shear_Test", 104)):
        d = os.path.join(EX, folder)
        np.savez_compressed(os.path.join(HERE, f"ref_{tag}.npz"),
                            bottom_reaction=table(os.path.join(d, "bottom.reaction"), n),
                            energy=table(os.path.join(d, "total.energy"), n),
                            conv=table(os.path.join(d, "phasefieldx.conv"), n),
                            dof=table(os.path.join(d, "top.dof"), n))
    fat = os.path.join(REF, "examples", "Fatigue", "1800_Fatigue_Single_Edge_Notched_Tension_Test")
    np.savez_compressed(os.path.join(HERE, "ref_1800.npz"), energy=table(os.path.join(fat, "total.energy"), 40),
                        conv=table(os.path.join(fat, "phasefieldx.conv"), 40), dof=table(os.path.join(fat, "top.dof"), 40))
    ec = os.path.join(REF, "examples", "PhaseFieldFractureEnergyControlled")
    np.savez_compressed(os.path.join(HERE, "ref_energy_controlled.npz"),
                        var_energy=table(os.path.join(ec, "results_three_point_var", "total.energy"), 12),
                        var_reaction=table(os.path.join(ec, "results_three_point_var", "bottom_left.reaction"), 12),
                        var_lambda=table(os.path.join(ec, "results_three_point_var", "lambda.dof"), 12),
                        nonvar_energy=table(os.path.join(ec, "results_three_point_non_var", "total.energy"), 12),
                        nonvar_reaction=table(os.path.join(ec, "results_three_point_non_var", "bottom_left.reaction"), 12),
                        cc_energy=table(os.path.join(ec, "results_central_cracked_non_var", "total.energy"), 12),
                        cc_reaction=table(os.path.join(ec, "results_central_cracked_non_var", "bottom.reaction"), 12))
    import gzip
    vtu = os.path.join(EX, "1711_Single_Edge_Notched_Tension_Test", "paraview-solutions_vtu", "phasefieldx_p0_000149.vtu")
    with open(vtu, "rb") as src, gzip.GzipFile(os.path.join(HERE, "ref_1711_final.vtu.gz"), "wb", mtime=0) as dst:
        dst.write(src.read())
    from ufl_shim import reference_constitutive
    np.savez_compressed(os.path.join(HERE, "ref_constitutive.npz"), **reference_constitutive())
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_g", os.path.join(REF, "src", "phasefieldx", "Element", "Phase_Field_Fracture", "g_degradation_functions.py"))
    ref_g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_g)
    phi = np.linspace(-0.25, 1.25, 61)
    out = {"phi": phi}
    for name in ("quadratic", "borden"):
        out["g_" + name] = np.array([ref_g.g(float(v), name) for v in phi])
        out["dg_" + name] = np.array([ref_g.dg(float(v), name) for v in phi])
    np.savez_compressed(os.path.join(HERE, "ref_degradation.npz"), **out)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    main()
