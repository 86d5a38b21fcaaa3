#!/usr/bin/env python
"""Multi-GPU run of the DROP-IN solve() (one process per GPU under torchrun, the reference's
`mpirun` analogue): config 1 partitioned over the ranks, result files written by rank 0 and checked
against the reference's stored example-1711 outputs (same checks as
tests/test_gpu_solve_dropin.py::test_config1_through_solve_matches_stored_reference_files).
Usage: torchrun --nproc-per-node 2 tools/mgpu_solve_check.py"""
import os
import sys
import tempfile

import numpy as np

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

import phasefieldx_b200.compat as compat  # noqa: E402
compat.install()
import dolfinx  # noqa: E402
from phasefieldx.Element.Phase_Field_Fracture.Input import Input  # noqa: E402
from phasefieldx.Element.Phase_Field_Fracture.solver.solver import solve  # noqa: E402
from phasefieldx.Boundary.boundary_conditions import bc_xy, bc_y  # noqa: E402
from phasefieldx.PostProcessing.ReferenceResult import AllResults  # noqa: E402
from phasefieldx_b200.mesh import Mesh  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "c1_mesh.npz"))
msh = Mesh(d["points"], d["cells"], "triangle", gdim=2)
tmp = [tempfile.mkdtemp(prefix="pfx_mgpu_") if rank == 0 else None]
dist.broadcast_object_list(tmp, src=0)
tmp = tmp[0]
Data = Input(E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no",
             degradation_function="quadratic", irreversibility="miehe", fatigue=False, k=0.0,
             save_solution_xdmf=False, save_solution_vtu=True, results_folder_name=os.path.join(tmp, "1711"))
fdim = 1
bottom = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], -0.5))
top = dolfinx.mesh.locate_entities_boundary(msh, fdim, lambda x: np.isclose(x[1], 0.5))
V_u = dolfinx.fem.functionspace(msh, ("Lagrange", 1, (2,)))
V_phi = dolfinx.fem.functionspace(msh, ("Lagrange", 1))
bcs = [bc_y(top, V_u, fdim), bc_xy(bottom, V_u, fdim)]


def update_boundary_conditions(bcs_, time):
    val = 1e-4 * time if time <= 50 else 5e-3 + 1e-5 * (time - 50)
    bcs_[0].g.value[...] = val
    return 0, val, 0


PIECES = os.environ.get("PFX_VTU_PIECES", "0") == "1"      # per-rank pieces + .pvtu instead of a gathered file
solve(Data, msh, 5.0, V_u, V_phi, bcs, [], update_boundary_conditions, None, None, None, None, 1.0, path=tmp,
      bcs_list_u_names=["top", "bottom"], min_stagger_iter=2, max_stagger_iter=500, stagger_error_tol=1e-8, vtu_every=4,
      vtu_parallel=PIECES)
dist.barrier()
if rank == 0:
    S = AllResults(Data.results_folder_name)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_1711.npz"))
    Ry = S.reaction_files["bottom.reaction"]["Ry"].to_numpy()
    np.testing.assert_allclose(Ry[1:5], ref["bottom_reaction"][1:5, 2], rtol=2e-9)
    en = S.energy_files["total.energy"]
    np.testing.assert_allclose(en["PSI_a"].to_numpy()[1:5], ref["energy"][1:5, 9], rtol=1e-11)
    assert S.convergence_files["phasefieldx.conv"]["stagger"].tolist() == ref["conv"][:5, 1].astype(int).tolist()
    from phasefieldx_b200.io_vtk import read_vtu_points_cells
    vdir = os.path.join(Data.results_folder_name, "paraview-solutions_vtu")
    if PIECES:          # merge the owned nodes of the per-rank pieces back into the global numbering
        assert os.path.exists(os.path.join(vdir, "phasefieldx000004.pvtu"))
        v = {"phi": np.zeros(msh.num_nodes), "u": np.zeros((msh.num_nodes, 3))}
        for r in range(world):
            p = read_vtu_points_cells(os.path.join(vdir, f"phasefieldx_p{r}_000004.vtu"))
            own = p["point_ghost"] == 0
            v["phi"][p["point_gid"][own]] = p["phi"][own]
            v["u"][p["point_gid"][own]] = p["u"][own]
    else:
        v = read_vtu_points_cells(os.path.join(vdir, "phasefieldx_p0_000004.vtu"))
    print("vtu:", v["phi"].shape, v["u"].shape, float(v["u"][:, 1].max()), float(v["u"][:, 1].min()), bool(np.isfinite(v["u"]).all()), flush=True)
    assert v["phi"].shape[0] == msh.num_nodes and np.isfinite(v["u"]).all()
    assert np.allclose(v["u"][d["top_nodes"], 1], 4e-4, rtol=0, atol=1e-15) and np.all(v["u"][d["bottom_nodes"], :2] == 0.0)
    print(f"multi-GPU solve() ok on {world} ranks: Ry = {Ry[1:5]}", flush=True)
dist.barrier()
dist.destroy_process_group()
