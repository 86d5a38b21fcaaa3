#!/usr/bin/env python
"""GPU probe on the config-1/3 mesh (5 089 nodes): load steps/s in the launch-latency regime."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import workloads as W
from phasefieldx_b200.mesh import Mesh
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_mesh.npz"))
msh = Mesh(d["points"], d["cells"], "triangle", gdim=2)
fat = len(sys.argv) > 1 and sys.argv[1] == "fatigue"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
w = W.sen_tension(msh, d["bottom_nodes"], d["top_nodes"], fatigue=fat)
e = W.make_engine(w)
print(w.name, e.block_stats())
t0 = time.time(); tot_st = tot_cg = 0
for s in range(nsteps):
    for i, v in enumerate(w.bc_values(float(s))):
        e.set_bc_values_u(i, v)
    r = e.load_step(min_stagger_iter=2, max_stagger_iter=500, stagger_tol=1e-8)
    tot_st += r["stagger"]; tot_cg += r["cg_its_u"] + r["cg_its_phi"] + r["cg_its_proj"]
    if s % 10 == 0 or s == nsteps - 1:
        print(f"step {s}: stagger {r['stagger']} cg_u {r['cg_its_u']} cg_phi {r['cg_its_phi']} gpu_ms {r['gpu_ms']:.1f} Ry {e.reaction(1)[1]:.6e}")
dt = time.time() - t0
print(f"{nsteps} steps in {dt:.2f}s = {nsteps / dt:.2f} steps/s; {tot_st} stagger its, {tot_cg} CG its, {dt / max(tot_cg, 1) * 1e6:.1f} us per CG iteration")
