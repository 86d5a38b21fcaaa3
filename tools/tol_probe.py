#!/usr/bin/env python
"""How far is a load step solved with cg_rtol from one solved much tighter?  (Jacobi vs two-level PCG)
Usage: python tools/tol_probe.py [shear|beam] [size]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "beam"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 96
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, size // 3, size // 6)
print(f"workload {w.name}: nodes {w.msh.num_nodes} cells {w.msh.num_cells}")


def run(coarse, rtol):
    os.environ["PFX_NO_COARSE"] = "0" if coarse else "1"
    e = W.make_engine(w, cg_rtol=rtol)
    its = 0
    for s in range(2):
        for i, vals in enumerate(w.bc_values(float(s))):
            e.set_bc_values_u(i, vals)
        r = e.load_step(min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, stagger_tol=w.stagger_tol)
        its += r["cg_its_u"]
    out = dict(u=e.get_field("u"), phi=e.get_field("phi"), H=e.get_field("H"), its=its, stagger=r["stagger"])
    e.close()
    return out


ref = run(True, 1e-14)
print("reference: two-level, cg_rtol 1e-14, its", ref["its"])
for coarse in (True, False):
    for rtol in (1e-11, 1e-12, 1e-13):
        o = run(coarse, rtol)
        d = {f: np.linalg.norm(o[f] - ref[f]) / np.linalg.norm(ref[f]) for f in ("u", "phi", "H")}
        print(f"{'two-level' if coarse else 'jacobi   '} rtol {rtol:.0e}: its {o['its']:6d} stagger {o['stagger']}  "
              f"err u {d['u']:.2e} phi {d['phi']:.2e} H {d['H']:.2e}")
