#!/usr/bin/env python
"""GPU check of the aggregate coarse space: the same load steps with and without it
(PFX_NO_COARSE=1), PCG iteration counts, step times and field agreement.
Usage: python tools/coarse_probe.py [shear|beam] [size] [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shear"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, size // 3, size // 6)
print(f"workload {w.name}: nodes {w.msh.num_nodes} cells {w.msh.num_cells}")
out = {}
for tag, env in (("coarse", "0"), ("jacobi", "1"))[: 1 if os.environ.get("ONLY_COARSE") else 2]:
    os.environ["PFX_NO_COARSE"] = env
    e = W.make_engine(w)
    res = []
    for s in range(steps):
        for i, vals in enumerate(w.bc_values(float(s))):
            e.set_bc_values_u(i, vals)
        t = time.time()
        r = e.load_step(min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, stagger_tol=w.stagger_tol)
        print(f"[{tag}] step {s}: gpu {r['gpu_ms']:.1f} ms wall {1e3 * (time.time() - t):.1f} stagger {r['stagger']} "
              f"cg u/phi/proj {r['cg_its_u']}/{r['cg_its_phi']}/{r['cg_its_proj']} newton u/phi {r['u_its']}/{r['phi_its']}")
        res.append((r["gpu_ms"], r["stagger"]))
    out[tag] = dict(u=e.get_field("u"), phi=e.get_field("phi"), H=e.get_field("H"), R=e.reaction(1), ms=[x[0] for x in res],
                    en=e.energies())
    if tag == "coarse" and kind == "shear":
        uu, pp, vv = W.operator_state(w)
        e.apply_u(vv, use_bc=True)
        ms = e.time_operator(4, reps=12, flush_l2=True)[2:]
        print(f"[coarse] pcg iteration (flushed L2): {ms.mean() * 1e3:.1f} us")
    e.close()
if "jacobi" not in out:
    sys.exit(0)
a, b = out["coarse"], out["jacobi"]
for f in ("u", "phi", "H"):
    print(f"rel L2 diff {f}: {np.linalg.norm(a[f] - b[f]) / max(np.linalg.norm(b[f]), 1e-300):.3e}")
print("reaction", a["R"], b["R"], "rel", np.linalg.norm(a["R"] - b["R"]) / max(np.linalg.norm(b["R"]), 1e-300))
print("speed-up per step:", [round(y / x, 2) for x, y in zip(a["ms"], b["ms"])])
