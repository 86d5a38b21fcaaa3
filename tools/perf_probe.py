#!/usr/bin/env python
"""GPU probe: operator timings (achieved algorithmic GB/s) and a few load steps of a workload.
Usage: python tools/perf_probe.py [shear|beam] [size] [steps] [block_nodes]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shear"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.time()
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, size // 3, size // 6)
t1 = time.time()
e = W.make_engine(w, block_nodes=bn)
t2 = time.time()
print(f"workload {w.name}: nodes {w.msh.num_nodes} cells {w.msh.num_cells}; mesh {t1 - t0:.1f}s engine {t2 - t1:.1f}s")
print("blocks", e.block_stats())
u, phi, v = W.operator_state(w)
e.set_field("u", u)
e.set_field("phi", phi)
e.set_field("H", np.abs(np.random.default_rng(1).normal(size=len(phi))) * 1e-3)
y = e.apply_u(v, use_bc=True)
peak = 6538.9
for name, k in (("apply_u", 0), ("apply_phi", 1), ("apply_mass", 2), ("residual_u", 3), ("pcg_iter_u", 4)):
    for flush in (1, 0):
        ms = e.time_operator(k, reps=12, flush_l2=bool(flush))[2:]
        b = e.algorithmic_bytes(k)
        print(f"{name:12s} flush={flush} ms avg {ms.mean():.4f} min {ms.min():.4f}  alg {b / 1e6:.1f} MB  "
              f"{b / ms.mean() / 1e6:.0f} GB/s ({b / ms.mean() / 1e6 / peak:.2%} of measured peak)")
e.set_field("u", np.zeros_like(u))
e.set_field("phi", np.zeros_like(phi))
e.set_field("H", np.zeros_like(phi))
for s in range(steps):
    for i, vals in enumerate(w.bc_values(float(s))):
        e.set_bc_values_u(i, vals)
    t = time.time()
    r = e.load_step(min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, stagger_tol=w.stagger_tol)
    dt = time.time() - t
    r.pop("history")
    print(f"step {s}: wall {dt:.2f}s gpu {r['gpu_ms'] / 1e3:.2f}s", json.dumps(r))
    print("   R(bottom)", e.reaction(1), "energies", {k: float('%.6g' % val) for k, val in e.energies().items() if k in ('E', 'gamma', 'PSI_a', 'PSI_b')})
print("launches", e.launch_count())
