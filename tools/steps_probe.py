#!/usr/bin/env python
"""GPU probe: run load steps of a workload and print one JSON line per step (stagger / PCG counts, device
time per phase, reactions, energies).
Usage: python tools/steps_probe.py [shear|beam] [size] [first_step] [last_step] [key=value solver options ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shear"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
s0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s1 = int(sys.argv[4]) if len(sys.argv) > 4 else 3
opts = {}
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
t0 = time.time()
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, size // 3, size // 6)
t1 = time.time()
e = W.make_engine(w, **opts)
t2 = time.time()
print(json.dumps({"workload": w.name, "nodes": int(w.msh.num_nodes), "cells": int(w.msh.num_cells), "mesh_s": t1 - t0,
                  "engine_s": t2 - t1, "blocks": e.block_stats(), "coarse": e.coarse_stats()}), flush=True)
for s in range(s0, s1 + 1):
    for i, vals in enumerate(w.bc_values(float(s))):
        e.set_bc_values_u(i, vals)
    t = time.time()
    r = e.load_step(min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, stagger_tol=w.stagger_tol)
    r.pop("history")
    r["wall_s"] = time.time() - t
    r["step"] = s
    r["phase_ms"] = {k: round(v, 2) for k, v in r["phase_ms"].items()}
    r["reaction"] = [e.reaction(i).tolist() for i in range(len(w.bc_sets))]
    r["energies"] = e.energies()
    print(json.dumps(r), flush=True)
print(json.dumps({"launches": e.launch_count(), "coarse": e.coarse_stats()}))
