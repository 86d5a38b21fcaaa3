#!/usr/bin/env python
"""GPU probe: energy-controlled three-point bending (BASELINE config 4) — marches the d-tau loop for a number of steps
and prints one JSON line per step (Newton / GMRES / inner PCG iterations, device ms, lambda, gamma).
Usage: python tools/ec_probe.py [var|nonvar] [divx] [steps] [key=value pfx_ec_opts ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import _lib, workloads as W  # noqa: E402
from phasefieldx_b200.Materials.conversion import get_lambda_lame, get_mu_lame  # noqa: E402

var = (sys.argv[1] if len(sys.argv) > 1 else "var") == "var"
divx = int(sys.argv[2]) if len(sys.argv) > 2 else 400
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
opts = {}
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
msh, p, sets, (nodes, w, T), kw = W.three_point_bending(divx, divx // 4, var)
t0 = time.time()
e = _lib.Engine(msh.geometry.x, msh.cells, "quadrilateral", get_lambda_lame(p["E"], p["nu"]), get_mu_lame(p["E"], p["nu"]), p["Gc"], p["l"],
                k=0.0, model=0, coarse_region=int(os.environ.get("EC_COARSE_REGION", "6144")))
for i, (_, dofs) in enumerate(sets):
    e.set_bc_u(i, dofs)
    e.set_bc_values_u(i, np.zeros(len(dofs)))
e.set_traction(0, nodes, w)
e.set_traction_value(0, T)
print(json.dumps({"nodes": int(msh.num_nodes), "cells": int(msh.num_cells), "engine_s": time.time() - t0, "blocks": e.block_stats(),
                  "coarse": e.coarse_stats()}), flush=True)
golden = 2 / (1 + 5 ** 0.5)
dtau, tau, lam, step, gamma = kw["dtau"], kw["dtau"], 0.0, 0, 0.0
e.ec_checkpoint()
total_ms = 0.0
while gamma < kw["final_gamma"] and step < steps:
    tau += dtau
    t = time.time()
    r = e.ec_solve(tau, lam, c1=kw["c1"], c2=kw["c2"], variational=var, **opts)
    r["wall_s"] = time.time() - t
    r["tau"], r["dtau"] = tau, dtau
    if not r["converged"]:
        e.ec_checkpoint(restore=True)
        tau -= dtau
        dtau *= golden
        r["rejected"] = True
        print(json.dumps(r), flush=True)
        continue
    lam = r["lam"]
    e.ec_checkpoint()
    if r["its"] <= 2:
        dtau = min(dtau * (1 + golden), kw["dtau_max"])
    step += 1
    gamma = r["gamma"]
    total_ms += r["gpu_ms"]
    r["step"] = step
    r["reaction"] = e.reaction(0).tolist()
    print(json.dumps(r), flush=True)
print(json.dumps({"steps": step, "device_ms": total_ms, "steps_per_s": step / (total_ms * 1e-3), "launches": e.launch_count(),
                  "coarse": e.coarse_stats()}))
