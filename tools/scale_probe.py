#!/usr/bin/env python
"""Strong-scaling probe of one PCG iteration (apply + vector kernels + halo push + scalar
all-reduces) on a large partitioned mesh.  Run with N = 1, 2, 4, 8:
  torchrun --nproc-per-node N tools/scale_probe.py beam 384     (3D notched beam, P1 tets, spectral)
  torchrun --nproc-per-node N tools/scale_probe.py shear 2000   (2D shear, P1 triangles, spectral)
Every rank times the same 40 iterations (CUDA events on the library stream); the max over ranks
is reported together with the node/cell counts and the partition imbalance."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "beam"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 192
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t0 = time.time()
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, size // 3, size // 6)
t1 = time.time()
e, local_sets = bench.build_rank_engine(w, rank, world, lr, {})
t2 = time.time()
st = e.block_stats()
# a damaged, loaded state (operator micro-benchmark state of SURVEY §8d), restricted to this rank
u, phi, v = W.operator_state(w)
if world > 1:
    from phasefieldx_b200 import _lib
    R = _lib.Partition(w.cell_type, w.msh.geometry.x, w.msh.cells, world).rank(rank)
    l2g = R["l2g"]
    d = w.msh.geometry.dim
    u = u.reshape(-1, d)[l2g].ravel()
    phi = phi[l2g]
    v = v.reshape(-1, d)[l2g].ravel()
e.set_field("u", u)
e.set_field("phi", phi)
e.apply_u(v, use_bc=True)                      # uploads v as the right-hand side of the timed PCG
if world > 1:
    dist.barrier()
ms_it = e.time_operator(4, reps=48, flush_l2=False)[8:]
ms_ap = e.time_operator(0, reps=24, flush_l2=False)[4:] if world == 1 else None
tt = torch.tensor([ms_it.mean(), float(st["n_owned"]), float(st["elems_with_overlap"])], dtype=torch.float64, device="cuda")
if world > 1:
    mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    mx, sm = tt, tt
if rank == 0:
    out = {"workload": w.name, "n_gpus": world, "nodes": int(w.msh.num_nodes), "cells": int(w.msh.num_cells),
           "dofs_u": int(w.msh.num_nodes * w.msh.geometry.dim), "ms_per_pcg_iteration_max_over_ranks": float(mx[0]),
           "owned_nodes_max": int(mx[1]), "owned_nodes_mean": float(sm[1]) / world,
           "cells_with_overlap_total": int(sm[2]), "setup_s": {"mesh": t1 - t0, "engine": t2 - t1},
           "ms_apply_u": (float(ms_ap.mean()) if ms_ap is not None else None)}
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
e.close()
if world > 1:
    dist.destroy_process_group()
