#!/usr/bin/env python
"""Multi-GPU parity probe (run under torchrun): N ranks each own a partition of the workload;
after every load step the owned parts of u/phi/H are gathered and compared with a single-GPU
engine run by rank 0 on the whole mesh; reactions and energies are compared too.
Usage: torchrun --nproc-per-node N tools/mgpu_check.py [shear|beam] [size] [steps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phasefieldx_b200 import _lib, workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shear"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 200
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, max(2, size // 3), max(2, size // 6))
e, local_sets = bench.build_rank_engine(w, rank, world, lr, {})
P = _lib.Partition(w.cell_type, w.msh.geometry.x, w.msh.cells, world)
R = P.rank(rank)
d = w.msh.geometry.dim
ref = W.make_engine(w, device=lr) if rank == 0 else None
tstep = 20.0 if kind == "shear" else 4.0
for s in range(1, steps + 1):
    t = tstep * s
    bench.set_bc_values(e, w, t, local_sets)
    r = e.load_step(min_stagger_iter=2, max_stagger_iter=60, stagger_tol=1e-8)
    Rr = [e.reaction(i) for i in range(len(w.bc_sets))]
    en = e.energies()
    gathered = {}
    for name, nc in (("u", d), ("phi", 1), ("H", 1)):
        loc = e.get_field(name).reshape(-1, nc)[: R["n_owned"]]
        full = torch.zeros((w.msh.num_nodes, nc), dtype=torch.float64, device="cuda")
        full[torch.from_numpy(R["l2g"][: R["n_owned"]]).cuda()] = torch.from_numpy(loc).cuda()
        dist.all_reduce(full)
        gathered[name] = full.cpu().numpy().ravel()
    if rank == 0:
        bench.set_bc_values(ref, w, t, None)
        rr = ref.load_step(min_stagger_iter=2, max_stagger_iter=60, stagger_tol=1e-8)
        line = [f"step {s}: stagger {r['stagger']}/{rr['stagger']} cg_u {r['cg_its_u']}/{rr['cg_its_u']} ms {r['gpu_ms']:.0f}/{rr['gpu_ms']:.0f}"]
        for name in ("u", "phi", "H"):
            a, b = gathered[name], ref.get_field(name)
            line.append(f"{name} {np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300):.2e}")
        Rs = [ref.reaction(i) for i in range(len(w.bc_sets))]
        line.append("R " + " ".join(f"{np.abs(x - y).max() / max(np.abs(y).max(), 1e-300):.1e}" for x, y in zip(Rr, Rs)))
        es = ref.energies()
        line.append("E " + f"{max(abs(en[k] - es[k]) / max(abs(es[k]), 1e-300) for k in es if es[k] != 0):.1e}")
        print(" | ".join(line), flush=True)
dist.barrier()
e.close()
dist.destroy_process_group()
