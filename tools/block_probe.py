#!/usr/bin/env python
"""PCG iteration time of a workload for several CTA block sizes.  Usage: python tools/block_probe.py [shear|beam] [size] [block_nodes ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "beam"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 192
sizes = [int(v) for v in sys.argv[3:]] or [0]
w = W.sen_shear(size) if kind == "shear" else W.notched_beam_3d(size, size // 3, size // 6)
u, phi, v = W.operator_state(w)
for bn in sizes:
    e = W.make_engine(w, block_nodes=bn)
    e.set_field("u", u)
    e.set_field("phi", phi)
    e.apply_u(v, use_bc=True)
    ms_a = e.time_operator(0, reps=10, flush_l2=True)[3:]
    ms_i = e.time_operator(4, reps=10, flush_l2=True)[3:]
    st = e.block_stats()
    print(f"block_nodes {bn:4d}: blocks {st['n_blocks']} max_elems {st['max_elems']} overlap x{st['elems_with_overlap'] / st['n_cells']:.3f}  "
          f"apply_u {ms_a.mean() * 1e3:8.1f} us  pcg iteration {ms_i.mean() * 1e3:8.1f} us")
    e.close()
