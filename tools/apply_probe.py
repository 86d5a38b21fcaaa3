#!/usr/bin/env python
"""Time the J_u / J_phi applies and one PCG iteration on a workload (L2 flushed); small enough to run under ncu.
Usage: python tools/apply_probe.py [shear|beam] [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from phasefieldx_b200 import workloads as W  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shear"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w = W.sen_shear(n) if kind == "shear" else W.notched_beam_3d(n, n // 3, n // 6)
e = W.make_engine(w)
u, phi, v = W.operator_state(w)
e.set_field("u", u)
e.set_field("phi", phi)
e.set_field("H", np.abs(np.random.default_rng(1).normal(size=len(phi))) * 1e-3)
e.apply_u(v, use_bc=True)
ms = e.time_operator(0, reps=12, flush_l2=True)[4:]
ms_p = e.time_operator(1, reps=12, flush_l2=True)[4:]
ms_it = e.time_operator(4, reps=12, flush_l2=True)[4:]
b0, b5 = e.algorithmic_bytes(0), e.algorithmic_bytes(5)
print(f"{w.name}: apply_u {ms.mean() * 1e3:.2f} us (min {ms.min() * 1e3:.2f}; algorithmic {b0 / 1e6:.1f} MB = {b0 / ms.mean() / 1e6:.0f} GB/s, "
      f"streamed {b5 / 1e6:.1f} MB = {b5 / ms.mean() / 1e6:.0f} GB/s), apply_phi {ms_p.mean() * 1e3:.2f} us, pcg iteration {ms_it.mean() * 1e3:.2f} us")
