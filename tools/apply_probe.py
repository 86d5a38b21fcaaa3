#!/usr/bin/env python
"""Time the J_u apply variants on the bench workload (L2 flushed).  Usage: PFX_APPLY_VARIANT=3|4 python tools/apply_probe.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from phasefieldx_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = W.sen_shear(n)
e = W.make_engine(w)
u, phi, v = W.operator_state(w)
e.set_field("u", u)
e.set_field("phi", phi)
y_ref = e.apply_u(v, use_bc=True)          # v1 kernel (arbitrary input)
ms = e.time_operator(0, reps=18, flush_l2=True)[4:]
ms_it = e.time_operator(4, reps=14, flush_l2=True)[4:]
print(f"PFX_APPLY_VARIANT={os.environ.get('PFX_APPLY_VARIANT', 'default')}: apply_u {ms.mean() * 1e3:.2f} us (min {ms.min() * 1e3:.2f}), "
      f"pcg iteration {ms_it.mean() * 1e3:.2f} us")
