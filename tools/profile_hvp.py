#!/usr/bin/env python
"""Config 5 (reverse-on-forward Hessian-vector product, n = 1e6, float64) a few times: target of `ncu` launch lists
(python tools/profile_hvp.py [n])."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200 import workloads as wl  # noqa: E402
from sigma.diffsharp_b200.config import CudaBackendProvider, GlobalConfig  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
prov = CudaBackendProvider(device=0)
GlobalConfig.BackendProvider = prov
x, v, d = wl.hvp_inputs(n)
dx, dv, dd = wl.to_DV(x), wl.to_DV(v), wl.to_DV(d)
for _ in range(3):
    s0 = prov.ctx.stats()["kernel_launches"]
    hv = wl.hvp(dd, dx, dv)
    prov.ctx.sync()
    print("launches per HVP:", prov.ctx.stats()["kernel_launches"] - s0)
ref = wl.hvp_closed_form(x, v, d)
print("max rel err", float(np.abs(hv.ToArray() - ref).max() / np.abs(ref).max()))
