#!/usr/bin/env python
"""Runs the config-3 step eagerly a few times at a given number of batch rows (python tools/profile_step.py [rows]): the target
of the ncu capture of the step's memory-bound kernels (profiles/r02_ncu_step_kernels.*)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200 import workloads as wl  # noqa: E402
from sigma.diffsharp_b200.config import CudaBackendProvider, GlobalConfig  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
provider = CudaBackendProvider(device=0)
GlobalConfig.BackendProvider = provider
X, Y, Ws, bs = wl.mlp_inputs([784, 1024, 1024, 10], rows)
dX, dY = wl.to_DM(X), wl.to_DM(Y)
dWs, dbs = [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs]
for _ in range(4):
    provider.ctx.invalidate_caches()
    r = wl.mlp_grad(dX, dY, dWs, dbs)
    provider.ctx.sync()
print("done", provider.ctx.stats())
