#!/usr/bin/env python
"""Launches the f32 GEMM shapes of the config-3 step (rows x 1024 x 1024 forward, weight adjoint with
K = rows) a few times: the target of ncu captures (python tools/profile_gemm.py [rows])."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200.backend import CudaBackend  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext, ShapedDataBufferView  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = CudaContext(0)
b = CudaBackend(np.float32, ctx=ctx)
rng = np.random.default_rng(0)
mat = lambda m, n: ShapedDataBufferView(b.CreateDataBuffer((rng.random(m * n) * 2 - 1).astype(np.float32)), m, n)  # noqa: E731
H, dZ, W2 = mat(rows, 1024), mat(rows, 1024), mat(1024, 1024)
for _ in range(3):
    b.Mul_M_M(H, W2)                     # forward
    b.Mul_M_M(b.Transpose_M(H), dZ)      # weight adjoint
ctx.sync()
print("done", ctx.stats())
