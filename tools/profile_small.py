#!/usr/bin/env python
"""Launches the small / memory-bound kernels of the config-3 step (skinny GEMMs, column sums, tanh
adjoint) a few times at a given number of batch rows: the target of targeted ncu captures
(python tools/profile_small.py [rows])."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200.backend import CudaBackend, MapOp  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext, ShapedDataBufferView  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = CudaContext(0)
b = CudaBackend(np.float32, ctx=ctx)
rng = np.random.default_rng(0)
mat = lambda m, n: ShapedDataBufferView(b.CreateDataBuffer((rng.random(m * n) * 2 - 1).astype(np.float32)), m, n)  # noqa: E731
H, dZ, Z3, W3 = mat(rows, 1024), mat(rows, 1024), mat(rows, 10), mat(1024, 10)
b1 = b.CreateDataBuffer(np.zeros(1024, np.float32))
b3 = b.CreateDataBuffer(np.zeros(10, np.float32))
for _ in range(3):
    keep = [b.Mul_M_M(H, W3),                      # skinny rows
            b.Mul_M_M(Z3, b.Transpose_M(W3)),      # tiny k
            b.Mul_M_M(b.Transpose_M(H), Z3)]       # skinny cols  (products are deferred: held until the flush below)
    b.Add_M_Colwise_V_InPlace(dZ, b1)
    b.Add_M_Colwise_V_InPlace(Z3, b3)
    keep.append(b._fused(0, dZ.DataBuffer, H.DataBuffer, (rows, 1024)))
    keep.append(b._out(b._f["sum_dev"], 1, H.DataBuffer.handle))
    ctx.flush()
ctx.sync()
print("done", ctx.stats())
