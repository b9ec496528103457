#!/usr/bin/env python
"""Launches the five float32 tensor-core products of the config-3 step (forward 784 / 1024, input adjoint, the two weight
adjoints with K = rows) a few times, in step order and with the operand-split companions handled as in the step (dropped at
the start of every iteration): the target of ncu captures (python tools/profile_gemm.py [rows])."""
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
X, H, dZ, W1, W2 = mat(rows, 784), mat(rows, 1024), mat(rows, 1024), mat(784, 1024), mat(1024, 1024)
for _ in range(3):
    ctx.invalidate_caches()
    keep = []
    for f in (lambda: b.Mul_M_M(X, W1),                       # forward, layer 1
              lambda: b.Mul_M_M(H, W2),                       # forward, layer 2
              lambda: b.Mul_M_M(dZ, b.Transpose_M(W2)),       # input adjoint
              lambda: b.Mul_M_M(b.Transpose_M(H), dZ),        # weight adjoint, layer 2
              lambda: b.Mul_M_M(b.Transpose_M(X), dZ)):       # weight adjoint, layer 1
        keep.append(f())
        ctx.flush()                                           # products are deferred: run this one now
ctx.sync()
print("done", ctx.stats())
