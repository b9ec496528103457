#!/usr/bin/env python
"""Launches each hot kernel a few times on config-2 / config-3 shapes; the target of the ncu captures
committed under profiles/ (python tools/profile_kernels.py [f32|f64])."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200.backend import CudaBackend, MapOp  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext, ShapedDataBufferView  # noqa: E402

dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
ctx = CudaContext(0, 8)  # 8 = DSC_FLAG_NO_FUSE: every call launches its own kernel(s)
b = CudaBackend(dt, ctx=ctx)
rng = np.random.default_rng(0)
n = 4096
A = ShapedDataBufferView(b.CreateDataBuffer((rng.random(n * n) * 2 - 1).astype(dt)), n, n)
B = ShapedDataBufferView(b.CreateDataBuffer((rng.random(n * n) * 2 - 1).astype(dt)), n, n)
H = ShapedDataBufferView(b.CreateDataBuffer((rng.random(8192 * 1024) * 2 - 1).astype(dt)), 8192, 1024)
W = ShapedDataBufferView(b.CreateDataBuffer((rng.random(1024 * 1024) * 2 - 1).astype(dt)), 1024, 1024)
for _ in range(3):
    b.Mul_M_M(H, W)          # hidden-layer GEMM of config 3
    b.Mul_M_M(A, B)          # config 2 GEMM
    b.Mul_Had_M_M(A, B)
    b.Map_F_M(MapOp.Tanh, None, A)
    b.Map_F_M(MapOp.Exp, None, A)
    b._out(b._f["sum_dev"], 1, A.DataBuffer.handle)
    b.Transpose_M(A)
ctx.sync()
print("done", ctx.stats())
