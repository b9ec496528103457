#!/usr/bin/env python
"""Launch a few GEMM shapes (for ncu launch lists): python tools/gemm_shapes.py f32 m,n,k,ta,tb [...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200.backend import CudaBackend  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext  # noqa: E402

ctx = CudaContext(0, 8)  # 8 = DSC_FLAG_NO_FUSE: every call launches its own kernel(s)
b = CudaBackend(np.float32 if sys.argv[1] == "f32" else np.float64, ctx=ctx)
for spec in sys.argv[2:]:
    m, n, k, ta, tb = [int(x) for x in spec.split(",")]
    rng = np.random.default_rng(0)
    dA = b.CreateDataBuffer((rng.random(m * k) * 2 - 1).astype(b.dtype))
    dB = b.CreateDataBuffer((rng.random(k * n) * 2 - 1).astype(b.dtype))
    e0, e1 = ctx.event(), ctx.event()
    for it in range(4):
        ctx.sync()
        t0 = time.perf_counter()
        ctx.record(e0)
        o = b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, 0)
        t1 = time.perf_counter()
        ctx.record(e1)
        ms = ctx.elapsed_ms(e0, e1)
        print(f"{spec} iter {it}: device {ms:.4f} ms, host enqueue {1e3 * (t1 - t0):.4f} ms", flush=True)
