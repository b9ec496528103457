#!/usr/bin/env python
"""Standalone accuracy + speed check of dsc_sgemm / dsc_dgemm (run under `timeout` on the GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200.backend import CudaBackend  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext  # noqa: E402


def check(b, m, n, k, ta, tb, positive=False, time_it=False):
    dt = b.dtype
    rng = np.random.default_rng(m + 3 * n + 7 * k)
    A = rng.random((m, k)).astype(dt) if positive else (rng.random((m, k)) * 2 - 1).astype(dt)
    B = rng.random((k, n)).astype(dt) if positive else (rng.random((k, n)) * 2 - 1).astype(dt)
    As = np.ascontiguousarray(A.T if ta else A).reshape(-1)
    Bs = np.ascontiguousarray(B.T if tb else B).reshape(-1)
    dA, dB = b.CreateDataBuffer(As), b.CreateDataBuffer(Bs)
    out = b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, 0)
    got = out.SubData.reshape(m, n).astype(np.float64)
    del out
    rows = np.arange(m) if m * n * k < 5e8 else np.sort(rng.choice(m, 64, replace=False))
    ref = A[rows].astype(np.float64) @ B.astype(np.float64)
    mag = np.abs(A[rows]).astype(np.float64) @ np.abs(B).astype(np.float64)
    err = np.abs(got[rows] - ref)
    ratio = float((err / mag).max())
    msg = f"{np.dtype(dt).name} m={m} n={n} k={k} ta={ta} tb={tb} pos={int(positive)}: max err/(|A||B|) = {ratio:.3e}"
    if time_it:
        ctx = b.ctx
        e0, e1 = ctx.event(), ctx.event()
        for _ in range(3):
            b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, 0)
        ctx.record(e0)
        iters = 10
        for _ in range(iters):
            b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, 0)
        ctx.record(e1)
        ms = ctx.elapsed_ms(e0, e1) / iters
        msg += f"   {ms:.4f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    return ratio


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "f32"
    ctx = CudaContext(0, 8)  # 8 = DSC_FLAG_NO_FUSE: every call launches its own kernel(s)
    b = CudaBackend(np.float32 if which == "f32" else np.float64, ctx=ctx)
    tol = 1e-5 if which == "f32" else 1e-12
    worst = 0.0
    shapes = [(128, 64, 32), (128, 64, 64), (128, 256, 128), (256, 256, 256), (257, 129, 65), (128, 300, 784), (512, 512, 512),
              (1024, 2048, 96), (1000, 1000, 1000)]
    for m, n, k in shapes:
        for ta in (0, 1):
            for tb in (0, 1):
                worst = max(worst, check(b, m, n, k, ta, tb))
    worst = max(worst, check(b, 1024, 1024, 8192, 0, 0, positive=True))
    worst = max(worst, check(b, 1024, 1024, 8192, 1, 0, positive=True))
    for m, n, k in [(128, 300, 784), (1024, 1024, 1024), (2048, 1024, 1024), (4096, 1024, 1024), (1024, 1024, 2048), (8192, 1024, 1024), (8192, 1024, 784), (1024, 1024, 8192), (4096, 4096, 4096), (8192, 8192, 8192) if which == "f32" else (4096, 4096, 4096)]:
        worst = max(worst, check(b, m, n, k, 0, 0, time_it=True))
    worst = max(worst, check(b, 1024, 1024, 8192, 1, 0, time_it=True))
    worst = max(worst, check(b, 8192, 1024, 1024, 0, 1, time_it=True))
    print("stats", ctx.stats())
    print("WORST", worst, "PASS" if worst <= tol else "FAIL")
    return 0 if worst <= tol else 1


if __name__ == "__main__":
    sys.exit(main())
