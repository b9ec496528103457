#!/usr/bin/env python
"""Sweep the plans of the float32 tensor-core GEMM (dsc_gemm_tune: tile width, 1- or 2-CTA kernel,
split-K, in-kernel operand split) over the five GEMM shapes of the config-3 step at the given batch
rows; device time per call from a CUDA graph of 6 calls rotating over 3 operand sets.

    python tools/gemm_sweep.py [rows ...]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200._lib import check, lib  # noqa: E402
from sigma.diffsharp_b200.backend import CudaBackend  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext  # noqa: E402

ctx = CudaContext(0, 8)  # 8 = DSC_FLAG_NO_FUSE: every call launches its own kernel(s)
b = CudaBackend(np.float32, ctx=ctx)
rng = np.random.default_rng(0)


def time_graph(fns, reps=5):
    for f in fns:
        f()
    ctx.sync()
    ctx.graph_begin()
    try:
        keep = [f() for f in fns]
    finally:
        g = ctx.graph_end()
    e0, e1 = ctx.event(), ctx.event()
    ctx.graph_launch(g)
    ctx.sync()
    best = 1e30
    for _ in range(reps):
        ctx.record(e0)
        ctx.graph_launch(g)
        ctx.record(e1)
        best = min(best, ctx.elapsed_ms(e0, e1))
    out = keep[0].SubData.copy()
    del keep
    ctx.graph_destroy(g)
    return best / len(fns) * 1e3, out


def sweep(name, m, n, k, ta, tb, configs):
    A = [b.CreateDataBuffer((rng.random(m * k) * 2 - 1).astype(np.float32)) for _ in range(3)]
    B = [b.CreateDataBuffer((rng.random(k * n) * 2 - 1).astype(np.float32)) for _ in range(3)]
    ref = None
    rows = []
    for cfg in configs:
        check(lib.dsc_gemm_tune(ctx.handle, *cfg))
        try:
            us, out = time_graph([(lambda i=i: b._out(b._f["gemm"], m * n, ta, tb, m, n, k, A[i % 3].handle, B[i % 3].handle, 0)) for i in range(6)])
        except Exception as e:  # a plan that does not apply to the shape
            rows.append((cfg, None, str(e)[:60]))
            continue
        if ref is None:
            ref = out
        diff = float(np.abs(out - ref).max() / np.abs(ref).max())
        rows.append((cfg, us, diff))
    check(lib.dsc_gemm_tune(ctx.handle, -1, -1, -1, -1))
    print(f"{name}: m={m} n={n} k={k} ta={ta} tb={tb}")
    for cfg, us, d in rows:
        tag = "planner" if cfg == (-1, -1, -1, -1) else f"bn={cfg[0]:3d} pair={cfg[1]} split={cfg[2]:2d} fuse={cfg[3]}"
        if us is None:
            print(f"    {tag:32s}  n/a  {d}")
        else:
            print(f"    {tag:32s} {us:8.2f} us  {2.0 * m * n * k / us / 1e6:7.1f} TFLOP/s   maxdiff {d:.1e}")
    sys.stdout.flush()


def main():
    rows_list = [int(a) for a in sys.argv[1:]] or [8192, 1024]
    for R in rows_list:
        print(f"===== rows = {R} =====")
        big = [(-1, -1, -1, -1), (256, 1, -1, 0), (256, 1, -1, 1), (128, 1, -1, 0), (256, 0, -1, 0), (128, 0, -1, 0), (128, 0, -1, 1), (64, 0, -1, 1), (64, 1, -1, 0)]
        deep = [(-1, -1, -1, -1)] + [(256, 1, s, 0) for s in (1, 2, 4)] + [(128, 1, s, 0) for s in (1, 2)] + \
               [(128, 0, s, 1) for s in (1, 2)] + [(128, 0, 2, 0)] + [(256, 0, s, 0) for s in (1, 2, 4)]
        small = R <= 4096
        sweep("fwd1  X*W1", R, 1024, 784, 0, 0, deep if small else big)
        sweep("fwd2  H*W2", R, 1024, 1024, 0, 0, deep if small else big)
        sweep("dH1   dZ*W2^T", R, 1024, 1024, 0, 1, deep if small else big)
        sweep("dW2   H^T*dZ", 1024, 1024, R, 1, 0, deep)
        sweep("dW1   X^T*dZ", 784, 1024, R, 1, 0, deep)


if __name__ == "__main__":
    main()
