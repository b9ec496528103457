#!/usr/bin/env python
"""Per-op device time of every backend call of the config-3 step (784-1024-1024-10 tanh MLP) at a
given number of batch rows, each op launched R times from one CUDA graph rotating over input sets
(python tools/step_ops_bench.py [rows ...]).  Prints microseconds per call and the HBM / tensor
fraction where it applies."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200.backend import CudaBackend, MapOp  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext, ShapedDataBufferView  # noqa: E402

ctx = CudaContext(0)                      # deferred evaluation on: products are held and run when the graph capture ends
b = bf = CudaBackend(np.float32, ctx=ctx)
rng = np.random.default_rng(0)


def mat(m, n, sets=3):
    return [ShapedDataBufferView(b.CreateDataBuffer((rng.random(m * n) * 2 - 1).astype(np.float32)), m, n) for _ in range(sets)]


def vec(n, sets=3):
    return [b.CreateDataBuffer((rng.random(n) * 2 - 1).astype(np.float32)) for _ in range(sets)]


def time_graph(fns, reps=5):
    for f in fns:
        keep0 = f()
        ctx.flush()
    ctx.sync()
    ctx.graph_begin()
    try:
        keep = []
        for f in fns:
            keep.append(f())
            ctx.flush()                   # one call's kernels after the other, as in the step
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
    del keep
    ctx.graph_destroy(g)
    return best / len(fns) * 1e3


def run(rows):
    R = 9
    X, H, dZ = mat(rows, 784), mat(rows, 1024), mat(rows, 1024)
    Z3, Y = mat(rows, 10), mat(rows, 10)
    W1, W2, W3 = mat(784, 1024), mat(1024, 1024), mat(1024, 10)
    b1, b3 = vec(1024), vec(10)
    T = lambda v: bf.Transpose_M(v)  # noqa: E731  lazy transpose consumed by the GEMM
    cases = [
        ("Mul_M_M X*W1 (fwd1)", lambda i: bf.Mul_M_M(X[i % 3], W1[i % 3]), 2.0 * rows * 784 * 1024, 0),
        ("Mul_M_M H*W2 (fwd2)", lambda i: bf.Mul_M_M(H[i % 3], W2[i % 3]), 2.0 * rows * 1024 * 1024, 0),
        ("Mul_M_M H*W3 (fwd3, n=10)", lambda i: bf.Mul_M_M(H[i % 3], W3[i % 3]), 0, rows * 1024 * 4),
        ("Mul_M_M dZ3*W3^T (k=10)", lambda i: bf.Mul_M_M(Z3[i % 3], T(W3[i % 3])), 0, rows * 1024 * 4),
        ("Mul_M_M H^T*dZ3 (dW3, n=10)", lambda i: bf.Mul_M_M(T(H[i % 3]), Z3[i % 3]), 0, rows * 1024 * 4),
        ("Mul_M_M dZ*W2^T (dH1)", lambda i: bf.Mul_M_M(dZ[i % 3], T(W2[i % 3])), 2.0 * rows * 1024 * 1024, 0),
        ("Mul_M_M H^T*dZ (dW2)", lambda i: bf.Mul_M_M(T(H[i % 3]), dZ[i % 3]), 2.0 * rows * 1024 * 1024, 0),
        ("Mul_M_M X^T*dZ (dW1)", lambda i: bf.Mul_M_M(T(X[i % 3]), dZ[i % 3]), 2.0 * rows * 784 * 1024, 0),
        ("bias+tanh (fused rows_act)", lambda i: bf.Map_F_M(MapOp.Tanh, None, bf.Add_M_M(H[i % 3], bf.RepeatReshapeCopy_V_MRows(rows, b1[i % 3]))), 0, 3 * rows * 1024 * 4),
        ("tanh adjoint (fused_bwd)", lambda i: bf._fused(0, dZ[i % 3].DataBuffer, H[i % 3].DataBuffer, (rows, 1024)), 0, 3 * rows * 1024 * 4),
        ("Add_M_Colwise_V_InPlace 1024", lambda i: bf.Add_M_Colwise_V_InPlace(dZ[i % 3], b1[i % 3]), 0, rows * 1024 * 4),
        ("Add_M_Colwise_V_InPlace 10", lambda i: bf.Add_M_Colwise_V_InPlace(Z3[i % 3], b3[i % 3]), 0, rows * 10 * 4),
        ("Sub_M_M rows x 10", lambda i: b.Sub_M_M(Z3[i % 3], Y[i % 3]), 0, 3 * rows * 10 * 4),
        ("Sum_M(dev) rows x 10", lambda i: b._out(b._f["sum_dev"], 1, Z3[i % 3].DataBuffer.handle), 0, rows * 10 * 4),
        ("Sum_M(dev) rows x 1024", lambda i: b._out(b._f["sum_dev"], 1, H[i % 3].DataBuffer.handle), 0, rows * 1024 * 4),
    ]
    print(f"--- rows = {rows} ---")
    tot = 0.0
    for name, fn, flop, nbytes in cases:
        us = time_graph([(lambda i=i: fn(i)) for i in range(R)])
        extra = ""
        if flop:
            extra = f"{flop / us / 1e6:8.1f} TFLOP/s ({flop / us / 1e6 / 244.16:.2f} of 3xTF32)"
        elif nbytes:
            extra = f"{nbytes / us / 1e3:8.1f} GB/s ({nbytes / us / 1e3 / 6545.6:.2f} of HBM)"
        print(f"{name:34s} {us:8.2f} us  {extra}")
        tot += us
    print(f"{'sum':34s} {tot:8.2f} us")


for r in ([int(a) for a in sys.argv[1:]] or [8192, 1024]):
    run(r)
