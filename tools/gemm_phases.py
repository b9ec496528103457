#!/usr/bin/env python
"""Where the time of one float32 tensor-core GEMM goes, as CTA 0 sees it (%globaltimer stamps written by the
kernel, DSC_GEMM_STAMPS=1): prologue, first TMA fill, main loop, drain, epilogue / fix-up.

    DSC_GEMM_STAMPS=1 python tools/gemm_phases.py
"""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("DSC_GEMM_STAMPS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200._lib import check, lib  # noqa: E402
from sigma.diffsharp_b200.backend import CudaBackend  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext  # noqa: E402

ctx = CudaContext(0, 8)  # 8 = DSC_FLAG_NO_FUSE: every call launches its own kernel(s)
b = CudaBackend(np.float32, ctx=ctx)
rng = np.random.default_rng(0)


def run(m, n, k, ta, tb, cfg):
    A = b.CreateDataBuffer((rng.random(m * k) * 2 - 1).astype(np.float32))
    B = b.CreateDataBuffer((rng.random(k * n) * 2 - 1).astype(np.float32))
    check(lib.dsc_gemm_tune(ctx.handle, *cfg))
    e0, e1 = ctx.event(), ctx.event()
    for _ in range(3):
        keep = b._out(b._f["gemm"], m * n, ta, tb, m, n, k, A.handle, B.handle, 0)
    ctx.sync()
    ctx.record(e0)
    keep = b._out(b._f["gemm"], m * n, ta, tb, m, n, k, A.handle, B.handle, 0)
    ctx.record(e1)
    ctx.sync()
    st = (C.c_uint64 * 8)()
    check(lib.dsc_gemm_debug_stamps(ctx.handle, st))
    d = [(st[i + 1] - st[i]) / 1e3 for i in range(5)]
    check(lib.dsc_gemm_tune(ctx.handle, -1, -1, -1, -1))
    print(f"m={m} n={n} k={k} ta={ta} tb={tb} bn={cfg[0]} pair={cfg[1]} split={cfg[2]} fuse={cfg[3]}: events {ctx.elapsed_ms(e0, e1) * 1e3:6.1f} us (pre-pass + kernel) | "
          f"prologue {d[0]:5.1f} | first fill {d[1]:5.1f} | main loop {d[2]:5.1f} | drain {d[3]:5.1f} | epilogue {d[4]:5.1f} | kernel {sum(d):5.1f}"
          + (f" | split-K: park {(st[6] - st[4]) / 1e3:4.1f}, wait {(st[7] - st[6]) / 1e3:4.1f}, reduce {(st[5] - st[7]) / 1e3:4.1f}" if st[7] > st[4] else ""), flush=True)
    del keep


for cfg in [(-1, -1, -1, -1), (128, 0, 2, 1), (128, 1, 2, 0), (128, 1, 1, 0), (64, 1, 1, 0), (64, 0, 1, 1), (256, 1, 4, 0)]:
    run(1024, 1024, 1024, 0, 0, cfg)
for cfg in [(-1, -1, -1, -1), (256, 0, -1, 0)]:
    run(8192, 1024, 1024, 0, 0, cfg)
run(1024, 1024, 8192, 1, 0, (-1, -1, -1, -1))
