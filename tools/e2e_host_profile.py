#!/usr/bin/env python
"""Where the host time of the end-to-end loop of bench.py goes (python tools/e2e_host_profile.py [rows]): the same loop —
upload of step i+2 and download of step i on the copy streams while step i+1 replays — with a wall-clock timer around every
C-ABI call group.  At 1024 rows per rank the replayed step takes ~0.14 ms on the device, so the loop is bound by the host."""
import ctypes as C
import os
import sys
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200 import dist as D, workloads as wl  # noqa: E402
from sigma.diffsharp_b200._lib import check, lib  # noqa: E402
from sigma.diffsharp_b200.ad import DNDArray  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaDataBuffer, ShapedDataBufferView  # noqa: E402
from sigma.diffsharp_b200.config import CudaBackendProvider, GlobalConfig  # noqa: E402

LAYERS = [784, 1024, 1024, 10]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
provider = CudaBackendProvider(device=0)
GlobalConfig.BackendProvider = provider
ctx = provider.ctx
backend = provider.GetBackend(np.float32(0)).BackendHandle
comm = D.NcclComm(ctx)
X, Y, Ws, bs = wl.mlp_inputs(LAYERS, rows)
dWs, dbs = [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs]


def pinned(shape):
    n = int(np.prod(shape))
    p = C.c_void_p()
    check(lib.dsc_host_alloc_pinned(n * 4, C.byref(p)))
    return np.frombuffer((C.c_char * (n * 4)).from_address(p.value), dtype=np.float32, count=n).reshape(shape)


hX, hY = pinned(X.shape), pinned(Y.shape)
hX[...] = X
hY[...] = Y
n_params = sum(w.size for w in Ws) + sum(b.size for b in bs)
sets = []
for _ in range(2):
    bX, bY = CudaDataBuffer.empty(ctx, np.float32, X.size), CudaDataBuffer.empty(ctx, np.float32, Y.size)
    st = {"bX": bX, "bY": bY, "eX": DNDArray(ShapedDataBufferView(bX, *X.shape)), "eY": DNDArray(ShapedDataBufferView(bY, *Y.shape)),
          "hG": pinned((n_params + 1,)), "ev": ctx.event(), "ev_up": ctx.event()}
    bX.upload(X.reshape(-1))
    bY.upload(Y.reshape(-1))
    for _ in range(2):
        r = D.sharded_mlp_grad(comm, st["eX"], st["eY"], dWs, dbs)
        ctx.sync()
    backend.capture_scalars = []
    ctx.graph_begin()
    try:
        _, cW, cb = D.sharded_mlp_grad(comm, st["eX"], st["eY"], dWs, dbs)
    finally:
        st["graph"] = ctx.graph_end()
        st["loss"], backend.capture_scalars = backend.capture_scalars[0], None
    st["outs"] = [w.Buffer.DataBuffer for w in cW] + [v.Buffer for v in cb]
    st["keep"] = (cW, cb)
    sets.append(st)

T = defaultdict(float)


class timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.t = time.perf_counter()

    def __exit__(self, *a):
        T[self.name] += time.perf_counter() - self.t


def upload(st):
    check(lib.dsc_copy_upload(st["bX"].handle, 0, hX.ctypes.data, X.size))
    check(lib.dsc_copy_upload(st["bY"].handle, 0, hY.ctypes.data, Y.size))
    check(lib.dsc_event_record_copy(ctx.handle, st["ev_up"]))


def download(st):
    off = 0
    for a in st["outs"]:
        check(lib.dsc_copy_download(a.handle, 0, st["hG"].ctypes.data + 4 * off, a.Length))
        off += a.Length
    check(lib.dsc_copy_download(st["loss"].handle, 0, st["hG"].ctypes.data + 4 * n_params, 1))
    check(lib.dsc_event_record_copy_down(ctx.handle, st["ev"]))


def loop(n):
    upload(sets[0])
    upload(sets[1])
    for i in range(n):
        st = sets[i % 2]
        with timed("wait events (2 calls)"):
            check(lib.dsc_compute_wait_event(ctx.handle, st["ev_up"]))
            if i >= 2:
                check(lib.dsc_compute_wait_event(ctx.handle, st["ev"]))
        with timed("graph launch"):
            ctx.graph_launch(st["graph"])
        with timed("copy_wait_compute"):
            check(lib.dsc_copy_wait_compute(ctx.handle))
        with timed("queue downloads (7 copies + event)"):
            download(st)
        with timed("queue uploads (2 copies + event)"):
            upload(st)
        if i > 0:
            with timed("event sync + read loss"):
                check(lib.dsc_event_sync(ctx.handle, sets[(i - 1) % 2]["ev"]))
                float(sets[(i - 1) % 2]["hG"][n_params])
    check(lib.dsc_compute_wait_copy(ctx.handle))


loop(10)
ctx.sync()
check(lib.dsc_copy_sync(ctx.handle))
T.clear()
N = 400
t0 = time.perf_counter()
loop(N)
ctx.sync()
check(lib.dsc_copy_sync(ctx.handle))
wall = time.perf_counter() - t0
print(f"rows {rows}: {wall / N * 1e6:.1f} us per step (wall), host time per step by call group:")
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f"  {v / N * 1e6:7.1f} us  {k}")
