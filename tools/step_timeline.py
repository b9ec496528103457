#!/usr/bin/env python
"""Device timeline of ONE replay of the config-3 step's CUDA graph (python tools/step_timeline.py [rows]; under torchrun the
step includes the adjoint all-reduce and rank 0 prints).  CUPTI activity records (through torch.profiler: plumbing only)
give every kernel's start and duration as the GPU ran it — concurrent lanes and gaps included — which the serialised
`ncu` launch list cannot show.  Prints start offset, duration and stream per kernel of the median replay."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from sigma.diffsharp_b200 import dist as D, workloads as wl  # noqa: E402
from sigma.diffsharp_b200.config import CudaBackendProvider, GlobalConfig  # noqa: E402

LAYERS = [784, 1024, 1024, 10]


def main():
    rank, world = D.init_process_group()
    local = int(os.environ.get("LOCAL_RANK", 0))
    batch = int(sys.argv[1]) * world if len(sys.argv) > 1 else 8192
    torch.cuda.set_device(local)
    provider = CudaBackendProvider(device=local)
    GlobalConfig.BackendProvider = provider
    ctx = provider.ctx
    backend = provider.GetBackend(np.float32(0)).BackendHandle
    comm = D.NcclComm(ctx)
    r0, r1 = D.shard_range(batch, rank, world)
    X, Y, Ws, bs = wl.mlp_inputs(LAYERS, batch, row_range=(r0, r1))
    dX, dY = wl.to_DM(X), wl.to_DM(Y)
    dWs, dbs = [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs]
    for _ in range(3):
        ctx.invalidate_caches()
        res = D.sharded_mlp_grad(comm, dX, dY, dWs, dbs)
        comm.join()
        ctx.sync()        # the deferred products run while their handles are alive
        del res
    backend.capture_scalars = []
    ctx.graph_begin()
    try:
        keep = D.sharded_mlp_grad(comm, dX, dY, dWs, dbs)
    finally:
        graph = ctx.graph_end()
        backend.capture_scalars = None
    for _ in range(20):
        ctx.graph_launch(graph)
    ctx.sync()
    D.barrier()
    reps = 9
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            ctx.graph_launch(graph)
        ctx.sync()
    D.barrier()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // reps
    if rank == 0 and per * reps == len(evs):
        groups = [evs[i * per:(i + 1) * per] for i in range(reps)]
        spans = [g[-1].time_range.end - g[0].time_range.start for g in groups]
        g = groups[int(np.argsort(spans)[len(spans) // 2])]
        t0 = g[0].time_range.start
        print(f"rows per rank {r1 - r0}, world {world}: {per} kernels per replay, span {(g[-1].time_range.end - t0):.1f} us "
              f"(replays: {', '.join(f'{s:.0f}' for s in spans)})")
        busy = 0.0
        for e in g:
            busy += e.time_range.end - e.time_range.start
            print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:110]}")
        print(f"sum of kernel durations {busy:.1f} us")
    elif rank == 0:
        print(f"{len(evs)} kernel records over {reps} replays: not a multiple; dumping all")
        for e in evs[:200]:
            print(f"{e.time_range.start:12.1f} +{e.time_range.end - e.time_range.start:7.1f}  {e.name[:100]}")
    if world > 1:
        import ctypes as C
        from sigma.diffsharp_b200._lib import check, lib
        names = ["pack", "barrier", "reduce+broadcast", "barrier", "unpack"]
        for back, label in ((1, "the all-reduce before the last"), (0, "the last all-reduce")):
            st = (C.c_uint64 * 6)()
            check(lib.dsc_comm_debug_stamps_at(ctx.handle, back, st))
            if rank == 0:
                print(f"{label}, CTA 0: " + " | ".join(f"{n} {(st[i + 1] - st[i]) / 1e3:.1f}" for i, n in enumerate(names)) + f" | total {(st[5] - st[0]) / 1e3:.1f} us")
    del keep
    ctx.graph_destroy(graph)
    comm.close()


if __name__ == "__main__":
    main()
