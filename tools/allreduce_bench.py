#!/usr/bin/env python
"""Device time of the adjoint all-reduce of the config-3 step (six buffers, 1 863 690 f32 = 7.45 MB)
in isolation: R all-reduces captured in one CUDA graph, max over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_bench.py
  DSC_P2P=0 ... (same)     # NCCL path (pack -> ncclAllReduce -> unpack) for comparison
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma.diffsharp_b200 import dist as D  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaContext, CudaDataBuffer  # noqa: E402

rank, world = D.init_process_group()
ctx = CudaContext(int(os.environ.get("LOCAL_RANK", 0)))
comm = D.NcclComm(ctx)
sizes = (784 * 1024, 1024 * 1024, 1024 * 10, 1024, 1024, 10)
rng = np.random.default_rng(rank)
host = [rng.standard_normal(n).astype(np.float32) for n in sizes]
bufs = [CudaDataBuffer.from_host(ctx, h) for h in host]
# correctness first: the sum over ranks, added in rank order in float32, bit for bit
comm.allreduce_sum(bufs)
comm.join()
ctx.sync()
for n, b in zip(sizes, bufs):
    want = None
    for r in range(world):
        rr = np.random.default_rng(r)
        parts = [rr.standard_normal(m).astype(np.float32) for m in sizes]
        x = parts[sizes.index(n)] if sizes.count(n) == 1 else parts[[i for i, m in enumerate(sizes) if m == n][[bb is b for bb in bufs if bb.Length == n].index(True)]]
        want = x.copy() if want is None else (want + x).astype(np.float32)
    got = b.SubData
    assert np.array_equal(got, want), (rank, n, float(np.abs(got - want).max()))
if rank == 0:
    print(f"world {world}: all-reduce result equals the rank-ordered float32 sum, bit for bit", flush=True)
for _ in range(3):
    comm.allreduce_sum(bufs)
    comm.join()
ctx.sync()
R = 20
ctx.graph_begin()
try:
    for _ in range(R):
        comm.allreduce_sum(bufs)
        comm.join()
finally:
    g = ctx.graph_end()
e0, e1 = ctx.event(), ctx.event()
ctx.graph_launch(g)
ctx.sync()
best = 1e30
for _ in range(5):
    D.barrier()
    ctx.record(e0)
    ctx.graph_launch(g)
    ctx.record(e1)
    ctx.sync()
    best = min(best, D.max_over_ranks(ctx.elapsed_ms(e0, e1)))
st = ctx.stats()
if st["allreduce_p2p"]:
    import ctypes as C
    from sigma.diffsharp_b200._lib import check, lib
    stamps = (C.c_uint64 * 6)()
    check(lib.dsc_comm_debug_stamps(ctx.handle, stamps))
    d = [(stamps[i + 1] - stamps[i]) / 1e3 for i in range(5)]
    print(f"rank {rank}: last call, CTA 0: pack {d[0]:.1f} us | barrier {d[1]:.1f} | reduce+broadcast {d[2]:.1f} | barrier {d[3]:.1f} | unpack {d[4]:.1f}", flush=True)
D.barrier()
if rank == 0:
    nbytes = sum(sizes) * 4
    print(f"world {world}: all-reduce of {nbytes / 1e6:.2f} MB: {best / R * 1e3:.1f} us per call "
          f"(peer-memory kernel calls {st['allreduce_p2p']}, NCCL calls {st['allreduce_nccl']})", flush=True)
ctx.graph_destroy(g)
D.barrier()
comm.close()
