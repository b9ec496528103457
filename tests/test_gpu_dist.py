"""N > 1 on real GPUs: batch-row shards + ONE all-reduce of the adjoints (peer-memory kernel over
NVLink, NCCL when P2P is unavailable) reproduce the single-GPU gradient; tangent-row shards + all-gather reproduce the Jacobian.  Needs >= 2 GPUs
(skipped on the 1-GPU box; the same host logic runs on CPU/gloo in test_dist_cpu.py)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        from sigma.diffsharp_b200 import _lib
        n = C.c_int()
        return n.value if _lib.lib.dsc_device_count(C.byref(n)) == 0 else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from sigma.diffsharp_b200 import dist as D, workloads as wl
    from sigma.diffsharp_b200.buffers import CudaDataBuffer
    from sigma.diffsharp_b200.config import CudaBackendProvider, GlobalConfig
    D.init_process_group()
    prov = CudaBackendProvider(device=rank)
    GlobalConfig.BackendProvider = prov
    comm = D.NcclComm(prov.ctx)
    layers, B = [96, 80, 64, 10], 300
    r0, r1 = D.shard_range(B, rank, world)
    X, Y, Ws, bs = wl.mlp_inputs(layers, B, row_range=(r0, r1))
    L, dW, db = D.sharded_mlp_grad(comm, wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    comm.join()
    grads = [a.ToArray().copy() for a in dW + db]
    # the same at a size whose products run on the tensor cores and whose weight adjoints, bias column sums and first
    # all-reduce go to the second compute lane (csrc/dsc_lane.cu): eagerly, and as a replayed CUDA graph of the step
    layers2, B2 = [160, 256, 192, 10], 512
    q0, q1 = D.shard_range(B2, rank, world)
    X2, Y2, Ws2, bs2 = wl.mlp_inputs(layers2, B2, seed=7, row_range=(q0, q1))
    dX2, dY2, dWs2, dbs2 = wl.to_DM(X2), wl.to_DM(Y2), [wl.to_DM(w) for w in Ws2], [wl.to_DV(b) for b in bs2]
    for _ in range(2):
        _, dW2, db2 = D.sharded_mlp_grad(comm, dX2, dY2, dWs2, dbs2)
        comm.join()
        grads2 = [a.ToArray().copy() for a in dW2 + db2]
    backend = prov.GetBackend(np.float32(0)).BackendHandle
    backend.capture_scalars = []
    prov.ctx.graph_begin()
    try:
        _, gW2, gb2 = D.sharded_mlp_grad(comm, dX2, dY2, dWs2, dbs2)
    finally:
        g2 = prov.ctx.graph_end()
        backend.capture_scalars = None
    for _ in range(3):
        prov.ctx.graph_launch(g2)
    prov.ctx.sync()
    grads2g = [a.ToArray().copy() for a in gW2 + gb2]
    del gW2, gb2
    prov.ctx.graph_destroy(g2)
    # the all-reduce is deterministic and replayable: repeat it eagerly and from a CUDA graph on fresh
    # copies of known data and compare against world * data (exact: every rank contributes the same values)
    st = prov.ctx.stats()
    rng = np.random.default_rng(5)
    host = [rng.standard_normal(n).astype(np.float32) for n in (1000, 37, 4096 * 3 + 1)]
    for use_graph in (False, True):
        bufs = [CudaDataBuffer.from_host(prov.ctx, h) for h in host]
        if use_graph:
            comm.allreduce_sum(bufs); comm.join(); prov.ctx.sync()          # warm (sizes the staging buffers)
            for b, h in zip(bufs, host):
                b.upload(h)
            prov.ctx.graph_begin()
            try:
                comm.allreduce_sum(bufs); comm.join()
            finally:
                g = prov.ctx.graph_end()
            for rep in range(3):
                for b, h in zip(bufs, host):
                    b.upload(h)
                prov.ctx.graph_launch(g)
                prov.ctx.sync()
                for b, h in zip(bufs, host):
                    assert np.array_equal(b.SubData, h * np.float32(world)), ("graph", rep)
            prov.ctx.graph_destroy(g)
        else:
            for rep in range(3):
                comm.allreduce_sum(bufs); comm.join()
            for b, h in zip(bufs, host):
                assert np.array_equal(b.SubData, h * np.float32(world ** 3)), "eager"
    # float64 buffers, a window that starts at an odd element (not 8- or 16-byte aligned: the element-wise paths of the
    # kernel) and a one-element buffer, in ONE call per dtype
    h64 = [rng.standard_normal(n) for n in (5, 1025, 1)]
    b64 = [CudaDataBuffer.from_host(prov.ctx, h) for h in h64]
    comm.allreduce_sum(b64); comm.join()
    for b, h in zip(b64, h64):
        assert np.array_equal(b.SubData, h * world), "float64"
    base = rng.standard_normal(4099).astype(np.float32)
    whole = CudaDataBuffer.from_host(prov.ctx, base)
    odd = whole.GetValues(3, 4001)                      # shares storage with `whole`, starts 12 bytes into it
    one = CudaDataBuffer.from_host(prov.ctx, np.array([1.5], np.float32))
    comm.allreduce_sum([odd, one]); comm.join()
    want = base.copy()
    want[3:4004] *= np.float32(world)
    assert np.array_equal(whole.SubData, want) and one.SubData[0] == np.float32(1.5 * world), "odd window"
    st1 = prov.ctx.stats()
    used_p2p = st1["allreduce_p2p"] - st["allreduce_p2p"]
    # config-4 shape in miniature: tangent rows sharded, all-gather at the end
    n = 64
    W1, b1, W2, b2, x = wl.jacobian_inputs(n)
    c0, c1 = D.shard_range(n, rank, world)
    J = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), c0, c1 - c0)
    full = CudaDataBuffer.empty(prov.ctx, np.float64, n * n)
    comm.allgather(J.Buffer.DataBuffer, full)
    comm.join()
    Jfull = full.SubData.reshape(n, n)
    if rank == 0:
        out.put((float(L), grads, Jfull, used_p2p, st1["allreduce_nccl"], grads2, grads2g))
    D.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_two_gpu_sharded_gradient_and_jacobian():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    L0, grads, Jfull, used_p2p, used_nccl, grads2, grads2g = q.get()
    print(f"all-reduces: peer-memory {used_p2p}, NCCL {used_nccl}")
    assert used_p2p + used_nccl > 0
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    from oracle.backend import OracleBackend
    from sigma.diffsharp_b200 import workloads as wl
    from sigma.diffsharp_b200.config import DictBackendProvider, GlobalConfig
    old = GlobalConfig.BackendProvider
    GlobalConfig.BackendProvider = DictBackendProvider(OracleBackend(np.float32), OracleBackend(np.float64))
    try:
        X, Y, Ws, bs = wl.mlp_inputs([96, 80, 64, 10], 300)
        L, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
        for g, a in zip(grads, dW + db):
            ref = a.ToArray()
            assert np.abs(g - ref).max() <= 1e-5 * np.abs(ref).max()
        X, Y, Ws, bs = wl.mlp_inputs([160, 256, 192, 10], 512, seed=7)
        _, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
        for g, gg, a in zip(grads2, grads2g, dW + db):
            ref = a.ToArray()
            assert np.abs(g - ref).max() <= 1e-5 * np.abs(ref).max()
            assert np.array_equal(g, gg)   # the replayed graph (lane branches included) gives the bits of the eager step
        W1, b1, W2, b2, x = wl.jacobian_inputs(64)
        ref = wl.jacobian_closed_form(W1, b1, W2, b2, x)
        assert np.abs(Jfull - ref).max() <= 1e-12 * np.abs(ref).max()
    finally:
        GlobalConfig.BackendProvider = old
