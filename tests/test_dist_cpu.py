"""Host-side logic of the N>1 path on CPU: gloo, world_size 2 (and 3), CPU oracle backend.
Row-block sharding + one all-reduce of the adjoints reproduces the single-process gradient; the
tangent-row sharding + all-gather reproduces the Jacobian."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from sigma.diffsharp_b200.dist import shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 8192, 4096, 1000003):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from oracle.backend import OracleBackend, set_num_threads
    from sigma.diffsharp_b200 import dist as D, workloads as wl
    from sigma.diffsharp_b200.buffers import NativeDataBuffer
    from sigma.diffsharp_b200.config import DictBackendProvider, GlobalConfig
    set_num_threads(1)
    D.init_process_group()
    GlobalConfig.BackendProvider = DictBackendProvider(OracleBackend(np.float32), OracleBackend(np.float64))
    comm = D.GlooHostComm()
    layers, B = [20, 16, 12, 5], 37
    r0, r1 = D.shard_range(B, rank, world)
    X, Y, Ws, bs = wl.mlp_inputs(layers, B, row_range=(r0, r1))
    L, dW, db = D.sharded_mlp_grad(comm, wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    grads = [a.ToArray().copy() for a in dW + db]
    # Jacobian rows
    W1, b1, W2, b2, x = wl.jacobian_inputs(24)
    c0, c1 = D.shard_range(24, rank, world)
    J = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), c0, c1 - c0)
    if 24 % world == 0:
        full = NativeDataBuffer(np.zeros(24 * 24))
        comm.allgather(NativeDataBuffer(J.ToArray().reshape(-1).copy()), full)
        Jfull = full.window().reshape(24, 24).copy()
    else:
        Jfull = None
    if rank == 0:
        out.put((float(L), grads, Jfull))
    D.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gradient_and_jacobian_equal_single_process(world, oracle_provider, use_provider):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    L0, grads, Jfull = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    from sigma.diffsharp_b200 import workloads as wl
    use_provider(oracle_provider)
    X, Y, Ws, bs = wl.mlp_inputs([20, 16, 12, 5], 37)
    L, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    for g, a in zip(grads, dW + db):
        ref = a.ToArray()
        assert np.abs(g - ref).max() <= 1e-5 * np.abs(ref).max()
    if Jfull is not None:
        W1, b1, W2, b2, x = wl.jacobian_inputs(24)
        ref = wl.jacobian_closed_form(W1, b1, W2, b2, x)
        assert np.abs(Jfull - ref).max() <= 1e-12 * np.abs(ref).max()
