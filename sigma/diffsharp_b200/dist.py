"""Sharding of the two batched-gradient workloads over the GPUs of one node.

One process per GPU.  The reference has no distributed code; this is new surface:

  * reverse-mode MLP loss gradient: rank r owns the contiguous batch-row block
    [r*B/g, (r+1)*B/g); weights are replicated; every rank runs the identical AD program on its
    block; the weight/bias adjoints (``W.A`` / ``b.A``) are summed over the ranks with ONE exchange step,
    ``dsc_?allreduce_sum`` (one kernel over peer memory, flag-in-data lines over NVLink; NCCL when peer
    access is unavailable).
  * forward-mode Jacobian: rank r owns the identity-tangent rows [r*n/g, (r+1)*n/g); no
    communication until the final all-gather (``dsc_?allgather``).

``torch.distributed`` is plumbing only: rendezvous, the broadcast of the NCCL unique id, barriers and
the max-over-ranks of timings (gloo, CPU tensors).  The data path never goes through torch.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np



def _native():
    """the ctypes binding, imported on first use so that the gloo-only helpers (shard_range, GlooHostComm) work without the CUDA library"""
    from . import _lib
    return _lib


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of `total` units owned by `rank`; blocks differ by at most one unit."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def env_rank() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group():
    """gloo control plane (rendezvous / barrier / small CPU reductions)."""
    import torch.distributed as dist
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


class NcclComm:
    """NCCL communicator bound to one CudaContext (dsc_comm_*)."""

    def __init__(self, ctx):
        import torch
        import torch.distributed as dist
        self.ctx = ctx
        self.rank, self.world, _ = env_rank()
        _lib = _native()
        check, lib = _lib.check, _lib.lib
        if self.world == 1:
            check(lib.dsc_comm_init(ctx.handle, 1, 0, None))
            return
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            check(lib.dsc_comm_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        dist.broadcast(ident, src=0)
        raw = (C.c_uint8 * 128)(*ident.tolist())
        check(lib.dsc_comm_init(ctx.handle, self.world, self.rank, raw))

    def allreduce_sum(self, buffers):
        """In-place sum over ranks of a list of CudaDataBuffers of one dtype (one NCCL group)."""
        if not buffers:
            return
        _lib = _native()
        p = "s" if buffers[0].dtype == np.float32 else "d"
        for b in buffers:
            b._flush_host_writes()
        arr = (_lib.buf_t * len(buffers))(*[b.handle for b in buffers])
        _lib.check(getattr(_lib.lib, f"dsc_{p}allreduce_sum")(self.ctx.handle, len(buffers), arr))
        for b in buffers:
            b._invalidate_host()

    def allgather(self, send, recv):
        _lib = _native()
        p = "s" if send.dtype == np.float32 else "d"
        send._flush_host_writes()
        _lib.check(getattr(_lib.lib, f"dsc_{p}allgather")(self.ctx.handle, send.handle, recv.handle))
        recv._invalidate_host()

    def join(self):
        _lib = _native()
        _lib.check(_lib.lib.dsc_comm_join(self.ctx.handle))

    def close(self):
        _lib = _native()
        _lib.check(_lib.lib.dsc_comm_destroy(self.ctx.handle))


class GlooHostComm:
    """Host twin of NcclComm for NativeDataBuffers — used by the CPU (gloo, world_size 2) tests of the
    sharding logic; not a product path."""

    def __init__(self):
        self.rank, self.world, _ = env_rank()

    def allreduce_sum(self, buffers):
        import torch
        import torch.distributed as dist
        for b in buffers:
            w = b.window()
            t = torch.from_numpy(w)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def allgather(self, send, recv):
        import torch
        import torch.distributed as dist
        s = torch.from_numpy(send.window())
        parts = [torch.empty_like(s) for _ in range(self.world)]
        dist.all_gather(parts, s)
        recv.window()[:] = torch.cat(parts).numpy()

    def join(self):
        pass


def sharded_mlp_grad(comm, X_shard, Y_shard, Ws, bs):
    """One data-parallel gradient step: local reverse sweep on this rank's rows, then one all-reduce
    of every weight and bias adjoint.  Returns (local loss D, [dW], [db]) with GLOBAL adjoints."""
    from . import workloads as wl
    L, dW, db = wl.mlp_grad(X_shard, Y_shard, Ws, bs)
    # ONE collective for all adjoints after the sweep.  Measured alternative (profiles/README.md): summing the later layers'
    # adjoints on a second stream while the first layer's backward products run — an all-reduce confined to the SMs a GEMM
    # leaves free is NVLink-starved and ended later than this one.
    comm.allreduce_sum([a.Buffer.DataBuffer for a in dW] + [a.Buffer for a in db])
    return L, dW, db
