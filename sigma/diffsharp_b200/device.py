"""Device side of the plugin boundary: ``CudaContext`` (dsc_ctx_t) and ``CudaDataBuffer``, the device-resident
``ISigmaDiffDataBuffer<'T>`` (reference src/DiffSharp/Util.fs:133-141; SURVEY §8b ``CudaDataBuffer<'T>``).

Storage lives in HBM behind a ``dsc_buf_t`` handle; nothing on the AD path touches host memory, so reverse sweeps never
round-trip through it.  ``Data`` is the interface's host array (Util.fs:136): a CACHED mirror of the whole underlying
allocation, shared by every view of it (consumers index it at ``Offset``, Util.fs:190-197), downloaded on first access and
dropped when the device side is written; host-side writes into it (``aa.Data.[i] <- ...``) are detected and uploaded before
the next device use of the buffer or of any of its views.  ``SubData`` is a copy of the window (``data.[o..o+l-1]``,
Util.fs:158).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib
from .buffers import ISigmaDiffDataBuffer, _np_dtype


class CudaContext:
    """One context = one GPU = one compute stream (+ one communication stream).  Wraps dsc_ctx_t."""

    def __init__(self, device: int = 0, flags: int = 0):
        h = _lib.ctx_t()
        check(lib.dsc_create(device, flags, C.byref(h)))
        self.handle = h
        self.device = device
        self.flags = flags
        self._finalizer = weakref.finalize(self, lib.dsc_destroy, h)

    def sync(self):
        check(lib.dsc_sync(self.handle))

    def flush(self):
        """run every deferred operation nobody has consumed yet (dsc_flush)"""
        check(lib.dsc_flush(self.handle))

    def set_fuse(self, on: bool):
        check(lib.dsc_set_fuse(self.handle, 1 if on else 0))

    def invalidate_caches(self):
        check(lib.dsc_invalidate_caches(self.handle))

    def stats(self) -> dict:
        s = _lib.Stats()
        check(lib.dsc_get_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def flush_l2(self):
        check(lib.dsc_flush_l2(self.handle))

    def trim_pool(self):
        check(lib.dsc_trim_pool(self.handle))

    # CUDA events on the context's stream
    def event(self) -> int:
        e = C.c_uint64()
        check(lib.dsc_event_create(self.handle, C.byref(e)))
        return e.value

    def record(self, ev: int):
        check(lib.dsc_event_record(self.handle, ev))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        check(lib.dsc_event_elapsed_ms(self.handle, a, b, C.byref(ms)))
        return ms.value

    def event_destroy(self, ev: int):
        check(lib.dsc_event_destroy(self.handle, ev))

    # CUDA graphs
    def graph_begin(self):
        check(lib.dsc_graph_begin(self.handle))

    def graph_end(self) -> int:
        g = C.c_uint64()
        check(lib.dsc_graph_end(self.handle, C.byref(g)))
        return g.value

    def graph_launch(self, g: int):
        check(lib.dsc_graph_launch(self.handle, g))

    def graph_destroy(self, g: int):
        check(lib.dsc_graph_destroy(self.handle, g))

    def close(self):
        self._finalizer()


def _free_handle(h):
    lib.dsc_buf_free(h)


class CudaDataBuffer(ISigmaDiffDataBuffer):
    """Device-resident ``ISigmaDiffDataBuffer``.

    Owns one ``dsc_buf_t``; the handle is released when the object is collected (the managed twin is a SafeHandle)."""

    def __init__(self, ctx: CudaContext, dtype, handle: int, length: int, offset: int = 0, root: "CudaDataBuffer | None" = None):
        self.ctx = ctx
        self.dtype = _np_dtype(dtype)
        self.handle = handle
        self._length = int(length)
        self._offset = int(offset)  # offset of the window inside the underlying allocation
        self._root = root           # the buffer whose allocation this view shares (None: this buffer is the allocation)
        self._mirror = None         # roots only: host mirror of the whole allocation, and its state as last seen on the device
        self._clean = None
        self._finalizer = weakref.finalize(self, _free_handle, handle)

    # -- construction ---------------------------------------------------------------------------
    @staticmethod
    def _code(dtype) -> int:
        return _lib.DSC_F32 if np.dtype(dtype) == np.float32 else _lib.DSC_F64

    @classmethod
    def empty(cls, ctx, dtype, n):
        h = _lib.buf_t()
        check(lib.dsc_buf_alloc(ctx.handle, cls._code(dtype), n, C.byref(h)))
        return cls(ctx, dtype, h.value, n)

    @classmethod
    def filled(cls, ctx, dtype, n, value):
        h = _lib.buf_t()
        check(lib.dsc_buf_alloc_fill(ctx.handle, cls._code(dtype), n, float(value), C.byref(h)))
        return cls(ctx, dtype, h.value, n)

    @classmethod
    def from_host(cls, ctx, array):
        a = np.ascontiguousarray(array).reshape(-1)
        b = cls.empty(ctx, a.dtype, a.size)
        if a.size:
            check(lib.dsc_buf_upload(b.handle, 0, a.ctypes.data, a.size))
        return b

    # -- host mirror ------------------------------------------------------------------------------
    def _owner(self) -> "CudaDataBuffer":
        return self._root if self._root is not None else self

    def _flush_host_writes(self):
        """Called before every device use: if the consumer wrote into the ``Data`` array, upload what changed."""
        r = self._owner()
        if r._mirror is None:
            return
        w = r._mirror[r._offset:r._offset + r._length]
        if not np.array_equal(w, r._clean, equal_nan=True):
            check(lib.dsc_buf_upload(r.handle, 0, w.ctypes.data, r._length))
            r._clean = w.copy()

    def _invalidate_host(self):
        """Called after the device side was written in place: the mirror no longer describes the buffer."""
        r = self._owner()
        r._mirror = r._clean = None

    # -- ISigmaDiffDataBuffer -------------------------------------------------------------------
    Length = property(lambda s: s._length)
    Offset = property(lambda s: s._offset)

    @property
    def SubData(self) -> np.ndarray:
        self._flush_host_writes()
        out = np.empty(self._length, self.dtype)
        if self._length:
            check(lib.dsc_buf_download(self.handle, 0, out.ctypes.data, self._length))
        return out

    @property
    def Data(self) -> np.ndarray:
        # The reference exposes the whole underlying array; consumers index it at Offset (Util.fs:190-197) and may write it.
        r = self._owner()
        if r._mirror is None:
            m = np.zeros(r._offset + r._length, r.dtype)
            if r._length:
                check(lib.dsc_buf_download(r.handle, 0, m[r._offset:].ctypes.data, r._length))
            r._mirror, r._clean = m, m[r._offset:].copy()
        return r._mirror

    def GetValues(self, startIndex, length):
        self._flush_host_writes()
        h = _lib.buf_t()
        check(lib.dsc_buf_view(self.handle, startIndex, length, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, length, self._offset + startIndex, root=self._owner())

    def GetStackedValues(self, rows, cols, rowStart, rowFinish, colStart, colFinish):
        self._flush_host_writes()
        h = _lib.buf_t()
        check(lib.dsc_buf_gather2d(self.handle, rows, cols, rowStart, rowFinish, colStart, colFinish, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, (rowFinish - rowStart + 1) * (colFinish - colStart + 1))

    def DeepCopy(self):
        self._flush_host_writes()
        h = _lib.buf_t()
        check(lib.dsc_buf_copy(self.handle, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, self._length)

    def ShallowCopy(self):
        return self.GetValues(0, self._length)

    def upload(self, array, offset: int = 0):
        a = np.ascontiguousarray(array, dtype=self.dtype).reshape(-1)
        if a.size:
            check(lib.dsc_buf_upload(self.handle, offset, a.ctypes.data, a.size))
        self._invalidate_host()

    def __repr__(self):
        return f"CudaDataBuffer-{self._length} ({self.dtype}, device {self.ctx.device})"
