"""Data-buffer types of the plugin boundary.

Mirrors (same names, same members) the reference's
  * ``ISigmaDiffDataBuffer<'T>``   src/DiffSharp/Util.fs:133-141
  * ``NativeDataBuffer<'T>``       src/DiffSharp/Util.fs:147-165   (host array)
  * ``ShapedDataBufferView<'T>``   src/DiffSharp/Util.fs:167-215   (buffer + int64[] shape, row-major)
and adds ``CudaDataBuffer``: the device-resident implementation whose storage lives in HBM behind
a ``dsc_buf_t`` handle of libdiffsharp_cuda.so.  ``Data``/``SubData`` are lazy downloads; nothing on
the AD path touches them, so reverse sweeps never round-trip through host memory.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib


def _np_dtype(dtype) -> np.dtype:
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("Unsupported type, try float32 or float")  # Config.fs:120
    return dt


class ISigmaDiffDataBuffer:
    """Interface (Util.fs:133-141).  ``dtype`` is the Python stand-in for the generic 'T."""

    dtype: np.dtype

    @property
    def Length(self) -> int: raise NotImplementedError
    @property
    def Offset(self) -> int: raise NotImplementedError
    @property
    def Data(self) -> np.ndarray: raise NotImplementedError
    @property
    def SubData(self) -> np.ndarray: raise NotImplementedError
    def GetValues(self, startIndex: int, length: int) -> "ISigmaDiffDataBuffer": raise NotImplementedError
    def GetStackedValues(self, rows, cols, rowStart, rowFinish, colStart, colFinish) -> "ISigmaDiffDataBuffer": raise NotImplementedError
    def DeepCopy(self) -> "ISigmaDiffDataBuffer": raise NotImplementedError
    def ShallowCopy(self) -> "ISigmaDiffDataBuffer": raise NotImplementedError


class NativeDataBuffer(ISigmaDiffDataBuffer):
    """Host buffer (Util.fs:147-165).  The AD layer creates these itself (DVector.ZeroN,
    AD.Float32.fs:707; DNDArray.Zero, :1833), so every backend receives them as *foreign* buffers.

    Deliberate differences from the reference class, whose slicing members are broken
    (SURVEY App. D-10): ``GetValues`` honours ``startIndex`` (the reference adds the constructor's
    ``offset`` instead, Util.fs:159) and ``GetStackedValues`` is implemented (the reference throws,
    Util.fs:160)."""

    def __init__(self, data, offset: int = 0, length: int | None = None):
        self._data = np.ascontiguousarray(data)
        if self._data.ndim != 1:
            self._data = self._data.reshape(-1)
        self.dtype = _np_dtype(self._data.dtype)
        self._offset = int(offset)
        self._length = int(self._data.size if length is None else length)

    Length = property(lambda s: s._length)
    Offset = property(lambda s: s._offset)
    Data = property(lambda s: s._data)

    @property
    def SubData(self):
        return self._data[self._offset:self._offset + self._length].copy()  # a COPY, like data.[o..o+l-1]

    def window(self) -> np.ndarray:
        """The live window (not part of the reference interface; host backends use it)."""
        return self._data[self._offset:self._offset + self._length]

    def GetValues(self, startIndex, length):
        return NativeDataBuffer(self._data, self._offset + startIndex, length)

    def GetStackedValues(self, rows, cols, rowStart, rowFinish, colStart, colFinish):
        m = self.window()[: rows * cols].reshape(rows, cols)
        return NativeDataBuffer(m[rowStart:rowFinish + 1, colStart:colFinish + 1].reshape(-1).copy())

    def ShallowCopy(self):
        return NativeDataBuffer(self._data, self._offset, self._length)

    def DeepCopy(self):
        return NativeDataBuffer(self._data.copy(), self._offset, self._length)

    def __repr__(self):
        return f"DataBuffer-{self._length} {self.SubData!r}"


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
    """Device-resident ``ISigmaDiffDataBuffer`` (SURVEY §8b ``CudaDataBuffer<'T>``).

    Owns one ``dsc_buf_t``; the handle is released when the object is collected (the managed
    twin is a SafeHandle)."""

    def __init__(self, ctx: CudaContext, dtype, handle: int, length: int, offset: int = 0):
        self.ctx = ctx
        self.dtype = _np_dtype(dtype)
        self.handle = handle
        self._length = int(length)
        self._offset = int(offset)  # offset of the window inside the underlying allocation
        self._lazy_t = None
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

    # -- ISigmaDiffDataBuffer -------------------------------------------------------------------
    Length = property(lambda s: s._length)
    Offset = property(lambda s: s._offset)

    @property
    def SubData(self) -> np.ndarray:
        out = np.empty(self._length, self.dtype)
        if self._length:
            check(lib.dsc_buf_download(self.handle, 0, out.ctypes.data, self._length))
        return out

    @property
    def Data(self) -> np.ndarray:
        # The reference exposes the whole underlying array; consumers index it at Offset.  We hand
        # back an array that is valid at [Offset, Offset+Length) (the only range the reference's own
        # indexers touch: Util.fs:190-197).
        out = np.zeros(self._offset + self._length, self.dtype)
        out[self._offset:] = self.SubData
        return out

    def GetValues(self, startIndex, length):
        h = _lib.buf_t()
        check(lib.dsc_buf_view(self.handle, startIndex, length, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, length, self._offset + startIndex)

    def GetStackedValues(self, rows, cols, rowStart, rowFinish, colStart, colFinish):
        h = _lib.buf_t()
        check(lib.dsc_buf_gather2d(self.handle, rows, cols, rowStart, rowFinish, colStart, colFinish, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, (rowFinish - rowStart + 1) * (colFinish - colStart + 1))

    def DeepCopy(self):
        h = _lib.buf_t()
        check(lib.dsc_buf_copy(self.handle, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, self._length)

    def ShallowCopy(self):
        return self.GetValues(0, self._length)

    def upload(self, array, offset: int = 0):
        a = np.ascontiguousarray(array, dtype=self.dtype).reshape(-1)
        if a.size:
            check(lib.dsc_buf_upload(self.handle, offset, a.ctypes.data, a.size))

    def __repr__(self):
        return f"CudaDataBuffer-{self._length} ({self.dtype}, device {self.ctx.device})"


class ShapedDataBufferView:
    """Buffer + shape, row-major from the buffer's window start (Util.fs:167-215)."""

    __slots__ = ("_buffer", "_shape")

    def __init__(self, buffer: ISigmaDiffDataBuffer, *shape):
        if len(shape) == 1 and not np.isscalar(shape[0]):
            shape = tuple(shape[0])
        self._buffer = buffer
        self._shape = self._checkshape(tuple(int(s) for s in shape))

    def _checkshape(self, s):
        if len(s) <= 1:
            raise RuntimeError(f"Internal error: Cannot create shaped data buffer view without columns dimension (rank <= 1, _shape: {s})")
        total = 1
        for e in s:
            total *= e
        if total > self._buffer.Length:
            raise RuntimeError(f"Internal error: Cannot create shaped data buffer view with smaller buffer than total shape length (buffer length = {self._buffer.Length}, shape = {s})")
        return s

    DataBuffer = property(lambda s: s._buffer)
    Rows = property(lambda s: s._shape[0])
    Cols = property(lambda s: s._shape[1])
    Length = property(lambda s: s._buffer.Length)
    dtype = property(lambda s: s._buffer.dtype)

    @property
    def Shape(self):
        return self._shape

    @Shape.setter
    def Shape(self, v):
        self._shape = self._checkshape(tuple(int(e) for e in v))

    def __getitem__(self, ij):
        i, j = ij
        b = self._buffer
        return b.GetValues(i * self.Cols + j, 1).SubData[0]

    def FlatItem(self, i):
        return self._buffer.GetValues(i, 1).SubData[0]

    def GetRowSlice(self, row, colStart=None, colFinish=None):
        """``d.[row, c0..c1]`` (Util.fs:198-202).  The reference passes ``row * Cols`` as the start and
        ignores colStart; we honour it."""
        colStart = 0 if colStart is None else colStart
        colFinish = self.Cols - 1 if colFinish is None else colFinish
        n = colFinish - colStart + 1
        return ShapedDataBufferView(self._buffer.GetValues(row * self.Cols + colStart, n), 1, n)

    def GetSlice(self, rowStart=None, rowFinish=None, colStart=None, colFinish=None):
        """``d.[r0..r1, c0..c1]`` (Util.fs:204-210)."""
        rowStart = 0 if rowStart is None else rowStart
        rowFinish = self.Rows - 1 if rowFinish is None else rowFinish
        colStart = 0 if colStart is None else colStart
        colFinish = self.Cols - 1 if colFinish is None else colFinish
        return ShapedDataBufferView(
            self._buffer.GetStackedValues(self.Rows, self.Cols, rowStart, rowFinish, colStart, colFinish),
            rowFinish - rowStart + 1, colFinish - colStart + 1)

    def ShallowCopy(self):
        return ShapedDataBufferView(self._buffer.ShallowCopy(), *self._shape)

    def DeepCopy(self):
        return ShapedDataBufferView(self._buffer.DeepCopy(), *self._shape)

    def to_numpy(self) -> np.ndarray:
        n = 1
        for e in self._shape:
            n *= e
        return self._buffer.SubData[:n].reshape(self._shape)

    def __repr__(self):
        return f"ShapedDataBuffer view {self._buffer!r} as {self._shape}"
