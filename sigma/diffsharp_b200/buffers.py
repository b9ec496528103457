"""Data-buffer types of the plugin boundary.

Mirrors (same names, same members) the reference's
  * ``ISigmaDiffDataBuffer<'T>``   src/DiffSharp/Util.fs:133-141
  * ``NativeDataBuffer<'T>``       src/DiffSharp/Util.fs:147-165   (host array)
  * ``ShapedDataBufferView<'T>``   src/DiffSharp/Util.fs:167-215   (buffer + int64[] shape, row-major)
The device-resident implementation, ``CudaDataBuffer`` (storage in HBM behind a ``dsc_buf_t`` handle of
libdiffsharp_cuda.so), and ``CudaContext`` live in device.py; they are reachable from here by name
(``from .buffers import CudaDataBuffer``) but importing THIS module never loads the CUDA library, so the CPU oracle and the
reference arm of bench.py run without it.
"""
from __future__ import annotations

import numpy as np


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


def __getattr__(name):
    # device types, resolved on first use (PEP 562): keeps this module free of the CUDA library
    if name in ("CudaContext", "CudaDataBuffer"):
        from . import device
        return getattr(device, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
