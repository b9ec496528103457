"""``Backend<'T>`` for B200: the host-side mirror of the reference plugin interface.

``Backend`` lists the 55 abstract members of ``DiffSharp.Backend.Backend<'T>``
(reference src/DiffSharp/Backend.fs:74-139) with the same names, argument order and meaning;
``CudaBackend`` implements them over libdiffsharp_cuda.so (one C-ABI call per member — the Python
twin of the F# P/Invoke shim in INTEGRATION.md).  ``MapOp`` carries the ordinals of the reference
union (Backend.fs:42-70).

Closures: ``Map_F_*``/``Map2_F_*`` receive ``(MapOp, closure, ...)`` like the reference; only the
op-code is forwarded.  An op-code the device does not implement raises ``NotSupportedError`` —
there is no CPU fallback (north_star).

Deferred evaluation behind the unchanged call sequence, all bit-identical to the
eager sequence and switchable off with ``CudaBackend(fuse=False)``:
  * ``Transpose_M`` returns a lazy view; a following ``Mul_M_M`` consumes it as a transposed GEMM
    operand (AD.Float32.fs:4127-4143) and the transpose kernel never runs; any other consumer
    materialises it.
  * the tanh adjoint chain Map(Cosh) -> Map_F_S(1, Div) -> Had -> Had (AD.Float32.fs:4238-4240) is
    recognised and issued as one fused kernel with the same sequence of roundings.
  * the fork's bias broadcast ``RepeatReshapeCopy_V_MRows`` -> ``Add_M_M`` [-> ``Map_F_M``]
    (AD.Float32.fs:1860-1865,2357,2704) is deferred and issued as ONE pass ``dsc_?bias_rows_act``
    (z = A + rows(v); h = f(z)) instead of materialising the m copies of the bias.
  * zero adjoints are lazy: ``CreateDataBuffer(CreateZeroArray n)`` (two per tape node,
    AD.Float32.fs:1834-1835,3746) allocates nothing; the first ``Add_M_M_InPlace``/``Add_V_V`` into
    a known-zero buffer is a COPY-ON-WRITE alias of the addend (``dsc_buf_share``: 0 + v = v; only
    the sign of a zero can differ; the copy happens only if either side is later written in place),
    any other use materialises the zeros with a memset.
"""
from __future__ import annotations

import ctypes as C
import enum
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib
from .buffers import (CudaContext, CudaDataBuffer, ISigmaDiffDataBuffer, NativeDataBuffer,
                      ShapedDataBufferView)


class MapOp(enum.IntEnum):
    """Backend.fs:42-70, declaration order."""
    Mul = 0; Div = 1; Add = 2; Sub = 3; Pow_Ab = 4; Pow_Ba = 5; Atan2_Ab = 6; Atan2_Ba = 7
    Log = 8; Log10 = 9; Exp = 10; Sin = 11; Cos = 12; Tan = 13; Sqrt = 14; Sinh = 15; Cosh = 16
    Tanh = 17; Asin = 18; Acos = 19; Atan = 20; Abs = 21; Sign = 22; Floor = 23; Ceiling = 24
    Round = 25; ReL = 26; Sigmoid = 27


class ConstArray:
    """What ``CreateZeroArray``/``CreateValueArray``/``CreateUninitialisedArray`` hand back: the
    managed shim returns a ``'T[]`` registered in a weak table as "n copies of v" so that
    ``CreateDataBuffer`` can fill on the device instead of uploading.  Here the token
    itself plays the part of the array; it materialises only if somebody indexes it."""

    __slots__ = ("n", "value", "dtype", "_arr")

    def __init__(self, n, value, dtype):
        self.n, self.value, self.dtype, self._arr = int(n), value, np.dtype(dtype), None

    size = property(lambda s: s.n)
    Length = property(lambda s: s.n)

    def __len__(self):
        return self.n

    def materialize(self) -> np.ndarray:
        if self._arr is None:
            self._arr = (np.zeros(self.n, self.dtype) if self.value is None
                         else np.full(self.n, self.value, self.dtype))
        return self._arr

    def __getitem__(self, i):
        return self.materialize()[i]

    def __setitem__(self, i, v):
        self.materialize()[i] = v


class Backend:
    """Interface for DiffSharp backends (Backend.fs:74-139).  'T is ``self.dtype``."""
    dtype: np.dtype
    MEMBERS = (
        "CreateDataBuffer", "CreateUninitialisedArray", "CreateZeroArray", "CreateValueArray",
        "Mul_Dot_V_V", "L1Norm_V", "L2Norm_V", "SupNorm_V", "Sum_V", "Sum_M", "MaxIndex_V", "MinIndex_V",
        "Add_V_V", "Add_V_V_InPlace", "Add_S_V", "Sub_V_V", "Sub_S_V", "Sub_V_S", "Mul_S_V", "Mul_M_V",
        "Mul_M_V_Add_V", "Mul_V_M", "Solve_M_V", "SolveSymmetric_M_V", "Diagonal_M", "Map_F_V", "Map_F_S_V",
        "Map2_F_V_V", "ReshapeCopy_MRows_V", "Mul_Out_V_V", "Add_M_M", "Add_M_M_InPlace", "Add_S_M",
        "Add_V_MCols", "Sub_M_M", "Sub_M_S", "Sub_S_M", "Mul_M_M", "Mul_S_M", "Mul_M_M_Add_V_MCols",
        "Add_M_Colwise_V_InPlace", "Mul_Had_M_M", "Inverse_M", "Det_M", "Transpose_M", "Permute_M",
        "Reshape_M", "Map_F_M", "Map_F_S_M", "Map2_F_M_M", "ReshapeCopy_V_MRows",
        "RepeatReshapeCopy_V_MRows", "RepeatReshapeCopy_V_MCols", "CustomOp_DM_Forward",
        "CustomOp_DM_Backward")


def _size(view: ShapedDataBufferView) -> int:
    n = 1
    for e in view.Shape:
        n *= e
    return n


class _Lazy:
    """A deferred device value: ('T', src_buffer, m, n) for a transpose, or an elementwise chain."""
    __slots__ = ("kind", "args")

    def __init__(self, kind, *args):
        self.kind, self.args = kind, args


class LazyCudaBuffer(CudaDataBuffer):
    """CudaDataBuffer whose device value is computed on first use (see module docstring)."""
    def __init__(self, backend: "CudaBackend", length: int, lazy: _Lazy):
        # no handle yet; `handle` is a property here
        self.ctx = backend.ctx
        self.dtype = backend.dtype
        self._length = int(length)
        self._offset = 0
        self._lazy_t = None
        self._finalizer = None
        self._h = None
        self._lazy = lazy
        self._backend = backend

    @property
    def handle(self):
        if self._h is None:
            self._backend._materialize(self)
        return self._h

    @handle.setter
    def handle(self, v):
        self._h = v


class CudaBackend(Backend):
    """``Backend<'T>`` over libdiffsharp_cuda.so for 'T = float32 or float64."""

    def __init__(self, dtype=np.float32, ctx: CudaContext | None = None, device: int = 0,
                 flags: int = 0, fuse: bool = True):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("Unsupported type, try float32 or float")
        self.ctx = ctx if ctx is not None else CudaContext(device, flags)
        self._p = "s" if self.dtype == np.float32 else "d"
        self._code = _lib.DSC_F32 if self.dtype == np.float32 else _lib.DSC_F64
        self._scalar = C.c_float if self.dtype == np.float32 else C.c_double
        self.fuse = fuse
        self.capture_scalars = None  # list of (device buffer) when Sum_M runs under graph capture
        f = lambda name: getattr(lib, f"dsc_{self._p}{name}")  # noqa: E731
        self._f = {n: f(n) for n in (
            "gemm", "gemv", "ger", "dot", "asum", "nrm2", "amax", "sum", "argmax", "argmin", "sum_dev", "dot_dev",
            "map", "map_s", "map2", "scalar_op", "add_inplace", "add_inplace_range", "fused_bwd", "bias_rows_act", "transpose",
            "permute", "repeat_rows", "repeat_cols", "colsum_inplace", "add_v_mcols", "diag", "gesv", "sysv",
            "getri", "det", "allreduce_sum", "allgather")}
        self._empty = CudaDataBuffer.empty(self.ctx, self.dtype, 0)
        self._h = self.ctx.handle

    # ---- plumbing ---------------------------------------------------------------------------
    def _dev(self, b: ISigmaDiffDataBuffer) -> CudaDataBuffer:
        """Accept any ISigmaDiffDataBuffer: foreign host buffers are uploaded on first touch
        (DVector.ZeroN / DNDArray.Zero reach the backend)."""
        if isinstance(b, CudaDataBuffer):
            if b.ctx is not self.ctx:
                raise _lib.InvalidArgError(_lib.DSC_E_INVALID, "buffer belongs to another context")
            return b
        if b.Length == 0:
            return self._empty
        return CudaDataBuffer.from_host(self.ctx, np.asarray(b.SubData, dtype=self.dtype))

    def _new(self, handle: int, n: int) -> CudaDataBuffer:
        return CudaDataBuffer(self.ctx, self.dtype, handle, n)

    def _out(self, fn, n, *args) -> CudaDataBuffer:
        h = _lib.buf_t()
        check(fn(self._h, *args, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, n)

    def _touch(self, b: CudaDataBuffer):
        """Called before an in-place write to `b`: lazy views of `b` must see the old value."""
        kids = b._lazy_t
        if kids:
            for k in list(kids):
                k.handle  # materialise
            b._lazy_t = None

    def _materialize(self, lb: LazyCudaBuffer):
        lz = lb._lazy
        if lz.kind == "Z":
            h = _lib.buf_t()
            check(lib.dsc_buf_alloc_fill(self._h, self._code, lb._length, 0.0, C.byref(h)))
        elif lz.kind == "R":      # ('R', v, m, n): m copies of v as rows
            v, m, n = lz.args
            h = _lib.buf_t()
            check(self._f["repeat_rows"](self._h, m, v.handle, C.byref(h)))
        elif lz.kind == "AR":     # ('AR', A, v, m, n): A + rows(v)
            A, v, m, n = lz.args
            h = _lib.buf_t()
            check(self._f["bias_rows_act"](self._h, -1, m, n, A.handle, v.handle, C.byref(h), None))
        elif lz.kind == "T":
            src, m, n = lz.args
            h = _lib.buf_t()
            check(self._f["transpose"](self._h, m, n, src.handle, C.byref(h)))
        else:
            h = self._eval_chain(lz)
        lb._h = h.value
        lb._lazy = None
        lb._finalizer = weakref.finalize(lb, lib.dsc_buf_free, h.value)

    @staticmethod
    def _lz(b):
        return b._lazy if isinstance(b, LazyCudaBuffer) and b._h is None else None

    # ---- Create buffer (Backend.fs:76-79) -----------------------------------------------------
    def CreateDataBuffer(self, a) -> ISigmaDiffDataBuffer:
        if isinstance(a, ConstArray) and a._arr is None:
            if a.value is None:
                return CudaDataBuffer.empty(self.ctx, self.dtype, a.n)
            if self.fuse and a.value == 0.0 and a.n > 0:
                return LazyCudaBuffer(self, a.n, _Lazy("Z"))
            return CudaDataBuffer.filled(self.ctx, self.dtype, a.n, a.value)
        if isinstance(a, ConstArray):
            a = a._arr
        return CudaDataBuffer.from_host(self.ctx, np.asarray(a, dtype=self.dtype))

    def CreateUninitialisedArray(self, n):
        return ConstArray(n, None, self.dtype)

    def CreateZeroArray(self, n):
        return ConstArray(n, 0.0, self.dtype)

    def CreateValueArray(self, n, v):
        return ConstArray(n, float(v), self.dtype)

    # ---- Scalar valued (Backend.fs:81-88) -------------------------------------------------------
    def _scalar_red(self, name, *bufs):
        if self.capture_scalars is not None and name in ("sum", "dot"):
            # under CUDA-graph capture the scalar stays on the device (a host scalar needs a sync);
            # the AD layer gets a placeholder: adjoints never depend on the value of a reduction
            devs = [self._dev(b) for b in bufs]
            self.capture_scalars.append(self._out(self._f[name + "_dev"], 1, *[d.handle for d in devs]))
            return self.dtype.type(0)
        out = self._scalar()
        check(self._f[name](self._h, *[self._dev(b).handle for b in bufs], C.byref(out)))
        return self.dtype.type(out.value)

    def Mul_Dot_V_V(self, a, b): return self._scalar_red("dot", a, b)
    def L1Norm_V(self, a): return self._scalar_red("asum", a)
    def L2Norm_V(self, a): return self._scalar_red("nrm2", a)
    def SupNorm_V(self, a): return self._scalar_red("amax", a)
    def Sum_V(self, a): return self._scalar_red("sum", a)

    def Sum_M(self, a): return self._scalar_red("sum", a)

    def _index_red(self, name, a):
        out = C.c_int32()
        check(self._f[name](self._h, self._dev(a).handle, C.byref(out)))
        return int(out.value)

    def MaxIndex_V(self, a): return self._index_red("argmax", a)
    def MinIndex_V(self, a): return self._index_red("argmin", a)

    # ---- Vector valued (Backend.fs:90-109) ----------------------------------------------------
    def _is_zero(self, b):
        lz = self._lz(b)
        return lz is not None and lz.kind == "Z"

    def _copy_of(self, src: CudaDataBuffer) -> CudaDataBuffer:
        h = _lib.buf_t()
        check(lib.dsc_buf_copy(src.handle, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, src.Length)

    def _map2(self, op, a, b):
        da, db = self._dev(a), self._dev(b)
        if op in (MapOp.Add, MapOp.Sub) and da.Length == db.Length:
            if self._is_zero(db):      # x +/- 0 = x
                return LazyCudaBuffer(self, da.Length, _Lazy("Z")) if self._is_zero(da) else self._copy_of(da)
            if op == MapOp.Add and self._is_zero(da):  # 0 + x = x
                return self._copy_of(db)
        return self._out(self._f["map2"], max(da.Length, db.Length), int(op), da.handle, db.handle)

    def _scalar_op(self, kind, s, v):
        d = self._dev(v)
        return self._out(self._f["scalar_op"], d.Length, kind, self._scalar(s), d.handle)

    def Add_V_V(self, a, b): return self._map2(MapOp.Add, a, b)
    def Sub_V_V(self, a, b): return self._map2(MapOp.Sub, a, b)

    def Add_V_V_InPlace(self, src, srcOffset, dst, dstOffset, length):
        d = self._dev(dst)
        self._touch(d)
        check(self._f["add_inplace_range"](self._h, self._dev(src).handle, srcOffset, d.handle, dstOffset, length))
        return dst

    def Add_S_V(self, s, v): return self._scalar_op(0, s, v)
    def Sub_S_V(self, s, v): return self._scalar_op(1, s, v)
    def Sub_V_S(self, v, s): return self._scalar_op(2, s, v)
    def Mul_S_V(self, s, v): return self._scalar_op(3, s, v)

    def Mul_M_V(self, A: ShapedDataBufferView, x):
        dA, dx = self._dev(A.DataBuffer), self._dev(x)
        return self._out(self._f["gemv"], A.Rows, 0, A.Rows, A.Cols, dA.handle, dx.handle, 0)

    def Mul_M_V_Add_V(self, A, x, y):
        dA, dx, dy = self._dev(A.DataBuffer), self._dev(x), self._dev(y)
        return self._out(self._f["gemv"], A.Rows, 0, A.Rows, A.Cols, dA.handle, dx.handle, dy.handle)

    def Mul_V_M(self, x, A):
        dA, dx = self._dev(A.DataBuffer), self._dev(x)
        return self._out(self._f["gemv"], A.Cols, 1, A.Rows, A.Cols, dA.handle, dx.handle, 0)

    def _solve(self, name, A, b):
        dA, db = self._dev(A.DataBuffer), self._dev(b)
        try:
            return self._out(self._f[name], A.Rows, A.Rows, dA.handle, db.handle)
        except _lib.SingularMatrixError:
            return None  # F# option: None

    def Solve_M_V(self, A, b): return self._solve("gesv", A, b)
    def SolveSymmetric_M_V(self, A, b): return self._solve("sysv", A, b)

    def Diagonal_M(self, A):
        d = self._dev(A.DataBuffer)
        return self._out(self._f["diag"], min(A.Rows, A.Cols), A.Rows, A.Cols, d.handle)

    def Map_F_V(self, op, f, v):
        d = self._dev(v)
        return self._out(self._f["map"], d.Length, int(op), d.handle)

    def Map_F_S_V(self, s, op, f, v):
        d = self._dev(v)
        return self._out(self._f["map_s"], d.Length, int(op), self._scalar(s), d.handle)

    def Map2_F_V_V(self, op, f, a, b): return self._map2(op, a, b)

    def ReshapeCopy_MRows_V(self, A):
        d = self._dev(A.DataBuffer)
        n = _size(A)
        src = d if d.Length == n else d.GetValues(0, n)
        return self._out(lambda ctx, h, out: lib.dsc_buf_copy(h, out), n, src.handle)

    # ---- Matrix valued (Backend.fs:111-139) ---------------------------------------------------
    def _view(self, buf, *shape):
        return ShapedDataBufferView(buf, *shape)

    def Mul_Out_V_V(self, a, b):
        da, db = self._dev(a), self._dev(b)
        return self._view(self._out(self._f["ger"], da.Length * db.Length, da.handle, db.handle), da.Length, db.Length)

    def Add_M_M(self, a, b):
        if self.fuse:
            la, lb_ = self._lz(a.DataBuffer), self._lz(b.DataBuffer)
            for x, lx, y, ly in ((a, la, b, lb_), (b, lb_, a, la)):  # fp add commutes bit for bit
                if ly is not None and ly.kind == "R" and (lx is None or lx.kind != "R") and x.DataBuffer.Length == y.DataBuffer.Length \
                        and tuple(x.Shape) == (ly.args[1], ly.args[2]) and isinstance(x.DataBuffer, CudaDataBuffer):
                    return self._view(LazyCudaBuffer(self, x.DataBuffer.Length, _Lazy("AR", x.DataBuffer, ly.args[0], ly.args[1], ly.args[2])), *x.Shape)
        # shape of the non-empty operand (empty = additive identity)
        shp = a.Shape if a.DataBuffer.Length else b.Shape
        return self._view(self._map2(MapOp.Add, a.DataBuffer, b.DataBuffer), *shp)

    def Add_M_M_InPlace(self, a, b):
        da, db = self._dev(a.DataBuffer), self._dev(b.DataBuffer)
        if self._is_zero(da):
            # first accumulation into a known-zero adjoint (AD.Float32.fs:4118 after :3746)
            if db.Length == 0 or self._is_zero(db):
                return a
            if db.Length != da.Length:
                raise _lib.InvalidArgError(_lib.DSC_E_SHAPE, "Matrices must have same shape.")
            h = _lib.buf_t()
            check(lib.dsc_buf_share(db.handle, C.byref(h)))  # copy-on-write alias of the addend
            da._h, da._lazy = h.value, None
            da._finalizer = weakref.finalize(da, lib.dsc_buf_free, h.value)
            return a
        self._touch(da)
        check(self._f["add_inplace"](self._h, da.handle, db.handle))
        return a

    def Add_S_M(self, s, a): return self._view(self._scalar_op(0, s, a.DataBuffer), *a.Shape)

    def Add_V_MCols(self, v, a):
        dv, da = self._dev(v), self._dev(a.DataBuffer)
        return self._view(self._out(self._f["add_v_mcols"], a.Rows * a.Cols, a.Rows, a.Cols, dv.handle, da.handle), *a.Shape)

    def Sub_M_M(self, a, b):
        shp = a.Shape if a.DataBuffer.Length else b.Shape
        return self._view(self._map2(MapOp.Sub, a.DataBuffer, b.DataBuffer), *shp)

    def Sub_M_S(self, a, s): return self._view(self._scalar_op(2, s, a.DataBuffer), *a.Shape)
    def Sub_S_M(self, s, a): return self._view(self._scalar_op(1, s, a.DataBuffer), *a.Shape)

    def _gemm(self, a, b, bias):
        m, k, k2, n = a.Rows, a.Cols, b.Rows, b.Cols
        if k != k2:
            raise _lib.InvalidArgError(_lib.DSC_E_SHAPE, "Number of colums of first matrix must match number of rows of second matrix.")
        ta = tb = 0
        ab, bb = a.DataBuffer, b.DataBuffer
        la, lb_ = self._lz(ab), self._lz(bb)
        if la is not None and la.kind == "T":
            ab, ta = la.args[0], 1  # stored k x m
        if lb_ is not None and lb_.kind == "T":
            bb, tb = lb_.args[0], 1  # stored n x k
        da, db = self._dev(ab), self._dev(bb)
        hb = self._dev(bias).handle if bias is not None else 0
        return self._view(self._out(self._f["gemm"], m * n, ta, tb, m, n, k, da.handle, db.handle, hb), m, n)

    def Mul_M_M(self, a, b): return self._gemm(a, b, None)
    def Mul_S_M(self, s, a): return self._view(self._scalar_op(3, s, a.DataBuffer), *a.Shape)

    def Mul_M_M_Add_V_MCols(self, a, b, v): return self._gemm(a, b, v)

    def Add_M_Colwise_V_InPlace(self, a, v):
        da, dv = self._dev(a.DataBuffer), self._dev(v)
        self._touch(dv)
        check(self._f["colsum_inplace"](self._h, a.Rows, a.Cols, da.handle, dv.handle))
        return v

    def Mul_Had_M_M(self, a, b):
        if self.fuse:
            r = self._try_fused_had(a, b)
            if r is not None:
                return r
        return self._view(self._map2(MapOp.Mul, a.DataBuffer, b.DataBuffer), *a.Shape)

    def Inverse_M(self, a):
        d = self._dev(a.DataBuffer)
        try:
            return self._view(self._out(self._f["getri"], a.Rows * a.Rows, a.Rows, d.handle), a.Rows, a.Rows)
        except _lib.SingularMatrixError:
            return None

    def Det_M(self, a):
        out = self._scalar()
        try:
            check(self._f["det"](self._h, a.Rows, self._dev(a.DataBuffer).handle, C.byref(out)))
        except _lib.SingularMatrixError:
            return None
        return self.dtype.type(out.value)

    def Transpose_M(self, a):
        d = self._dev(a.DataBuffer)
        m, n = a.Rows, a.Cols
        if self.fuse and m * n > 0:
            lb = LazyCudaBuffer(self, m * n, _Lazy("T", d, m, n))
            if d._lazy_t is None:
                d._lazy_t = weakref.WeakSet()
            d._lazy_t.add(lb)
            return self._view(lb, n, m)
        return self._view(self._out(self._f["transpose"], m * n, m, n, d.handle), n, m)

    def Permute_M(self, a, dims):
        d = self._dev(a.DataBuffer)
        rank = len(a.Shape)
        shape = (C.c_int64 * rank)(*a.Shape)
        perm = (C.c_int32 * rank)(*[int(x) for x in dims])
        out = self._out(self._f["permute"], _size(a), rank, shape, perm, d.handle)
        return self._view(out, *[a.Shape[int(p)] for p in dims])

    def Reshape_M(self, a, newShape):
        return ShapedDataBufferView(a.DataBuffer, *[int(s) for s in newShape])

    def Map_F_M(self, op, f, a):
        lz = self._lz(a.DataBuffer) if self.fuse else None
        if lz is not None and lz.kind == "AR" and MapOp.Log <= int(op) <= MapOp.Sigmoid:
            # z = A + rows(v) is still pending: produce z and h = f(z) in one pass
            A, v, m, n = lz.args
            z, h = _lib.buf_t(), _lib.buf_t()
            check(self._f["bias_rows_act"](self._h, int(op), m, n, A.handle, v.handle, C.byref(z), C.byref(h)))
            zb = a.DataBuffer
            zb._h, zb._lazy = z.value, None
            zb._finalizer = weakref.finalize(zb, lib.dsc_buf_free, z.value)
            return self._view(CudaDataBuffer(self.ctx, self.dtype, h.value, m * n), *a.Shape)
        if self.fuse and int(op) == MapOp.Cosh:
            # possible head of the tanh adjoint chain (AD.Float32.fs:4238): defer
            d = self._dev(a.DataBuffer)
            return self._view(LazyCudaBuffer(self, d.Length, _Lazy("map", MapOp.Cosh, d)), *a.Shape)
        return self._view(self.Map_F_V(op, f, a.DataBuffer), *a.Shape)

    def Map_F_S_M(self, s, op, f, a):
        lz = self._lz(a.DataBuffer) if self.fuse else None
        if lz is not None and int(op) == MapOp.Div and float(s) == 1.0 and lz.kind == "map" and lz.args[0] == MapOp.Cosh:
            return self._view(LazyCudaBuffer(self, a.DataBuffer.Length, _Lazy("sech", lz.args[1], a.DataBuffer)), *a.Shape)
        return self._view(self.Map_F_S_V(s, op, f, a.DataBuffer), *a.Shape)

    def Map2_F_M_M(self, op, f, a, b): return self._view(self._map2(op, a.DataBuffer, b.DataBuffer), *a.Shape)

    def ReshapeCopy_V_MRows(self, m, v):
        d = self._dev(v)
        out = self._out(lambda ctx, h, o: lib.dsc_buf_copy(h, o), d.Length, d.handle)
        return self._view(out, m, d.Length // m)

    def RepeatReshapeCopy_V_MRows(self, m, v):
        d = self._dev(v)
        if self.fuse and m * d.Length > 0:
            return self._view(LazyCudaBuffer(self, m * d.Length, _Lazy("R", d, m, d.Length)), m, d.Length)
        return self._view(self._out(self._f["repeat_rows"], m * d.Length, m, d.handle), m, d.Length)

    def RepeatReshapeCopy_V_MCols(self, n, v):
        d = self._dev(v)
        return self._view(self._out(self._f["repeat_cols"], n * d.Length, n, d.handle), d.Length, n)

    def CustomOp_DM_Forward(self, a, info):
        # consumer hook (Backend.fs:137): `info` is a callable taking (backend, view)
        return info.forward(self, a)

    def CustomOp_DM_Backward(self, origin, adjoint, primal, info):
        return info.backward(self, origin, adjoint, primal)

    # ---- peephole for the adjoint chains ------------------------------------------------------
    def _fused(self, kind, dA, p, shape):
        da, dp = self._dev(dA), self._dev(p)
        return self._view(self._out(self._f["fused_bwd"], da.Length, kind, da.handle, dp.handle), *shape)

    def _try_fused_had(self, a, b):
        """tanh adjoint: (dA .* secha) .* secha with secha = 1/cosh(aP)  -> DSC_FUSED_TANH_BWD"""
        la, lb_ = self._lz(a.DataBuffer), self._lz(b.DataBuffer)
        if lb_ is None or lb_.kind != "sech":
            return None
        x = lb_.args[0]
        if la is not None and la.kind == "had_sech" and la.args[1] is x:
            return self._fused(0, la.args[0], x, a.Shape)       # second Hadamard: emit the fused kernel
        if la is None:                                          # first Hadamard: defer
            return self._view(LazyCudaBuffer(self, a.DataBuffer.Length, _Lazy("had_sech", a.DataBuffer, x, b.DataBuffer)), *a.Shape)
        return None

    def _eval_chain(self, lz):
        """Unfused evaluation of a deferred node whose consumer is not the fused pattern."""
        h = _lib.buf_t()
        k = lz.kind
        if k == "map":         # ('map', op, x)
            check(self._f["map"](self._h, int(lz.args[0]), lz.args[1].handle, C.byref(h)))
        elif k == "sech":      # ('sech', x, cosh_buffer): 1 / cosh(x)
            check(self._f["map_s"](self._h, int(MapOp.Div), self._scalar(1), lz.args[1].handle, C.byref(h)))
        elif k == "had_sech":  # ('had_sech', dA, x, sech_buffer)
            check(self._f["map2"](self._h, int(MapOp.Mul), self._dev(lz.args[0]).handle, lz.args[2].handle, C.byref(h)))
        else:
            raise RuntimeError(f"unknown deferred node {k}")
        return h
