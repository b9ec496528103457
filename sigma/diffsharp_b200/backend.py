"""``Backend<'T>`` for B200: the host-side mirror of the reference plugin interface.

``CudaBackend`` implements the 55 members of ``DiffSharp.Backend.Backend<'T>`` (reference src/DiffSharp/Backend.fs:74-139,
listed in ``interface.Backend``) over libdiffsharp_cuda.so with **one C-ABI call per member** — the Python twin of the F#
P/Invoke shim in INTEGRATION.md / integration/Backend.Cuda.fs.  Nothing is decided here: argument unpacking, one call, wrap
the returned handle.

Closures: ``Map_F_*``/``Map2_F_*`` receive ``(MapOp, closure, ...)`` like the reference; only the op-code is forwarded.  An
op-code the device does not implement raises ``NotSupportedError`` — there is no CPU fallback (north_star).

Deferred evaluation behind the unchanged call sequence lives INSIDE the library (csrc/dsc_lazy.cu): the handle
a call returns may stand for a value that is computed only when a consumer needs it, so that the chains the unchanged AD layer
issues (Transpose_M -> Mul_M_M, RepeatReshapeCopy_V_MRows -> Add_M_M -> Map_F_M, the tanh / sigmoid / ReLU adjoint chains,
zero adjoints, had(E, E).Sum()) run as single kernels with bit-identical results.  It is a property of the context
(``DSC_FLAG_NO_FUSE`` switches it off), not of this shim.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib
from .buffers import ISigmaDiffDataBuffer, NativeDataBuffer, ShapedDataBufferView
from .device import CudaContext, CudaDataBuffer
from .interface import Backend, ConstArray, MapOp  # noqa: F401  (re-exported)


def _size(view: ShapedDataBufferView) -> int:
    n = 1
    for e in view.Shape:
        n *= e
    return n


class CudaBackend(Backend):
    """``Backend<'T>`` over libdiffsharp_cuda.so for 'T = float32 or float64."""

    def __init__(self, dtype=np.float32, ctx: CudaContext | None = None, device: int = 0, flags: int = 0):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("Unsupported type, try float32 or float")
        self.ctx = ctx if ctx is not None else CudaContext(device, flags)
        self._p = "s" if self.dtype == np.float32 else "d"
        self._code = _lib.DSC_F32 if self.dtype == np.float32 else _lib.DSC_F64
        self._scalar = C.c_float if self.dtype == np.float32 else C.c_double
        # Under CUDA-graph capture a reduction cannot hand a host scalar back (that needs a synchronisation): when this is a
        # list, Sum_V / Sum_M / Mul_Dot_V_V leave their scalar in a 1-element device buffer appended to it and return NaN,
        # so that a program which goes on to USE the value (rather than just return it) is visibly poisoned.
        self.capture_scalars = None
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
            b._flush_host_writes()
            return b
        if b.Length == 0:
            return self._empty
        return CudaDataBuffer.from_host(self.ctx, np.asarray(b.SubData, dtype=self.dtype))

    def _out(self, fn, n, *args) -> CudaDataBuffer:
        h = _lib.buf_t()
        check(fn(self._h, *args, C.byref(h)))
        return CudaDataBuffer(self.ctx, self.dtype, h.value, n)

    def _view(self, buf, *shape):
        return ShapedDataBufferView(buf, *shape)

    # ---- Create buffer (Backend.fs:76-79) -----------------------------------------------------
    def CreateDataBuffer(self, a) -> ISigmaDiffDataBuffer:
        if isinstance(a, ConstArray) and a._arr is None:
            # the managed shim recognises the array CreateZeroArray / CreateValueArray returned (weak table) and fills on the
            # device instead of uploading
            if a.value is None:
                return CudaDataBuffer.empty(self.ctx, self.dtype, a.n)
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
            devs = [self._dev(b) for b in bufs]
            self.capture_scalars.append(self._out(self._f[name + "_dev"], 1, *[d.handle for d in devs]))
            return self.dtype.type(np.nan)
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
    def _map2(self, op, a, b):
        da, db = self._dev(a), self._dev(b)
        return self._out(self._f["map2"], max(da.Length, db.Length), int(op), da.handle, db.handle)

    def _scalar_op(self, kind, s, v):
        d = self._dev(v)
        return self._out(self._f["scalar_op"], d.Length, kind, self._scalar(s), d.handle)

    def Add_V_V(self, a, b): return self._map2(MapOp.Add, a, b)
    def Sub_V_V(self, a, b): return self._map2(MapOp.Sub, a, b)

    def Add_V_V_InPlace(self, src, srcOffset, dst, dstOffset, length):
        d = self._dev(dst)
        check(self._f["add_inplace_range"](self._h, self._dev(src).handle, srcOffset, d.handle, dstOffset, length))
        d._invalidate_host()
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
    def Mul_Out_V_V(self, a, b):
        da, db = self._dev(a), self._dev(b)
        return self._view(self._out(self._f["ger"], da.Length * db.Length, da.handle, db.handle), da.Length, db.Length)

    def Add_M_M(self, a, b):
        shp = a.Shape if a.DataBuffer.Length else b.Shape   # shape of the non-empty operand (empty = additive identity)
        return self._view(self._map2(MapOp.Add, a.DataBuffer, b.DataBuffer), *shp)

    def Add_M_M_InPlace(self, a, b):
        da, db = self._dev(a.DataBuffer), self._dev(b.DataBuffer)
        if db.Length and db.Length != da.Length:
            raise _lib.InvalidArgError(_lib.DSC_E_SHAPE, "Matrices must have same shape.")
        check(self._f["add_inplace"](self._h, da.handle, db.handle))
        da._invalidate_host()
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
        da, db = self._dev(a.DataBuffer), self._dev(b.DataBuffer)
        hb = self._dev(bias).handle if bias is not None else 0
        return self._view(self._out(self._f["gemm"], m * n, 0, 0, m, n, k, da.handle, db.handle, hb), m, n)

    def Mul_M_M(self, a, b): return self._gemm(a, b, None)
    def Mul_S_M(self, s, a): return self._view(self._scalar_op(3, s, a.DataBuffer), *a.Shape)
    def Mul_M_M_Add_V_MCols(self, a, b, v): return self._gemm(a, b, v)

    def Add_M_Colwise_V_InPlace(self, a, v):
        da, dv = self._dev(a.DataBuffer), self._dev(v)
        check(self._f["colsum_inplace"](self._h, a.Rows, a.Cols, da.handle, dv.handle))
        dv._invalidate_host()
        return v

    def Mul_Had_M_M(self, a, b):
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

    def Map_F_M(self, op, f, a): return self._view(self.Map_F_V(op, f, a.DataBuffer), *a.Shape)
    def Map_F_S_M(self, s, op, f, a): return self._view(self.Map_F_S_V(s, op, f, a.DataBuffer), *a.Shape)
    def Map2_F_M_M(self, op, f, a, b): return self._view(self._map2(op, a.DataBuffer, b.DataBuffer), *a.Shape)

    def ReshapeCopy_V_MRows(self, m, v):
        d = self._dev(v)
        out = self._out(lambda ctx, h, o: lib.dsc_buf_copy(h, o), d.Length, d.handle)
        return self._view(out, m, d.Length // m)

    def RepeatReshapeCopy_V_MRows(self, m, v):
        d = self._dev(v)
        return self._view(self._out(self._f["repeat_rows"], m * d.Length, m, d.handle), m, d.Length)

    def RepeatReshapeCopy_V_MCols(self, n, v):
        d = self._dev(v)
        return self._view(self._out(self._f["repeat_cols"], n * d.Length, n, d.handle), d.Length, n)

    def _fused(self, kind, dA, p, shape):
        """dsc_?fused_bwd directly (tools / micro-benchmarks): out = chain_kind(dA, p)"""
        da, dp = self._dev(dA), self._dev(p)
        return self._view(self._out(self._f["fused_bwd"], da.Length, kind, da.handle, dp.handle), *shape)

    def CustomOp_DM_Forward(self, a, info):
        # consumer hook (Backend.fs:137): the opaque payload is the consumer's own object; it is handed the backend and the view
        return info.forward(self, a)

    def CustomOp_DM_Backward(self, origin, adjoint, primal, info):
        return info.backward(self, origin, adjoint, primal)
