"""TEST INFRASTRUCTURE — the CPU oracle as a ``Backend<'T>`` implementation.

``OracleBackend`` implements the same 55-member interface as the product's ``CudaBackend`` but on
host ``NativeDataBuffer``s, by calling the C restatement in ``oracle/oracle_backend.c`` (genuine
OpenBLAS for the BLAS-backed members, sequential loops and double-precision libm elsewhere).
It plays the role of the reference deployment's OpenBLAS backend, which is not in the reference tree
(SURVEY §0.2): PARITY UNPINNED — see the header of oracle_backend.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  Nothing under sigma/ does.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import sysconfig

import numpy as np

from sigma.diffsharp_b200.interface import Backend, ConstArray, MapOp  # interface types only: no device code is imported
from sigma.diffsharp_b200.buffers import ISigmaDiffDataBuffer, NativeDataBuffer, ShapedDataBufferView

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_backend.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("oracle_backend.c", "oracle_backend_impl.h")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_backend.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def find_openblas() -> tuple[str, str]:
    """(path, symbol prefix) of a wheel-bundled OpenBLAS: scipy's LP64 build ('scipy_' prefix)."""
    site = sysconfig.get_paths()["purelib"]
    hits = sorted(glob.glob(os.path.join(site, "scipy.libs", "libscipy_openblas-*.so")))
    if hits:
        return hits[0], "scipy_"
    hits = sorted(glob.glob(os.path.join(site, "opencv_python_headless.libs", "libopenblas*.so")))
    if hits:
        return hits[0], ""
    raise ImportError("no bundled OpenBLAS found (looked in scipy.libs and opencv_python_headless.libs)")


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        path, prefix = find_openblas()
        l.oracle_init.argtypes = [C.c_char_p, C.c_char_p]
        if l.oracle_init(path.encode(), prefix.encode()) != 0:
            raise ImportError(f"oracle_init failed for {path}")
        l.oracle_openblas_config.restype = C.c_char_p
        for p, T in (("s", C.c_float), ("d", C.c_double)):
            for name in ("dot", "asum", "nrm2", "amax", "sum"):
                getattr(l, f"oracle_{p}{name}").restype = T
            getattr(l, f"oracle_{p}map_s").argtypes = [C.c_int, T, C.c_int, C.c_void_p, C.c_void_p]
            getattr(l, f"oracle_{p}scalar_op").argtypes = [C.c_int, T, C.c_int, C.c_void_p, C.c_void_p]
        _lib = l
        l.openblas_path = path
    return _lib


def set_num_threads(n: int):
    lib().oracle_set_num_threads(int(n))


def get_num_threads() -> int:
    return int(lib().oracle_get_num_threads())


def openblas_info() -> str:
    l = lib()
    return f"{os.path.basename(l.openblas_path)}: {l.oracle_openblas_config().decode()}"


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class OracleBackend(Backend):
    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self._p = "s" if self.dtype == np.float32 else "d"
        self._T = C.c_float if self.dtype == np.float32 else C.c_double
        self._l = lib()

    def _f(self, name):
        return getattr(self._l, f"oracle_{self._p}{name}")

    # host window of any buffer (foreign buffers included)
    def _w(self, b: ISigmaDiffDataBuffer) -> np.ndarray:
        if isinstance(b, NativeDataBuffer):
            w = b.window()
        else:
            w = b.SubData
        if w.dtype != self.dtype:
            w = w.astype(self.dtype)
        return np.ascontiguousarray(w)

    def _new(self, n) -> np.ndarray:
        return np.empty(int(n), self.dtype)

    @staticmethod
    def _buf(a: np.ndarray) -> NativeDataBuffer:
        return NativeDataBuffer(a)

    def _check(self, status):
        if status == -6:
            raise NotImplementedError("op-code not supported")
        if status != 0:
            raise RuntimeError(f"oracle status {status}")

    # ---- create (Backend.fs:76-79) --------------------------------------------------------------
    def CreateDataBuffer(self, a):
        if isinstance(a, ConstArray):
            a = a.materialize()
        return NativeDataBuffer(np.array(a, dtype=self.dtype, copy=True))

    def CreateUninitialisedArray(self, n): return np.empty(n, self.dtype)
    def CreateZeroArray(self, n): return np.zeros(n, self.dtype)
    def CreateValueArray(self, n, v): return np.full(n, v, self.dtype)

    # ---- scalar valued --------------------------------------------------------------------------
    def Mul_Dot_V_V(self, a, b):
        x, y = self._w(a), self._w(b)
        if x.size != y.size:
            raise ValueError("Vectors must have same length.")
        return self.dtype.type(self._f("dot")(x.size, _ptr(x), _ptr(y)))

    def L1Norm_V(self, a): x = self._w(a); return self.dtype.type(self._f("asum")(x.size, _ptr(x)))
    def L2Norm_V(self, a): x = self._w(a); return self.dtype.type(self._f("nrm2")(x.size, _ptr(x)))
    def SupNorm_V(self, a): x = self._w(a); return self.dtype.type(self._f("amax")(x.size, _ptr(x)))
    def Sum_V(self, a): x = self._w(a); return self.dtype.type(self._f("sum")(x.size, _ptr(x)))
    Sum_M = Sum_V
    def MaxIndex_V(self, a): x = self._w(a); return int(self._f("argmax")(x.size, _ptr(x)))
    def MinIndex_V(self, a): x = self._w(a); return int(self._f("argmin")(x.size, _ptr(x)))

    # ---- elementwise ----------------------------------------------------------------------------
    def _map2(self, op, a, b):
        x, y = self._w(a), self._w(b)
        if int(op) in (MapOp.Add, MapOp.Sub) and x.size != y.size and (x.size == 0 or y.size == 0):
            # empty operand = additive identity (AD.Float32.fs:3919,4257)
            if x.size:
                return self._buf(x.copy())
            return self._buf(y.copy() if int(op) == MapOp.Add else -y)
        if x.size != y.size:
            raise ValueError("Vectors must have same length.")
        out = self._new(x.size)
        self._check(self._f("map2")(int(op), x.size, _ptr(x), _ptr(y), _ptr(out)))
        return self._buf(out)

    def _scalar_op(self, kind, s, v):
        x = self._w(v)
        out = self._new(x.size)
        self._check(self._f("scalar_op")(kind, self._T(s), x.size, _ptr(x), _ptr(out)))
        return self._buf(out)

    def Add_V_V(self, a, b): return self._map2(MapOp.Add, a, b)
    def Sub_V_V(self, a, b): return self._map2(MapOp.Sub, a, b)

    def Add_V_V_InPlace(self, src, srcOffset, dst, dstOffset, length):
        s = self._w(src)
        if not isinstance(dst, NativeDataBuffer):
            raise TypeError("oracle in-place ops need a NativeDataBuffer destination")
        d = dst.window()
        self._f("add_inplace_range")(_ptr(s), srcOffset, _ptr(d), dstOffset, length)
        return dst

    def Add_S_V(self, s, v): return self._scalar_op(0, s, v)
    def Sub_S_V(self, s, v): return self._scalar_op(1, s, v)
    def Sub_V_S(self, v, s): return self._scalar_op(2, s, v)
    def Mul_S_V(self, s, v): return self._scalar_op(3, s, v)

    def Map_F_V(self, op, f, v):
        x = self._w(v)
        out = self._new(x.size)
        self._check(self._f("map")(int(op), x.size, _ptr(x), _ptr(out)))
        return self._buf(out)

    def Map_F_S_V(self, s, op, f, v):
        x = self._w(v)
        out = self._new(x.size)
        self._check(self._f("map_s")(int(op), self._T(s), x.size, _ptr(x), _ptr(out)))
        return self._buf(out)

    def Map2_F_V_V(self, op, f, a, b): return self._map2(op, a, b)

    def fused_bwd(self, kind, dA, p):
        x, y = self._w(dA), self._w(p)
        out = self._new(x.size)
        self._check(self._f("fused_bwd")(kind, x.size, _ptr(x), _ptr(y), _ptr(out)))
        return self._buf(out)

    # ---- BLAS-2 ---------------------------------------------------------------------------------
    def _mat(self, A: ShapedDataBufferView) -> np.ndarray:
        n = 1
        for e in A.Shape:
            n *= e
        return self._w(A.DataBuffer)[:n]

    def Mul_M_V(self, A, x, y=None):
        a, xv = self._mat(A), self._w(x)
        if xv.size != A.Cols:
            raise ValueError("Length of vector must match number of columns of matrix.")
        yv = self._w(y) if y is not None else None
        out = self._new(A.Rows)
        self._f("gemv")(0, A.Rows, A.Cols, _ptr(a), _ptr(xv), _ptr(yv) if yv is not None else None, _ptr(out))
        return self._buf(out)

    def Mul_M_V_Add_V(self, A, x, y): return self.Mul_M_V(A, x, y)

    def Mul_V_M(self, x, A):
        a, xv = self._mat(A), self._w(x)
        if xv.size != A.Rows:
            raise ValueError("Length of vector must match number of rows of matrix.")
        out = self._new(A.Cols)
        self._f("gemv")(1, A.Rows, A.Cols, _ptr(a), _ptr(xv), None, _ptr(out))
        return self._buf(out)

    def Mul_Out_V_V(self, a, b):
        x, y = self._w(a), self._w(b)
        out = self._new(x.size * y.size)
        self._f("ger")(x.size, y.size, _ptr(x), _ptr(y), _ptr(out))
        return ShapedDataBufferView(self._buf(out), x.size, y.size)

    # ---- LAPACK class ---------------------------------------------------------------------------
    def Solve_M_V(self, A, b):
        a, bv = self._mat(A), self._w(b)
        out = self._new(A.Rows)
        st = self._f("gesv")(A.Rows, _ptr(a), _ptr(bv), _ptr(out))
        return None if st != 0 else self._buf(out)

    def SolveSymmetric_M_V(self, A, b):
        a, bv = self._mat(A), self._w(b)
        out = self._new(A.Rows)
        st = self._f("sysv")(A.Rows, _ptr(a), _ptr(bv), _ptr(out))
        return None if st != 0 else self._buf(out)

    def Inverse_M(self, A):
        a = self._mat(A)
        out = self._new(A.Rows * A.Rows)
        st = self._f("getri")(A.Rows, _ptr(a), _ptr(out))
        return None if st != 0 else ShapedDataBufferView(self._buf(out), A.Rows, A.Rows)

    def Det_M(self, A):
        a = self._mat(A)
        out = self._T()
        st = self._f("det")(A.Rows, _ptr(a), C.byref(out))
        return None if st != 0 else self.dtype.type(out.value)

    def Diagonal_M(self, A):
        a = self._mat(A)
        out = self._new(min(A.Rows, A.Cols))
        self._f("diag")(A.Rows, A.Cols, _ptr(a), _ptr(out))
        return self._buf(out)

    def ReshapeCopy_MRows_V(self, A): return self._buf(self._mat(A).copy())

    # ---- matrix valued --------------------------------------------------------------------------
    def _v(self, buf, *shape): return ShapedDataBufferView(buf, *shape)

    def Add_M_M(self, a, b):
        shp = a.Shape if a.DataBuffer.Length else b.Shape
        return self._v(self._map2(MapOp.Add, a.DataBuffer, b.DataBuffer), *shp)

    def Add_M_M_InPlace(self, a, b):
        if b.DataBuffer.Length == 0:
            return a
        if not isinstance(a.DataBuffer, NativeDataBuffer):
            raise TypeError("oracle in-place ops need a NativeDataBuffer destination")
        d, s = a.DataBuffer.window(), self._w(b.DataBuffer)
        if d.size != s.size:
            raise ValueError("Matrices must have same shape.")
        self._f("add_inplace_range")(_ptr(s), 0, _ptr(d), 0, d.size)
        return a

    def Add_S_M(self, s, a): return self._v(self._scalar_op(0, s, a.DataBuffer), *a.Shape)

    def Add_V_MCols(self, v, a):
        m, vv = self._mat(a), self._w(v)
        out = self._new(m.size)
        self._f("add_v_mcols")(a.Rows, a.Cols, _ptr(vv), _ptr(m), _ptr(out))
        return self._v(self._buf(out), *a.Shape)

    def Sub_M_M(self, a, b):
        shp = a.Shape if a.DataBuffer.Length else b.Shape
        return self._v(self._map2(MapOp.Sub, a.DataBuffer, b.DataBuffer), *shp)

    def Sub_M_S(self, a, s): return self._v(self._scalar_op(2, s, a.DataBuffer), *a.Shape)
    def Sub_S_M(self, s, a): return self._v(self._scalar_op(1, s, a.DataBuffer), *a.Shape)

    def gemm(self, ta, tb, m, n, k, A, B, bias=None):
        """raw row-major GEMM with transposed-operand flags (mirrors dsc_?gemm)"""
        out = self._new(m * n)
        self._f("gemm")(ta, tb, m, n, k, _ptr(A), _ptr(B), _ptr(bias) if bias is not None else None, _ptr(out))
        return out

    def Mul_M_M(self, a, b, v=None):
        if a.Cols != b.Rows:
            raise ValueError("Number of colums of first matrix must match number of rows of second matrix.")
        bias = self._w(v) if v is not None else None
        out = self.gemm(0, 0, a.Rows, b.Cols, a.Cols, self._mat(a), self._mat(b), bias)
        return self._v(self._buf(out), a.Rows, b.Cols)

    def Mul_S_M(self, s, a): return self._v(self._scalar_op(3, s, a.DataBuffer), *a.Shape)
    def Mul_M_M_Add_V_MCols(self, a, b, v): return self.Mul_M_M(a, b, v)

    def Add_M_Colwise_V_InPlace(self, a, v):
        if not isinstance(v, NativeDataBuffer):
            raise TypeError("oracle in-place ops need a NativeDataBuffer destination")
        m, d = self._mat(a), v.window()
        if d.size != a.Cols:
            raise ValueError("Length of vector must match number of columns of matrix.")
        self._f("colsum_inplace")(a.Rows, a.Cols, _ptr(m), _ptr(d))
        return v

    def Mul_Had_M_M(self, a, b): return self._v(self._map2(MapOp.Mul, a.DataBuffer, b.DataBuffer), *a.Shape)

    def Transpose_M(self, a):
        m = self._mat(a)
        out = self._new(m.size)
        self._f("transpose")(a.Rows, a.Cols, _ptr(m), _ptr(out))
        return self._v(self._buf(out), a.Cols, a.Rows)

    def Permute_M(self, a, dims):
        m = self._mat(a)
        rank = len(a.Shape)
        out = self._new(m.size)
        shape = (C.c_longlong * rank)(*a.Shape)
        perm = (C.c_int * rank)(*[int(d) for d in dims])
        self._f("permute")(rank, shape, perm, _ptr(m), _ptr(out))
        return self._v(self._buf(out), *[a.Shape[int(d)] for d in dims])

    def Reshape_M(self, a, newShape): return ShapedDataBufferView(a.DataBuffer, *[int(s) for s in newShape])
    def Map_F_M(self, op, f, a): return self._v(self.Map_F_V(op, f, a.DataBuffer), *a.Shape)
    def Map_F_S_M(self, s, op, f, a): return self._v(self.Map_F_S_V(s, op, f, a.DataBuffer), *a.Shape)
    def Map2_F_M_M(self, op, f, a, b): return self._v(self._map2(op, a.DataBuffer, b.DataBuffer), *a.Shape)

    def ReshapeCopy_V_MRows(self, m, v):
        x = self._w(v)
        return self._v(self._buf(x.copy()), m, x.size // m)

    def RepeatReshapeCopy_V_MRows(self, m, v):
        x = self._w(v)
        out = self._new(m * x.size)
        self._f("repeat_rows")(m, x.size, _ptr(x), _ptr(out))
        return self._v(self._buf(out), m, x.size)

    def RepeatReshapeCopy_V_MCols(self, n, v):
        x = self._w(v)
        out = self._new(n * x.size)
        self._f("repeat_cols")(x.size, n, _ptr(x), _ptr(out))
        return self._v(self._buf(out), x.size, n)

    def CustomOp_DM_Forward(self, a, info): return info.forward(self, a)
    def CustomOp_DM_Backward(self, origin, adjoint, primal, info): return info.backward(self, origin, adjoint, primal)
