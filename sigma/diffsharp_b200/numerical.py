"""Numerical differentiation over a ``Backend<'T>`` — mirror of ``DiffSharp.Numerical.Float32`` / ``.Float64``
(reference src/DiffSharp/Numerical.Float32.fs:66-233; the float64 file is the same text over ``float``).

Every operation is finite differences of a consumer function over *buffers* (``f : IDataBuffer -> number`` or
``IDataBuffer -> IDataBuffer``) whose vector arithmetic goes through the backend handed in as the last argument —
``Add_V_V``, ``Sub_V_V``, ``Mul_S_V``, ``Sub_M_M``, ``Mul_S_M``, ``Transpose_M``, ``Diagonal_M``, ``Sum_V`` — so it runs
unchanged on the CUDA backend and gives the AD layer an independent check (tests/test_gpu_numerical.py: AD
against finite differences, the property suite the fork deleted from upstream).  Host arrays reach the backend as foreign
``NativeDataBuffer``s, exactly as in the reference (``standardBasisVal``, Util.fs:74-77).

QUIRK (kept, switchable): the fork perturbs by ``FixedPointEpsilon()`` = ``GlobalConfig.Float32FixedPointEpsilon`` = 0.01
(Numerical.Float32.fs:46, Config.fs:70) but rescales by ``EpsilonRec`` / ``EpsilonRec2`` = 1 / 1e-5 and 0.5 / 1e-5 (:47-48,
Config.fs:73-74): its first derivatives come out 1000x too large and ``diff2`` divides by the step squared (consistent).
``Steps.reference(dtype)`` reproduces that text; ``Steps.consistent(dtype, h)`` uses one step everywhere and is what the
property tests use.
"""
from __future__ import annotations

import numpy as np

from .buffers import NativeDataBuffer, ShapedDataBufferView


class Steps:
    """(step, 1/step-like factor for one-sided differences, 0.5/step-like factor for central differences)"""

    def __init__(self, dtype, step, rec, rec2):
        t = np.dtype(dtype).type
        self.dtype = np.dtype(dtype)
        self.eps, self.rec, self.rec2 = t(step), t(rec), t(rec2)

    @classmethod
    def reference(cls, dtype):
        """the fork's constants: step 0.01, factors 1/1e-5 and 0.5/1e-5 (QUIRK above)"""
        t = np.dtype(dtype).type
        return cls(dtype, t(0.01), t(1) / t(0.00001), t(0.5) / t(0.00001))

    @classmethod
    def consistent(cls, dtype, h=None):
        t = np.dtype(dtype).type
        h = t(h if h is not None else (1e-2 if np.dtype(dtype) == np.float32 else 1e-5))
        return cls(dtype, h, t(1) / h, t(0.5) / h)


def _basis(n, i, v, dtype):
    """``standardBasisVal n i v`` (Util.fs:74-77) as a foreign host buffer"""
    s = np.zeros(n, dtype)
    s[i] = v
    return NativeDataBuffer(s)


# ---- scalar-to-scalar (:66-72, :83-92) ---------------------------------------------------------------------------------
def diff(f, x, st: Steps):
    """``diff`` (:66-67): central difference"""
    return (f(x + st.eps) - f(x - st.eps)) * st.rec2


def diff_(f, x, st: Steps):
    """``diff'`` (:70)"""
    return f(x), diff(f, x, st)


def diff2(f, x, st: Steps):
    """``diff2`` (:83-85)"""
    return (f(x + st.eps) - st.dtype.type(2) * f(x) + f(x - st.eps)) / (st.eps * st.eps)


def diff2_(f, x, st: Steps):
    return f(x), diff2(f, x, st)


def diff2__(f, x, st: Steps):
    return f(x), diff(f, x, st), diff2(f, x, st)


# ---- vector-to-scalar (:73-80, :95-157) ----------------------------------------------------------------------------------
def gradv(f, x, v, backend, st: Steps):
    """``gradv`` (:73-76): directional derivative, central difference along v"""
    veps = backend.Mul_S_V(st.eps, v)
    return (f(backend.Add_V_V(x, veps)) - f(backend.Sub_V_V(x, veps))) * st.rec2


def gradv_(f, x, v, backend, st: Steps):
    return f(x), gradv(f, x, v, backend, st)


def grad_(f, x, backend, st: Steps):
    """``grad'`` (:95-104): one-sided differences, one evaluation of f per coordinate; the results are collected on the host
    and come back through the backend as ``Mul_S_V(1/eps, Sub_V_V(gg, g))``"""
    fx = f(x)
    n = x.Length
    g = NativeDataBuffer(np.full(n, fx, st.dtype))
    gg = NativeDataBuffer(np.array([f(backend.Add_V_V(x, _basis(n, i, st.eps, st.dtype))) for i in range(n)], st.dtype))
    return fx, backend.Mul_S_V(st.rec, backend.Sub_V_V(gg, g))


def grad(f, x, backend, st: Steps):
    return grad_(f, x, backend, st)[1]


def gradhessian_(f, x, backend, st: Steps):
    """``gradhessian'`` (:110-131): row i of the Hessian = (grad f (x + eps e_i) - grad f x) / eps"""
    fx, g = grad_(f, x, backend, st)
    n = x.Length
    h = ShapedDataBufferView(NativeDataBuffer(np.tile(np.asarray(g.SubData, st.dtype), n)), n, n)
    rows = [np.asarray(grad(f, backend.Add_V_V(x, _basis(n, i, st.eps, st.dtype)), backend, st).SubData, st.dtype) for i in range(n)]
    hh = ShapedDataBufferView(NativeDataBuffer(np.concatenate(rows) if rows else np.zeros(0, st.dtype)), n, n)
    return fx, g, backend.Mul_S_M(st.rec, backend.Sub_M_M(hh, h))


def gradhessian(f, x, backend, st: Steps):
    r = gradhessian_(f, x, backend, st)
    return r[1], r[2]


def hessian_(f, x, backend, st: Steps):
    r = gradhessian_(f, x, backend, st)
    return r[0], r[2]


def hessian(f, x, backend, st: Steps):
    return gradhessian_(f, x, backend, st)[2]


def hessianv_(f, x, v, backend, st: Steps):
    """``hessianv'`` (:146-150): central difference of the gradient along v"""
    veps = backend.Mul_S_V(st.eps, v)
    fx, gg1 = grad_(f, backend.Add_V_V(x, veps), backend, st)
    gg2 = grad(f, backend.Sub_V_V(x, veps), backend, st)
    return fx, backend.Mul_S_V(st.rec2, backend.Sub_V_V(gg1, gg2))


def hessianv(f, x, v, backend, st: Steps):
    return hessianv_(f, x, v, backend, st)[1]


def gradhessianv_(f, x, v, backend, st: Steps):
    fx, gv = gradv_(f, x, v, backend, st)
    return fx, gv, hessianv(f, x, v, backend, st)


def gradhessianv(f, x, v, backend, st: Steps):
    r = gradhessianv_(f, x, v, backend, st)
    return r[1], r[2]


def laplacian_(f, x, backend, st: Steps):
    """``laplacian'`` (:166-168): trace of the Hessian through ``Diagonal_M`` and ``Sum_V``"""
    v, h = hessian_(f, x, backend, st)
    return v, backend.Sum_V(backend.Diagonal_M(h))


def laplacian(f, x, backend, st: Steps):
    return laplacian_(f, x, backend, st)[1]


# ---- vector-to-vector (:173-233) ---------------------------------------------------------------------------------------
def jacobianT_(f, x, backend, st: Steps):
    """``jacobianT'`` (:173-193): row i = (f(x + eps e_i) - f x) / eps.  The reference builds BOTH stacked matrices with
    ``x.Length`` rows of ``fx`` (a square Jacobian is assumed, :176-178); the same here."""
    fx = f(x)
    n = x.Length
    fxh = np.asarray(fx.SubData, st.dtype)
    j = ShapedDataBufferView(NativeDataBuffer(np.tile(fxh, n)), n, fxh.size)
    rows = [np.asarray(f(backend.Add_V_V(x, _basis(n, i, st.eps, st.dtype))).SubData, st.dtype) for i in range(n)]
    jj = ShapedDataBufferView(NativeDataBuffer(np.concatenate(rows) if rows else np.zeros(0, st.dtype)), n, fxh.size)
    return fx, backend.Mul_S_M(st.rec, backend.Sub_M_M(jj, j))


def jacobianT(f, x, backend, st: Steps):
    return jacobianT_(f, x, backend, st)[1]


def jacobian_(f, x, backend, st: Steps):
    r, j = jacobianT_(f, x, backend, st)
    return r, backend.Transpose_M(j)


def jacobian(f, x, backend, st: Steps):
    return jacobian_(f, x, backend, st)[1]


def jacobianv(f, x, v, backend, st: Steps):
    """``jacobianv`` (:205-208)"""
    veps = backend.Mul_S_V(st.eps, v)
    return backend.Mul_S_V(st.rec2, backend.Sub_V_V(f(backend.Add_V_V(x, veps)), f(backend.Sub_V_V(x, veps))))


def jacobianv_(f, x, v, backend, st: Steps):
    return f(x), jacobianv(f, x, v, backend, st)


def curl_(f, x, backend, st: Steps):
    """``curl'`` (:214-220)"""
    v, j = jacobianT_(f, x, backend, st)
    if (j.Rows, j.Cols) != (3, 3):
        raise ValueError("Curl is supported only for functions with a three-by-three Jacobian matrix.")
    return v, np.array([j[1, 2] - j[2, 1], j[2, 0] - j[0, 2], j[0, 1] - j[1, 0]], st.dtype)


def curl(f, x, backend, st: Steps):
    return curl_(f, x, backend, st)[1]


def div_(f, x, backend, st: Steps):
    """``div'`` (:226-229)"""
    v, j = jacobianT_(f, x, backend, st)
    if j.Rows != j.Cols:
        raise ValueError("Div is defined only for functions with a square Jacobian matrix.")
    return v, backend.Sum_V(backend.Diagonal_M(j))


def div(f, x, backend, st: Steps):
    return div_(f, x, backend, st)[1]


def curldiv_(f, x, backend, st: Steps):
    v, j = jacobianT_(f, x, backend, st)
    if (j.Rows, j.Cols) != (3, 3):
        raise ValueError("Curldiv is supported only for functions with a three-by-three Jacobian matrix.")
    return v, np.array([j[1, 2] - j[2, 1], j[2, 0] - j[0, 2], j[0, 1] - j[1, 0]], st.dtype), j[0, 0] + j[1, 1] + j[2, 2]


def curldiv(f, x, backend, st: Steps):
    r = curldiv_(f, x, backend, st)
    return r[1], r[2]
