"""Host-side mirror of the reference's AD layer — the CALLER of the Backend<'T> hot path.

In a deployment this layer is the unchanged F# of src/DiffSharp/AD.Float32.fs / AD.Float64.fs; no
.NET toolchain exists in this image, so this module restates it, backend-agnostic, to issue the
*same sequence of Backend<'T> calls* (same ops, same order, same zero-adjoint allocations) against
whatever provider is installed in ``GlobalConfig.BackendProvider`` — the CUDA backend in the product
path, the CPU oracle backend in tests.  Citations are to AD.Float32.fs unless noted.

  DNumber / DVector / DNDArray      :77 / :545 / :1678   primal | forward dual | reverse tape node
  Op_* templates                    :185-236, :757-1081, :1892-2354  (one generic template here)
  operators                         :238-543, :1083-1598, :2356-2938
  TraceOp                           :3009-3250
  reverseReset / reversePush        :3572-3845 / :3848-4287  (same work-list discipline: depth
                                    first, children consed on the front, node fires when its
                                    fan-out counter returns to 0)
  DiffOps                           :4296-4379 (+ the commented-out hessianv :4416-4427)

Reference quirks are kept where they change numbers or call sequences and are marked QUIRK.
"""
from __future__ import annotations

import math

import numpy as np

from .interface import MapOp
from .buffers import NativeDataBuffer, ShapedDataBufferView
from .config import GlobalConfig

PRIMAL, FORWARD, REVERSE = 0, 1, 2


def Backend(a):
    """``let inline Backend a = GlobalConfig.BackendProvider.GetBackend(a).BackendHandle`` (:70)"""
    return GlobalConfig.BackendProvider.GetBackend(a).BackendHandle


class GlobalTagger:
    """Util.fs:218-229 (a bare mutable counter; not thread-safe in the reference either)."""
    LastTag = 0

    @classmethod
    def Next(cls):
        cls.LastTag += 1
        return cls.LastTag

    @classmethod
    def Reset(cls):
        cls.LastTag = 0


class TraceOp:
    __slots__ = ("name", "args")

    def __init__(self, name, *args):
        self.name, self.args = name, args

    def __repr__(self):
        return f"TraceOp.{self.name}"


Noop = TraceOp("Noop")


class _AD:
    """Shared shape of the three unions: kind, payload/primal ``p``, tangent ``t``, adjoint ``a``
    (mutable ref), parent op, fan-out ``f`` (mutable ref), tag."""
    __slots__ = ("kind", "p", "t", "a", "op", "f", "tag")
    __array_ufunc__ = None  # NumPy scalars on the left defer to our reflected operators

    def _init(self, kind, p, t=None, a=None, op=None, tag=0):
        self.kind, self.p, self.t, self.a, self.op, self.f, self.tag = kind, p, t, a, op, 0, tag

    @property
    def P(self):
        return self if self.kind == PRIMAL else self.p

    @property
    def PD(self):
        x = self
        while x.kind != PRIMAL:
            x = x.p
        return x

    @property
    def F(self):
        if self.kind != REVERSE:
            raise RuntimeError("Cannot get fan-out value of a non-reverse value.")
        return self.f

    @F.setter
    def F(self, v):
        if self.kind != REVERSE:
            raise RuntimeError("Cannot set fan-out value of a non-reverse value.")
        self.f = v


# ================================================================================================
# DNumber (:77-543)
# ================================================================================================
class DNumber(_AD):
    __slots__ = ()

    def __init__(self, value, dtype=None):
        """``D of number`` — primal.  `value` is a Python/NumPy scalar; dtype fixes 'T."""
        if dtype is None:
            dtype = value.dtype if hasattr(value, "dtype") else np.float64
        self._init(PRIMAL, np.dtype(dtype).type(value))

    @staticmethod
    def DF(p, t, tag):
        d = DNumber.__new__(DNumber)
        d._init(FORWARD, p, t=t, tag=tag)
        return d

    @staticmethod
    def DR(p, op, tag):
        d = DNumber.__new__(DNumber)
        d._init(REVERSE, p, a=DNumber(0, p.dtype), op=op, tag=tag)
        return d

    # factory protocol used by the op template
    primal = staticmethod(lambda v: DNumber(v))
    forward = staticmethod(lambda cp, t, tag: DNumber.DF(cp, t, tag))
    reverse = staticmethod(lambda cp, op, tag: DNumber.DR(cp, op, tag))

    @property
    def dtype(self):
        return self.PD.p.dtype

    @property
    def Value(self):
        return self.PD.p

    def __float__(self):
        return float(self.PD.p)

    @property
    def T(self):
        if self.kind == PRIMAL:
            return DNumber(0, self.dtype)
        if self.kind == FORWARD:
            return self.t
        raise RuntimeError("Cannot get tangent value of DR.")

    @property
    def A(self):
        if self.kind == PRIMAL:
            return DNumber(0, self.dtype)
        if self.kind == FORWARD:
            raise RuntimeError("Cannot get adjoint value of DF.")
        return self.a

    @A.setter
    def A(self, v):
        if self.kind == FORWARD:
            raise RuntimeError("Cannot set adjoint value of DF.")
        if self.kind == REVERSE:
            self.a = v

    def GetForward(self, t, i):
        return DNumber.DF(self, t, i)

    def GetReverse(self, i):
        return DNumber.DR(self, Noop, i)

    def __repr__(self):
        return f"{('D', 'DF', 'DR')[self.kind]} {float(self): e}"

    # comparisons compare primal values (:162-172)
    def __eq__(self, o):
        return float(self) == float(o)

    def __hash__(self):
        return hash(float(self))

    def __lt__(self, o): return float(self) < float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __ge__(self, o): return float(self) >= float(o)

    # ---- binary (:238-330) ------------------------------------------------------------------
    def __add__(a, b):
        b = _coerce(b, a)
        if not isinstance(b, DNumber):
            return NotImplemented
        return _op2(a, b, lambda x, y: x + y, lambda x, y: x + y, lambda cp, ap, at: at, lambda cp, bp, bt: bt,
                    lambda cp, ap, at, bp, bt: at + bt,
                    lambda x, y: TraceOp("Add_D_D", x, y), lambda x, y: TraceOp("Add_D_DCons", x),
                    lambda x, y: TraceOp("Add_D_DCons", y), DNumber)

    def __radd__(b, a):
        return _coerce(a, b) + b

    def __sub__(a, b):
        b = _coerce(b, a)
        if not isinstance(b, DNumber):
            return NotImplemented
        return _op2(a, b, lambda x, y: x - y, lambda x, y: x - y, lambda cp, ap, at: at, lambda cp, bp, bt: -bt,
                    lambda cp, ap, at, bp, bt: at - bt,
                    lambda x, y: TraceOp("Sub_D_D", x, y), lambda x, y: TraceOp("Sub_D_DCons", x),
                    lambda x, y: TraceOp("Sub_DCons_D", y), DNumber)

    def __rsub__(b, a):
        return _coerce(a, b) - b

    def __mul__(a, b):
        b = _coerce(b, a)
        if not isinstance(b, DNumber):
            return NotImplemented
        return _op2(a, b, lambda x, y: x * y, lambda x, y: x * y, lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                    lambda cp, ap, at, bp, bt: at * bp + ap * bt,
                    lambda x, y: TraceOp("Mul_D_D", x, y), lambda x, y: TraceOp("Mul_D_DCons", x, y),
                    lambda x, y: TraceOp("Mul_D_DCons", y, x), DNumber)

    def __rmul__(b, a):
        return _coerce(a, b) * b

    def __truediv__(a, b):
        b = _coerce(b, a)
        if not isinstance(b, DNumber):
            return NotImplemented
        return _op2(a, b, lambda x, y: x / y, lambda x, y: x / y, lambda cp, ap, at: at / b,
                    lambda cp, bp, bt: -bt * cp / bp, lambda cp, ap, at, bp, bt: (at - bt * cp) / bp,
                    lambda x, y: TraceOp("Div_D_D", x, y), lambda x, y: TraceOp("Div_D_DCons", x, y),
                    lambda x, y: TraceOp("Div_DCons_D", x, y), DNumber)

    def __rtruediv__(b, a):
        return _coerce(a, b) / b

    def __pow__(a, b):
        b = _coerce(b, a)
        if not isinstance(b, DNumber):
            return NotImplemented
        one = DNumber(1, a.dtype)
        return _op2(a, b, _npow, lambda x, y: x ** y, lambda cp, ap, at: at * (ap ** (b - one)) * b,
                    lambda cp, bp, bt: bt * cp * log(a),
                    lambda cp, ap, at, bp, bt: (ap ** (bp - one)) * (at * bp + ap * bt * log(ap)),
                    lambda x, y: TraceOp("Pow_D_D", x, y), lambda x, y: TraceOp("Pow_D_DCons", x, y),
                    lambda x, y: TraceOp("Pow_DCons_D", x, y), DNumber)

    def __rpow__(b, a):
        return _coerce(a, b) ** b

    def __neg__(a):
        return _op1(a, lambda x: -x, lambda x: -x, lambda cp, ap, at: -at, lambda x: TraceOp("Neg_D", x), DNumber)

    def __abs__(a):
        return Abs(a)

    @staticmethod
    def FixedPoint(g, a0, b):
        """:476-528 — fixed point of a = g a b, differentiable in b (Christianson 1994)."""
        imax = GlobalConfig.FixedPointMaxIterations
        eps = DNumber(GlobalConfig.FixedPointEpsilon(b.dtype), b.dtype)
        a, i = a0, 0
        if b.kind == PRIMAL:
            while i < imax:
                i += 1
                if i < imax:
                    aa = g(a, b)
                    if abs(aa - a) <= eps:
                        i = imax
                    a = aa
            return DNumber(a.Value, b.dtype)
        if b.kind == FORWARD:
            while i < imax:
                i += 1
                if i < imax:
                    aa = g(a, b)
                    if abs(aa.P - a.P) <= eps and abs(aa.T - a.T) <= eps:
                        i = imax
                    a = aa
            return DNumber.DF(a.P, a.T, b.tag)
        bfirst = DNumber.DR(b.p, Noop, b.tag)  # cut the connection between b and bfirst
        while i < imax:
            i += 1
            if i < imax:
                aa = g(a, bfirst)
                if abs(aa - a) <= eps:
                    i = imax
                a = aa
        aprev = DNumber.DR(a.P, Noop, b.tag)
        alast = g(aprev, bfirst)
        return DNumber.DR(a.P, TraceOp("FixedPoint_D", b, bfirst, aprev, alast), b.tag)


D = DNumber


def _npow(x, y):
    """F# ``x ** y`` on 'T: Math.Pow in double, rounded to 'T."""
    with np.errstate(all="ignore"):
        return x.dtype.type(np.power(np.float64(x), np.float64(y)))


def _coerce(x, like):
    """number/int -> D of the other operand's precision (the `DNumber - number` overloads, :316-330)."""
    if isinstance(x, _AD):
        return x
    if isinstance(x, (int, float, np.floating, np.integer)):
        return DNumber(x, like.dtype)
    return x


def _np1(fn):
    """System.Math semantics for a scalar: evaluate in double, round to 'T."""
    def f(x):
        with np.errstate(all="ignore"):
            return x.dtype.type(fn(np.float64(x)))
    return f


def _signummod(x):
    return x.dtype.type(-1 if x < 0 else (1 if x > 0 else 0))


# ================================================================================================
# generic op templates (Op_D_D :185, Op_D_D_D :193, Op_DV_* :757-1081, Op_DM_* :1892-2354)
# ================================================================================================
def _op1(a, ff, fd, df, r, mk):
    k = a.kind
    if k == PRIMAL:
        return mk.primal(ff(a.p))
    if k == FORWARD:
        cp = fd(a.p)
        return mk.forward(cp, df(cp, a.p, a.t), a.tag)
    cp = fd(a.p)
    return mk.reverse(cp, r(a), a.tag)


_SAME_LEVEL = "Forward and reverse AD cannot run on the same level."


def _op2(a, b, ff, fd, df_da, df_db, df_dab, r_d_d, r_d_c, r_c_d, mk):
    ka, kb = a.kind, b.kind
    if ka == PRIMAL:
        if kb == PRIMAL:
            return mk.primal(ff(a.p, b.p))
        cp = fd(a, b.p)
        if kb == FORWARD:
            return mk.forward(cp, df_db(cp, b.p, b.t), b.tag)
        return mk.reverse(cp, r_c_d(a, b), b.tag)
    if ka == FORWARD:
        if kb == PRIMAL:
            cp = fd(a.p, b)
            return mk.forward(cp, df_da(cp, a.p, a.t), a.tag)
        if kb == FORWARD:
            if a.tag == b.tag:
                cp = fd(a.p, b.p)
                return mk.forward(cp, df_dab(cp, a.p, a.t, b.p, b.t), a.tag)
            if a.tag < b.tag:
                cp = fd(a, b.p)
                return mk.forward(cp, df_db(cp, b.p, b.t), b.tag)
            cp = fd(a.p, b)
            return mk.forward(cp, df_da(cp, a.p, a.t), a.tag)
        if a.tag < b.tag:
            cp = fd(a, b.p)
            return mk.reverse(cp, r_c_d(a, b), b.tag)
        if a.tag > b.tag:
            cp = fd(a.p, b)
            return mk.forward(cp, df_da(cp, a.p, a.t), a.tag)
        raise RuntimeError(_SAME_LEVEL)
    # a is REVERSE
    if kb == PRIMAL:
        cp = fd(a.p, b)
        return mk.reverse(cp, r_d_c(a, b), a.tag)
    if kb == FORWARD:
        if a.tag < b.tag:
            cp = fd(a, b.p)
            return mk.forward(cp, df_db(cp, b.p, b.t), b.tag)
        if a.tag > b.tag:
            cp = fd(a.p, b)
            return mk.reverse(cp, r_d_c(a, b), a.tag)
        raise RuntimeError(_SAME_LEVEL)
    if a.tag == b.tag:
        cp = fd(a.p, b.p)
        return mk.reverse(cp, r_d_d(a, b), a.tag)
    if a.tag < b.tag:
        cp = fd(a, b.p)
        return mk.reverse(cp, r_c_d(a, b), b.tag)
    cp = fd(a.p, b)
    return mk.reverse(cp, r_d_c(a, b), a.tag)


# ================================================================================================
# DVector (:545-1675)
# ================================================================================================
class DVector(_AD):
    __slots__ = ()

    def __init__(self, buffer):
        """``DV of IDataBuffer`` — primal."""
        self._init(PRIMAL, buffer)

    @staticmethod
    def DVF(p, t, tag):
        d = DVector.__new__(DVector)
        d._init(FORWARD, p, t=t, tag=tag)
        return d

    @staticmethod
    def DVR(p, op, tag, adjoint=None):
        d = DVector.__new__(DVector)
        # every DVR starts life with a HOST zero adjoint (DVector.ZeroN, :707,765,...); reverseReset
        # replaces it with a backend buffer (:3640-3641)
        d._init(REVERSE, p, a=adjoint if adjoint is not None else DVector.ZeroN(p.Length, p.dtype), op=op, tag=tag)
        return d

    primal = staticmethod(lambda buf: DVector(buf))
    forward = staticmethod(lambda cp, t, tag: DVector.DVF(cp, t, tag))
    reverse = staticmethod(lambda cp, op, tag: DVector.DVR(cp, op, tag))

    @property
    def dtype(self):
        return self.PD.p.dtype

    @property
    def Buffer(self):
        return self.PD.p

    @property
    def Length(self):
        return self.PD.p.Length

    @property
    def T(self):
        if self.kind == PRIMAL:
            return DVector.ZeroN(self.Length, self.dtype)
        if self.kind == FORWARD:
            return self.t
        raise RuntimeError("Cannot get tangent value of DVR.")

    @property
    def A(self):
        if self.kind == PRIMAL:
            return DVector.ZeroN(self.Length, self.dtype)
        if self.kind == FORWARD:
            raise RuntimeError("Cannot get adjoint value of DVF.")
        return self.a

    @A.setter
    def A(self, v):
        if self.kind == FORWARD:
            raise RuntimeError("Cannot set adjoint value of DVF.")
        if self.kind == REVERSE:
            self.a = v

    def GetForward(self, t, i):
        return DVector.DVF(self, t, i)

    def GetReverse(self, i):
        return DVector.DVR(self, Noop, i)

    def DeepCopy(self):
        if self.kind == PRIMAL:
            return DVector(self.p.DeepCopy())
        if self.kind == FORWARD:
            return DVector.DVF(self.p.DeepCopy(), self.t.DeepCopy(), self.tag)
        d = DVector.DVR(self.p.DeepCopy(), self.op, self.tag, adjoint=self.a.DeepCopy())
        d.f = self.f
        return d

    def ToArray(self) -> np.ndarray:
        """``DVector.op_Explicit`` (:711-720): the deepest primal as a host array."""
        return self.Buffer.SubData

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.GetSlice(i.start, None if i.stop is None else i.stop - 1)
        if self.kind == PRIMAL:
            return DNumber(self.p.GetValues(i, 1).SubData[0], self.dtype)  # D(ap.SubData.[i]) (:644)
        if self.kind == FORWARD:
            return DNumber.DF(self.p[i], self.t[i], self.tag)
        return DNumber.DR(self.p[i], TraceOp("Item_DV", self, i), self.tag)

    def GetSlice(self, lower=None, upper=None):
        """:668-677 (`upper` inclusive)"""
        l = 0 if lower is None else lower
        u = self.Length - 1 if upper is None else upper
        if self.kind == PRIMAL:
            return DVector(self.p.GetValues(l, u - l + 1))
        if self.kind == FORWARD:
            return DVector.DVF(self.p.GetSlice(l, u), self.t.GetSlice(l, u), self.tag)
        return DVector.DVR(self.p.GetSlice(l, u), TraceOp("Slice_DV", self, l), self.tag)

    def __repr__(self):
        return f"{('DV', 'DVF', 'DVR')[self.kind]} : {self.Length}"

    # constructors (:706-707, :722-735)
    @staticmethod
    def Zero(dtype=np.float32):
        return DVector(NativeDataBuffer(np.empty(0, dtype)))

    @staticmethod
    def ZeroN(n, dtype=np.float32):
        return DVector(NativeDataBuffer(np.zeros(n, dtype)))

    @staticmethod
    def OfArray(a):
        a0 = a[0]
        if a0.kind == PRIMAL:
            return DVector(Backend(a0).CreateDataBuffer(np.array([x.p for x in a], dtype=a0.dtype)))
        if a0.kind == FORWARD:
            return DVector.DVF(DVector.OfArray([x.P for x in a]), DVector.OfArray([x.T for x in a]), a0.tag)
        return DVector.DVR(DVector.OfArray([x.P for x in a]), TraceOp("Make_DV_ofDs", list(a)), a0.tag)

    # ---- operators ----------------------------------------------------------------------------
    def __add__(a, b):
        if isinstance(b, DVector):  # :1083
            return _op2(a, b, lambda x, y: Backend(x).Add_V_V(x, y), lambda x, y: x + y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: bt, lambda cp, ap, at, bp, bt: at + bt,
                        lambda x, y: TraceOp("Add_DV_DV", x, y), lambda x, y: TraceOp("Add_DV_DVCons", x),
                        lambda x, y: TraceOp("Add_DV_DVCons", y), DVector)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :1227
            return _op2(a, b, lambda x, s: Backend(x).Add_S_V(s, x), lambda x, y: x + y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: DVector.OfArray([bt] * a.Length),
                        lambda cp, ap, at, bp, bt: at + bt,
                        lambda x, y: TraceOp("Add_DV_D", x, y), lambda x, y: TraceOp("Add_DV_DCons", x),
                        lambda x, y: TraceOp("Add_DVCons_D", y), DVector)
        return NotImplemented

    def __radd__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # :1239
            return _op2(a, b, lambda s, x: Backend(s).Add_S_V(s, x), lambda x, y: x + y,
                        lambda cp, ap, at: DVector.OfArray([at] * b.Length), lambda cp, bp, bt: bt,
                        lambda cp, ap, at, bp, bt: at + bt,
                        lambda x, y: TraceOp("Add_DV_D", y, x), lambda x, y: TraceOp("Add_DVCons_D", x),
                        lambda x, y: TraceOp("Add_DV_DCons", y), DVector)
        return NotImplemented

    def __sub__(a, b):
        if isinstance(b, DVector):  # :1095
            return _op2(a, b, lambda x, y: Backend(x).Sub_V_V(x, y), lambda x, y: x - y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: -bt, lambda cp, ap, at, bp, bt: at - bt,
                        lambda x, y: TraceOp("Sub_DV_DV", x, y), lambda x, y: TraceOp("Sub_DV_DVCons", x),
                        lambda x, y: TraceOp("Sub_DVCons_DV", y), DVector)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :1251
            return _op2(a, b, lambda x, s: Backend(x).Sub_V_S(x, s), lambda x, y: x - y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: DVector.OfArray([-bt] * a.Length),
                        lambda cp, ap, at, bp, bt: at - bt,
                        lambda x, y: TraceOp("Sub_DV_D", x, y), lambda x, y: TraceOp("Sub_DV_DCons", x),
                        lambda x, y: TraceOp("Sub_DVCons_D", y), DVector)
        return NotImplemented

    def __rsub__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # :1263
            return _op2(a, b, lambda s, x: Backend(s).Sub_S_V(s, x), lambda x, y: x - y,
                        lambda cp, ap, at: DVector.OfArray([at] * b.Length), lambda cp, bp, bt: -bt,
                        lambda cp, ap, at, bp, bt: at - bt,
                        lambda x, y: TraceOp("Sub_D_DV", x, y), lambda x, y: TraceOp("Sub_D_DVCons", x),
                        lambda x, y: TraceOp("Sub_DCons_DV", y), DVector)
        return NotImplemented

    def __mul__(a, b):
        if isinstance(b, DVector):  # inner product :1107
            return _op2(a, b, lambda x, y: Backend(x).Mul_Dot_V_V(x, y), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_Dot_DV_DV", x, y), lambda x, y: TraceOp("Mul_Dot_DV_DVCons", x, y),
                        lambda x, y: TraceOp("Mul_Dot_DV_DVCons", y, x), DNumber)
        if isinstance(b, DNDArray):  # DV * DM :2417
            return _op2(a, b, lambda x, y: Backend(y).Mul_V_M(x, y), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DV_DM", x, y), lambda x, y: TraceOp("Mul_DV_DMCons", x, y),
                        lambda x, y: TraceOp("Mul_DVCons_DM", x, y), DVector)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # DV * D :1179
            return _op2(a, b, lambda x, s: Backend(x).Mul_S_V(s, x), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DV_D", x, y), lambda x, y: TraceOp("Mul_DV_DCons", x, y),
                        lambda x, y: TraceOp("Mul_DVCons_D", x, y), DVector)
        return NotImplemented

    def __rmul__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # D * DV :1191
            return _op2(a, b, lambda s, x: Backend(s).Mul_S_V(s, x), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DV_D", y, x), lambda x, y: TraceOp("Mul_DVCons_D", y, x),
                        lambda x, y: TraceOp("Mul_DV_DCons", y, x), DVector)
        return NotImplemented

    def __truediv__(a, b):
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # DV / D :1203 — QUIRK: multiply by the reciprocal computed in 'T
            one = a.dtype.type(1)
            return _op2(a, b, lambda x, s: Backend(x).Mul_S_V(one / s, x), lambda x, y: x / y,
                        lambda cp, ap, at: at / b, lambda cp, bp, bt: -bt * cp / bp,
                        lambda cp, ap, at, bp, bt: (at - bt * cp) / bp,
                        lambda x, y: TraceOp("Div_DV_D", x, y), lambda x, y: TraceOp("Div_DV_DCons", x, y),
                        lambda x, y: TraceOp("Div_DVCons_D", x, y), DVector)
        return NotImplemented

    def __rtruediv__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # D / DV :1215
            return _op2(a, b, lambda s, x: Backend(s).Map_F_S_V(s, MapOp.Div, None, x), lambda x, y: x / y,
                        lambda cp, ap, at: at / b, lambda cp, bp, bt: had(-bt, hdiv(cp, bp)),
                        lambda cp, ap, at, bp, bt: (at - bt * cp) / bp,
                        lambda x, y: TraceOp("Div_D_DV", x, y), lambda x, y: TraceOp("Div_D_DVCons", x, y),
                        lambda x, y: TraceOp("Div_DCons_DV", x, y), DVector)
        return NotImplemented

    def __neg__(a):  # :1432
        m1 = a.dtype.type(-1)
        return _op1(a, lambda x: Backend(x).Mul_S_V(m1, x), lambda x: -x, lambda cp, ap, at: -at,
                    lambda x: TraceOp("Neg_DV", x), DVector)

    def __pow__(a, b):
        one = DNumber(1, a.dtype)
        if isinstance(b, DVector):  # :1155
            return _op2(a, b, lambda x, y: Backend(x).Map2_F_V_V(MapOp.Pow_Ab, None, x, y), lambda x, y: x ** y,
                        lambda cp, ap, at: had(had(at, ap ** (b - one)), b),
                        lambda cp, bp, bt: had(had(bt, cp), log(a)),
                        lambda cp, ap, at, bp, bt: had(ap ** (bp - one), had(at, bp) + had(had(ap, bt), log(ap))),
                        lambda x, y: TraceOp("Pow_DV_DV", x, y), lambda x, y: TraceOp("Pow_DV_DVCons", x, y),
                        lambda x, y: TraceOp("Pow_DVCons_DV", x, y), DVector)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :1275
            return _op2(a, b, lambda x, s: Backend(x).Map_F_S_V(s, MapOp.Pow_Ab, None, x), lambda x, y: x ** y,
                        lambda cp, ap, at: had(at, ap ** (b - one)) * b,
                        lambda cp, bp, bt: had(bt * cp, log(a)),
                        lambda cp, ap, at, bp, bt: had(ap ** (bp - one), (at * bp) + had(ap * bt, log(ap))),
                        lambda x, y: TraceOp("Pow_DV_D", x, y), lambda x, y: TraceOp("Pow_DV_DCons", x, y),
                        lambda x, y: TraceOp("Pow_DVCons_D", x, y), DVector)
        return NotImplemented

    def __rpow__(b, a):
        a = _coerce(a, b)
        one = DNumber(1, b.dtype)
        if isinstance(a, DNumber):  # :1288
            return _op2(a, b, lambda s, x: Backend(s).Map_F_S_V(s, MapOp.Pow_Ba, None, x), lambda x, y: x ** y,
                        lambda cp, ap, at: had(at * (ap ** (b - one)), b),
                        lambda cp, bp, bt: had(bt, cp) * log(a),
                        lambda cp, ap, at, bp, bt: had(ap ** (bp - one), (at * bp) + (ap * bt * log(ap))),
                        lambda x, y: TraceOp("Pow_D_DV", x, y), lambda x, y: TraceOp("Pow_D_DVCons", x, y),
                        lambda x, y: TraceOp("Pow_DCons_DV", x, y), DVector)
        return NotImplemented

    # Hadamard product / division / outer product (:1119, :1143, :1131)
    def Had(a, b):
        return _op2(a, b, lambda x, y: Backend(x).Map2_F_V_V(MapOp.Mul, None, x, y), lambda x, y: had(x, y),
                    lambda cp, ap, at: had(at, b), lambda cp, bp, bt: had(a, bt),
                    lambda cp, ap, at, bp, bt: had(at, bp) + had(ap, bt),
                    lambda x, y: TraceOp("Mul_Had_DV_DV", x, y), lambda x, y: TraceOp("Mul_Had_DV_DVCons", x, y),
                    lambda x, y: TraceOp("Mul_Had_DV_DVCons", y, x), DVector)

    def HDiv(a, b):
        return _op2(a, b, lambda x, y: Backend(x).Map2_F_V_V(MapOp.Div, None, x, y), lambda x, y: hdiv(x, y),
                    lambda cp, ap, at: hdiv(at, b), lambda cp, bp, bt: hdiv(had(-bt, cp), bp),
                    lambda cp, ap, at, bp, bt: hdiv(at - had(bt, cp), bp),
                    lambda x, y: TraceOp("Div_Had_DV_DV", x, y), lambda x, y: TraceOp("Div_Had_DV_DVCons", x, y),
                    lambda x, y: TraceOp("Div_Had_DVCons_DV", x, y), DVector)

    def Outer(a, b):
        return _op2(a, b, lambda x, y: Backend(x).Mul_Out_V_V(x, y), lambda x, y: outer(x, y),
                    lambda cp, ap, at: outer(at, b), lambda cp, bp, bt: outer(a, bt),
                    lambda cp, ap, at, bp, bt: outer(at, bp) + outer(ap, bt),
                    lambda x, y: TraceOp("Mul_Out_DV_DV", x, y), lambda x, y: TraceOp("Mul_Out_DV_DVCons", x, y),
                    lambda x, y: TraceOp("Mul_Out_DVCons_DV", x, y), DNDArray)

    # reductions (:1526-1561)
    def L1Norm(a):
        return _op1(a, lambda x: Backend(x).L1Norm_V(x), lambda x: x.L1Norm(), lambda cp, ap, at: at * Sign(ap),
                    lambda x: TraceOp("L1Norm_DV", x), DNumber)

    def L2NormSq(a):
        def ff(x):
            n = Backend(x).L2Norm_V(x)  # QUIRK: nrm2 squared, not a dot product (:1538-1540)
            return n * n
        two = DNumber(2, a.dtype)
        return _op1(a, ff, lambda x: x.L2NormSq(), lambda cp, ap, at: two * (ap * at),
                    lambda x: TraceOp("L2NormSq_DV", x), DNumber)

    def L2Norm(a):
        return _op1(a, lambda x: Backend(x).L2Norm_V(x), lambda x: x.L2Norm(), lambda cp, ap, at: (ap * at) / cp,
                    lambda x: TraceOp("L2Norm_DV", x), DNumber)

    def Sum(a):
        return _op1(a, lambda x: Backend(x).Sum_V(x), lambda x: x.Sum(), lambda cp, ap, at: at.Sum(),
                    lambda x: TraceOp("Sum_DV", x), DNumber)

    def Mean(a):
        return a.Sum() / a.Length

    def ReshapeToDM(a, m):  # :1578
        return _op1(a, lambda x: Backend(x).ReshapeCopy_V_MRows(m, x), lambda x: x.ReshapeToDM(m),
                    lambda cp, ap, at: at.ReshapeToDM(m), lambda x: TraceOp("ReshapeCopy_DV_DM", x), DNDArray)

    def MaxIndex(a):  # host loop on purpose (:1629-1638)
        v = a.ToArray()
        maxi, maxv = 0, v[0]
        for i in range(v.size):
            if v[i] > maxv:
                maxi, maxv = i, v[i]
        return maxi

    def MinIndex(a):
        v = a.ToArray()
        mini, minv = 0, v[0]
        for i in range(v.size):
            if v[i] < minv:
                mini, minv = i, v[i]
        return mini

    def Max(a): return a[a.MaxIndex()]
    def Min(a): return a[a.MinIndex()]

    def SoftMax(a):  # :1654-1657
        e = exp(a - a.Max())
        return e / e.Sum()


DV = DVector


# ================================================================================================
# DNDArray (:1678-3006)
# ================================================================================================
class DNDArray(_AD):
    __slots__ = ()

    def __init__(self, view: ShapedDataBufferView):
        """``DM of ShapedDataBufferView`` — primal."""
        self._init(PRIMAL, view)

    @staticmethod
    def DMF(p, t, tag):
        d = DNDArray.__new__(DNDArray)
        d._init(FORWARD, p, t=t, tag=tag)
        return d

    @staticmethod
    def DMR(p, op, tag, adjoint=None):
        d = DNDArray.__new__(DNDArray)
        # zero adjoint through the BACKEND at node creation (ZeroMN, :1834-1835) ...
        d._init(REVERSE, p, a=adjoint if adjoint is not None else DNDArray.ZeroMN(p.Rows, p.Cols, Backend(p)),
                op=op, tag=tag)
        return d

    primal = staticmethod(lambda view: DNDArray(view))
    forward = staticmethod(lambda cp, t, tag: DNDArray.DMF(cp, t, tag))
    reverse = staticmethod(lambda cp, op, tag: DNDArray.DMR(cp, op, tag))

    @property
    def dtype(self):
        return self.PD.p.dtype

    @property
    def Buffer(self) -> ShapedDataBufferView:
        return self.PD.p

    Rows = property(lambda s: s.PD.p.Rows)
    Cols = property(lambda s: s.PD.p.Cols)
    Length = property(lambda s: s.PD.p.Length)

    @property
    def T(self):
        if self.kind == PRIMAL:
            return DNDArray.ZeroMN(self.Rows, self.Cols, Backend(self))
        if self.kind == FORWARD:
            return self.t
        raise RuntimeError("Cannot get tangent value of DMR.")

    @property
    def A(self):
        if self.kind == PRIMAL:
            return DNDArray.ZeroMN(self.Rows, self.Cols, Backend(self))
        if self.kind == FORWARD:
            raise RuntimeError("Cannot get adjoint value of DMF.")
        return self.a

    @A.setter
    def A(self, v):
        if self.kind == FORWARD:
            raise RuntimeError("Cannot set adjoint value of DMF.")
        if self.kind == REVERSE:
            self.a = v

    def GetForward(self, t, i):
        return DNDArray.DMF(self, t, i)

    def GetReverse(self, i):  # :1743-1744
        b = Backend(self)
        z = DNDArray(ShapedDataBufferView(b.CreateDataBuffer(b.CreateZeroArray(self.Buffer.Length)), *self.Buffer.Shape))
        return DNDArray.DMR(self, Noop, i, adjoint=z)

    def DeepCopy(self):
        if self.kind == PRIMAL:
            return DNDArray(self.p.DeepCopy())
        if self.kind == FORWARD:
            return DNDArray.DMF(self.p.DeepCopy(), self.t.DeepCopy(), self.tag)
        d = DNDArray.DMR(self.p.DeepCopy(), self.op, self.tag, adjoint=self.a.DeepCopy())
        d.f = self.f
        return d

    def CustomOp(self, customInfo):  # :1758-1763 — opaque consumer hook (Backend.fs:137-139)
        if self.kind == PRIMAL:
            return DNDArray(Backend(self.p).CustomOp_DM_Forward(self.p, customInfo))
        if self.kind == FORWARD:
            return DNDArray.DMF(DNDArray(Backend(self.p).CustomOp_DM_Forward(self.p.Buffer, customInfo)),
                                DNDArray(Backend(self.t).CustomOp_DM_Forward(self.t.Buffer, customInfo)), self.tag)
        return DNDArray.DMR(DNDArray(Backend(self.p).CustomOp_DM_Forward(self.p.Buffer, customInfo)),
                            TraceOp("Custom_DM", self, customInfo), self.tag)

    def ToArray(self) -> np.ndarray:
        return self.Buffer.to_numpy()

    def FlatItem(self, i):
        if self.kind == PRIMAL:
            return DNumber(self.p.FlatItem(i), self.dtype)
        if self.kind == FORWARD:
            return DNumber.DF(self.p.FlatItem(i), self.t.FlatItem(i), self.tag)
        return DNumber.DR(self.p.FlatItem(i), TraceOp("Item_DM", self, i // self.Cols, i % self.Cols), self.tag)

    def __getitem__(self, ij):
        i, j = ij
        if isinstance(i, slice) or isinstance(j, slice):
            return self._slice(i, j)
        if self.kind == PRIMAL:
            return DNumber(self.p[i, j], self.dtype)
        if self.kind == FORWARD:
            return DNumber.DF(self.p[i, j], self.t[i, j], self.tag)
        return DNumber.DR(self.p[i, j], TraceOp("Item_DM", self, i, j), self.tag)

    def _slice(self, i, j):
        def bounds(s, n):
            if isinstance(s, slice):
                return (0 if s.start is None else s.start), (n - 1 if s.stop is None else s.stop - 1)
            return s, s
        c0, c1 = bounds(j, self.Cols)
        if not isinstance(i, slice):  # d.[row, c0..c1] -> DVector (:1813-1821)
            row = i
            if self.kind == PRIMAL:
                return DVector(self.p.GetRowSlice(row, c0, c1).DataBuffer)
            if self.kind == FORWARD:
                return DVector.DVF(self.p[row, c0:c1 + 1], self.t[row, c0:c1 + 1], self.tag)
            cp = self.p[row, c0:c1 + 1]
            b = Backend(self)
            z = DVector(b.CreateDataBuffer(b.CreateZeroArray(cp.Length)))
            return DVector.DVR(cp, TraceOp("SliceRow_DM", self, row, c0), self.tag, adjoint=z)
        r0, r1 = bounds(i, self.Rows)  # d.[r0..r1, c0..c1] (:1801-1811)
        if self.kind == PRIMAL:
            return DNDArray(self.p.GetSlice(r0, r1, c0, c1))
        if self.kind == FORWARD:
            return DNDArray.DMF(self.p[r0:r1 + 1, c0:c1 + 1], self.t[r0:r1 + 1, c0:c1 + 1], self.tag)
        cp = self.p[r0:r1 + 1, c0:c1 + 1]
        return DNDArray.DMR(cp, TraceOp("Slice_DM", self, r0, c0), self.tag,
                            adjoint=DNDArray.ZeroMN(cp.Rows, cp.Cols, Backend(self)))

    def GetRows(self):
        return (self[i, :] for i in range(self.Rows))

    def __repr__(self):
        return f"{('DM', 'DMF', 'DMR')[self.kind]} : {self.Rows} x {self.Cols}"

    # constructors (:1833-1890)
    @staticmethod
    def Zero(dtype=np.float32):
        return DNDArray(ShapedDataBufferView(NativeDataBuffer(np.empty(0, dtype)), 0, 0))

    @staticmethod
    def ZeroMN(m, n, b):
        return DNDArray(ShapedDataBufferView(b.CreateDataBuffer(b.CreateZeroArray(m * n)), m, n))

    @staticmethod
    def OfRows(m, a: DVector, b):
        """``DNDArray.OfRows(m, DVector, backend)`` (:1860-1865): m copies of `a` as rows — how the fork
        broadcasts a bias."""
        if a.kind == PRIMAL:
            return DNDArray(b.RepeatReshapeCopy_V_MRows(m, a.p))
        if a.kind == FORWARD:
            return DNDArray.DMF(DNDArray.OfRows(m, a.p, b), DNDArray.OfRows(m, a.t, b), a.tag)
        cp = DNDArray.OfRows(m, a.p, b)
        return DNDArray.DMR(cp, TraceOp("Make_DMRows_ofDV", a), a.tag, adjoint=DNDArray.ZeroMN(cp.Rows, cp.Cols, b))

    @staticmethod
    def OfDNumberArray(m, n, a: DNumber, backend):
        """:1868-1880 (float32 signature; the float64 file takes a DNumber[] — same device result)"""
        if a.kind == PRIMAL:
            data = backend.CreateValueArray(m * n, a.p)
            return DNDArray(ShapedDataBufferView(backend.CreateDataBuffer(data), m, n))
        if a.kind == FORWARD:
            return DNDArray.DMF(DNDArray.OfDNumberArray(m, n, a.P, backend), DNDArray.OfDNumberArray(m, n, a.T, backend), a.tag)
        cp = DNDArray.OfDNumberArray(m, n, a.P, backend)
        return DNDArray.DMR(cp, TraceOp("Make_DM_ofDs", a), a.tag)

    @staticmethod
    def OfNumberArray(m, a, backend):
        """:1882-1890 (QUIRK App. D-14: the padding branch discards its result; harmless)"""
        n = len(a) // m
        return DNDArray(ShapedDataBufferView(backend.CreateDataBuffer(a), m, n))

    @staticmethod
    def create(m, n, v, b):
        """``DM.create`` (:3427-3432)"""
        if isinstance(v, DNumber):
            return DNDArray.OfDNumberArray(m, n, v, b)
        return DNDArray.OfNumberArray(m, b.CreateValueArray(m * n, v), b)

    # ---- operators ----------------------------------------------------------------------------
    def __add__(a, b):
        if isinstance(b, DNDArray):  # :2357
            return _op2(a, b, lambda x, y: Backend(x).Add_M_M(x, y), lambda x, y: x + y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: bt, lambda cp, ap, at, bp, bt: at + bt,
                        lambda x, y: TraceOp("Add_DM_DM", x, y), lambda x, y: TraceOp("Add_DM_DMCons", x),
                        lambda x, y: TraceOp("Add_DM_DMCons", y), DNDArray)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :2507
            backend = Backend(a)
            return _op2(a, b, lambda x, s: backend.Add_S_M(s, x), lambda x, y: x + y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: DNDArray.OfDNumberArray(a.Rows, a.Cols, bt, backend),
                        lambda cp, ap, at, bp, bt: at + bt,
                        lambda x, y: TraceOp("Add_DM_D", x, y), lambda x, y: TraceOp("Add_DM_DCons", x),
                        lambda x, y: TraceOp("Add_DMCons_D", y), DNDArray)
        return NotImplemented

    def __radd__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # :2519
            backend = Backend(b)
            return _op2(a, b, lambda s, x: backend.Add_S_M(s, x), lambda x, y: x + y,
                        lambda cp, ap, at: DNDArray.OfDNumberArray(b.Rows, b.Cols, at, backend), lambda cp, bp, bt: bt,
                        lambda cp, ap, at, bp, bt: at + bt,
                        lambda x, y: TraceOp("Add_DM_D", y, x), lambda x, y: TraceOp("Add_DMCons_D", x),
                        lambda x, y: TraceOp("Add_DM_DCons", y), DNDArray)
        return NotImplemented

    def __sub__(a, b):
        if isinstance(b, DNDArray):  # :2369
            return _op2(a, b, lambda x, y: Backend(x).Sub_M_M(x, y), lambda x, y: x - y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: -bt, lambda cp, ap, at, bp, bt: at - bt,
                        lambda x, y: TraceOp("Sub_DM_DM", x, y), lambda x, y: TraceOp("Sub_DM_DMCons", x),
                        lambda x, y: TraceOp("Sub_DMCons_DM", y), DNDArray)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :2531
            backend = Backend(a)
            return _op2(a, b, lambda x, s: backend.Sub_M_S(x, s), lambda x, y: x - y,
                        lambda cp, ap, at: at, lambda cp, bp, bt: DNDArray.OfDNumberArray(a.Rows, a.Cols, -bt, backend),
                        lambda cp, ap, at, bp, bt: at - bt,
                        lambda x, y: TraceOp("Sub_DM_D", x, y), lambda x, y: TraceOp("Sub_DM_DCons", x),
                        lambda x, y: TraceOp("Sub_DMCons_D", y), DNDArray)
        return NotImplemented

    def __rsub__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # :2543
            backend = Backend(b)
            return _op2(a, b, lambda s, x: backend.Sub_S_M(s, x), lambda x, y: x - y,
                        lambda cp, ap, at: DNDArray.OfDNumberArray(b.Rows, b.Cols, at, backend), lambda cp, bp, bt: -bt,
                        lambda cp, ap, at, bp, bt: at - bt,
                        lambda x, y: TraceOp("Sub_D_DM", x, y), lambda x, y: TraceOp("Sub_D_DMCons", x),
                        lambda x, y: TraceOp("Sub_DCons_DM", y), DNDArray)
        return NotImplemented

    def __mul__(a, b):
        if isinstance(b, DNDArray):  # matrix product :2381
            return _op2(a, b, lambda x, y: Backend(x).Mul_M_M(x, y), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DM_DM", x, y), lambda x, y: TraceOp("Mul_DM_DMCons", x, y),
                        lambda x, y: TraceOp("Mul_DMCons_DM", x, y), DNDArray)
        if isinstance(b, DVector):  # DM * DV :2405
            return _op2(a, b, lambda x, y: Backend(x).Mul_M_V(x, y), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DM_DV", x, y), lambda x, y: TraceOp("Mul_DM_DVCons", x, y),
                        lambda x, y: TraceOp("Mul_DMCons_DV", x, y), DVector)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :2463
            return _op2(a, b, lambda x, s: Backend(x).Mul_S_M(s, x), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DM_D", x, y), lambda x, y: TraceOp("Mul_DM_DCons", x, y),
                        lambda x, y: TraceOp("Mul_DMCons_D", x, y), DNDArray)
        return NotImplemented

    def __rmul__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # :2474
            return _op2(a, b, lambda s, x: Backend(x).Mul_S_M(s, x), lambda x, y: x * y,
                        lambda cp, ap, at: at * b, lambda cp, bp, bt: a * bt,
                        lambda cp, ap, at, bp, bt: (at * bp) + (ap * bt),
                        lambda x, y: TraceOp("Mul_DM_D", y, x), lambda x, y: TraceOp("Mul_DMCons_D", y, x),
                        lambda x, y: TraceOp("Mul_DM_DCons", y, x), DNDArray)
        return NotImplemented

    def __truediv__(a, b):
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :2485 — QUIRK: reciprocal on the host in 'T
            one = a.dtype.type(1)
            return _op2(a, b, lambda x, s: Backend(x).Mul_S_M(one / s, x), lambda x, y: x / y,
                        lambda cp, ap, at: at / b, lambda cp, bp, bt: -bt * cp / bp,
                        lambda cp, ap, at, bp, bt: (at - bt * cp) / bp,
                        lambda x, y: TraceOp("Div_DM_D", x, y), lambda x, y: TraceOp("Div_DM_DCons", x, y),
                        lambda x, y: TraceOp("Div_DMCons_D", x, y), DNDArray)
        return NotImplemented

    def __rtruediv__(b, a):
        a = _coerce(a, b)
        if isinstance(a, DNumber):  # :2496
            return _op2(a, b, lambda s, x: Backend(x).Map_F_S_M(s, MapOp.Div, None, x), lambda x, y: x / y,
                        lambda cp, ap, at: at / b, lambda cp, bp, bt: had(-bt, hdiv(cp, bp)),
                        lambda cp, ap, at, bp, bt: hdiv(at - had(bt, cp), bp),
                        lambda x, y: TraceOp("Div_D_DM", x, y), lambda x, y: TraceOp("Div_D_DMCons", x, y),
                        lambda x, y: TraceOp("Div_DCons_DM", x, y), DNDArray)
        return NotImplemented

    def __neg__(a):  # :2675
        m1 = a.dtype.type(-1)
        return _op1(a, lambda x: Backend(x).Mul_S_M(m1, x), lambda x: -x, lambda cp, ap, at: -at,
                    lambda x: TraceOp("Neg_DM", x), DNDArray)

    def __pow__(a, b):
        one = DNumber(1, a.dtype)
        if isinstance(b, DNDArray):  # :2440
            return _op2(a, b, lambda x, y: Backend(x).Map2_F_M_M(MapOp.Pow_Ab, None, x, y), lambda x, y: x ** y,
                        lambda cp, ap, at: had(had(at, ap ** (b - one)), b),
                        lambda cp, bp, bt: had(had(bt, cp), log(a)),
                        lambda cp, ap, at, bp, bt: had(ap ** (bp - one), had(at, bp) + had(had(ap, bt), log(ap))),
                        lambda x, y: TraceOp("Pow_DM_DM", x, y), lambda x, y: TraceOp("Pow_DM_DMCons", x, y),
                        lambda x, y: TraceOp("Pow_DMCons_DM", x, y), DNDArray)
        b = _coerce(b, a)
        if isinstance(b, DNumber):  # :2555
            return _op2(a, b, lambda x, s: Backend(x).Map_F_S_M(s, MapOp.Pow_Ab, None, x), lambda x, y: x ** y,
                        lambda cp, ap, at: had(at, ap ** (b - one)) * b,
                        lambda cp, bp, bt: had(bt * cp, log(a)),
                        lambda cp, ap, at, bp, bt: had(ap ** (bp - one), (at * bp) + had(ap * bt, log(ap))),
                        lambda x, y: TraceOp("Pow_DM_D", x, y), lambda x, y: TraceOp("Pow_DM_DCons", x, y),
                        lambda x, y: TraceOp("Pow_DMCons_D", x, y), DNDArray)
        return NotImplemented

    def __rpow__(b, a):
        a = _coerce(a, b)
        one = DNumber(1, b.dtype)
        if isinstance(a, DNumber):  # :2567
            return _op2(a, b, lambda s, x: Backend(x).Map_F_S_M(s, MapOp.Pow_Ba, None, x), lambda x, y: x ** y,
                        lambda cp, ap, at: had(at * (ap ** (b - one)), b),
                        lambda cp, bp, bt: had(bt, cp) * log(a),
                        lambda cp, ap, at, bp, bt: had(ap ** (bp - one), (at * bp) + (ap * bt * log(ap))),
                        lambda x, y: TraceOp("Pow_D_DM", x, y), lambda x, y: TraceOp("Pow_D_DMCons", x, y),
                        lambda x, y: TraceOp("Pow_DCons_DM", x, y), DNDArray)
        return NotImplemented

    def Had(a, b):  # :2393
        return _op2(a, b, lambda x, y: Backend(x).Mul_Had_M_M(x, y), lambda x, y: had(x, y),
                    lambda cp, ap, at: had(at, b), lambda cp, bp, bt: had(a, bt),
                    lambda cp, ap, at, bp, bt: had(at, bp) + had(ap, bt),
                    lambda x, y: TraceOp("Mul_Had_DM_DM", x, y), lambda x, y: TraceOp("Mul_Had_DM_DMCons", x, y),
                    lambda x, y: TraceOp("Mul_Had_DM_DMCons", y, x), DNDArray)

    def HDiv(a, b):  # :2429
        return _op2(a, b, lambda x, y: Backend(x).Map2_F_M_M(MapOp.Div, None, x, y), lambda x, y: hdiv(x, y),
                    lambda cp, ap, at: hdiv(at, b), lambda cp, bp, bt: hdiv(had(-bt, cp), bp),
                    lambda cp, ap, at, bp, bt: hdiv(at - had(bt, cp), bp),
                    lambda x, y: TraceOp("Div_Had_DM_DM", x, y), lambda x, y: TraceOp("Div_Had_DM_DMCons", x, y),
                    lambda x, y: TraceOp("Div_Had_DMCons_DM", x, y), DNDArray)

    def Transpose(a):  # :2786
        return _op1(a, lambda x: Backend(x).Transpose_M(x), lambda x: x.Transpose(), lambda cp, ap, at: at.Transpose(),
                    lambda x: TraceOp("Transpose_DM", x), DNDArray)

    def Permute(a, dims):  # :2771
        return _op1(a, lambda x: Backend(x).Permute_M(x, dims), lambda x: x.Permute(dims),
                    lambda cp, ap, at: at.Permute(dims), lambda x: TraceOp("Permute_DM", x, dims), DNDArray)

    def Reshape(a, newShape):  # :2778
        return _op1(a, lambda x: Backend(x).Reshape_M(x, newShape), lambda x: x.Reshape(newShape),
                    lambda cp, ap, at: at.Reshape(newShape), lambda x: TraceOp("Reshape_DM", x, newShape), DNDArray)

    def Diagonal(a):  # :2794
        return _op1(a, lambda x: Backend(x).Diagonal_M(x), lambda x: x.Diagonal(), lambda cp, ap, at: at.Diagonal(),
                    lambda x: TraceOp("Diagonal_DM", x), DVector)

    def Trace(a):
        return a.Diagonal().Sum()

    def Sum(a):  # :2805 — Sum_M takes the *buffer*, not the view
        return _op1(a, lambda x: Backend(x).Sum_M(x.DataBuffer), lambda x: x.Sum(), lambda cp, ap, at: at.Sum(),
                    lambda x: TraceOp("Sum_DM", x), DNumber)

    def Mean(a):
        return a.Sum() / a.Length

    def ReshapeToDV(a):  # :2894
        return _op1(a, lambda x: Backend(x).ReshapeCopy_MRows_V(x), lambda x: x.ReshapeToDV(),
                    lambda cp, ap, at: at.ReshapeToDV(), lambda x: TraceOp("ReshapeCopy_DM_DV", x), DVector)

    @staticmethod
    def _solve(name):
        def Solve(a, b):  # :2814, :2830
            def ff(x, y):
                r = getattr(Backend(x), name)(x, y)
                if r is None:
                    raise ValueError("Cannot solve the given system of linear equations.")
                return r
            me = lambda x, y: getattr(x, "Solve" if name == "Solve_M_V" else "SolveSymmetric")(y)  # noqa: E731
            return _op2(a, b, ff, me, lambda cp, ap, at: me(ap, -at * cp), lambda cp, bp, bt: me(a, bt),
                        lambda cp, ap, at, bp, bt: me(ap, bt - at * cp),
                        lambda x, y: TraceOp("Solve_DM_DV", x, y), lambda x, y: TraceOp("Solve_DM_DVCons", x, y),
                        lambda x, y: TraceOp("Solve_DMCons_DV", x, y), DVector)
        return Solve

    def Inverse(a):  # :2903
        def ff(x):
            r = Backend(x).Inverse_M(x)
            if r is None:
                raise ValueError("Cannot compute the inverse of given matrix.")
            return r
        return _op1(a, ff, lambda x: x.Inverse(), lambda cp, ap, at: -cp * at * cp,
                    lambda x: TraceOp("Inverse_DM", x), DNDArray)

    def Det(a):  # :2915
        def ff(x):
            r = Backend(x).Det_M(x)
            if r is None:
                raise ValueError("Cannot compute the determinant of given matrix.")
            return r
        return _op1(a, ff, lambda x: x.Det(), lambda cp, ap, at: cp * (ap.Inverse() * at).Trace(),
                    lambda x: TraceOp("Det_DM", x), DNumber)

    def MaxIndex(a):  # :2974 — through the backend for DM (float32 file)
        return Backend(a).MaxIndex_V(a.Buffer.DataBuffer)

    def MinIndex(a):
        return Backend(a).MinIndex_V(a.Buffer.DataBuffer)

    def Max(a): return a.FlatItem(a.MaxIndex())
    def Min(a): return a.FlatItem(a.MinIndex())

    def SoftMax(a):  # :2938-2941
        e = exp(a - a.Max())
        return e / e.Sum()


DNDArray.Solve = DNDArray._solve("Solve_M_V")
DNDArray.SolveSymmetric = DNDArray._solve("SolveSymmetric_M_V")
DM = DNDArray


# ================================================================================================
# polymorphic free functions (the F# operators .* ./ &* and the math functions)
# ================================================================================================
def had(a, b):
    """``a .* b``"""
    return a.Had(b)


def hdiv(a, b):
    """``a ./ b``"""
    return a.HDiv(b)


def outer(a, b):
    """``a &* b``"""
    return a.Outer(b)


def _zeros_like_tangent(a, cp):
    """df of the piecewise-constant ops: DV -> host zeros (:1505), DM -> backend zeros (:2746)."""
    if isinstance(a, DVector):
        return DVector.ZeroN(a.Length, a.dtype)
    return DNDArray.ZeroMN(a.Rows, a.Cols, Backend(cp))


def _unary(name, mapop, scalar_fn, df, scalar_df):
    """Build the D / DV / DM overloads of one elementwise function.
    df(cp, ap, at) is written once with had/hdiv and works for DV and DM; scalar_df for D."""
    trace = {"Ceiling": "Ceil", "ReLU": "ReLU"}.get(name, name)

    def fn(a):
        if isinstance(a, DNumber):
            return _op1(a, scalar_fn, fn, scalar_df, lambda x: TraceOp(f"{trace}_D", x), DNumber)
        if isinstance(a, DVector):
            return _op1(a, lambda x: Backend(x).Map_F_V(mapop, None, x), fn, df,
                        lambda x: TraceOp(f"{trace}_DV", x), DVector)
        if isinstance(a, DNDArray):
            return _op1(a, lambda x: Backend(x).Map_F_M(mapop, None, x), fn, df,
                        lambda x: TraceOp(f"{trace}_DM", x), DNDArray)
        return fn(DNumber(a))
    fn.__name__ = name.lower()
    return fn


def _one(a):
    return DNumber(1, a.dtype)


def _log10val(a):
    return a.dtype.type(math.log(10.0)) if a.dtype == np.float64 else np.float32(np.log(np.float32(10.0)))


log = _unary("Log", MapOp.Log, _np1(np.log), lambda cp, ap, at: hdiv(at, ap), lambda cp, ap, at: at / ap)
log10 = _unary("Log10", MapOp.Log10, _np1(np.log10), lambda cp, ap, at: hdiv(at, ap * _log10val(ap)),
               lambda cp, ap, at: at / (ap * _log10val(ap)))
exp = _unary("Exp", MapOp.Exp, _np1(np.exp), lambda cp, ap, at: had(at, cp), lambda cp, ap, at: at * cp)
sin = _unary("Sin", MapOp.Sin, _np1(np.sin), lambda cp, ap, at: had(at, cos(ap)), lambda cp, ap, at: at * cos(ap))
cos = _unary("Cos", MapOp.Cos, _np1(np.cos), lambda cp, ap, at: had(-at, sin(ap)), lambda cp, ap, at: -at * sin(ap))


def _tan_df(cp, ap, at):
    cosa = cos(ap)
    return hdiv(at, had(cosa, cosa))


def _tan_sdf(cp, ap, at):
    cosa = cos(ap)
    return at / (cosa * cosa)


tan = _unary("Tan", MapOp.Tan, _np1(np.tan), _tan_df, _tan_sdf)
sqrt = _unary("Sqrt", MapOp.Sqrt, _np1(np.sqrt), lambda cp, ap, at: hdiv(at, DNumber(2, cp.dtype) * cp),
              lambda cp, ap, at: at / (DNumber(2, cp.dtype) * cp))
sinh = _unary("Sinh", MapOp.Sinh, _np1(np.sinh), lambda cp, ap, at: had(at, cosh(ap)), lambda cp, ap, at: at * cosh(ap))
cosh = _unary("Cosh", MapOp.Cosh, _np1(np.cosh), lambda cp, ap, at: had(at, sinh(ap)), lambda cp, ap, at: at * sinh(ap))


def _tanh_df(cp, ap, at):  # :1464-1466, :2707-2709
    cosha = cosh(ap)
    return hdiv(at, had(cosha, cosha))


def _tanh_sdf(cp, ap, at):
    cosha = cosh(ap)
    return at / (cosha * cosha)


tanh = _unary("Tanh", MapOp.Tanh, _np1(np.tanh), _tanh_df, _tanh_sdf)
asin = _unary("Asin", MapOp.Asin, _np1(np.arcsin), lambda cp, ap, at: hdiv(at, sqrt(_one(ap) - had(ap, ap))),
              lambda cp, ap, at: at / sqrt(_one(ap) - ap * ap))
acos = _unary("Acos", MapOp.Acos, _np1(np.arccos), lambda cp, ap, at: hdiv(-at, sqrt(_one(ap) - had(ap, ap))),
              lambda cp, ap, at: -at / sqrt(_one(ap) - ap * ap))
# QUIRK (App. B): the DV/DM forward tangent of atan divides by sqrt(1 + ap^2) (:1489, :2732);
# the scalar rule (:437) and every reverse rule (:4243) use 1 + ap^2.
atan = _unary("Atan", MapOp.Atan, _np1(np.arctan), lambda cp, ap, at: hdiv(at, sqrt(_one(ap) + had(ap, ap))),
              lambda cp, ap, at: at / (_one(ap) + ap * ap))
Abs = _unary("Abs", MapOp.Abs, lambda x: abs(x), lambda cp, ap, at: had(at, Sign(ap)), lambda cp, ap, at: at * Sign(ap))
Sign = _unary("Sign", MapOp.Sign, _signummod, lambda cp, ap, at: _zeros_like_tangent(ap, cp),
              lambda cp, ap, at: DNumber(0, ap.dtype))
floor = _unary("Floor", MapOp.Floor, _np1(np.floor), lambda cp, ap, at: _zeros_like_tangent(ap, cp),
               lambda cp, ap, at: DNumber(0, ap.dtype))
ceil = _unary("Ceiling", MapOp.Ceiling, _np1(np.ceil), lambda cp, ap, at: _zeros_like_tangent(ap, cp),
              lambda cp, ap, at: DNumber(0, ap.dtype))
round_ = _unary("Round", MapOp.Round, _np1(np.rint), lambda cp, ap, at: _zeros_like_tangent(ap, cp),
                lambda cp, ap, at: DNumber(0, ap.dtype))


def _relu_scalar(x):
    return x if x != x else (x if x > 0 else x.dtype.type(0))


def _sigmoid_scalar(x):
    one = x.dtype.type(1)
    with np.errstate(all="ignore"):
        return one / (one + x.dtype.type(np.exp(np.float64(-x))))


reLU = _unary("ReLU", MapOp.ReL, _relu_scalar, lambda cp, ap, at: had(at, (ap.dtype.type(1) + Sign(ap)) / ap.dtype.type(2)),
              lambda cp, ap, at: at * (ap.dtype.type(1) + Sign(ap)) / ap.dtype.type(2))
sigmoid = _unary("Sigmoid", MapOp.Sigmoid, _sigmoid_scalar, lambda cp, ap, at: had(had(at, cp), cp.dtype.type(1) - cp),
                 lambda cp, ap, at: at * cp * (cp.dtype.type(1) - cp))


# ================================================================================================
# TraceOp tables for the reverse sweep
#   _RESET[name](args)    -> operands that received a tape edge (resetRec, :3572-3845)
#   _PUSH[name](d, args)  -> [(adjoint contribution, operand), ...] in the reference's cons order
# ================================================================================================
_RESET, _PUSH = {}, {}


def _reg(name, reset, push):
    _RESET[name], _PUSH[name] = reset, push


def _zero_like(a):
    if isinstance(a, DNumber):
        return DNumber(0, a.dtype)
    if isinstance(a, DVector):
        return DVector.Zero(a.dtype)
    return DNDArray.Zero(a.dtype)


def _reg_family(S, a0=lambda a: [a[0]], both=lambda a: [a[0], a[1]], second=lambda a: [a[1]]):
    """Elementwise / scalar-broadcast families whose adjoint text is identical for DV and DM up to the
    suffix S ('DV' or 'DM') — :3958-4057 and :4122-4244."""
    red = (lambda x: x.Sum())
    # + -
    _reg(f"Add_{S}_{S}", both, lambda d, a: [(d.A, a[0]), (d.A, a[1])])
    _reg(f"Add_{S}_{S}Cons", a0, lambda d, a: [(d.A, a[0])])
    _reg(f"Add_{S}_D", both, lambda d, a: [(d.A, a[0]), (red(d.A), a[1])])
    _reg(f"Add_{S}_DCons", a0, lambda d, a: [(d.A, a[0])])
    _reg(f"Add_{S}Cons_D", a0, lambda d, a: [(red(d.A), a[0])])
    _reg(f"Sub_{S}_{S}", both, lambda d, a: [(d.A, a[0]), (-d.A, a[1])])
    _reg(f"Sub_{S}_{S}Cons", a0, lambda d, a: [(d.A, a[0])])
    _reg(f"Sub_{S}Cons_{S}", a0, lambda d, a: [(-d.A, a[0])])
    _reg(f"Sub_{S}_D", both, lambda d, a: [(d.A, a[0]), (-(red(d.A)), a[1])])
    _reg(f"Sub_{S}_DCons", a0, lambda d, a: [(d.A, a[0])])
    _reg(f"Sub_{S}Cons_D", a0, lambda d, a: [(-(red(d.A)), a[0])])
    _reg(f"Sub_D_{S}", both, lambda d, a: [(red(d.A), a[0]), (-d.A, a[1])])
    _reg(f"Sub_D_{S}Cons", a0, lambda d, a: [(red(d.A), a[0])])
    _reg(f"Sub_DCons_{S}", a0, lambda d, a: [(-d.A, a[0])])
    # .* ./
    _reg(f"Mul_Had_{S}_{S}", both, lambda d, a: [(had(d.A, a[1].P), a[0]), (had(d.A, a[0].P), a[1])])
    _reg(f"Mul_Had_{S}_{S}Cons", a0, lambda d, a: [(had(d.A, a[1]), a[0])])
    _reg(f"Div_Had_{S}_{S}", both,
         lambda d, a: [(hdiv(d.A, a[1].P), a[0]), (had(d.A, hdiv(-a[0].P, had(a[1].P, a[1].P))), a[1])])
    _reg(f"Div_Had_{S}_{S}Cons", a0, lambda d, a: [(hdiv(d.A, a[1]), a[0])])
    _reg(f"Div_Had_{S}Cons_{S}", second, lambda d, a: [(had(d.A, hdiv(-a[0], had(a[1].P, a[1].P))), a[1])])
    # scalar scaling
    if S == "DV":
        _reg("Mul_DV_D", both, lambda d, a: [(d.A * a[1].P, a[0]), (d.A * a[0].P, a[1])])
        _reg("Mul_DV_DCons", a0, lambda d, a: [(d.A * a[1], a[0])])
        _reg("Mul_DVCons_D", second, lambda d, a: [(d.A * a[0], a[1])])
        _reg("Div_DV_D", both, lambda d, a: [(d.A / a[1].P, a[0]), (d.A * (-a[0].P / (a[1].P * a[1].P)), a[1])])
        _reg("Div_DVCons_D", second, lambda d, a: [(d.A * (-a[0] / (a[1].P * a[1].P)), a[1])])
    else:
        _reg("Mul_DM_D", both, lambda d, a: [(d.A * a[1].P, a[0]), (had(d.A, a[0].P).Sum(), a[1])])
        _reg("Mul_DM_DCons", a0, lambda d, a: [(d.A * a[1], a[0])])
        _reg("Mul_DMCons_D", second, lambda d, a: [(had(d.A, a[0]).Sum(), a[1])])
        _reg("Div_DM_D", both, lambda d, a: [(d.A / a[1].P, a[0]), (had(d.A, -a[0].P / (a[1].P * a[1].P)).Sum(), a[1])])
        # QUIRK: pushes a *matrix* into a scalar node (:4184); kept verbatim
        _reg("Div_DMCons_D", second, lambda d, a: [(had(d.A, -a[0] / (a[1].P * a[1].P)), a[1])])
    _reg(f"Div_{S}_DCons", a0, lambda d, a: [(d.A / a[1], a[0])])
    _reg(f"Div_D_{S}", both,
         lambda d, a: [(hdiv(d.A, a[1].P).Sum(), a[0]), (had(d.A, -a[0].P / had(a[1].P, a[1].P)), a[1])])
    _reg(f"Div_D_{S}Cons", a0, lambda d, a: [(hdiv(d.A, a[1]).Sum(), a[0])])
    _reg(f"Div_DCons_{S}", second, lambda d, a: [(had(d.A, -a[0] / had(a[1].P, a[1].P)), a[1])])
    # pow
    one = lambda x: DNumber(1, x.dtype)  # noqa: E731
    _reg(f"Pow_{S}_{S}", both, lambda d, a: [(had(had(d.A, a[0].P ** (a[1].P - one(d))), a[1].P), a[0]),
                                             (had(had(d.A, a[0].P ** a[1].P), log(a[0].P)), a[1])])
    _reg(f"Pow_{S}_{S}Cons", a0, lambda d, a: [(had(had(d.A, a[0].P ** (a[1] - one(d))), a[1]), a[0])])
    _reg(f"Pow_{S}Cons_{S}", second, lambda d, a: [(had(had(d.A, a[0] ** a[1].P), log(a[0])), a[1])])
    _reg(f"Pow_{S}_D", both, lambda d, a: [(had(d.A, a[0].P ** (a[1].P - one(d))) * a[1].P, a[0]),
                                           (had(had(d.A, a[0].P ** a[1].P), log(a[0].P)).Sum(), a[1])])
    _reg(f"Pow_{S}_DCons", a0, lambda d, a: [(had(d.A, a[0].P ** (a[1] - one(d))) * a[1], a[0])])
    _reg(f"Pow_{S}Cons_D", second, lambda d, a: [(had(had(d.A, a[0] ** a[1].P), log(a[0])).Sum(), a[1])])
    _reg(f"Pow_D_{S}", both, lambda d, a: [(had(had(d.A, a[0].P ** (a[1].P - one(d))), a[1].P).Sum(), a[0]),
                                           (had(d.A, a[0].P ** a[1].P) * log(a[0].P), a[1])])
    _reg(f"Pow_D_{S}Cons", a0, lambda d, a: [(had(had(d.A, a[0].P ** (a[1] - one(d))), a[1]).Sum(), a[0])])
    _reg(f"Pow_DCons_{S}", second, lambda d, a: [(had(d.A, a[0] ** a[1].P) * log(a[0]), a[1])])
    # unary maps
    _reg(f"Log_{S}", a0, lambda d, a: [(hdiv(d.A, a[0].P), a[0])])
    _reg(f"Log10_{S}", a0, lambda d, a: [(hdiv(d.A, a[0].P * _log10val(d)), a[0])])
    _reg(f"Exp_{S}", a0, lambda d, a: [(had(d.A, d.P), a[0])])
    _reg(f"Sin_{S}", a0, lambda d, a: [(had(d.A, cos(a[0].P)), a[0])])
    _reg(f"Cos_{S}", a0, lambda d, a: [(had(-d.A, sin(a[0].P)), a[0])])

    def tan_push(d, a):
        seca = one(d) / cos(a[0].P)
        return [(had(had(d.A, seca), seca), a[0])]
    _reg(f"Tan_{S}", a0, tan_push)
    _reg(f"Neg_{S}", a0, lambda d, a: [(-d.A, a[0])])
    _reg(f"Sqrt_{S}", a0, lambda d, a: [(hdiv(d.A, d.dtype.type(2) * d.P), a[0])])
    _reg(f"Sinh_{S}", a0, lambda d, a: [(had(d.A, cosh(a[0].P)), a[0])])
    _reg(f"Cosh_{S}", a0, lambda d, a: [(had(d.A, sinh(a[0].P)), a[0])])

    def tanh_push(d, a):  # :4055-4057, :4238-4240
        secha = one(d) / cosh(a[0].P)
        return [(had(had(d.A, secha), secha), a[0])]
    _reg(f"Tanh_{S}", a0, tanh_push)
    _reg(f"Asin_{S}", a0, lambda d, a: [(hdiv(d.A, sqrt(one(d) - had(a[0].P, a[0].P))), a[0])])
    _reg(f"Acos_{S}", a0, lambda d, a: [(hdiv(-d.A, sqrt(one(d) - had(a[0].P, a[0].P))), a[0])])
    _reg(f"Atan_{S}", a0, lambda d, a: [(hdiv(d.A, one(d) + had(a[0].P, a[0].P)), a[0])])
    _reg(f"Abs_{S}", a0, lambda d, a: [(had(d.A, Sign(a[0].P)), a[0])])
    for nm in ("Sign", "Floor", "Ceil", "Round"):
        _reg(f"{nm}_{S}", a0, lambda d, a: [(_zero_like(d), a[0])])
    _reg(f"ReLU_{S}", a0, lambda d, a: [(had(d.A, (Sign(a[0].P) + d.dtype.type(1)) / d.dtype.type(2)), a[0])])
    _reg(f"Sigmoid_{S}", a0, lambda d, a: [(had(had(d.A, d.P), d.dtype.type(1) - d.P), a[0])])


_reg_family("DV")
_reg_family("DM")

# ---- scalar nodes (:3866-3950) ----------------------------------------------------------------
_a0 = lambda a: [a[0]]            # noqa: E731
_both = lambda a: [a[0], a[1]]    # noqa: E731
_second = lambda a: [a[1]]        # noqa: E731
_one_d = lambda d: DNumber(1, d.dtype)  # noqa: E731
_reg("Add_D_D", _both, lambda d, a: [(d.A, a[0]), (d.A, a[1])])
_reg("Add_D_DCons", _a0, lambda d, a: [(d.A, a[0])])
_reg("Sub_D_D", _both, lambda d, a: [(d.A, a[0]), (-d.A, a[1])])
_reg("Sub_D_DCons", _a0, lambda d, a: [(d.A, a[0])])
_reg("Sub_DCons_D", _a0, lambda d, a: [(-d.A, a[0])])
_reg("Mul_D_D", _both, lambda d, a: [(d.A * a[1].P, a[0]), (d.A * a[0].P, a[1])])
_reg("Mul_D_DCons", _a0, lambda d, a: [(d.A * a[1], a[0])])
_reg("Div_D_D", _both, lambda d, a: [(d.A / a[1].P, a[0]), (d.A * (-a[0].P / (a[1].P * a[1].P)), a[1])])
_reg("Div_D_DCons", _a0, lambda d, a: [(d.A / a[1], a[0])])
_reg("Div_DCons_D", _second, lambda d, a: [(d.A * (-a[0] / (a[1].P * a[1].P)), a[1])])
_reg("Pow_D_D", _both, lambda d, a: [(d.A * (a[0].P ** (a[1].P - _one_d(d))) * a[1].P, a[0]),
                                     (d.A * (a[0].P ** a[1].P) * log(a[0].P), a[1])])
_reg("Pow_D_DCons", _a0, lambda d, a: [(d.A * (a[0].P ** (a[1] - _one_d(d))) * a[1], a[0])])
_reg("Pow_DCons_D", _second, lambda d, a: [(d.A * (a[0] ** a[1].P) * log(a[0]), a[1])])
_reg("Log_D", _a0, lambda d, a: [(d.A / a[0].P, a[0])])
_reg("Log10_D", _a0, lambda d, a: [(d.A / (a[0].P * _log10val(d)), a[0])])
_reg("Exp_D", _a0, lambda d, a: [(d.A * d.P, a[0])])
_reg("Sin_D", _a0, lambda d, a: [(d.A * cos(a[0].P), a[0])])
_reg("Cos_D", _a0, lambda d, a: [(d.A * (-sin(a[0].P)), a[0])])


def _tan_d_push(d, a):
    seca = _one_d(d) / cos(a[0].P)
    return [(d.A * seca * seca, a[0])]


def _tanh_d_push(d, a):
    secha = _one_d(d) / cosh(a[0].P)
    return [(d.A * secha * secha, a[0])]


_reg("Tan_D", _a0, _tan_d_push)
_reg("Neg_D", _a0, lambda d, a: [(-d.A, a[0])])
_reg("Sqrt_D", _a0, lambda d, a: [(d.A / (DNumber(2, d.dtype) * d.P), a[0])])
_reg("Sinh_D", _a0, lambda d, a: [(d.A * cosh(a[0].P), a[0])])
_reg("Cosh_D", _a0, lambda d, a: [(d.A * sinh(a[0].P), a[0])])
_reg("Tanh_D", _a0, _tanh_d_push)
_reg("Asin_D", _a0, lambda d, a: [(d.A / sqrt(_one_d(d) - a[0].P * a[0].P), a[0])])
_reg("Acos_D", _a0, lambda d, a: [(-d.A / sqrt(_one_d(d) - a[0].P * a[0].P), a[0])])
_reg("Atan_D", _a0, lambda d, a: [(d.A / (_one_d(d) + a[0].P * a[0].P), a[0])])
_reg("Abs_D", _a0, lambda d, a: [(d.A * Sign(a[0].P), a[0])])
for _nm in ("Sign_D", "Floor_D", "Ceil_D", "Round_D"):
    _reg(_nm, _a0, lambda d, a: [(DNumber(0, d.dtype), a[0])])
_reg("ReLU_D", _a0, lambda d, a: [(d.A * ((Sign(a[0].P) + d.dtype.type(1)) / d.dtype.type(2)), a[0])])
_reg("Sigmoid_D", _a0, lambda d, a: [(d.A * d.P * (d.dtype.type(1) - d.P), a[0])])
_reg("Mul_Dot_DV_DV", _both, lambda d, a: [(d.A * a[1].P, a[0]), (d.A * a[0].P, a[1])])
_reg("Mul_Dot_DV_DVCons", _a0, lambda d, a: [(d.A * a[1], a[0])])
# QUIRK: the fork's text is `DV.create a.Length d.A` with the backend argument missing (:3914) — a
# partial application that cannot be pushed; restated with the intended backend argument.
_reg("Sum_DV", _a0, lambda d, a: [(DVector(Backend(a[0]).CreateDataBuffer(Backend(a[0]).CreateValueArray(a[0].Length, d.A.Value))), a[0])])
_reg("L1Norm_DV", _a0, lambda d, a: [(d.A * Sign(a[0].P), a[0])])
_reg("L2NormSq_DV", _a0, lambda d, a: [(d.A * DNumber(2, d.dtype) * a[0].P, a[0])])
_reg("L2Norm_DV", _a0, lambda d, a: [((d.A / d.P) * a[0].P, a[0])])
_reg("Sum_DM", _a0, lambda d, a: [(DNDArray.create(a[0].Rows, a[0].Cols, d.A, Backend(a[0])), a[0])])  # :3920
_reg("Make_DM_ofDs", _a0, lambda d, a: [(d.A.Sum(), a[0])])


def _fixedpoint_push(d, a):  # :3927-3947
    b, bfirst, aprev, alast = a
    imax = GlobalConfig.FixedPointMaxIterations
    eps = DNumber(GlobalConfig.FixedPointEpsilon(d.dtype), d.dtype)
    i, r = 0, d.A
    reverseReset(alast)
    reversePush(r, alast)
    while i < imax:
        i += 1
        if i >= imax:
            pass
        elif abs(aprev.A + r - alast.A) <= eps:
            i = imax
        else:
            reverseReset(alast)
            reversePush(r + aprev.A, alast)
    return [(bfirst.A, b)]


_reg("FixedPoint_D", _a0, _fixedpoint_push)

# ---- vector / matrix structural nodes ------------------------------------------------------------
_reg("Mul_DM_DV", _both, lambda d, a: [(outer(d.A, a[1].P), a[0]), (a[0].P.Transpose() * d.A, a[1])])   # :3980
_reg("Mul_DM_DVCons", _a0, lambda d, a: [(outer(d.A, a[1]), a[0])])
_reg("Mul_DMCons_DV", _second, lambda d, a: [(a[0].Transpose() * d.A, a[1])])
_reg("Mul_DV_DM", _both, lambda d, a: [(d.A * a[1].P.Transpose(), a[0]), (outer(a[0].P, d.A), a[1])])   # :3984
_reg("Mul_DV_DMCons", _a0, lambda d, a: [(d.A * a[1].Transpose(), a[0])])
_reg("Mul_DVCons_DM", _second, lambda d, a: [(outer(a[0], d.A), a[1])])
_reg("Mul_Out_DV_DV", _both, lambda d, a: [(d.A * a[1].P, a[0]), (d.A.Transpose() * a[0].P, a[1])])       # :4152
_reg("Mul_Out_DV_DVCons", _a0, lambda d, a: [(d.A * a[1], a[0])])
_reg("Mul_Out_DVCons_DV", _second, lambda d, a: [(d.A.Transpose() * a[0], a[1])])


def _mul_dm_dm_push(d, a):  # :4127-4133
    pushLeft = d.A * a[1].P.Transpose()
    pushRight = a[0].P.Transpose() * d.A
    return [(pushLeft, a[0]), (pushRight, a[1])]


_reg("Mul_DM_DM", _both, _mul_dm_dm_push)
_reg("Mul_DM_DMCons", _a0, lambda d, a: [(d.A * a[1].Transpose(), a[0])])                                # :4134-4136
_reg("Mul_DMCons_DM", _second, lambda d, a: [(a[0].Transpose() * d.A, a[1])])                             # :4137-4143
_reg("Transpose_DM", _a0, lambda d, a: [(d.A.Transpose(), a[0])])
_reg("Permute_DM", _a0, lambda d, a: [(d.A.Permute(a[1]), a[0])])  # QUIRK App. D-11: re-applies dims, not the inverse
_reg("Reshape_DM", _a0, lambda d, a: [(d.A.Reshape(a[0].Buffer.Shape), a[0])])
_reg("ReshapeCopy_DM_DV", _a0, lambda d, a: [(d.A.ReshapeToDM(a[0].Rows), a[0])])
_reg("ReshapeCopy_DV_DM", _a0, lambda d, a: [(d.A.ReshapeToDV(), a[0])])


def _make_dmrows_ofdv_push(d, a):  # :4254-4257 (float32) / AD.Float64.fs:4264-4266 (row by row)
    v = a[0]
    if d.dtype == np.float32:
        Backend(d.A.Buffer).Add_M_Colwise_V_InPlace(d.A.Buffer, v.A.Buffer)
    else:
        for row in d.A.GetRows():
            v.A = v.A + row
    return [(DVector.Zero(d.dtype), v)]


_reg("Make_DMRows_ofDV", _a0, _make_dmrows_ofdv_push)
_reg("Make_DV_ofDs", lambda a: list(a[0]), lambda d, a: [(d.A[i], v) for i, v in enumerate(a[0])])


def _inverse_push(d, a):
    dpt = d.P.Transpose()
    return [(-dpt * d.A * dpt, a[0])]


_reg("Inverse_DM", _a0, _inverse_push)
_reg("Custom_DM", _a0, lambda d, a: [(DNDArray(Backend(a[0]).CustomOp_DM_Backward(a[0].P.Buffer, d.A.Buffer, d.P.Buffer, a[1])), a[0])])  # :4282
_reg("Diagonal_DM", _a0, None)   # needs AddDiagonal (not restated)
_reg("Noop", lambda a: [], lambda d, a: [])


def _solve_push(which):
    def push(d, a):
        A, b = a
        ba = A.Transpose().Solve(d.A)  # :4069-4077 (the node itself, not its primal, as in the reference)
        if which == "DM_DV":
            return [(outer(-ba, d.A), A), (ba, b)]
        if which == "DM_DVCons":
            return [(outer(-ba, d.A), A)]
        return [(ba, b)]
    return push


_reg("Solve_DM_DV", _both, _solve_push("DM_DV"))
_reg("Solve_DM_DVCons", _a0, _solve_push("DM_DVCons"))
_reg("Solve_DMCons_DV", _second, _solve_push("Cons_DV"))


# ================================================================================================
# reverse sweep (:3572-4292)
# ================================================================================================
def reverseReset(d):
    """Zero the adjoint of every node in the trace of `d` and count fan-out.  A node's inputs are
    visited only on its first visit (d.F = 1) but its adjoint is re-zeroed on every visit (:3746)."""
    stack = [d]
    while stack:
        d = stack.pop()  # LIFO == the reference's cons-on-front list
        if not isinstance(d, _AD) or d.kind != REVERSE:
            continue
        if isinstance(d, DNumber):
            d.a = DNumber(0, d.dtype)
        elif isinstance(d, DVector):
            backend = Backend(d.Buffer)
            d.a = DVector(backend.CreateDataBuffer(backend.CreateZeroArray(d.Buffer.Length)))
        else:
            d.a = DNDArray.ZeroMN(d.Rows, d.Cols, Backend(d))
        d.f += 1
        if d.f == 1:
            kids = _RESET[d.op.name](d.op.args)
            # resetRec (box a :: box b :: t): a is processed first => push in reverse onto the LIFO
            for k in reversed(kids):
                stack.append(k)


def reversePush(v, d):
    """Push adjoint `v` backwards through the trace of `d` (:3848-4287)."""
    stack = [(v, d)]
    while stack:
        v, d = stack.pop()
        if not isinstance(d, _AD) or d.kind != REVERSE:
            continue
        if isinstance(d, DNumber):
            d.a = d.a + v
        elif isinstance(d, DVector):
            d.a = d.a + v                                                      # allocating Add_V_V (:3956)
        else:
            d.a = DNDArray(Backend(d).Add_M_M_InPlace(d.a.Buffer, v.Buffer))   # in place (:4118)
        d.f -= 1
        if d.f == 0:
            push = _PUSH.get(d.op.name)
            if push is None:
                raise NotImplementedError(f"adjoint of {d.op.name} is not restated in this mirror")
            contributions = push(d, d.op.args)
            for c in reversed(contributions):
                stack.append(c)


def reverseProp(v, d):
    reverseReset(d)
    reversePush(v, d)


# ================================================================================================
# DiffOps (:4296-4379)
# ================================================================================================
def makeForward(i, t, p):
    return p.GetForward(t, i)


def makeReverse(i, p):
    return p.GetReverse(i)


def primal(d): return d.P
def tangent(d): return d.T
def adjoint(d): return d.A
def primalTangent(d): return d.P, d.T


def diff_(f, x):
    """``diff'`` (:4298)"""
    return primalTangent(f(makeForward(GlobalTagger.Next(), DNumber(1, x.dtype), x)))


def diff(f, x):
    return diff_(f, x)[1]


def grad_(f, x):
    """``grad'`` (:4334-4339): original value and gradient, reverse AD."""
    xa = makeReverse(GlobalTagger.Next(), x)
    z = f(xa)
    reverseReset(z)
    reversePush(DNumber(1, z.dtype), z)
    return primal(z), adjoint(xa)


def grad(f, x):
    return grad_(f, x)[1]


def jacobianv_(f, x, v):
    """``jacobianv'`` (:4345-4349): original value and Jacobian-vector product, forward AD."""
    return primalTangent(f(makeForward(GlobalTagger.Next(), v, x)))


def jacobianv(f, x, v):
    return jacobianv_(f, x, v)[1]


gradv, gradv_ = jacobianv, jacobianv_


def jacobianTv__(f, x):
    """``jacobianTv''`` (:4361-4371)"""
    xa = makeReverse(GlobalTagger.Next(), x)
    z = f(xa)
    r1 = primal(z)

    def r2(v):
        reverseReset(z)
        reversePush(v, z)
        return adjoint(xa)
    return r1, r2


def jacobianTv_(f, x, v):
    r1, r2 = jacobianTv__(f, x)
    return r1, r2(v)


def jacobianTv(f, x, v):
    return jacobianTv_(f, x, v)[1]


def gradhessianv_(f, x, v):
    """``gradhessianv'`` — commented out in the fork but both ingredients are live (:4416-4419):
    ``grad' (fun xx -> jacobianv f xx v) x``.  Reverse-on-forward."""
    gv, hv = grad_(lambda xx: jacobianv(f, xx, v), x)
    return f(x), gv, hv


def hessianv(f, x, v):
    return grad_(lambda xx: jacobianv(f, xx, v), x)[1]
