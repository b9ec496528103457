"""The plugin interface of the hot path, free of any device code (importing this module never loads libdiffsharp_cuda.so):

  * ``MapOp``      the op-codes of ``DiffSharp.Backend.MapOp`` (reference src/DiffSharp/Backend.fs:42-70), declaration order;
  * ``Backend``    the 55 abstract members of ``DiffSharp.Backend.Backend<'T>`` (Backend.fs:74-139);
  * ``ConstArray`` what ``CreateZeroArray`` / ``CreateValueArray`` / ``CreateUninitialisedArray`` hand back.

``CudaBackend`` (backend.py) implements it over the C ABI; the CPU oracle (oracle/backend.py, test infrastructure) implements
it over OpenBLAS; the AD mirror (ad.py) calls it.
"""
from __future__ import annotations

import enum

import numpy as np


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
