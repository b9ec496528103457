"""Backend dispatch: mirror of src/DiffSharp/Config.fs.

  * ``BackendConfig``      Config.fs:43-50  (record: handle + epsilons)
  * ``IBackendProvider``   Config.fs:53-54  (``GetBackend<'T>(obj)``)
  * ``GlobalConfig``       Config.fs:63-110 (process-wide static; provider is a mutable static)

The reference's ``DefaultBackendProvider`` hands out null handles (and swaps the precisions,
Config.fs:118-119), so a consumer must install its own provider; ``CudaBackendProvider`` is that
provider for this backend.  ``GetBackend`` is called once per primitive op with whatever the call
site has in hand — a buffer, a view, a boxed scalar, a DNumber/DVector/DNDArray —
so the provider keys on the value's dtype only.
"""
from __future__ import annotations

import numpy as np


class BackendConfig:
    def __init__(self, handle, dtype):
        eps, fpeps = 0.00001, 0.01
        t = np.dtype(dtype).type
        self.BackendHandle = handle
        self.Epsilon = t(eps)
        self.EpsilonRec = t(1) / t(eps)
        self.EpsilonRec2 = t(0.5) / t(eps)
        self.FixedPointEpsilon = t(fpeps)
        self.FixedPointMaxIterations = 100
        self.VisualisationContrast = t(1.2)


class IBackendProvider:
    def GetBackend(self, value) -> BackendConfig:
        raise NotImplementedError


def dtype_of(value) -> np.dtype:
    """dtype of anything a call site may pass to ``Backend(x)`` (AD.Float32.fs:70)."""
    dt = getattr(value, "dtype", None)
    if dt is not None:
        return np.dtype(dt)
    if isinstance(value, float):
        return np.dtype(np.float64)
    raise TypeError("Unsupported type, try float32 or float")  # Config.fs:120


class DictBackendProvider(IBackendProvider):
    """Provider over explicit per-precision backends (used for the CUDA backend and, in tests, for
    the CPU oracle backend)."""

    def __init__(self, float32_backend=None, float64_backend=None):
        self._cfg = {}
        if float32_backend is not None:
            self._cfg[np.dtype(np.float32)] = BackendConfig(float32_backend, np.float32)
        if float64_backend is not None:
            self._cfg[np.dtype(np.float64)] = BackendConfig(float64_backend, np.float64)

    def GetBackend(self, value) -> BackendConfig:
        try:
            return self._cfg[dtype_of(value)]
        except KeyError:
            raise TypeError("Unsupported type, try float32 or float") from None


class CudaBackendProvider(DictBackendProvider):
    """One GPU context shared by a float32 and a float64 ``CudaBackend``."""

    def __init__(self, device: int = 0, flags: int = 0, fuse: bool = True, ctx=None):
        """`fuse=False` creates the context with DSC_FLAG_NO_FUSE: the library launches every call's own kernel at once
        (the deferred evaluation of csrc/dsc_lazy.cu is a property of the context, not of this shim)."""
        from . import _lib
        from .backend import CudaBackend
        from .device import CudaContext
        if ctx is None:
            ctx = CudaContext(device, flags | (0 if fuse else _lib.DSC_FLAG_NO_FUSE))
        self.ctx = ctx
        super().__init__(CudaBackend(np.float32, ctx=self.ctx), CudaBackend(np.float64, ctx=self.ctx))


class _NullProvider(IBackendProvider):
    """DefaultBackendProvider: handles are null until a consumer installs a provider (Config.fs:71,79)."""

    def GetBackend(self, value):
        raise RuntimeError("GlobalConfig.BackendProvider has not been set: the reference ships no built-in backend "
                           "(Config.fs:108-110 'FSharp backends not supported.')")


class _GlobalConfig:
    FixedPointMaxIterations = 100
    GrayscalePalette = [" ", ".", ":", "x", "T", "Y", "V", "X", "H", "N", "M"]

    def __init__(self):
        self._provider: IBackendProvider = _NullProvider()

    @property
    def BackendProvider(self) -> IBackendProvider:
        return self._provider

    @BackendProvider.setter
    def BackendProvider(self, value: IBackendProvider):
        self._provider = value

    def FixedPointEpsilon(self, dtype):
        return np.dtype(dtype).type(0.01)

    @staticmethod
    def SetDefaultBackend(backend: str):
        raise ValueError("FSharp backends not supported.")  # Config.fs:108-110


GlobalConfig = _GlobalConfig()
