"""sigma.diffsharp_b200 — B200-native compute backend behind Sigma.DiffSharp's ``Backend<'T>``.

Layout
  csrc/        hand-written sm_100a kernels + the C ABI (include/diffsharp_cuda.h)
  _lib.py      ctypes binding of libdiffsharp_cuda.so (the P/Invoke stand-in)
  buffers.py   ISigmaDiffDataBuffer / NativeDataBuffer / ShapedDataBufferView mirrors (host side)
  interface.py MapOp, Backend (55 members), ConstArray — no device code
  device.py    CudaContext, CudaDataBuffer
  backend.py   CudaBackend: one C-ABI call per Backend<'T> member
  config.py    GlobalConfig / IBackendProvider / CudaBackendProvider
  ad.py        mirror of the reference's AD layer (the caller of the hot path)
  workloads.py the five BASELINE.json configurations, written in the fork's idiom
  dist.py      batch-row / tangent-column sharding over NCCL

Importing the device pieces (CudaBackend, CudaDataBuffer, CudaContext, CudaBackendProvider — resolved on first use)
requires the built shared object; there is no CPU fallback.  The interface, the AD mirror and the workloads import without it.
"""
from .config import GlobalConfig, IBackendProvider, BackendConfig, DictBackendProvider  # noqa: F401
from .buffers import ISigmaDiffDataBuffer, NativeDataBuffer, ShapedDataBufferView  # noqa: F401
from .interface import MapOp, Backend  # noqa: F401

_DEVICE = {"CudaDataBuffer": "device", "CudaContext": "device", "CudaBackend": "backend", "CudaBackendProvider": "config"}


def __getattr__(name):
    if name in _DEVICE:
        import importlib
        return getattr(importlib.import_module(f".{_DEVICE[name]}", __name__), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
