"""sigma.diffsharp_b200 — B200-native compute backend behind Sigma.DiffSharp's ``Backend<'T>``.

Layout
  csrc/        hand-written sm_100a kernels + the C ABI (include/diffsharp_cuda.h)
  _lib.py      ctypes binding of libdiffsharp_cuda.so (the P/Invoke stand-in)
  buffers.py   ISigmaDiffDataBuffer / NativeDataBuffer / ShapedDataBufferView mirrors + CudaDataBuffer
  backend.py   MapOp, Backend (55 members), CudaBackend
  config.py    GlobalConfig / IBackendProvider / CudaBackendProvider
  ad.py        mirror of the reference's AD layer (the caller of the hot path)
  workloads.py the five BASELINE.json configurations, written in the fork's idiom
  dist.py      batch-row / tangent-column sharding over NCCL

Importing the device pieces requires the built shared object; there is no CPU fallback.
"""
from .config import GlobalConfig, IBackendProvider, BackendConfig, DictBackendProvider  # noqa: F401
from .buffers import (ISigmaDiffDataBuffer, NativeDataBuffer, ShapedDataBufferView,  # noqa: F401
                      CudaDataBuffer, CudaContext)
from .backend import MapOp, Backend, CudaBackend  # noqa: F401
from .config import CudaBackendProvider  # noqa: F401
