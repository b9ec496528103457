"""ctypes binding of libdiffsharp_cuda.so — the Python stand-in for the P/Invoke layer a managed
``Backend<'T>`` implementation would use (see INTEGRATION.md for the F# twin of this file).

The library is built in-tree (``sigma/diffsharp_b200/csrc/Makefile`` -> ``sigma/diffsharp_b200/
libdiffsharp_cuda.so``).  There is no CPU fallback: if the shared object is missing the import of
this module raises, and if no B200 is present ``dsc_create`` fails with DSC_E_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffsharp_cuda.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "diffsharp_cuda.h"))

DSC_OK = 0
DSC_E_INVALID, DSC_E_SHAPE, DSC_E_OOM, DSC_E_CUDA, DSC_E_NCCL = -1, -2, -3, -4, -5
DSC_E_UNSUPPORTED_OP, DSC_E_SINGULAR, DSC_E_DTYPE = -6, -7, -8
DSC_F32, DSC_F64 = 0, 1
DSC_FLAG_GEMM_FP32_SIMT, DSC_FLAG_GEMM_F64_SIMT, DSC_FLAG_NO_POOL, DSC_FLAG_NO_FUSE, DSC_FLAG_NVTX = 1, 2, 4, 8, 16


class DscError(RuntimeError):
    """Base of every error raised across the C ABI (status < 0)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[dsc {code}] {msg}")
        self.code = code


class InvalidArgError(DscError, ValueError):
    """.NET ``invalidArg`` twin: shape mismatches (reference Util.fs:117-131)."""


class NotSupportedError(DscError, NotImplementedError):
    """.NET ``NotSupportedException`` twin: op-code unknown to the device (closures never run)."""


class SingularMatrixError(DscError, ArithmeticError):
    """The managed side maps this to ``None`` (Backend.fs:102-103,125-126)."""


class DeviceError(DscError):
    """CUDA / NCCL / out-of-memory failures."""


_ERR_CLASS = {
    DSC_E_SHAPE: InvalidArgError,
    DSC_E_UNSUPPORTED_OP: NotSupportedError,
    DSC_E_SINGULAR: SingularMatrixError,
    DSC_E_OOM: DeviceError,
    DSC_E_CUDA: DeviceError,
    DSC_E_NCCL: DeviceError,
}


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "kernel_launches", "memcpy_h2d_bytes", "memcpy_d2h_bytes", "pool_bytes_reserved",
        "pool_bytes_in_use", "pool_hits", "pool_misses", "gemm_tcgen05", "gemm_dmma", "gemm_simt",
        "allreduce_p2p", "allreduce_nccl", "lane_ops")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def declared_symbols(header_path: str = HEADER_PATH) -> list[str]:
    """Every entry point include/diffsharp_cuda.h declares (typed macro expanded for s and d)."""
    text = open(header_path).read()
    # drop comments
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names: list[str] = []
    typed = re.search(r"#define DSC_DECLARE_TYPED\(P, T\)(.*?)\n\n", text, flags=re.S)
    typed_body = typed.group(1) if typed else ""
    for m in re.finditer(r"dsc_##P##(\w+)\s*\(", typed_body):
        names += [f"dsc_s{m.group(1)}", f"dsc_d{m.group(1)}"]
    rest = text.replace(typed_body, "")
    for m in re.finditer(r"DSC_API\s+[\w\s\*]+?\b(dsc_\w+)\s*\(", rest):
        names.append(m.group(1))
    return sorted(set(names))


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C sigma/diffsharp_b200/csrc` (or "
        "`python -c 'import __graft_entry__ as g; g.build()'`). This backend has no CPU fallback.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

ctx_t = C.c_void_p
buf_t = C.c_uint64
i32 = C.c_int32
P = C.POINTER


def _sig(name, *argtypes, restype=C.c_int):
    fn = getattr(lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = restype
    return fn


_sig("dsc_version")
_sig("dsc_device_count", P(C.c_int))
_sig("dsc_create", C.c_int, C.c_uint32, P(ctx_t))
_sig("dsc_destroy", ctx_t)
_sig("dsc_sync", ctx_t)
_sig("dsc_last_error", restype=C.c_char_p)
_sig("dsc_get_stats", ctx_t, P(Stats))
_sig("dsc_get_stream", ctx_t, P(C.c_void_p))
_sig("dsc_trim_pool", ctx_t)
_sig("dsc_set_fuse", ctx_t, C.c_int)
_sig("dsc_flush", ctx_t)
_sig("dsc_invalidate_caches", ctx_t)
_sig("dsc_event_create", ctx_t, P(C.c_uint64))
_sig("dsc_event_record", ctx_t, C.c_uint64)
_sig("dsc_event_elapsed_ms", ctx_t, C.c_uint64, C.c_uint64, P(C.c_float))
_sig("dsc_event_destroy", ctx_t, C.c_uint64)
_sig("dsc_event_record_copy", ctx_t, C.c_uint64)
_sig("dsc_event_record_copy_down", ctx_t, C.c_uint64)
_sig("dsc_event_sync", ctx_t, C.c_uint64)
_sig("dsc_compute_wait_event", ctx_t, C.c_uint64)
_sig("dsc_graph_begin", ctx_t)
_sig("dsc_graph_end", ctx_t, P(C.c_uint64))
_sig("dsc_graph_launch", ctx_t, C.c_uint64)
_sig("dsc_graph_destroy", ctx_t, C.c_uint64)
_sig("dsc_flush_l2", ctx_t)
_sig("dsc_gemm_tune", ctx_t, C.c_int, C.c_int, C.c_int, C.c_int)
_sig("dsc_gemm_epilogue", ctx_t, C.c_int)
_sig("dsc_gemm_debug_stamps", ctx_t, P(C.c_uint64))
_sig("dsc_gemm_plan", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P(i32))
_sig("dsc_buf_alloc", ctx_t, C.c_int, i32, P(buf_t))
_sig("dsc_buf_alloc_fill", ctx_t, C.c_int, i32, C.c_double, P(buf_t))
_sig("dsc_buf_free", buf_t)
_sig("dsc_buf_upload", buf_t, i32, C.c_void_p, i32)
_sig("dsc_buf_download", buf_t, i32, C.c_void_p, i32)
_sig("dsc_buf_upload_async", buf_t, i32, C.c_void_p, i32)
_sig("dsc_buf_download_async", buf_t, i32, C.c_void_p, i32)
_sig("dsc_buf_fill", buf_t, i32, i32, C.c_double)
_sig("dsc_copy_upload", buf_t, i32, C.c_void_p, i32)
_sig("dsc_copy_download", buf_t, i32, C.c_void_p, i32)
_sig("dsc_copy_wait_compute", ctx_t)
_sig("dsc_compute_wait_copy", ctx_t)
_sig("dsc_copy_sync", ctx_t)
_sig("dsc_buf_view", buf_t, i32, i32, P(buf_t))
_sig("dsc_buf_copy", buf_t, P(buf_t))
_sig("dsc_buf_share", buf_t, P(buf_t))
_sig("dsc_buf_gather2d", buf_t, i32, i32, i32, i32, i32, i32, P(buf_t))
_sig("dsc_buf_info", buf_t, P(C.c_int), P(i32), P(i32), P(i32), P(C.c_void_p))
_sig("dsc_host_alloc_pinned", C.c_size_t, P(C.c_void_p))
_sig("dsc_host_free_pinned", C.c_void_p)
_sig("dsc_comm_unique_id", C.c_void_p)
_sig("dsc_comm_init", ctx_t, C.c_int, C.c_int, C.c_void_p)
_sig("dsc_comm_join", ctx_t)
_sig("dsc_comm_debug_stamps", ctx_t, P(C.c_uint64))
_sig("dsc_comm_debug_stamps_at", ctx_t, C.c_int, P(C.c_uint64))
_sig("dsc_comm_destroy", ctx_t)

for _p, _T in (("s", C.c_float), ("d", C.c_double)):
    _sig(f"dsc_{_p}gemm", ctx_t, C.c_int, C.c_int, i32, i32, i32, buf_t, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}gemv", ctx_t, C.c_int, i32, i32, buf_t, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}ger", ctx_t, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}dot", ctx_t, buf_t, buf_t, P(_T))
    for _r in ("asum", "nrm2", "amax", "sum"):
        _sig(f"dsc_{_p}{_r}", ctx_t, buf_t, P(_T))
    _sig(f"dsc_{_p}argmax", ctx_t, buf_t, P(i32))
    _sig(f"dsc_{_p}argmin", ctx_t, buf_t, P(i32))
    _sig(f"dsc_{_p}sum_dev", ctx_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}dot_dev", ctx_t, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}map", ctx_t, C.c_int, buf_t, P(buf_t))
    _sig(f"dsc_{_p}map_s", ctx_t, C.c_int, _T, buf_t, P(buf_t))
    _sig(f"dsc_{_p}map2", ctx_t, C.c_int, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}scalar_op", ctx_t, C.c_int, _T, buf_t, P(buf_t))
    _sig(f"dsc_{_p}add_inplace", ctx_t, buf_t, buf_t)
    _sig(f"dsc_{_p}add_inplace_range", ctx_t, buf_t, i32, buf_t, i32, i32)
    _sig(f"dsc_{_p}fused_bwd", ctx_t, C.c_int, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}bias_rows_act", ctx_t, C.c_int, i32, i32, buf_t, buf_t, P(buf_t), C.c_void_p)
    _sig(f"dsc_{_p}transpose", ctx_t, i32, i32, buf_t, P(buf_t))
    _sig(f"dsc_{_p}permute", ctx_t, C.c_int, P(C.c_int64), P(i32), buf_t, P(buf_t))
    _sig(f"dsc_{_p}repeat_rows", ctx_t, i32, buf_t, P(buf_t))
    _sig(f"dsc_{_p}repeat_cols", ctx_t, i32, buf_t, P(buf_t))
    _sig(f"dsc_{_p}colsum_inplace", ctx_t, i32, i32, buf_t, buf_t)
    _sig(f"dsc_{_p}add_v_mcols", ctx_t, i32, i32, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}diag", ctx_t, i32, i32, buf_t, P(buf_t))
    _sig(f"dsc_{_p}gesv", ctx_t, i32, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}sysv", ctx_t, i32, buf_t, buf_t, P(buf_t))
    _sig(f"dsc_{_p}getri", ctx_t, i32, buf_t, P(buf_t))
    _sig(f"dsc_{_p}det", ctx_t, i32, buf_t, P(_T))
    _sig(f"dsc_{_p}allreduce_sum", ctx_t, i32, P(buf_t))
    _sig(f"dsc_{_p}allgather", ctx_t, buf_t, buf_t)


def last_error() -> str:
    return (lib.dsc_last_error() or b"").decode(errors="replace")


def check(status: int) -> None:
    """Turn a negative status into the matching exception (error behaviour of the managed shim)."""
    if status == DSC_OK:
        return
    raise _ERR_CLASS.get(status, DscError)(status, last_error())
