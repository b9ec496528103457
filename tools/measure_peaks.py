#!/usr/bin/env python
"""Measures the GEMM roofline denominators that MEASURED_PEAKS.json lacks: cuBLAS TF32 and
FP64 GEMM at 8192^3 through torch.matmul (library call, used ONLY as a yardstick, never on the
product path), plus a device-to-device copy for HBM.  Writes gpurun_out/gemm_peaks.json; commit a copy
under profiles/gemm_peaks.json."""
import json
import os
import sys

import torch


def bench(fn, iters):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = 8192
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "n": n}
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    ms = bench(lambda: torch.matmul(a, b), 10)
    out["tf32_tflops"] = 2 * n ** 3 / ms / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    ms = bench(lambda: torch.matmul(a, b), 5)
    out["fp32_cublas_tflops"] = 2 * n ** 3 / ms / 1e9
    a4, b4 = a[:4096, :4096].contiguous(), b[:4096, :4096].contiguous()
    ms = bench(lambda: torch.matmul(a4, b4), 5)
    out["fp32_cublas_tflops_4096"] = 2 * 4096 ** 3 / ms / 1e9
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = bench(lambda: torch.matmul(a4, b4), 10)
    out["tf32_tflops_4096"] = 2 * 4096 ** 3 / ms / 1e9
    del a, b, a4, b4
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    ms = bench(lambda: torch.matmul(a, b), 3)
    out["fp64_tflops"] = 2 * n ** 3 / ms / 1e9
    a4, b4 = a[:4096, :4096].contiguous(), b[:4096, :4096].contiguous()
    ms = bench(lambda: torch.matmul(a4, b4), 5)
    out["fp64_tflops_4096"] = 2 * 4096 ** 3 / ms / 1e9
    del a, b, a4, b4
    x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    y = torch.empty_like(x)
    ms = bench(lambda: y.copy_(x), 10)
    out["hbm_copy_gbs"] = 2 * x.numel() * 2 / ms / 1e6
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_peaks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
