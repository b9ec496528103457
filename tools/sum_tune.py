import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
import bench
from sigma.diffsharp_b200.backend import CudaBackend
from sigma.diffsharp_b200.buffers import CudaContext
ctx = CudaContext(0, 8)  # 8 = DSC_FLAG_NO_FUSE: every call launches its own kernel(s)
for dt in (np.float32, np.float64):
    b = CudaBackend(dt, ctx=ctx)
    host = np.random.default_rng(0).random(4096 * 4096).astype(dt)
    As = [b.CreateDataBuffer(np.roll(host, i)) for i in range(4)]
    ms = bench.time_graph(ctx, [(lambda i=i: b._out(b._f["sum_dev"], 1, As[i % 4].handle)) for i in range(12)])
    print(os.environ.get("DSC_RED_MAXP"), np.dtype(dt).name, f"{ms*1e3:.2f} us", f"{host.nbytes/ms/1e6:.0f} GB/s", flush=True)
