"""Generates tests/golden/golden.npz.

The reference cannot be built or run in this image (F#/.NET) and holds no Backend<'T>
known-answer vectors, so the fixtures are:
  * the reference's own constants: tests/DiffSharp.Tests/Script.fsx:7-16 (RepeatReshapeCopy_V_MRows)
    and tests/DiffSharp.Tests/AD.Float32.fs:47-57 (FixedPoint sqrt(25) = 5, d/db = 0.1);
  * frozen outputs of the CPU oracle (oracle/, OpenBLAS 0.3.31.dev from the scipy wheel) on seeded
    inputs, for every Backend<'T> op and for configs C1 / C4 / C5 at small sizes — these pin the
    oracle against drift and give the GPU path size-independent known answers.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.backend import OracleBackend  # noqa: E402
from sigma.diffsharp_b200 import workloads as wl  # noqa: E402
from sigma.diffsharp_b200.backend import MapOp  # noqa: E402
from sigma.diffsharp_b200.buffers import NativeDataBuffer, ShapedDataBufferView  # noqa: E402
from sigma.diffsharp_b200.config import DictBackendProvider, GlobalConfig  # noqa: E402

UNARY = [m for m in MapOp if m >= MapOp.Log]


def op_inputs(dtype, seed=7, n=37, m=5, k=7):
    rng = np.random.default_rng(seed)
    return {
        "x": rng.standard_normal(n).astype(dtype),
        "y": rng.standard_normal(n).astype(dtype),
        "pos": (rng.random(n) + 0.25).astype(dtype),       # log / sqrt / pow domain
        "unit": (rng.random(n) * 1.8 - 0.9).astype(dtype),  # asin / acos domain
        "A": rng.standard_normal((m, k)).astype(dtype),
        "B": rng.standard_normal((k, 6)).astype(dtype),
        "S": (rng.standard_normal((6, 6)) + 6 * np.eye(6)).astype(dtype),
        "vm": rng.standard_normal(m).astype(dtype),
        "vk": rng.standard_normal(k).astype(dtype),
        "v6": rng.standard_normal(6).astype(dtype),
        # drawn after everything above, so the earlier fixtures keep their values
        "Sym": (lambda g: (g + g.T) / 2 + np.diag([3, -3, 2.5, -2, 4, -4]))(rng.standard_normal((6, 6))).astype(dtype),   # symmetric INDEFINITE
    }


class ScaledSquare:
    """A consumer-defined custom op (the `customInfo : obj` payload of CustomOp_DM_Forward / _Backward, Backend.fs:137-139;
    call sites AD.Float32.fs:1760-1763, 4282): y = s * a .* a, adjoint = 2 s * origin .* adjoint — written against the
    backend's own members, so it runs on whichever backend it is handed."""

    def __init__(self, s):
        self.s = s

    def forward(self, backend, a):
        return backend.Mul_S_M(backend.dtype.type(self.s), backend.Mul_Had_M_M(a, a))

    def backward(self, backend, origin, adjoint, primal):
        return backend.Mul_S_M(backend.dtype.type(2 * self.s), backend.Mul_Had_M_M(origin, adjoint))


def run_ops(b, inp):
    """Every Backend<'T> op on the fixed inputs -> dict of host arrays."""
    buf = lambda a: b.CreateDataBuffer(np.ascontiguousarray(a).reshape(-1))  # noqa: E731
    view = lambda a: ShapedDataBufferView(buf(a), *a.shape)  # noqa: E731
    out = {}
    x, y, pos, unit = inp["x"], inp["y"], inp["pos"], inp["unit"]
    for op in UNARY:
        src = pos if op in (MapOp.Log, MapOp.Log10, MapOp.Sqrt) else unit if op in (MapOp.Asin, MapOp.Acos) else x
        out[f"map_{op.name}"] = b.Map_F_V(op, None, buf(src)).SubData
    s = b.dtype.type(1.75)
    for op in (MapOp.Div, MapOp.Pow_Ab, MapOp.Pow_Ba, MapOp.Atan2_Ab, MapOp.Atan2_Ba):
        out[f"map_s_{op.name}"] = b.Map_F_S_V(s, op, None, buf(pos)).SubData
    for op in (MapOp.Mul, MapOp.Div, MapOp.Pow_Ab, MapOp.Atan2_Ab):
        out[f"map2_{op.name}"] = b.Map2_F_V_V(op, None, buf(pos), buf(np.abs(y) + 0.5)).SubData
    out["add"] = b.Add_V_V(buf(x), buf(y)).SubData
    out["sub"] = b.Sub_V_V(buf(x), buf(y)).SubData
    out["add_s"] = b.Add_S_V(s, buf(x)).SubData
    out["sub_s_v"] = b.Sub_S_V(s, buf(x)).SubData
    out["sub_v_s"] = b.Sub_V_S(buf(x), s).SubData
    out["mul_s"] = b.Mul_S_V(s, buf(x)).SubData
    out["dot"] = np.array([b.Mul_Dot_V_V(buf(x), buf(y))])
    out["l1"] = np.array([b.L1Norm_V(buf(x))])
    out["l2"] = np.array([b.L2Norm_V(buf(x))])
    out["sup"] = np.array([b.SupNorm_V(buf(x))])
    out["sum"] = np.array([b.Sum_V(buf(x))])
    out["argmax"] = np.array([b.MaxIndex_V(buf(x))])
    out["argmin"] = np.array([b.MinIndex_V(buf(x))])
    A, B, S = inp["A"], inp["B"], inp["S"]
    out["gemm"] = b.Mul_M_M(view(A), view(B)).DataBuffer.SubData
    out["gemm_bias"] = b.Mul_M_M_Add_V_MCols(view(A), view(B), buf(inp["vm"])).DataBuffer.SubData
    out["gemv"] = b.Mul_M_V(view(A), buf(inp["vk"])).SubData
    out["gemv_add"] = b.Mul_M_V_Add_V(view(A), buf(inp["vk"]), buf(inp["vm"])).SubData
    out["gemv_t"] = b.Mul_V_M(buf(inp["vm"]), view(A)).SubData
    out["ger"] = b.Mul_Out_V_V(buf(inp["vm"]), buf(inp["vk"])).DataBuffer.SubData
    out["had"] = b.Mul_Had_M_M(view(A), view(A)).DataBuffer.SubData
    out["transpose"] = b.Transpose_M(view(A)).DataBuffer.SubData
    out["repeat_rows"] = b.RepeatReshapeCopy_V_MRows(3, buf(inp["vk"])).DataBuffer.SubData
    out["repeat_cols"] = b.RepeatReshapeCopy_V_MCols(3, buf(inp["vk"])).DataBuffer.SubData
    out["add_v_mcols"] = b.Add_V_MCols(buf(inp["vm"]), view(A)).DataBuffer.SubData
    acc = buf(inp["vk"])
    b.Add_M_Colwise_V_InPlace(view(A), acc)
    out["colsum"] = acc.SubData
    out["diag"] = b.Diagonal_M(view(A)).SubData
    out["solve"] = b.Solve_M_V(view(S), buf(inp["v6"])).SubData
    out["inverse"] = b.Inverse_M(view(S)).DataBuffer.SubData
    out["det"] = np.array([b.Det_M(view(S))])
    out["solve_symmetric"] = b.SolveSymmetric_M_V(view(inp["Sym"]), buf(inp["v6"])).SubData
    out["reshape_copy_mrows_v"] = b.ReshapeCopy_MRows_V(view(A)).SubData                         # Backend.fs:109
    rc = b.ReshapeCopy_V_MRows(5, buf(np.arange(35, dtype=b.dtype)))                             # Backend.fs:134
    out["reshape_copy_v_mrows"] = np.concatenate([np.array(rc.Shape, b.dtype), rc.DataBuffer.SubData])
    rs = b.Reshape_M(view(A), [7, 5])                                                            # Backend.fs:129: same data, new shape
    out["reshape"] = np.concatenate([np.array(rs.Shape, b.dtype), rs.DataBuffer.SubData[:35], b.Transpose_M(rs).DataBuffer.SubData[:35]])
    cop = ScaledSquare(0.75)
    fw = b.CustomOp_DM_Forward(view(A), cop)                                                     # Backend.fs:137
    out["custom_forward"] = fw.DataBuffer.SubData
    out["custom_backward"] = b.CustomOp_DM_Backward(view(A), view(A * 0.5 + 1), fw, cop).DataBuffer.SubData   # Backend.fs:138-139
    T3 = np.arange(24, dtype=b.dtype).reshape(2, 3, 4)
    out["permute"] = b.Permute_M(ShapedDataBufferView(buf(T3), 2, 3, 4), [2, 0, 1]).DataBuffer.SubData
    return out


C3_LAYERS, C3_BATCH = [784, 1024, 1024, 10], 8192


def c3_full_step():
    """config 3 at its full size (global batch 8192) on whichever backend is installed: (loss, [dW1, dW2, dW3, db1, db2, db3])."""
    X, Y, Ws, bs = wl.mlp_inputs(C3_LAYERS, C3_BATCH)
    L, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    return float(L), [a.ToArray().reshape(-1) for a in dW + db]


def c3_loss_float64():
    """The config-3 loss in float64 NumPy from the same inputs (independent of every backend).  The loss is a sum of 81 920
    same-sign terms: a sequential float32 sum (what the CPU oracle does) is itself ~1e-5 off, so both the oracle's and the
    GPU's loss are compared with THIS number."""
    X, Y, Ws, bs = wl.mlp_inputs(C3_LAYERS, C3_BATCH)
    H = X.astype(np.float64)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        H = H @ W.astype(np.float64) + b.astype(np.float64)
        if i < len(Ws) - 1:
            H = np.tanh(H)
    E = H - Y
    return float(np.sum(E * E))


def sample_index(n, count=64):
    """`count` positions spread over a buffer of n elements (first, last and a fixed stride in between)."""
    return np.unique(np.linspace(0, n - 1, min(count, n)).astype(np.int64))


def c3_digest(loss, grads):
    """What bench.py and the N-GPU tests compare the all-reduced adjoints with ("gradients equal the 1-GPU
    result <= 1e-5"): per adjoint buffer its L2 norm, its first 16 elements and 64 samples spread over the buffer."""
    g = {"c3_loss": np.array([loss]), "c3_loss64": np.array([c3_loss_float64()])}
    for i, v in enumerate(grads):
        g[f"c3_grad{i}_l2"] = np.array([np.sqrt(np.sum(v.astype(np.float64) ** 2))])
        g[f"c3_grad{i}_head"] = v[:16].copy()
        g[f"c3_grad{i}_samp"] = v[sample_index(v.size)].copy()
    return g


def c3_digest_max_rel(loss, grads, gold):
    """max over the digest entries of |got - gold| / scale (scale = the buffer's max magnitude among the stored samples /
    its L2 norm / the loss)."""
    worst = abs(loss - float(gold["c3_loss64"][0])) / abs(float(gold["c3_loss64"][0])) / 2   # the loss against float64, bar 2e-5: see c3_loss_float64
    for i, v in enumerate(grads):
        l2 = np.sqrt(np.sum(v.astype(np.float64) ** 2))
        worst = max(worst, abs(l2 - float(gold[f"c3_grad{i}_l2"][0])) / float(gold[f"c3_grad{i}_l2"][0]))
        ref = np.concatenate([gold[f"c3_grad{i}_head"], gold[f"c3_grad{i}_samp"]]).astype(np.float64)
        got = np.concatenate([v[:16], v[sample_index(v.size)]]).astype(np.float64)
        scale = max(np.abs(ref).max(), float(gold[f"c3_grad{i}_l2"][0]) / np.sqrt(v.size))
        worst = max(worst, float(np.abs(got - ref).max() / scale))
    return worst


def config_goldens():
    g = {}
    X, Y, Ws, bs = wl.mlp_inputs([784, 300, 10], 128)
    L, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    g["c1_loss"] = np.array([L.Value])
    for i, a in enumerate(dW + db):
        v = a.ToArray().reshape(-1)
        g[f"c1_grad{i}_head"] = v[:16].copy()
        g[f"c1_grad{i}_l2"] = np.array([np.sqrt(np.sum(v.astype(np.float64) ** 2))])
    g.update(c3_digest(*c3_full_step()))
    W1, b1, W2, b2, x = wl.jacobian_inputs(96)
    g["c4_rows"] = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), 16, 32).ToArray()
    x, v, d = wl.hvp_inputs(4096)
    g["c5_hv"] = wl.hvp(wl.to_DV(d), wl.to_DV(x), wl.to_DV(v)).ToArray()
    return g


def main():
    GlobalConfig.BackendProvider = DictBackendProvider(OracleBackend(np.float32), OracleBackend(np.float64))
    g = {
        # tests/DiffSharp.Tests/Script.fsx:7-16
        "script_fsx_v": np.array([-4.99094, -0.34702, 5.98291, -6.16668]),
        "script_fsx_m": np.array([2]),
        # tests/DiffSharp.Tests/AD.Float32.fs:47-57
        "fixedpoint_expected": np.array([5.0, 0.1]),
    }
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        for k, v in run_ops(OracleBackend(dt), op_inputs(dt)).items():
            g[f"{name}_{k}"] = v
    g.update(config_goldens())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **g)
    print(f"wrote {path}: {len(g)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
