"""Parity of every Backend<'T> member: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Tolerances (north_star): bit-exact for Transpose / reshape / indexing / broadcast
copies; relative error <= 1e-5 (float32) and <= 1e-12 (float64) elsewhere, measured norm-wise
(max |gpu - ref| / max |ref|); GEMM-class ops against the backward-error bound rtol * (|A| |B|)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402
from _util import rel_err  # noqa: E402

from oracle.backend import OracleBackend  # noqa: E402
from sigma.diffsharp_b200 import _lib  # noqa: E402
from sigma.diffsharp_b200.backend import MapOp  # noqa: E402
from sigma.diffsharp_b200.buffers import CudaDataBuffer, NativeDataBuffer, ShapedDataBufferView  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
EXACT = ("argmax", "argmin", "transpose", "repeat_rows", "repeat_cols", "diag", "permute",
         "reshape_copy_mrows_v", "reshape_copy_v_mrows", "reshape")
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


def backend(provider, dtype):
    return provider.GetBackend(np.dtype(dtype).type(0)).BackendHandle


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_op_against_oracle_and_goldens(cuda_provider, dtype):
    b = backend(cuda_provider, dtype)
    inp = mg.op_inputs(dtype)
    got = mg.run_ops(b, inp)
    ref = mg.run_ops(OracleBackend(dtype), inp)
    name = "f32" if dtype == np.float32 else "f64"
    for k in ref:
        if k in EXACT:
            assert np.array_equal(got[k], ref[k]), k
            assert np.array_equal(got[k], GOLD[f"{name}_{k}"]), k
        else:
            scale = 100 if k in ("solve", "solve_symmetric", "inverse", "det") else 1  # conditioning of the 6x6 solve
            assert rel_err(got[k], ref[k]) <= RTOL[np.dtype(dtype)] * scale, k
            assert rel_err(got[k], GOLD[f"{name}_{k}"]) <= RTOL[np.dtype(dtype)] * scale, k


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 31, 1000, 4097, 1 << 20])
def test_elementwise_sizes_and_misaligned_views(cuda_provider, dtype, n):
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n + 3).astype(dtype)
    y = rng.standard_normal(n + 3).astype(dtype)
    for off in (0, 1, 3):  # views at odd offsets are not 16-byte aligned: scalar path
        dx = b.CreateDataBuffer(x).GetValues(off, n)
        dy = b.CreateDataBuffer(y).GetValues(off, n)
        hx, hy = NativeDataBuffer(x, off, n), NativeDataBuffer(y, off, n)
        for op in (MapOp.Tanh, MapOp.Exp, MapOp.Cosh, MapOp.Sigmoid, MapOp.Abs, MapOp.Sign, MapOp.ReL):
            assert rel_err(b.Map_F_V(op, None, dx).SubData, o.Map_F_V(op, None, hx).SubData) <= RTOL[np.dtype(dtype)], (op, off)
        assert np.array_equal(b.Add_V_V(dx, dy).SubData, o.Add_V_V(hx, hy).SubData)
        assert np.array_equal(b.Sub_V_V(dx, dy).SubData, o.Sub_V_V(hx, hy).SubData)
        assert np.array_equal(b.Map2_F_V_V(MapOp.Mul, None, dx, dy).SubData, o.Map2_F_V_V(MapOp.Mul, None, hx, hy).SubData)
        assert np.array_equal(b.Mul_S_V(dtype(0.3), dx).SubData, o.Mul_S_V(dtype(0.3), hx).SubData)
        # reductions: compare both to an exact float64 sum
        exact = float(np.sum(x[off:off + n].astype(np.float64)))
        bound = RTOL[np.dtype(dtype)] * float(np.sum(np.abs(x[off:off + n].astype(np.float64))))
        assert abs(float(b.Sum_V(dx)) - exact) <= bound
        assert abs(float(b.Mul_Dot_V_V(dx, dy)) - float(x[off:off + n].astype(np.float64) @ y[off:off + n].astype(np.float64))) <= RTOL[np.dtype(dtype)] * float(np.abs(x[off:off + n].astype(np.float64)) @ np.abs(y[off:off + n].astype(np.float64)))
        assert abs(float(b.L2Norm_V(dx)) - float(np.linalg.norm(x[off:off + n].astype(np.float64)))) <= RTOL[np.dtype(dtype)] * float(np.linalg.norm(x[off:off + n].astype(np.float64)))
        assert float(b.SupNorm_V(dx)) == float(np.abs(x[off:off + n]).max())
        assert b.MaxIndex_V(dx) == int(np.argmax(x[off:off + n])) and b.MinIndex_V(dx) == int(np.argmin(x[off:off + n]))


def test_reductions_are_deterministic(cuda_provider):
    b = backend(cuda_provider, np.float32)
    x = b.CreateDataBuffer(np.random.default_rng(0).standard_normal(3_000_001).astype(np.float32))
    vals = {float(b.Sum_V(x)) for _ in range(5)}
    assert len(vals) == 1


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_edge_semantics_match_oracle(cuda_provider, dtype):
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    sp = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 0.5, 1.5, 2.5, -0.5, -1.5, 1e-40, -3.0], dtype)
    for op in (MapOp.Sign, MapOp.Round, MapOp.ReL, MapOp.Abs, MapOp.Floor, MapOp.Ceiling):
        g = b.Map_F_V(op, None, b.CreateDataBuffer(sp)).SubData
        r = o.Map_F_V(op, None, NativeDataBuffer(sp)).SubData
        assert np.array_equal(g, r, equal_nan=True), op
    # empty operands: foreign zero-length host buffers reach the backend
    x = b.CreateDataBuffer(np.array([1, 2, 3], dtype))
    empty = NativeDataBuffer(np.empty(0, dtype))
    assert np.array_equal(b.Add_V_V(x, empty).SubData, [1, 2, 3])
    assert np.array_equal(b.Add_V_V(empty, x).SubData, [1, 2, 3])
    assert np.array_equal(b.Sub_V_V(empty, x).SubData, [-1, -2, -3])
    m = ShapedDataBufferView(b.CreateDataBuffer(np.ones(6, dtype)), 2, 3)
    z = ShapedDataBufferView(empty, 0, 0)
    assert np.array_equal(b.Add_M_M_InPlace(m, z).DataBuffer.SubData, np.ones(6))
    assert b.Sum_V(empty) == 0 and b.Mul_Dot_V_V(empty, empty) == 0
    # foreign (host) buffers are uploaded on first touch (App. D-2)
    host = NativeDataBuffer(np.array([4, 5, 6], dtype))
    assert np.array_equal(b.Add_V_V(x, host).SubData, [5, 7, 9])
    # first occurrence; a[0] = NaN pins index 0 like the sequential twin
    assert b.MaxIndex_V(b.CreateDataBuffer(np.array([1, 5, 5, 2], dtype))) == 1
    assert b.MaxIndex_V(b.CreateDataBuffer(np.array([np.nan, 5], dtype))) == 0
    # singular -> None, closures never executed
    zz = ShapedDataBufferView(b.CreateDataBuffer(np.zeros(4, dtype)), 2, 2)
    assert b.Solve_M_V(zz, b.CreateDataBuffer(np.ones(2, dtype))) is None and b.Inverse_M(zz) is None
    with pytest.raises(_lib.NotSupportedError):
        b.Map_F_V(MapOp.Mul, lambda v: v * 2, x)
    with pytest.raises(_lib.NotSupportedError):
        b.Map_F_V(99, lambda v: v, x)
    with pytest.raises(_lib.InvalidArgError):
        b.Add_V_V(x, b.CreateDataBuffer(np.ones(2, dtype)))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (33, 65), (64, 64), (100, 257), (1024, 768), (10, 8192)])
def test_transpose_bit_exact_and_involutive(cuda_provider_nofuse, dtype, shape):
    b = backend(cuda_provider_nofuse, dtype)
    a = np.random.default_rng(1).standard_normal(shape).astype(dtype)
    v = ShapedDataBufferView(b.CreateDataBuffer(a.reshape(-1)), *shape)
    t = b.Transpose_M(v)
    assert t.Shape == (shape[1], shape[0])
    assert np.array_equal(t.to_numpy(), a.T)
    assert np.array_equal(b.Transpose_M(t).to_numpy(), a)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (5, 7, 3), (128, 784, 300), (128, 300, 10), (300, 128, 10), (257, 129, 65),
                                   (512, 512, 512), (1024, 96, 2048), (64, 4096, 64), (2048, 512, 10), (512, 4100, 10), (777, 333, 16), (2048, 10, 512), (300, 16, 4096), (129, 7, 1000)])
def test_gemm_all_transposes_against_openblas(cuda_provider, dtype, m, k, n):
    """Mul_M_M and the transposed-operand forms of dsc_?gemm (what Transpose_M + Mul_M_M computes)."""
    import ctypes as C
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    rng = np.random.default_rng(m * 1000 + n)
    A = (rng.random((m, k)) * 2 - 1).astype(dtype)
    B = (rng.random((k, n)) * 2 - 1).astype(dtype)
    bias = rng.standard_normal(m).astype(dtype)
    bound = RTOL[np.dtype(dtype)] * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64))
    for ta in (0, 1):
        for tb in (0, 1):
            As = np.ascontiguousarray(A.T if ta else A)
            Bs = np.ascontiguousarray(B.T if tb else B)
            ref = o.gemm(ta, tb, m, n, k, As.reshape(-1), Bs.reshape(-1), bias).reshape(m, n).astype(np.float64)
            dA, dB, dbias = b.CreateDataBuffer(As.reshape(-1)), b.CreateDataBuffer(Bs.reshape(-1)), b.CreateDataBuffer(bias)
            out = b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, dbias.handle)
            got = out.SubData.reshape(m, n).astype(np.float64)
            assert np.all(np.abs(got - ref) <= bound + 1e-30), (ta, tb, float(np.abs(got - ref).max()), float(bound.min()))
            exact = A.astype(np.float64) @ B.astype(np.float64) + bias.astype(np.float64)[:, None]
            assert np.all(np.abs(got - exact) <= bound + RTOL[np.dtype(dtype)] * np.abs(bias).max())


@pytest.mark.parametrize("bn,pair,split,fuse", [(256, 0, -1, 0), (128, 0, -1, 0), (128, 0, -1, 1), (64, 0, -1, 1), (64, 0, 2, 1),
                                                (256, 1, -1, 0), (256, 1, 2, 0), (128, 1, -1, 0), (64, 1, -1, 0), (256, 1, -1, 2), (256, 1, 4, 2)])
@pytest.mark.parametrize("m,k,n", [(512, 1024, 512), (300, 520, 260), (1024, 96, 2048)])
def test_f32_tensor_core_gemm_variants(cuda_provider, bn, pair, split, fuse, m, k, n):
    """Every plan of the tcgen05 kernel (tile width, 1-CTA / 2-CTA cta_group::2, split-K, in-kernel operand
    split) pinned with dsc_gemm_tune, all four operand layouts, against OpenBLAS sgemm; the same plan twice
    gives the same bits (deterministic fix-up order)."""
    from sigma.diffsharp_b200._lib import check, lib
    b, o = backend(cuda_provider, np.float32), OracleBackend(np.float32)
    rng = np.random.default_rng(m + n + bn)
    A = (rng.random((m, k)) * 2 - 1).astype(np.float32)
    B = (rng.random((k, n)) * 2 - 1).astype(np.float32)
    bias = rng.standard_normal(m).astype(np.float32)
    bound = RTOL[np.dtype(np.float32)] * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64))
    st0 = cuda_provider.ctx.stats()
    check(lib.dsc_gemm_tune(cuda_provider.ctx.handle, bn, pair, split, fuse))
    try:
        for ta in (0, 1):
            for tb in (0, 1):
                As = np.ascontiguousarray(A.T if ta else A)
                Bs = np.ascontiguousarray(B.T if tb else B)
                ref = o.gemm(ta, tb, m, n, k, As.reshape(-1), Bs.reshape(-1), bias).reshape(m, n).astype(np.float64)
                dA, dB, dbias = b.CreateDataBuffer(As.reshape(-1)), b.CreateDataBuffer(Bs.reshape(-1)), b.CreateDataBuffer(bias)
                got = [b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, dbias.handle).SubData.reshape(m, n) for _ in range(2)]
                assert np.array_equal(got[0], got[1])
                assert np.all(np.abs(got[0].astype(np.float64) - ref) <= bound + 1e-30), (ta, tb, float(np.abs(got[0] - ref).max()))
    finally:
        check(lib.dsc_gemm_tune(cuda_provider.ctx.handle, -1, -1, -1, -1))
    st1 = cuda_provider.ctx.stats()
    assert st1["gemm_tcgen05"] - st0["gemm_tcgen05"] == 8     # none of the calls fell back to the CUDA-core kernel


def _plan(m, n, k, bn=-1, pair=-1, split=-1, fuse=-1, sms=148):
    import ctypes as C
    out = (C.c_int32 * 20)()
    _lib.check(_lib.lib.dsc_gemm_plan(sms, m, n, k, 1, 1, bn, pair, split, fuse, out))
    keys = "elig bn pair workers cpt kb tm tn dp_waves dp_extra sk_units perm_s base rem dp_tiles perm_tiles kc a_f b_f".split()
    return dict(zip(keys, list(out)))


# (m, n, k, tune, how the partial tiles of the plan are combined) — the modes are asserted against dsc_gemm_plan below
DEEP_K = [
    (1024, 1024, 8192, (-1, -1, -1, -1), "workspace"),   # the weight-adjoint product of config 3 as the planner runs it
    (1024, 784, 8192, (-1, -1, -1, -1), "workspace"),    # dW1^T shape (n = 784)
    (1024, 1024, 8192, (128, 0, 2, 1), "cluster"),       # K ranges of a tile = the 2 CTAs of a cluster (DSM exchange), in-kernel split
    (1024, 1024, 8192, (256, 1, 2, 0), "cluster"),       # 2 CTA pairs per cluster
    (1024, 1024, 8192, (256, 0, -1, 0), "cluster"),      # 4 single CTAs per cluster
    (1024, 1024, 8192, (256, 1, 1, 0), "whole"),         # one accumulation chain of 32 chunks per tile, no split
    (2048, 1280, 8192, (-1, -1, -1, -1), "streamk"),     # a whole-tile wave + stream-K ranges over the rest
]


@pytest.mark.parametrize("m,n,k,tune,mode", DEEP_K)
def test_f32_tensor_core_gemm_deep_k_same_sign(cuda_provider, m, n, k, tune, mode):
    """K = 8192 with SAME-SIGN data — the case where a single TMEM accumulation chain was 9.1e-5 off (the tensor core's
    fp32 accumulator truncates) and the reason K is cut into 256-element chunks: every way of combining the chunks
    (registers only, cluster DSM exchange, workspace slots, stream-K fix-up) against OpenBLAS sgemm and float64, in the
    layouts config 3 uses (NN for forward, TN for the weight adjoints, NT for the input adjoints)."""
    from sigma.diffsharp_b200._lib import check, lib
    b, o = backend(cuda_provider, np.float32), OracleBackend(np.float32)
    pl = _plan(m, n, k, *tune)
    assert pl["elig"] == 1 and pl["cpt"] == 32
    got_mode = ("whole" if pl["sk_units"] == 0 or (pl["perm_s"] == 0 and pl["workers"] == pl["tm"] * pl["tn"]) else
                "cluster" if pl["kc"] > 1 else "workspace" if pl["perm_s"] > 1 else "streamk")
    assert got_mode == mode, pl
    rng = np.random.default_rng(k + m + n)
    A = rng.random((m, k)).astype(np.float32)            # all positive
    B = rng.random((k, n)).astype(np.float32)
    exact = A.astype(np.float64) @ B.astype(np.float64)
    st0 = cuda_provider.ctx.stats()
    check(lib.dsc_gemm_tune(cuda_provider.ctx.handle, *tune))
    try:
        for ta, tb in ((0, 0), (1, 0), (0, 1)):
            As = np.ascontiguousarray(A.T if ta else A)
            Bs = np.ascontiguousarray(B.T if tb else B)
            dA, dB = b.CreateDataBuffer(As.reshape(-1)), b.CreateDataBuffer(Bs.reshape(-1))
            got = b._out(b._f["gemm"], m * n, ta, tb, m, n, k, dA.handle, dB.handle, 0).SubData.reshape(m, n).astype(np.float64)
            ref = o.gemm(ta, tb, m, n, k, As.reshape(-1), Bs.reshape(-1)).reshape(m, n).astype(np.float64)
            # same-sign data: |A||B| = A B, so the bound is simply 1e-5 relative, element by element
            assert np.all(np.abs(got - ref) <= 1e-5 * exact), (ta, tb, float((np.abs(got - ref) / exact).max()))
            assert np.all(np.abs(got - exact) <= 1e-5 * exact), (ta, tb, float((np.abs(got - exact) / exact).max()))
    finally:
        check(lib.dsc_gemm_tune(cuda_provider.ctx.handle, -1, -1, -1, -1))
    assert cuda_provider.ctx.stats()["gemm_tcgen05"] - st0["gemm_tcgen05"] == 3


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 17, 64, 200])
def test_lapack_class_ops(cuda_provider, dtype, n):
    """Solve_M_V / SolveSymmetric_M_V / Inverse_M / Det_M (Backend.fs:102-103,125-126) against LAPACK ?gesv / ?sysv / ?getri /
    ?getrf through the oracle, beyond the 6 x 6 fixture: residual-based bounds (the problems are well conditioned by
    construction), `None` for singular input."""
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    rng = np.random.default_rng(100 + n)
    G = rng.standard_normal((n, n))
    A = (G / n + np.eye(n)).astype(dtype)                                   # well conditioned, |det| = O(1)
    S = ((G + G.T) / (2 * n) + np.diag(np.where(np.arange(n) % 2 == 0, 1.5, -1.5))).astype(dtype)   # symmetric INDEFINITE
    rhs = rng.standard_normal(n).astype(dtype)
    eps = np.finfo(dtype).eps
    for M, name in ((A, "Solve_M_V"), (S, "SolveSymmetric_M_V")):
        dv = ShapedDataBufferView(b.CreateDataBuffer(M.reshape(-1)), n, n)
        hv = ShapedDataBufferView(NativeDataBuffer(M.reshape(-1)), n, n)
        xg = getattr(b, name)(dv, b.CreateDataBuffer(rhs)).SubData.astype(np.float64)
        xr = getattr(o, name)(hv, NativeDataBuffer(rhs)).SubData.astype(np.float64)
        cond = np.linalg.cond(M.astype(np.float64))
        assert np.abs(xg - xr).max() <= 50 * n * eps * cond * np.abs(xr).max(), name
        assert np.abs(M.astype(np.float64) @ xg - rhs).max() <= 50 * n * eps * (np.abs(M).astype(np.float64) @ np.abs(xg)).max(), name
    dv, hv = ShapedDataBufferView(b.CreateDataBuffer(A.reshape(-1)), n, n), ShapedDataBufferView(NativeDataBuffer(A.reshape(-1)), n, n)
    Ig, Ir = b.Inverse_M(dv).to_numpy().astype(np.float64), o.Inverse_M(hv).to_numpy().astype(np.float64)
    assert np.abs(Ig - Ir).max() <= 50 * n * eps * np.linalg.cond(A.astype(np.float64)) * np.abs(Ir).max()
    dg, dr = float(b.Det_M(dv)), float(o.Det_M(hv))
    assert abs(dg - dr) <= 50 * n * eps * abs(dr)
    if n > 1:
        Z = A.copy(); Z[-1] = Z[0]                                            # two equal rows: singular
        zv = ShapedDataBufferView(b.CreateDataBuffer(Z.reshape(-1)), n, n)
        assert b.Solve_M_V(zv, b.CreateDataBuffer(rhs)) is None or dtype == np.float32   # exact zero pivot in float64; float32 may round
        assert abs(float(b.Det_M(zv))) <= 1e3 * n * eps * abs(dr)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_custom_op_through_the_ad_layer(cuda_provider, oracle_provider, use_provider, dtype):
    """DNDArray.CustomOp (AD.Float32.fs:1758-1763) and its adjoint Custom_DM (:4282): the consumer hook is forwarded with the
    backend, the primal / adjoint buffers and the opaque payload, in primal, forward and reverse mode; the same program on
    the CUDA backend and on the oracle, and against the same function written with native operators."""
    from sigma.diffsharp_b200 import ad, workloads as wl
    rng = np.random.default_rng(3)
    X = rng.standard_normal((9, 7)).astype(dtype)
    W = rng.standard_normal((7, 5)).astype(dtype)
    op = mg.ScaledSquare(0.75)
    res = {}
    for name, prov in (("cuda", cuda_provider), ("oracle", oracle_provider)):
        use_provider(prov)
        dX, dW = wl.to_DM(X), wl.to_DM(W)
        f_custom = lambda w: (dX * w).CustomOp(op).Sum()                              # noqa: E731
        f_native = lambda w: (ad.had(dX * w, dX * w) * ad.DNumber(0.75, dtype)).Sum()  # noqa: E731
        p1, g1 = ad.grad_(f_custom, dW)
        p2, g2 = ad.grad_(f_native, dW)
        pf, tf = ad.jacobianv_(lambda w: (dX * w).CustomOp(op), dW, wl.to_DM(np.ones_like(W)))
        res[name] = (float(p1), g1.ToArray(), float(p2), g2.ToArray(), pf.ToArray(), tf.ToArray())
    tol = RTOL[np.dtype(dtype)]
    c, r = res["cuda"], res["oracle"]
    assert abs(c[0] - r[0]) <= tol * abs(r[0]) and abs(c[0] - c[2]) <= 10 * tol * abs(c[2])
    assert rel_err(c[1], r[1]) <= tol and rel_err(c[1], c[3]) <= 10 * tol
    assert rel_err(c[4], r[4]) <= tol and rel_err(c[5], r[5]) <= tol
    assert rel_err(c[4], 0.75 * (X.astype(np.float64) @ W.astype(np.float64)) ** 2) <= 10 * tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gemv_ger_colsum_larger(cuda_provider, dtype):
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    rng = np.random.default_rng(9)
    for m, n in [(1, 1), (7, 13), (1000, 10), (10, 1000), (2048, 1024), (513, 1027)]:
        A = rng.standard_normal((m, n)).astype(dtype)
        x, y = rng.standard_normal(n).astype(dtype), rng.standard_normal(m).astype(dtype)
        dv = ShapedDataBufferView(b.CreateDataBuffer(A.reshape(-1)), m, n)
        hv = ShapedDataBufferView(NativeDataBuffer(A.reshape(-1)), m, n)
        tol = RTOL[np.dtype(dtype)]
        assert rel_err(b.Mul_M_V(dv, b.CreateDataBuffer(x)).SubData, o.Mul_M_V(hv, NativeDataBuffer(x)).SubData) <= tol
        assert rel_err(b.Mul_V_M(b.CreateDataBuffer(y), dv).SubData, o.Mul_V_M(NativeDataBuffer(y), hv).SubData) <= tol
        assert np.array_equal(b.Mul_Out_V_V(b.CreateDataBuffer(y), b.CreateDataBuffer(x)).to_numpy(), np.outer(y, x).astype(dtype))
        acc_d, acc_h = b.CreateDataBuffer(x), NativeDataBuffer(x.copy())
        b.Add_M_Colwise_V_InPlace(dv, acc_d)
        o.Add_M_Colwise_V_InPlace(hv, acc_h)
        exact = x.astype(np.float64) + A.astype(np.float64).sum(0)
        bound = tol * (np.abs(x).astype(np.float64) + np.abs(A).astype(np.float64).sum(0))
        assert np.all(np.abs(acc_d.SubData.astype(np.float64) - exact) <= bound)
        assert np.array_equal(b.RepeatReshapeCopy_V_MRows(m, b.CreateDataBuffer(x)).to_numpy(), np.tile(x, (m, 1)))
        assert np.array_equal(b.RepeatReshapeCopy_V_MCols(n, b.CreateDataBuffer(y)).to_numpy(), np.tile(y[:, None], (1, n)))


def test_buffer_interface(cuda_provider):
    """ISigmaDiffDataBuffer members on the device buffer (Util.fs:133-141)."""
    b = backend(cuda_provider, np.float32)
    a = np.arange(24, dtype=np.float32)
    buf = b.CreateDataBuffer(a)
    assert isinstance(buf, CudaDataBuffer) and buf.Length == 24 and buf.Offset == 0
    v = buf.GetValues(6, 5)                 # view: same storage
    assert v.Offset == 6 and np.array_equal(v.SubData, a[6:11])
    b.Add_V_V_InPlace(b.CreateDataBuffer(np.ones(5, np.float32)), 0, buf, 6, 5)
    assert np.array_equal(v.SubData, a[6:11] + 1)   # the view sees the in-place update
    assert np.array_equal(v.Data[v.Offset:v.Offset + v.Length], a[6:11] + 1)
    d = buf.DeepCopy()
    b.Add_V_V_InPlace(b.CreateDataBuffer(np.ones(5, np.float32)), 0, buf, 0, 5)
    assert np.array_equal(d.SubData[:5], a[:5])      # deep copy is detached
    s = ShapedDataBufferView(b.CreateDataBuffer(a), 4, 6).GetSlice(1, 2, 2, 4)   # GetStackedValues
    assert s.Shape == (2, 3) and np.array_equal(s.to_numpy(), a.reshape(4, 6)[1:3, 2:5])
    z = b.CreateDataBuffer(b.CreateZeroArray(1000))
    assert np.array_equal(z.SubData, np.zeros(1000, np.float32))
    w = b.CreateDataBuffer(b.CreateValueArray(1000, 2.5))
    assert np.array_equal(w.SubData, np.full(1000, 2.5, np.float32))
    st = b.ctx.stats()
    assert st["kernel_launches"] > 0 and st["pool_hits"] + st["pool_misses"] > 0


def test_copy_on_write_share(cuda_provider):
    """dsc_buf_share: DeepCopy semantics without the copy until somebody writes in place; true views
    (GetValues) keep sharing mutable storage (Util.fs:137, AD.Float32.fs:663)."""
    import ctypes as C
    b = backend(cuda_provider, np.float32)
    x = b.CreateDataBuffer(np.arange(8, dtype=np.float32))
    h = _lib.buf_t()
    _lib.check(_lib.lib.dsc_buf_share(x.handle, C.byref(h)))
    y = CudaDataBuffer(b.ctx, np.float32, h.value, 8)
    one = b.CreateDataBuffer(np.ones(8, np.float32))
    b.Add_V_V_InPlace(one, 0, y, 0, 8)                      # write through the alias: it detaches first
    assert np.array_equal(x.SubData, np.arange(8)) and np.array_equal(y.SubData, np.arange(8) + 1)
    _lib.check(_lib.lib.dsc_buf_share(x.handle, C.byref(h)))
    z = CudaDataBuffer(b.ctx, np.float32, h.value, 8)
    b.Add_V_V_InPlace(one, 0, x, 0, 8)                      # write through the original: it detaches
    assert np.array_equal(z.SubData, np.arange(8)) and np.array_equal(x.SubData, np.arange(8) + 1)
    v = x.GetValues(2, 3)                                   # a true view still sees in-place updates
    b.Add_V_V_InPlace(one, 0, x, 0, 8)
    assert np.array_equal(v.SubData, np.arange(2, 5) + 2)
    _lib.check(_lib.lib.dsc_buf_share(x.handle, C.byref(h)))  # a block with views is copied eagerly
    w = CudaDataBuffer(b.ctx, np.float32, h.value, 8)
    b.Add_V_V_InPlace(one, 0, x, 0, 8)
    assert np.array_equal(w.SubData, np.arange(8) + 2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("m,n", [(1, 1), (7, 10), (33, 64), (500, 1024)])
def test_fused_bias_activation_matches_the_unfused_sequence(cuda_provider, cuda_provider_nofuse, dtype, m, n):
    """RepeatReshapeCopy_V_MRows -> Add_M_M -> Map_F_M issued lazily as one pass == the three eager calls, bit for bit."""
    rng = np.random.default_rng(m * n)
    A, v = rng.standard_normal((m, n)).astype(dtype), rng.standard_normal(n).astype(dtype)
    outs = []
    for prov in (cuda_provider, cuda_provider_nofuse):
        b = backend(prov, dtype)
        dA = ShapedDataBufferView(b.CreateDataBuffer(A.reshape(-1)), m, n)
        rows = b.RepeatReshapeCopy_V_MRows(m, b.CreateDataBuffer(v))
        z = b.Add_M_M(dA, rows)
        hs = [b.Map_F_M(op, None, z).to_numpy() for op in (MapOp.Tanh,)]
        outs.append((z.to_numpy(), hs[0], b.Add_M_M(rows, dA).to_numpy(), rows.to_numpy()))
    for f, u in zip(outs[0], outs[1]):
        assert np.array_equal(f, u)
    assert np.array_equal(outs[0][3], np.tile(v, (m, 1)))


# ---- deferred evaluation inside the library (csrc/dsc_lazy.cu) ---------------------------------------------------------
def _act_loss_grads(act, X, Y, W, b):
    """grad of Sum(E .* E), E = act(X W + rows(b)) - Y, with respect to W and b, through the AD mirror on the installed backend"""
    from sigma.diffsharp_b200 import ad, workloads as wl
    dX, dY = wl.to_DM(X), wl.to_DM(Y)

    def f(Wb):
        Wm, bv = Wb
        Z = dX * Wm + ad.DNDArray.OfRows(dX.Rows, bv, ad.Backend(dX))
        E = act(Z) - dY
        return ad.had(E, E).Sum()
    i = ad.GlobalTagger.Next()
    Wr, br = ad.makeReverse(i, wl.to_DM(W)), ad.makeReverse(i, wl.to_DV(b))
    L = f((Wr, br))
    ad.reverseProp(ad.DNumber(1, L.dtype), L)
    return float(L), Wr.A.ToArray(), br.A.ToArray()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("actname", ["tanh", "sigmoid", "reLU"])
@pytest.mark.parametrize("shape", [(37, 19, 23), (256, 96, 64)])
def test_adjoint_chains_fused_equal_unfused(cuda_provider, cuda_provider_nofuse, oracle_provider, use_provider, dtype, actname, shape):
    """The tanh (AD.Float32.fs:4238-4240), sigmoid (:4281) and ReLU (:4280) adjoint chains, the bias broadcast + activation and
    had(E, E).Sum() issued by the UNCHANGED call sequence: recognised inside the library and run as single kernels, bit for bit
    what the same sequence gives with DSC_FLAG_NO_FUSE, and within tolerance of the oracle.  The second shape reaches the
    float32 tensor-core GEMM (plan pinned so that both contexts accumulate in the same order); there the epilogues are also
    forced INTO the kernel (dsc_gemm_epilogue 2)."""
    from sigma.diffsharp_b200 import ad
    from sigma.diffsharp_b200._lib import check, lib
    act = getattr(ad, actname)
    m, k, n = shape
    rng = np.random.default_rng(m + n)
    X = rng.standard_normal((m, k)).astype(dtype)
    Y = rng.standard_normal((m, n)).astype(dtype)
    W = (rng.standard_normal((k, n)) * 0.3).astype(dtype)
    b = rng.standard_normal(n).astype(dtype)
    use_provider(oracle_provider)
    ref = _act_loss_grads(act, X, Y, W, b)
    outs = {}
    for name, prov, epi in (("fused", cuda_provider, 1), ("fused_epilogue_in_kernel", cuda_provider, 2), ("unfused", cuda_provider_nofuse, 1)):
        check(lib.dsc_gemm_tune(prov.ctx.handle, 128, 0, 1, 0))
        check(lib.dsc_gemm_epilogue(prov.ctx.handle, epi))
        try:
            use_provider(prov)
            st0 = prov.ctx.stats()["kernel_launches"]
            outs[name] = _act_loss_grads(act, X, Y, W, b)
            outs[name + "_launches"] = prov.ctx.stats()["kernel_launches"] - st0
        finally:
            check(lib.dsc_gemm_tune(prov.ctx.handle, -1, -1, -1, -1))
            check(lib.dsc_gemm_epilogue(prov.ctx.handle, 1))
    tol = RTOL[np.dtype(dtype)]
    for name in ("fused", "fused_epilogue_in_kernel"):
        assert outs[name][0] == outs["unfused"][0], name
        assert np.array_equal(outs[name][1], outs["unfused"][1]) and np.array_equal(outs[name][2], outs["unfused"][2]), name
    assert abs(outs["fused"][0] - ref[0]) <= 10 * tol * abs(ref[0])      # a sum of squares: both sides carry their own summation error
    assert rel_err(outs["fused"][1], ref[1]) <= 10 * tol and rel_err(outs["fused"][2], ref[2]) <= 10 * tol
    assert outs["fused_launches"] < outs["unfused_launches"]              # the chains really ran as fewer kernels


def test_deferred_values_behave_like_values(cuda_provider, cuda_provider_nofuse):
    """A handle whose value is deferred is indistinguishable from a computed one: any consumer materialises it; an in-place
    write to an operand AFTER the deferring call does not leak into the deferred result; handles that end up sharing one
    result do not see each other's writes; pending work runs at the next synchronisation point."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((48, 20)).astype(np.float32)
    v = rng.standard_normal(20).astype(np.float32)
    one = np.ones(48 * 20, np.float32)
    got = []
    for prov in (cuda_provider, cuda_provider_nofuse):
        b = backend(prov, np.float32)
        dA = ShapedDataBufferView(b.CreateDataBuffer(A.reshape(-1)), 48, 20)
        T = b.Transpose_M(dA)                                        # deferred
        C = b.Map_F_M(MapOp.Cosh, None, dA)                          # deferred (head of the tanh chain)
        S = b.Map_F_S_M(np.float32(1), MapOp.Div, None, C)           # deferred: 1 / cosh
        R = b.RepeatReshapeCopy_V_MRows(48, b.CreateDataBuffer(v))   # deferred
        Z = b.Add_M_M(dA, R)                                         # deferred: A + rows(v)
        H = b.Mul_Had_M_M(dA, dA)                                    # deferred
        b.Add_V_V_InPlace(b.CreateDataBuffer(one), 0, dA.DataBuffer, 0, 48 * 20)   # A += 1 AFTER the calls above
        res = [T.to_numpy(), C.to_numpy(), S.to_numpy(), R.to_numpy(), Z.to_numpy(), H.to_numpy(), dA.to_numpy(),
               T.DataBuffer.GetValues(20, 5).SubData, float(b.Sum_M(H.DataBuffer)), float(b.Sum_M(b.Mul_Had_M_M(dA, dA).DataBuffer))]
        # zero adjoints: the first accumulation aliases the addend, later writes to either side stay private
        z = ShapedDataBufferView(b.CreateDataBuffer(b.CreateZeroArray(48 * 20)), 48, 20)
        b.Add_M_M_InPlace(z, dA)
        b.Add_M_M_InPlace(z, dA)                                     # z = 2 (A + 1): must not change A
        ones = ShapedDataBufferView(b.CreateDataBuffer(b.CreateValueArray(48 * 20, 1)), 48, 20)
        e1 = b.Mul_Had_M_M(ones, dA)                                 # 1 .* A: an alias of A
        b.Add_M_M_InPlace(e1, dA)                                    # written in place: A itself must not change
        res += [z.to_numpy(), e1.to_numpy(), dA.to_numpy()]
        got.append(res)
    for f, u in zip(got[0], got[1]):
        assert np.array_equal(f, u)
    f = got[0]
    assert np.array_equal(f[0], A.T) and np.array_equal(f[3], np.tile(v, (48, 1))) and np.array_equal(f[4], A + v) and np.array_equal(f[5], A * A)
    assert np.array_equal(f[6], A + 1) and np.array_equal(f[7], A.T.reshape(-1)[20:25])
    assert np.array_equal(f[10], 2 * (A + 1)) and np.array_equal(f[11], 2 * (A + 1)) and np.array_equal(f[12], A + 1)
    # pending compute work runs at the next synchronisation point, not never
    b = backend(cuda_provider, np.float32)
    dA = ShapedDataBufferView(b.CreateDataBuffer(A.reshape(-1)), 48, 20)
    cuda_provider.ctx.sync()
    n0 = cuda_provider.ctx.stats()["kernel_launches"]
    keep = b.Mul_Had_M_M(dA, dA)
    assert cuda_provider.ctx.stats()["kernel_launches"] == n0
    cuda_provider.ctx.sync()
    assert cuda_provider.ctx.stats()["kernel_launches"] == n0 + 1
    assert np.array_equal(keep.to_numpy(), A * A)


def test_copy_on_write_counts_live_aliases(cuda_provider):
    """Regression (round-1 advisor): share, free the alias, take a TRUE view, write — the view must keep sharing storage with
    its parent (copy-on-write only while aliases are alive); and dsc_copy_upload honours aliases."""
    import ctypes as C
    b = backend(cuda_provider, np.float32)
    x = b.CreateDataBuffer(np.arange(8, dtype=np.float32))
    h = _lib.buf_t()
    _lib.check(_lib.lib.dsc_buf_share(x.handle, C.byref(h)))
    _lib.check(_lib.lib.dsc_buf_free(h.value))                     # the alias is gone again
    v = x.GetValues(2, 3)
    one = b.CreateDataBuffer(np.ones(8, np.float32))
    b.Add_V_V_InPlace(one, 0, x, 0, 8)
    assert np.array_equal(v.SubData, np.arange(2, 5) + 1)          # written through the parent, seen by the view
    b.Add_V_V_InPlace(one, 0, v, 0, 3)
    assert np.array_equal(x.SubData, np.arange(8) + 1 + np.array([0, 0, 1, 1, 1, 0, 0, 0]))   # and the other way round
    y = b.CreateDataBuffer(np.arange(4, dtype=np.float32))
    _lib.check(_lib.lib.dsc_buf_share(y.handle, C.byref(h)))
    alias = CudaDataBuffer(b.ctx, np.float32, h.value, 4)
    pin = np.full(4, 7, np.float32)
    _lib.check(_lib.lib.dsc_copy_upload(y.handle, 0, pin.ctypes.data, 4))
    _lib.check(_lib.lib.dsc_copy_sync(b.ctx.handle))
    assert np.array_equal(alias.SubData, np.arange(4)) and np.array_equal(y.SubData, pin)


def test_host_mirror_of_a_device_buffer(cuda_provider):
    """ISigmaDiffDataBuffer.Data on the device buffer (Util.fs:136): a cached host array shared by the views of one allocation;
    host-side writes (`aa.Data.[i] <- ...`) reach the device before its next use; device-side writes drop the mirror."""
    b = backend(cuda_provider, np.float32)
    a = np.arange(12, dtype=np.float32)
    buf = b.CreateDataBuffer(a)
    d = buf.Data
    assert d is buf.Data and np.array_equal(d, a)                  # cached: the same array object, no second download
    d[3] = 100                                                     # host-side write
    assert np.array_equal(b.Mul_S_V(np.float32(2), buf).SubData, 2 * np.where(np.arange(12) == 3, 100, a))
    view = buf.GetValues(2, 4)
    assert view.Data is d and view.Offset == 2                     # the whole underlying array, indexed at Offset
    view.Data[view.Offset + 1] = -5                                # element 3 again, through the view
    assert float(b.Sum_V(view)) == 2 - 5 + 4 + 5
    b.Add_V_V_InPlace(b.CreateDataBuffer(np.ones(12, np.float32)), 0, buf, 0, 12)   # device-side write
    assert buf.Data is not d and buf.Data[3] == -4 and np.array_equal(buf.SubData, buf.Data)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_l2norm_at_the_extremes(cuda_provider, dtype):
    """L2Norm_V where sqrt(sum x^2) over- or underflows but ?nrm2 (scaled) does not: same answers as OpenBLAS through the oracle."""
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    big, small = (1e30, 1e-30) if dtype == np.float32 else (1e200, 1e-200)
    rng = np.random.default_rng(8)
    base = rng.standard_normal(5001).astype(dtype)
    for scale in (big, small, 1.0):
        x = (base * dtype(scale)).astype(dtype)
        g = float(b.L2Norm_V(b.CreateDataBuffer(x)))
        r = float(o.L2Norm_V(NativeDataBuffer(x)))
        exact = float(np.linalg.norm(base.astype(np.float64))) * scale
        assert np.isfinite(g) and g > 0
        assert abs(g - r) <= 4 * RTOL[np.dtype(dtype)] * r and abs(g - exact) <= 4 * RTOL[np.dtype(dtype)] * exact
    assert float(b.L2Norm_V(b.CreateDataBuffer(np.zeros(7, dtype)))) == 0.0
    assert np.isinf(float(b.L2Norm_V(b.CreateDataBuffer(np.array([1, np.inf, 2], dtype)))))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,perm", [((2, 3, 4), (2, 0, 1)), ((5, 7), (1, 0)), ((4, 1, 6, 3), (3, 1, 0, 2)), ((3, 4, 5, 2, 2), (4, 0, 3, 1, 2)),
                                        ((64, 33), (0, 1)), ((17, 9, 11), (1, 2, 0)), ((128, 96, 3), (2, 1, 0))])
def test_permute_and_reshape_bit_exact(cuda_provider, dtype, shape, perm):
    """Permute_M (Backend.fs:128, call site AD.Float32.fs:2773) against numpy.transpose for ranks 2..5, and Reshape_M
    (Backend.fs:129, :2780): same storage, new shape — bit exact, and the reshaped view feeds other ops correctly."""
    b, o = backend(cuda_provider, dtype), OracleBackend(dtype)
    a = np.random.default_rng(len(shape)).standard_normal(shape).astype(dtype)
    v = ShapedDataBufferView(b.CreateDataBuffer(a.reshape(-1)), *shape)
    p = b.Permute_M(v, list(perm))
    assert tuple(p.Shape) == tuple(shape[i] for i in perm)
    assert np.array_equal(p.DataBuffer.SubData.reshape(p.Shape), np.transpose(a, perm))
    hp = o.Permute_M(ShapedDataBufferView(NativeDataBuffer(a.reshape(-1)), *shape), list(perm))
    assert np.array_equal(p.DataBuffer.SubData, hp.DataBuffer.SubData)
    n = a.size
    rows = shape[0]
    r = b.Reshape_M(v, [n // rows, rows])
    assert r.DataBuffer is v.DataBuffer and tuple(r.Shape) == (n // rows, rows)            # the same buffer, no copy
    assert np.array_equal(b.Transpose_M(r).to_numpy(), a.reshape(n // rows, rows).T)
    flat = b.ReshapeCopy_MRows_V(r)                                                         # Backend.fs:109: a COPY
    back = b.ReshapeCopy_V_MRows(rows, flat)                                                # Backend.fs:134
    assert np.array_equal(flat.SubData, a.reshape(-1)) and tuple(back.Shape) == (rows, n // rows)
    b.Add_V_V_InPlace(b.CreateDataBuffer(np.ones(n, dtype)), 0, flat, 0, n)
    assert np.array_equal(v.DataBuffer.SubData, a.reshape(-1))                              # the source is untouched by writes to the copy


def test_second_lane_hazards(cuda_provider, cuda_provider_nofuse):
    """The sinks of a reverse sweep run on the library's second stream (csrc/dsc_lane.cu).  Whatever the host does next —
    read the weight adjoint at once, overwrite an operand in place, accumulate column sums into a non-empty vector, capture
    nothing — the results are those of the one-stream, unfused sequence, bit for bit."""
    rng = np.random.default_rng(11)
    m, k, n = 256, 384, 320          # small enough that two products share the GPU, large enough for the tensor-core kernel
    A, W, dC, X = (rng.standard_normal(s).astype(np.float32) for s in ((m, k), (k, n), (m, n), (m, n)))

    def run(prov):
        b = backend(prov, np.float32)
        mat = lambda a: ShapedDataBufferView(b.CreateDataBuffer(a.reshape(-1).copy()), *a.shape)  # noqa: E731
        dA, dW_, ddC, dX = mat(A), mat(W), mat(dC), mat(X)
        out = {}
        # the adjoint pair of C = A W, in the order the AD layer issues it (AD.Float32.fs:4127-4143): d.A W^T first (stays
        # pending: it may still get an epilogue), then A^T d.A (a sink: goes ahead on the second lane)
        gA = b.Mul_M_M(ddC, b.Transpose_M(dW_))
        gW = b.Mul_M_M(b.Transpose_M(dA), ddC)
        b.Add_M_M_InPlace(ddC, dX)                           # an operand of both products is overwritten right away
        out["gW"] = gW.DataBuffer.SubData.copy()             # ... and the sink is read at once
        out["gA"] = gA.DataBuffer.SubData.copy()
        out["dC"] = ddC.DataBuffer.SubData.copy()
        # column sums: into a fresh zero vector (stored) and into a non-empty accumulator (added), twice
        gA2 = b.Mul_M_M(ddC, b.Transpose_M(dW_))
        gW2 = b.Mul_M_M(b.Transpose_M(dA), ddC)
        acc = b.CreateDataBuffer(np.zeros(n, np.float32))
        b.Add_M_Colwise_V_InPlace(ddC, acc)
        b.Add_M_Colwise_V_InPlace(dX, acc)
        out["cs"] = acc.SubData.copy()
        # the weight adjoint accumulated twice into the same parameter adjoint (a shared weight)
        Wadj = b.CreateDataBuffer(np.zeros(k * n, np.float32))
        Wv = ShapedDataBufferView(Wadj, k, n)
        b.Add_M_M_InPlace(Wv, gW2)
        b.Add_M_M_InPlace(Wv, gW)
        out["Wadj"] = Wadj.SubData.copy()
        out["gA2"] = gA2.DataBuffer.SubData.copy()
        return out

    lane0 = cuda_provider.ctx.stats()["lane_ops"]
    got, want = run(cuda_provider), run(cuda_provider_nofuse)
    assert cuda_provider.ctx.stats()["lane_ops"] >= lane0 + 2      # both weight adjoints did go to the second lane
    assert cuda_provider_nofuse.ctx.stats()["lane_ops"] == 0
    for key in want:
        assert np.array_equal(got[key], want[key]), key
    ref = (A.T.astype(np.float64) @ dC.astype(np.float64))
    assert np.abs(got["gW"].reshape(k, n) - ref).max() <= 1e-5 * np.abs(ref).max()
