"""The five BASELINE.json configurations on the CUDA backend against the CPU oracle (same AD layer,
same call sequence, different Backend<'T>), plus size-independent properties at full size."""
import os

import numpy as np
import pytest

from _util import rel_err
from sigma.diffsharp_b200 import ad, workloads as wl
from sigma.diffsharp_b200.buffers import ShapedDataBufferView

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


def _mlp(layers, batch):
    X, Y, Ws, bs = wl.mlp_inputs(layers, batch)
    L, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    return float(L), [a.ToArray() for a in dW] + [a.ToArray() for a in db]


@pytest.mark.parametrize("fixture", ["cuda_provider", "cuda_provider_nofuse"])
def test_c1_mlp_gradient_parity(request, oracle_provider, use_provider, fixture):
    """config 1: 784-300-10 tanh MLP, B=128, float32 — loss and every adjoint within 1e-5 of the oracle."""
    use_provider(oracle_provider)
    Lr, gr = _mlp([784, 300, 10], 128)
    use_provider(request.getfixturevalue(fixture))
    Lg, gg = _mlp([784, 300, 10], 128)
    assert abs(Lg - Lr) <= 1e-5 * abs(Lr)
    assert abs(Lg - float(GOLD["c1_loss"][0])) <= 1e-5 * abs(Lr)
    for i, (a, b) in enumerate(zip(gg, gr)):
        assert rel_err(a, b) <= 1e-5, i
        assert rel_err(a.reshape(-1)[:16], GOLD[f"c1_grad{i}_head"]) <= 1e-4 or np.abs(GOLD[f"c1_grad{i}_head"]).max() < 1e-3


def test_fused_and_unfused_paths_are_bit_identical(cuda_provider, cuda_provider_nofuse, use_provider):
    """lazy Transpose_M -> transposed GEMM operand and the fused tanh adjoint chain perform the same
    roundings as the eager call sequence."""
    use_provider(cuda_provider)
    L1, g1 = _mlp([96, 64, 48, 10], 200)
    use_provider(cuda_provider_nofuse)
    L2, g2 = _mlp([96, 64, 48, 10], 200)
    assert L1 == L2
    for a, b in zip(g1, g2):
        assert np.array_equal(a, b)


def test_c3_shape_small_batch_parity(oracle_provider, cuda_provider, use_provider):
    """config 3's network (784-1024-1024-10) at a batch the oracle finishes in seconds."""
    use_provider(oracle_provider)
    Lr, gr = _mlp([784, 1024, 1024, 10], 256)
    use_provider(cuda_provider)
    Lg, gg = _mlp([784, 1024, 1024, 10], 256)
    assert abs(Lg - Lr) <= 1e-5 * abs(Lr)
    for a, b in zip(gg, gr):
        assert rel_err(a, b) <= 1e-5


@pytest.fixture(scope="module")
def c3_oracle_full(oracle_provider):
    """ONE full-size config-3 step on the CPU oracle (a few seconds), shared by the tests below."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    from sigma.diffsharp_b200.config import GlobalConfig
    old = GlobalConfig.BackendProvider
    GlobalConfig.BackendProvider = oracle_provider
    try:
        return mg.c3_full_step()
    finally:
        GlobalConfig.BackendProvider = old


@pytest.mark.parametrize("fixture", ["cuda_provider", "cuda_provider_nofuse"])
def test_c3_full_size_against_the_oracle(request, c3_oracle_full, use_provider, fixture):
    """config 3 AT THE SIZE THE METRIC IS QUOTED ON (784-1024-1024-10, batch 8192, float32): loss and all six adjoints
    within 1e-5 of the CPU oracle, with and without the deferred evaluation, and against the committed digest that
    bench.py and the N-GPU runs compare with.  The weight adjoints reduce over K = 8192 batch rows on the tensor cores."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    Lr, gr = c3_oracle_full
    use_provider(request.getfixturevalue(fixture))
    Lg, gg = mg.c3_full_step()
    L64 = float(GOLD["c3_loss64"][0])
    assert abs(Lg - L64) <= 1e-5 * L64                     # Sum_M against the float64 sum (App. D-6) ...
    assert abs(Lr - L64) <= 2e-5 * L64                     # ... next to the oracle's sequential float32 sum of 81 920 terms
    for i, (a, b) in enumerate(zip(gg, gr)):
        assert a.shape == b.shape
        assert rel_err(a, b) <= 1e-5, (i, rel_err(a, b))
    assert mg.c3_digest_max_rel(Lg, gg, GOLD) <= 1e-5
    assert mg.c3_digest_max_rel(Lr, gr, GOLD) <= 1e-5      # the oracle reproduces its own frozen digest


def test_c3_full_size_shard_additivity(cuda_provider, use_provider):
    """full batch 8192: the loss is a sum over rows, so gradients of two half-batches add up to the
    gradient of the whole batch (the property the multi-GPU sharding relies on)."""
    use_provider(cuda_provider)
    layers, B = [784, 1024, 1024, 10], 8192
    X, Y, Ws, bs = wl.mlp_inputs(layers, B)
    def run(r0, r1):
        L, dW, db = wl.mlp_grad(wl.to_DM(X[r0:r1]), wl.to_DM(Y[r0:r1]), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
        return float(L), [a.ToArray().astype(np.float64) for a in dW + db]
    Lf, gf = run(0, B)
    La, ga = run(0, B // 2)
    Lb, gb = run(B // 2, B)
    assert abs(Lf - (La + Lb)) <= 1e-5 * abs(Lf)
    for f, a, b in zip(gf, ga, gb):
        assert rel_err(a + b, f) <= 2e-5


def test_c4_jacobian_rows(oracle_provider, cuda_provider, use_provider):
    """config 4 at n=256: tangent rows vs oracle and closed form, float64 <= 1e-12."""
    W1, b1, W2, b2, x = wl.jacobian_inputs(256)
    def run():
        return wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), 64, 128).ToArray()
    use_provider(oracle_provider)
    Jr = run()
    use_provider(cuda_provider)
    Jg = run()
    assert rel_err(Jg, Jr) <= 1e-12
    assert rel_err(Jg, wl.jacobian_closed_form(W1, b1, W2, b2, x)[64:192]) <= 1e-12
    W1, b1, W2, b2, x = wl.jacobian_inputs(96)
    J96 = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), 16, 32).ToArray()
    assert rel_err(J96, GOLD["c4_rows"]) <= 1e-12


def test_c4_full_size_sampled_rows_against_the_oracle(oracle_provider, cuda_provider, use_provider):
    """config 4 at n = 4096 (float64): the GPU computes one rank's shard of tangent rows (512 of 4096, what a rank of the
    8-GPU run computes); 64 of those rows, sampled, are recomputed by the CPU oracle from the same identity tangents and must
    agree to 1e-12 — and both with the closed form."""
    n, c0, cnt = 4096, 1024, 512
    W1, b1, W2, b2, x = wl.jacobian_inputs(n)
    rows = np.sort(np.random.default_rng(11).choice(cnt, 64, replace=False))
    use_provider(cuda_provider)
    Jg = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), c0, cnt).ToArray()
    use_provider(oracle_provider)
    T = np.zeros((rows.size, n), np.float64)
    T[np.arange(rows.size), c0 + rows] = 1
    Jr = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), 0, rows.size, tangent=wl.to_DM(T)).ToArray()
    assert rel_err(Jg[rows], Jr) <= 1e-12
    ref = wl.jacobian_closed_form(W1, b1, W2, b2, x)[c0:c0 + cnt]
    assert rel_err(Jg, ref) <= 1e-12 and rel_err(Jr, ref[rows]) <= 1e-12


def test_c4_replicated_primal_is_bit_identical(cuda_provider, cuda_provider_nofuse, use_provider):
    """config 4 at n = 4096: with deferred evaluation the row-replicated primal (DNDArray.OfRows(m, x)) stays a
    vector — rows(x) W = rows(x W) through the same DMMA kernel, maps / sums of replicated rows on one row — and only the
    tangent GEMMs run at full size; the Jacobian rows are bit for bit those of the undeferred sequence, which multiplies the
    m identical rows out."""
    n, c0, cnt = 4096, 512, 128
    W1, b1, W2, b2, x = wl.jacobian_inputs(n)
    out = []
    for prov in (cuda_provider, cuda_provider_nofuse):
        use_provider(prov)
        st0 = prov.ctx.stats()["gemm_dmma"]
        J = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), c0, cnt).ToArray()
        out.append((J, prov.ctx.stats()["gemm_dmma"] - st0))
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1] == out[1][1] == 4            # four DMMA launches either way: two of them on ONE row when deferred
    assert rel_err(out[0][0], wl.jacobian_closed_form(W1, b1, W2, b2, x)[c0:c0 + cnt]) <= 1e-12


def test_c5_hvp(oracle_provider, cuda_provider, use_provider):
    """config 5: reverse-on-forward Hessian-vector product, float64, small vs oracle, full size vs closed form."""
    x, v, d = wl.hvp_inputs(4096)
    use_provider(oracle_provider)
    hr = wl.hvp(wl.to_DV(d), wl.to_DV(x), wl.to_DV(v)).ToArray()
    use_provider(cuda_provider)
    hg = wl.hvp(wl.to_DV(d), wl.to_DV(x), wl.to_DV(v)).ToArray()
    assert rel_err(hg, hr) <= 1e-12 and rel_err(hg, GOLD["c5_hv"]) <= 1e-12
    x, v, d = wl.hvp_inputs(1_000_000)
    hg = wl.hvp(wl.to_DV(d), wl.to_DV(x), wl.to_DV(v)).ToArray()
    assert rel_err(hg, wl.hvp_closed_form(x, v, d)) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c2_full_size_properties(cuda_provider, dtype):
    """config 2 at 4096x4096: properties that need no CPU pass over 16.7 M elements per check."""
    b = cuda_provider.GetBackend(np.dtype(dtype).type(0)).BackendHandle
    n = 4096
    rng = np.random.default_rng(2)
    A = (rng.random((n, n)) * 2 - 1).astype(dtype)
    dA = ShapedDataBufferView(b.CreateDataBuffer(A.reshape(-1)), n, n)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    # Transpose_M: involution, bit exact
    At = b.Transpose_M(dA)
    assert np.array_equal(b.Transpose_M(At).to_numpy(), A)
    assert np.array_equal(At.to_numpy()[:64], A.T[:64])
    # Sum_M vs exact float64 sum (App. D-6)
    exact = float(A.astype(np.float64).sum())
    assert abs(float(b.Sum_M(dA.DataBuffer)) - exact) <= tol * float(np.abs(A).astype(np.float64).sum())
    # Mul_Had_M_M and Map_F_M on a 1/64 sample of rows
    H = b.Mul_Had_M_M(dA, dA).to_numpy()
    assert np.array_equal(H[::64], (A * A)[::64])
    T = b.Map_F_M(17, None, dA).to_numpy()
    assert rel_err(T[::64], np.tanh(A[::64].astype(np.float64))) <= tol
    E = b.Map_F_M(10, None, dA).to_numpy()
    assert rel_err(E[::64], np.exp(A[::64].astype(np.float64))) <= tol
    # Mul_M_M: (A x) via GEMM equals GEMV-based check: C e_j and row sums against float64 on a sample
    Bm = (rng.random((n, n)) * 2 - 1).astype(dtype)
    dB = ShapedDataBufferView(b.CreateDataBuffer(Bm.reshape(-1)), n, n)
    Cm = b.Mul_M_M(dA, dB).to_numpy().astype(np.float64)
    rows = rng.integers(0, n, 48)
    ref = A[rows].astype(np.float64) @ Bm.astype(np.float64)
    bound = tol * (np.abs(A[rows]).astype(np.float64) @ np.abs(Bm).astype(np.float64))
    assert np.all(np.abs(Cm[rows] - ref) <= bound)
    # checksum of checksums: 1^T C 1 == (1^T A)(B 1)
    lhs = Cm.sum()
    rhs = A.astype(np.float64).sum(0) @ Bm.astype(np.float64).sum(1)
    assert abs(lhs - rhs) <= tol * float(np.abs(A).astype(np.float64).sum(0) @ np.abs(Bm).astype(np.float64).sum(1))


def test_graph_replay_equals_eager(cuda_provider, use_provider):
    """CUDA-graph capture of one MLP gradient step replays to the same adjoints as the eager step."""
    use_provider(cuda_provider)
    b = cuda_provider.GetBackend(np.float32(0)).BackendHandle
    X, Y, Ws, bs = wl.mlp_inputs([64, 48, 10], 96)
    dX, dY = wl.to_DM(X), wl.to_DM(Y)
    dWs, dbs = [wl.to_DM(w) for w in Ws], [wl.to_DV(v) for v in bs]
    L, gW, gb = wl.mlp_grad(dX, dY, dWs, dbs)
    eager = [a.ToArray() for a in gW + gb]
    ctx = cuda_provider.ctx
    ctx.sync()
    b.capture_scalars = []
    ctx.graph_begin()
    try:
        _, cW, cb = wl.mlp_grad(dX, dY, dWs, dbs)
    finally:
        g = ctx.graph_end()
        loss_bufs, b.capture_scalars = b.capture_scalars, None
    for _ in range(3):
        ctx.graph_launch(g)
    ctx.sync()
    for a, e in zip(cW + cb, eager):
        assert np.array_equal(a.ToArray(), e)
    assert abs(float(loss_bufs[0].SubData[0]) - float(L)) <= 1e-6 * abs(float(L))
    ctx.graph_destroy(g)


def test_eager_loop_without_sync_points_keeps_the_pool_bounded(cuda_provider, use_provider):
    """A host loop that never reads anything back (no loss read, no download, no dsc_sync) still issues sinks to the second
    compute lane every step; the blocks those operations hold must not pile up (regression: every allocation of the
    following steps then missed the pool and the step took 20x longer)."""
    use_provider(cuda_provider)
    ctx = cuda_provider.ctx
    b32 = cuda_provider.GetBackend(np.float32(0)).BackendHandle
    X, Y, Ws, bs = wl.mlp_inputs([160, 256, 192, 10], 512, seed=9)
    dX, dY = wl.to_DM(X), wl.to_DM(Y)
    dWs, dbs = [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs]
    b32.capture_scalars = []          # the loss stays on the device: no synchronisation point inside a step
    try:
        for _ in range(4):
            wl.mlp_grad(dX, dY, dWs, dbs)
        ctx.sync()
        s0 = ctx.stats()
        for _ in range(60):
            wl.mlp_grad(dX, dY, dWs, dbs)
            b32.capture_scalars.clear()
        s1 = ctx.stats()
        ctx.sync()
    finally:
        b32.capture_scalars = None
    assert s1["lane_ops"] - s0["lane_ops"] >= 60 * 3                      # the sinks did go to the lane
    assert s1["pool_misses"] - s0["pool_misses"] <= 40                    # ... and the pool reached a steady state
    assert s1["pool_bytes_reserved"] <= 4 * max(s0["pool_bytes_reserved"], 1 << 20)
