"""The host mirror of the AD layer (sigma/diffsharp_b200/ad.py) over the CPU oracle backend:
  * the Backend<'T> call SEQUENCE it issues for one tanh layer equals the trace documented from the
    reference (AD.Float32.fs:2381-2713 forward, :4118-4257 reverse);
  * gradients / Jacobians / Hessian-vector products agree with torch.autograd in float64 and with
    closed forms (independent derivations of the same maths)."""
import numpy as np
import pytest
import torch

from _util import TracingProvider, rel_err
from sigma.diffsharp_b200 import ad, workloads as wl
from sigma.diffsharp_b200.ad import DNDArray, DNumber, had, tanh


def test_one_tanh_layer_issues_the_reference_call_sequence(oracle_provider, use_provider):
    tp = use_provider(TracingProvider(oracle_provider))
    rng = np.random.default_rng(0)
    A = wl.to_DM(rng.standard_normal((6, 5)).astype(np.float32))
    W = wl.to_DM(rng.standard_normal((5, 4)).astype(np.float32))
    b = wl.to_DV(np.zeros(4, np.float32))
    tp.log.clear()
    i = ad.GlobalTagger.Next()
    Wr, br = ad.makeReverse(i, W), ad.makeReverse(i, b)
    fwd_start = len(tp.log)
    H = tanh(A * Wr + DNDArray.OfRows(A.Rows, br, ad.Backend(A)))
    L = H.Sum()
    fwd = tp.log[fwd_start:]
    zero = ["CreateZeroArray", "CreateDataBuffer"]
    assert tp.log[:fwd_start] == zero  # DM.GetReverse allocates through the backend (:1743); DV.GetReverse is a host ZeroN (:610)
    assert fwd == (["Mul_M_M"] + zero                       # (*) :2381 + ZeroMN :1834
                   + ["RepeatReshapeCopy_V_MRows"] + zero   # OfRows :1860-1865
                   + ["Add_M_M"] + zero                     # (+) :2357
                   + ["Map_F_M"] + zero                     # Tanh :2704
                   + ["Sum_M"])                             # Sum :2805
    tp.log.clear()
    ad.reverseReset(L)
    assert tp.log == zero * 6  # H, Z, A*W, W (:3746), rows(b) (:3746), b (:3640-3641) — every visit re-zeroes
    tp.log.clear()
    ad.reversePush(DNumber(1, np.float32), L)
    assert tp.log == (
        ["CreateValueArray", "CreateDataBuffer", "Add_M_M_InPlace"]                      # Sum_DM :3920 -> DM.create; push into H
        + ["Map_F_M", "Map_F_S_M", "Mul_Had_M_M", "Mul_Had_M_M", "Add_M_M_InPlace"]      # Tanh_DM :4238-4240; push into Z
        + ["Add_M_M_InPlace"]                                                            # Add_DM_DM :4122 first operand (A*W)
        + ["Transpose_M", "Mul_M_M", "Add_M_M_InPlace"]                                  # Mul_DMCons_DM :4137-4143; push into W
        + ["Add_M_M_InPlace"]                                                            # Add_DM_DM second operand (rows b)
        + ["Add_M_Colwise_V_InPlace", "Add_V_V"])                                        # Make_DMRows_ofDV :4254-4257; push Zero into b (:3956)


@pytest.mark.parametrize("layers,batch", [([784, 300, 10], 128), ([20, 16, 16, 5], 33)])
def test_mlp_gradient_matches_torch_autograd(oracle_provider, use_provider, layers, batch):
    use_provider(oracle_provider)
    X, Y, Ws, bs = wl.mlp_inputs(layers, batch)
    L, dW, db = wl.mlp_grad(wl.to_DM(X), wl.to_DM(Y), [wl.to_DM(w) for w in Ws], [wl.to_DV(b) for b in bs])
    tX, tY = torch.tensor(X, dtype=torch.float64), torch.tensor(Y, dtype=torch.float64)
    tW = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws]
    tb = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in bs]
    H = tX
    for k, (w, b) in enumerate(zip(tW, tb)):
        Z = H @ w + b
        H = Z if k == len(tW) - 1 else torch.tanh(Z)
    loss = ((H - tY) ** 2).sum()
    loss.backward()
    assert abs(float(L) - loss.item()) <= 1e-5 * abs(loss.item())
    for a, t in zip(dW + db, tW + tb):
        assert rel_err(a.ToArray().reshape(t.shape), t.grad.numpy()) <= 1e-5


def test_jacobian_rows_match_closed_form(oracle_provider, use_provider):
    use_provider(oracle_provider)
    W1, b1, W2, b2, x = wl.jacobian_inputs(64)
    J = wl.jacobian_rows(wl.to_DM(W1), wl.to_DV(b1), wl.to_DM(W2), wl.to_DV(b2), wl.to_DV(x), 8, 16).ToArray()
    assert rel_err(J, wl.jacobian_closed_form(W1, b1, W2, b2, x)[8:24]) <= 1e-12


def test_hvp_matches_closed_form(oracle_provider, use_provider):
    use_provider(oracle_provider)
    x, v, d = wl.hvp_inputs(2000)
    hv = wl.hvp(wl.to_DV(d), wl.to_DV(x), wl.to_DV(v)).ToArray()
    assert rel_err(hv, wl.hvp_closed_form(x, v, d)) <= 1e-12


def test_elementwise_adjoints_against_torch(oracle_provider, use_provider):
    """reverse rules of the MapOps north_star names (exp, log, tanh, sigmoid, ReLU, pow, abs, sign)"""
    use_provider(oracle_provider)
    rng = np.random.default_rng(5)
    x = (rng.random((7, 9)) + 0.3).astype(np.float64) * np.where(rng.random((7, 9)) < 0.5, -1, 1)
    pos = np.abs(x)
    cases = [
        (lambda a: ad.exp(a), torch.exp, x), (lambda a: ad.log(a), torch.log, pos), (lambda a: tanh(a), torch.tanh, x),
        (lambda a: ad.sigmoid(a), torch.sigmoid, x), (lambda a: ad.reLU(a), torch.relu, x),
        (lambda a: a ** DNumber(2.5, np.float64), lambda t: t ** 2.5, pos), (lambda a: ad.Abs(a), torch.abs, x),
        (lambda a: ad.sin(a), torch.sin, x), (lambda a: ad.cosh(a), torch.cosh, x), (lambda a: ad.sqrt(a), torch.sqrt, pos),
        (lambda a: ad.atan(a), torch.atan, x),
    ]
    w = rng.standard_normal((7, 9))
    for f, tf, inp in cases:
        g = ad.grad(lambda a: had(f(a), wl.to_DM(w)).Sum(), wl.to_DM(inp)).ToArray()
        t = torch.tensor(inp, requires_grad=True)
        (tf(t) * torch.tensor(w)).sum().backward()
        assert rel_err(g, t.grad.numpy()) <= 1e-12


def test_vector_ops_forward_and_reverse(oracle_provider, use_provider):
    use_provider(oracle_provider)
    rng = np.random.default_rng(6)
    A = rng.standard_normal((5, 4))
    x = rng.standard_normal(4)
    f = lambda v: ad.sin(wl.to_DM(A) * v).L2NormSq()  # noqa: E731
    g = ad.grad(f, wl.to_DV(x)).ToArray()
    t = torch.tensor(x, requires_grad=True)
    (torch.sin(torch.tensor(A) @ t) ** 2).sum().backward()
    assert rel_err(g, t.grad.numpy()) <= 1e-12
    v = rng.standard_normal(4)
    jv = ad.jacobianv(lambda z: ad.sin(wl.to_DM(A) * z), wl.to_DV(x), wl.to_DV(v)).ToArray()
    assert rel_err(jv, np.cos(A @ x) * (A @ v)) <= 1e-12
    # jacobianTv
    u = rng.standard_normal(5)
    jtv = ad.jacobianTv(lambda z: ad.sin(wl.to_DM(A) * z), wl.to_DV(x), wl.to_DV(u)).ToArray()
    assert rel_err(jtv, A.T @ (np.cos(A @ x) * u)) <= 1e-12
