"""The Numerical.* mirror (sigma/diffsharp_b200/numerical.py, reference Numerical.Float32.fs:66-233) on the CPU oracle backend:
every operation against closed forms, and the fork's epsilon QUIRK reproduced."""
import numpy as np
import pytest

from oracle.backend import OracleBackend
from sigma.diffsharp_b200 import numerical as N
from sigma.diffsharp_b200.buffers import NativeDataBuffer

B = OracleBackend(np.float64)
ST = N.Steps.consistent(np.float64, 1e-5)
buf = lambda a: NativeDataBuffer(np.asarray(a, np.float64))  # noqa: E731
host = lambda b: np.asarray(b.SubData if hasattr(b, "SubData") else b.DataBuffer.SubData, np.float64)  # noqa: E731

A = np.array([[2.0, -1.0, 0.5], [-1.0, 3.0, 0.25], [0.5, 0.25, 1.5]])


def f_vs(b):   # x . A x + sum(sin x)
    x = host(b)
    return np.float64(x @ A @ x + np.sin(x).sum())


def f_vv(b):   # (x1 x2, sin x0 + x2^2, x0 x1 x2)
    x = host(b)
    return buf([x[1] * x[2], np.sin(x[0]) + x[2] ** 2, x[0] * x[1] * x[2]])


X = np.array([0.3, -0.7, 1.1])
V = np.array([1.0, 2.0, -0.5])
GRAD = (A + A.T) @ X + np.cos(X)
HESS = (A + A.T) + np.diag(-np.sin(X))
JAC = np.array([[0, X[2], X[1]], [np.cos(X[0]), 0, 2 * X[2]], [X[1] * X[2], X[0] * X[2], X[0] * X[1]]])


def test_scalar_ops():
    assert abs(N.diff(np.sin, np.float64(0.4), ST) - np.cos(0.4)) < 1e-9
    assert abs(N.diff2(np.sin, np.float64(0.4), ST) + np.sin(0.4)) < 1e-4
    v, d, d2 = N.diff2__(np.exp, np.float64(0.2), ST)
    assert v == np.exp(0.2) and abs(d - v) < 1e-8 and abs(d2 - v) < 1e-4


def test_vector_to_scalar_ops():
    fx, g = N.grad_(f_vs, buf(X), B, ST)
    assert fx == f_vs(buf(X)) and np.allclose(host(g), GRAD, atol=1e-4)
    assert abs(N.gradv(f_vs, buf(X), buf(V), B, ST) - GRAD @ V) < 1e-8
    h = N.hessian(f_vs, buf(X), B, ST)
    assert h.Shape == (3, 3) and np.allclose(h.to_numpy(), HESS, atol=2e-2)     # one-sided differences of one-sided differences
    gg, hh = N.gradhessian(f_vs, buf(X), B, ST)
    assert np.allclose(host(gg), GRAD, atol=1e-4) and np.allclose(hh.to_numpy(), h.to_numpy())
    assert np.allclose(host(N.hessianv(f_vs, buf(X), buf(V), B, ST)), HESS @ V, atol=1e-3)
    gv, hv = N.gradhessianv(f_vs, buf(X), buf(V), B, ST)
    assert abs(gv - GRAD @ V) < 1e-8 and np.allclose(host(hv), HESS @ V, atol=1e-3)
    assert abs(N.laplacian(f_vs, buf(X), B, ST) - np.trace(HESS)) < 5e-2


def test_vector_to_vector_ops():
    assert np.allclose(N.jacobian(f_vv, buf(X), B, ST).to_numpy(), JAC, atol=1e-4)
    assert np.allclose(N.jacobianT(f_vv, buf(X), B, ST).to_numpy(), JAC.T, atol=1e-4)
    assert np.allclose(host(N.jacobianv(f_vv, buf(X), buf(V), B, ST)), JAC @ V, atol=1e-8)
    assert np.allclose(N.curl(f_vv, buf(X), B, ST), [JAC[2, 1] - JAC[1, 2], JAC[0, 2] - JAC[2, 0], JAC[1, 0] - JAC[0, 1]], atol=1e-4)
    assert abs(N.div(f_vv, buf(X), B, ST) - np.trace(JAC)) < 1e-4
    c, d = N.curldiv(f_vv, buf(X), B, ST)
    assert abs(d - np.trace(JAC)) < 1e-4 and np.allclose(c, N.curl(f_vv, buf(X), B, ST))
    with pytest.raises(ValueError):
        N.curl(lambda b: buf(host(b)[:2]), buf(X[:2]), B, ST)


def test_reference_epsilon_quirk_is_reproduced():
    """The fork perturbs by 0.01 but rescales by 1 / 1e-5 (Numerical.Float32.fs:46-48 over Config.fs:70-74): a first derivative
    comes out 1000 x too large; the second derivative (divided by the step squared) is consistent."""
    st = N.Steps.reference(np.float32)
    assert abs(float(st.eps) - 0.01) < 1e-9 and abs(float(st.rec) - 1e5) < 1 and abs(float(st.rec2) - 5e4) < 1
    d = N.diff(lambda t: np.float32(3) * t, np.float32(1.0), st)
    assert abs(d / 3000.0 - 1) < 1e-3
    assert abs(N.diff2(lambda t: t * t, np.float32(1.0), st) - 2) < 1e-2
