"""AD against finite differences on the CUDA backend — the property suite the fork deleted from upstream DiffSharp
(SURVEY §4, §8f-3), driven by hypothesis.  Both sides run through libdiffsharp_cuda.so: the AD mirror (ad.py) issues the
reference's call sequences, the Numerical mirror (numerical.py, Numerical.Float32.fs:66-233) perturbs buffers through the same
Backend<'T> members.  float64, one consistent step h = 1e-5 (central differences: error O(h^2 f''')); tolerances below."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from _util import rel_err
from sigma.diffsharp_b200 import ad, numerical as N, workloads as wl
from sigma.diffsharp_b200.ad import DNumber, DVector

pytestmark = pytest.mark.gpu
DT = np.float64
STEPS = N.Steps.consistent(DT, 1e-5)

# unary DV operators and a sampler of inputs inside their smooth domain
UNARY = {
    "tanh": (ad.tanh, lambda r, n: r.standard_normal(n)),
    "sigmoid": (ad.sigmoid, lambda r, n: r.standard_normal(n)),
    "exp": (ad.exp, lambda r, n: r.standard_normal(n)),
    "log": (ad.log, lambda r, n: r.random(n) + 0.5),
    "sqrt": (ad.sqrt, lambda r, n: r.random(n) + 0.5),
    "sin": (ad.sin, lambda r, n: r.standard_normal(n)),
    "cos": (ad.cos, lambda r, n: r.standard_normal(n)),
    "sinh": (ad.sinh, lambda r, n: r.standard_normal(n)),
    "cosh": (ad.cosh, lambda r, n: r.standard_normal(n)),
    "atan": (ad.atan, lambda r, n: r.standard_normal(n)),
    "asin": (ad.asin, lambda r, n: r.random(n) * 1.6 - 0.8),
    "reLU": (ad.reLU, lambda r, n: np.where(np.abs(x := r.standard_normal(n)) < 0.05, 0.5, x)),   # away from the kink
    "abs": (ad.Abs, lambda r, n: np.where(np.abs(x := r.standard_normal(n)) < 0.05, 0.5, x)),
}


def _objective(opname, w):
    """f(x) = sum(op(x) .* w) + 1/2 (x . x) as an AD program over DVectors"""
    op = UNARY[opname][0]

    def f(x):
        return ad.had(op(x), w).Sum() + DNumber(0.5, DT) * (x * x)
    return f


def _on_buffers(f):
    """the same program as a function of a raw buffer (what Numerical.* differentiates)"""
    return lambda b: DT(float(f(DVector(b))))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(opname=st.sampled_from(sorted(UNARY)), n=st.integers(1, 48), seed=st.integers(0, 2 ** 16))
def test_reverse_gradient_matches_finite_differences(cuda_provider, use_provider, opname, n, seed):
    """grad (reverse mode, AD.Float32.fs:4334-4339) = Numerical grad and gradv, for every unary operator"""
    use_provider(cuda_provider)
    b = cuda_provider.GetBackend(DT(0)).BackendHandle
    rng = np.random.default_rng(seed)
    x = UNARY[opname][1](rng, n).astype(DT)
    w, v = rng.standard_normal(n).astype(DT), rng.standard_normal(n).astype(DT)
    f = _objective(opname, wl.to_DV(w))
    g_ad = ad.grad(f, wl.to_DV(x)).ToArray()
    fb = _on_buffers(f)
    dx = b.CreateDataBuffer(x)
    gv = N.gradv(fb, dx, b.CreateDataBuffer(v), b, STEPS)
    scale = max(1.0, float(np.abs(g_ad).max()))
    assert abs(float(gv) - float(g_ad @ v)) <= 1e-6 * scale * max(1.0, float(np.abs(v).sum()))
    one_sided = N.Steps.consistent(DT, 1e-6)
    g_fd = np.asarray(N.grad(fb, dx, b, one_sided).SubData)
    assert np.abs(g_fd - g_ad).max() <= 2e-4 * scale


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(opname=st.sampled_from(["tanh", "sigmoid", "exp", "sin", "cosh", "log", "sqrt"]), n=st.integers(1, 24), seed=st.integers(0, 2 ** 16))   # not atan: the reference's forward-mode tangent of atan is wrong and the mirror keeps it
def test_forward_and_second_order_match_finite_differences(cuda_provider, use_provider, opname, n, seed):
    """jacobianv (forward mode, :4345-4349) and hessianv (reverse on forward, :4416-4427) = their Numerical twins"""
    use_provider(cuda_provider)
    b = cuda_provider.GetBackend(DT(0)).BackendHandle
    rng = np.random.default_rng(seed)
    x = UNARY[opname][1](rng, n).astype(DT)
    w, v = rng.standard_normal(n).astype(DT), rng.standard_normal(n).astype(DT)
    f = _objective(opname, wl.to_DV(w))
    dx, dv = wl.to_DV(x), wl.to_DV(v)
    jv_ad = float(ad.jacobianv(f, dx, dv))
    hv_ad = ad.hessianv(f, dx, dv).ToArray()
    fb = _on_buffers(f)
    bx, bv = b.CreateDataBuffer(x), b.CreateDataBuffer(v)
    assert abs(jv_ad - float(N.gradv(fb, bx, bv, b, STEPS))) <= 1e-6 * max(1.0, abs(jv_ad), float(np.abs(v).sum()))
    # the gradient of the AD program differentiated numerically along v: central differences of an exact gradient
    g = lambda p: ad.grad(f, wl.to_DV(p)).ToArray()  # noqa: E731
    hv_fd = (g(x + 1e-5 * v) - g(x - 1e-5 * v)) / 2e-5
    assert np.abs(hv_fd - hv_ad).max() <= 1e-6 * max(1.0, float(np.abs(hv_ad).max()))
    hv_num = np.asarray(N.hessianv(fb, bx, bv, b, N.Steps.consistent(DT, 1e-4)).SubData)   # finite differences of finite differences
    assert np.abs(hv_num - hv_ad).max() <= 5e-3 * max(1.0, float(np.abs(hv_ad).max()))


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(m=st.integers(1, 12), n=st.integers(1, 12), seed=st.integers(0, 2 ** 16))
def test_vector_function_jacobians_match_finite_differences(cuda_provider, use_provider, m, n, seed):
    """f(x) = tanh(W x + c): jacobianv (forward), jacobianTv (reverse) against Numerical jacobianv / jacobianT"""
    use_provider(cuda_provider)
    b = cuda_provider.GetBackend(DT(0)).BackendHandle
    rng = np.random.default_rng(seed)
    W, c = rng.standard_normal((m, n)).astype(DT), rng.standard_normal(m).astype(DT)
    x, v, u = rng.standard_normal(n).astype(DT), rng.standard_normal(n).astype(DT), rng.standard_normal(m).astype(DT)
    dW, dc = wl.to_DM(W), wl.to_DV(c)
    f = lambda p: ad.tanh(dW * p + dc)  # noqa: E731
    J = (1 - np.tanh(W @ x + c) ** 2)[:, None] * W
    jv = ad.jacobianv(f, wl.to_DV(x), wl.to_DV(v)).ToArray()
    jtu = ad.jacobianTv(f, wl.to_DV(x), wl.to_DV(u)).ToArray()
    assert rel_err(jv, J @ v) <= 1e-12 or np.abs(jv - J @ v).max() <= 1e-13
    assert rel_err(jtu, J.T @ u) <= 1e-12 or np.abs(jtu - J.T @ u).max() <= 1e-13
    fb = lambda buf: f(DVector(buf)).Buffer  # noqa: E731
    jv_fd = np.asarray(N.jacobianv(fb, b.CreateDataBuffer(x), b.CreateDataBuffer(v), b, STEPS).SubData)
    assert np.abs(jv_fd - jv).max() <= 1e-7 * max(1.0, float(np.abs(jv).max()))
    if m == n:   # the reference's jacobianT' assumes a square Jacobian (Numerical.Float32.fs:176-178)
        JT = N.jacobianT(fb, b.CreateDataBuffer(x), b, N.Steps.consistent(DT, 1e-6)).to_numpy()
        assert np.abs(JT - J.T).max() <= 1e-4 * max(1.0, float(np.abs(J).max()))
