"""The five BASELINE.json configurations, written against the AD layer in the fork's idiom
(batch in ROWS, bias broadcast with ``DNDArray.OfRows(m, DVector, backend)``,
AD.Float32.fs:1860-1865; parameters made reverse with one shared tag and read back through ``.A``
as in docs/input/examples-neuralnetworks.fsx:272-296).  Backend-agnostic: everything goes through
``GlobalConfig.BackendProvider``.

  C1  784-300-10 tanh MLP, B=128, f32        reverse-mode loss gradient
  C3  784-1024-1024-10 tanh MLP, B=8192, f32 the same, batch rows sharded over GPUs (dist.py)
  C4  two 4096->4096 tanh layers, f64        forward-mode Jacobian with a matrix tangent, columns sharded
  C5  tanh-quadratic scalar function, n=1e6  reverse-on-forward Hessian-vector product, f64
  (C2 is the per-op sweep; it needs no AD layer.)

SURVEY §8(d) fixes the synthetic inputs: NumPy ``default_rng(seed)``, seed = config index.
"""
from __future__ import annotations

import numpy as np

from . import ad
from .ad import DNDArray, DNumber, DVector, GlobalTagger, had, tanh
from .buffers import ShapedDataBufferView
from .config import GlobalConfig


def _backend(dtype):
    return GlobalConfig.BackendProvider.GetBackend(np.dtype(dtype).type(0)).BackendHandle


def to_DM(a: np.ndarray) -> DNDArray:
    """host matrix -> primal DM on the installed backend (CreateDataBuffer, Backend.fs:76)"""
    b = _backend(a.dtype)
    return DNDArray(ShapedDataBufferView(b.CreateDataBuffer(np.ascontiguousarray(a).reshape(-1)), *a.shape))


def to_DV(a: np.ndarray) -> DVector:
    b = _backend(a.dtype)
    return DVector(b.CreateDataBuffer(np.ascontiguousarray(a).reshape(-1)))


# ------------------------------------------------------------------------------------------------
# C1 / C3: reverse-mode gradient of a tanh MLP's summed squared error
# ------------------------------------------------------------------------------------------------
def mlp_inputs(layers, batch, dtype=np.float32, seed=1, row_range=None):
    """X ~ U[0,1), Y one-hot, W ~ U[-0.075,0.075) (examples-neuralnetworks.fsx:350,354), b = 0.
    `row_range=(r0, r1)` returns only that block of batch rows (the shard of one rank); the random
    stream is the same for every rank so shards are slices of one global batch."""
    rng = np.random.default_rng(seed)
    X = rng.random((batch, layers[0]), dtype=np.float64).astype(dtype)
    labels = rng.integers(layers[-1], size=batch)
    Y = np.zeros((batch, layers[-1]), dtype)
    Y[np.arange(batch), labels] = 1
    Ws = [(rng.random((layers[i], layers[i + 1])) * 0.15 - 0.075).astype(dtype) for i in range(len(layers) - 1)]
    bs = [np.zeros(layers[i + 1], dtype) for i in range(len(layers) - 1)]
    if row_range is not None:
        X, Y = X[row_range[0]:row_range[1]], Y[row_range[0]:row_range[1]]
    return X, Y, Ws, bs


def mlp_loss(X: DNDArray, Y: DNDArray, Ws, bs) -> DNumber:
    """Sum((tanh(...tanh(X·W1 + rows(b1))...)·Wn + rows(bn) − Y) .* (same))   (SURVEY §3.2 trace)"""
    backend = ad.Backend(X)
    H = X
    last = len(Ws) - 1
    for i, (W, b) in enumerate(zip(Ws, bs)):
        Z = H * W + DNDArray.OfRows(H.Rows, b, backend)
        H = Z if i == last else tanh(Z)
    E = H - Y
    return had(E, E).Sum()


def mlp_grad(X: DNDArray, Y: DNDArray, Ws, bs):
    """One gradient step's worth of AD: returns (loss D, [W.A ...], [b.A ...]).
    makeReverse with one tag for every parameter, forward, reverseProp (D 1)
    (examples-neuralnetworks.fsx:283-289)."""
    i = GlobalTagger.Next()
    Wr = [ad.makeReverse(i, W) for W in Ws]
    br = [ad.makeReverse(i, b) for b in bs]
    L = mlp_loss(X, Y, Wr, br)
    ad.reverseProp(DNumber(1, L.dtype), L)
    return L, [w.A for w in Wr], [b.A for b in br]


def sgd_update(params, grads, eta):
    """``l.W <- primal (l.W.P - eta * l.W.A)`` (examples-neuralnetworks.fsx:291-293): Mul_S + Sub."""
    return [p - eta * g for p, g in zip(params, grads)]


# ------------------------------------------------------------------------------------------------
# C4: forward-mode Jacobian of a two-layer tanh composition with a matrix tangent
# ------------------------------------------------------------------------------------------------
def jacobian_inputs(n=4096, dtype=np.float64, seed=3):
    rng = np.random.default_rng(seed)
    W1 = ((rng.random((n, n)) * 2 - 1) / np.sqrt(n)).astype(dtype)
    W2 = ((rng.random((n, n)) * 2 - 1) / np.sqrt(n)).astype(dtype)
    b1 = rng.standard_normal(n).astype(dtype)
    b2 = rng.standard_normal(n).astype(dtype)
    x = rng.standard_normal(n).astype(dtype)
    return W1, b1, W2, b2, x


def layers_fn(W1: DNDArray, b1: DVector, W2: DNDArray, b2: DVector):
    """f(X) = tanh(tanh(X·W1 + rows(b1))·W2 + rows(b2)), one input per ROW of X."""
    def f(X):
        backend = ad.Backend(W1)
        H1 = tanh(X * W1 + DNDArray.OfRows(X.Rows, b1, backend))
        return tanh(H1 * W2 + DNDArray.OfRows(X.Rows, b2, backend))
    return f


def identity_rows(n: int, col_start: int, col_count: int, dtype=np.float64) -> DNDArray:
    """Rows [col_start, col_start+col_count) of the n x n identity as a primal DM: the matrix tangent."""
    T = np.zeros((col_count, n), dtype)
    T[np.arange(col_count), col_start + np.arange(col_count)] = 1
    return to_DM(T)


def jacobian_rows(W1, b1, W2, b2, x: DVector, col_start: int, col_count: int, tangent: DNDArray | None = None) -> DNDArray:
    """Rows [col_start, col_start+col_count) of J^T, J = df/dx: ONE forward sweep whose tangent is a
    block of identity rows (same unchanged AD code, GEMM-bound instead of GEMV-bound).
    The faithful call sequence carries a row-replicated primal (RepeatReshapeCopy_V_MRows)."""
    backend = ad.Backend(W1)
    n = x.Length
    Xp = DNDArray.OfRows(col_count, x, backend)
    if tangent is None:
        tangent = identity_rows(n, col_start, col_count, x.dtype)
    primal, tan = ad.jacobianv_(layers_fn(W1, b1, W2, b2), Xp, tangent)
    return tan


def jacobian_closed_form(W1, b1, W2, b2, x):
    """J^T[c, :] = e_c W1 diag(1-h1^2) W2 diag(1-h2^2) in float64 NumPy (checker)."""
    h1 = np.tanh(x @ W1 + b1)
    h2 = np.tanh(h1 @ W2 + b2)
    return (W1 * (1 - h1 * h1)) @ (W2 * (1 - h2 * h2))


# ------------------------------------------------------------------------------------------------
# C5: reverse-on-forward Hessian-vector product of a tanh-quadratic scalar function
# ------------------------------------------------------------------------------------------------
def hvp_inputs(n=1_000_000, dtype=np.float64, seed=4):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype(dtype)
    v = rng.standard_normal(n).astype(dtype)
    d = (rng.random(n) + 0.5).astype(dtype)
    return x, v, d


def tanh_quadratic(d: DVector):
    """f(x) = 1/2 (tanh x)·(tanh x) + 1/2 x·(d .* x)"""
    def f(x):
        t = tanh(x)
        half = DNumber(0.5, x.dtype)
        return half * (t * t) + half * (x * had(d, x))
    return f


def hvp(d: DVector, x: DVector, v: DVector) -> DVector:
    """``hessianv f x v = grad' (fun xx -> jacobianv f xx v) x |> snd`` (AD.Float32.fs:4416-4427)"""
    return ad.hessianv(tanh_quadratic(d), x, v)


def hvp_closed_form(x, v, d):
    """H v = (sech^4 x - 2 tanh^2 x sech^2 x + d) .* v   (float64 NumPy checker)"""
    t = np.tanh(x)
    s2 = 1 - t * t
    return (s2 * s2 - 2 * t * t * s2 + d) * v
