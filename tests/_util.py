"""Shared helpers for the test-suite."""
import numpy as np

from sigma.diffsharp_b200.backend import Backend
from sigma.diffsharp_b200.config import BackendConfig, IBackendProvider, dtype_of


class TracingBackend:
    """Wraps a backend and records the name of every Backend<'T> member that is called."""

    def __init__(self, inner, log):
        self._inner, self._log, self.dtype = inner, log, inner.dtype

    def __getattr__(self, name):
        attr = getattr(self._inner, name)
        if name in Backend.MEMBERS:
            def wrapped(*a, **k):
                self._log.append(name)
                return attr(*a, **k)
            return wrapped
        return attr


class TracingProvider(IBackendProvider):
    def __init__(self, inner_provider):
        self.log = []
        self._cfg = {}
        for dt in (np.float32, np.float64):
            b = inner_provider.GetBackend(np.dtype(dt).type(0)).BackendHandle
            self._cfg[np.dtype(dt)] = BackendConfig(TracingBackend(b, self.log), dt)

    def GetBackend(self, value):
        return self._cfg[dtype_of(value)]


def rel_err(a, b):
    """max |a-b| / max |b| in float64 (norm-wise relative error)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0))
