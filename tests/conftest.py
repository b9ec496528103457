import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def _have_gpu():
    try:
        import ctypes as C
        from sigma.diffsharp_b200 import _lib
        n = C.c_int()
        return _lib.lib.dsc_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_provider():
    from oracle.backend import OracleBackend
    from sigma.diffsharp_b200.config import DictBackendProvider
    return DictBackendProvider(OracleBackend(np.float32), OracleBackend(np.float64))


@pytest.fixture(scope="session")
def cuda_provider():
    if not _have_gpu():
        pytest.skip("no CUDA device (gpu tests run on the GPU box)")
    from sigma.diffsharp_b200.config import CudaBackendProvider
    return CudaBackendProvider(device=0)


@pytest.fixture(scope="session")
def cuda_provider_nofuse(cuda_provider):
    """a second context on the same GPU created with DSC_FLAG_NO_FUSE: every call launches its own kernel(s) at once"""
    from sigma.diffsharp_b200.config import CudaBackendProvider
    return CudaBackendProvider(device=0, fuse=False)


@pytest.fixture()
def use_provider():
    """Install a provider in GlobalConfig for the duration of a test."""
    from sigma.diffsharp_b200.config import GlobalConfig
    old = GlobalConfig.BackendProvider

    def install(p):
        GlobalConfig.BackendProvider = p
        return p
    yield install
    GlobalConfig.BackendProvider = old
