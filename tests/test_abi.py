"""The C-ABI boundary: the library loads without a GPU, exports every symbol the header declares,
and fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import subprocess

import pytest

from sigma.diffsharp_b200 import _lib
from sigma.diffsharp_b200.backend import Backend, CudaBackend, MapOp


def test_library_exports_every_declared_symbol():
    declared = _lib.declared_symbols()
    assert len(declared) >= 100
    missing = [s for s in declared if not hasattr(_lib.lib, s)]
    assert missing == []


def test_no_undeclared_exports():
    out = subprocess.check_output(["nm", "-D", _lib.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T dsc_" in l}
    assert exported == set(_lib.declared_symbols())


def test_library_has_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_hot_kernels_use_the_blackwell_units():
    """SASS evidence that the hot paths are what DESIGN.md says they are: tcgen05 MMAs (1- and 2-CTA) fed by TMA with
    TMEM loads in the epilogue, thread-block cluster barriers, DMMA for float64, 128-bit global loads in the streams."""
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D", "LDTM", "UTCBAR", "UCGABAR_ARV", "DMMA", "LDG.E.128", "SYNCS.PHASECHK"):
        assert mnemonic in sass, mnemonic


def test_version():
    assert _lib.lib.dsc_version() == 100


def test_mapop_ordinals_follow_reference_declaration_order():
    # Backend.fs:42-70
    names = ["Mul", "Div", "Add", "Sub", "Pow_Ab", "Pow_Ba", "Atan2_Ab", "Atan2_Ba", "Log", "Log10", "Exp", "Sin",
             "Cos", "Tan", "Sqrt", "Sinh", "Cosh", "Tanh", "Asin", "Acos", "Atan", "Abs", "Sign", "Floor", "Ceiling",
             "Round", "ReL", "Sigmoid"]
    assert [m.name for m in MapOp] == names
    assert [int(m) for m in MapOp] == list(range(28))


def test_backend_interface_has_55_members():
    assert len(Backend.MEMBERS) == 55
    for m in Backend.MEMBERS:
        assert callable(getattr(CudaBackend, m)), m


def test_create_fails_loudly_without_a_gpu():
    n = C.c_int()
    if _lib.lib.dsc_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    h = _lib.ctx_t()
    st = _lib.lib.dsc_create(0, 0, C.byref(h))
    assert st == _lib.DSC_E_CUDA
    assert "no CPU fallback" in _lib.last_error()
    with pytest.raises(_lib.DeviceError):
        _lib.check(st)


def test_bad_handles_are_rejected_not_crashed():
    assert _lib.lib.dsc_sync(None) == _lib.DSC_E_INVALID
    assert _lib.lib.dsc_buf_free(0) == _lib.DSC_OK
    out = _lib.buf_t()
    assert _lib.lib.dsc_buf_alloc(None, 0, 4, C.byref(out)) == _lib.DSC_E_INVALID
