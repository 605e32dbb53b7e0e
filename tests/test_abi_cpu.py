"""CPU-only: the C-ABI library loads and exports every symbol include/isac_b200.h declares."""
import ctypes
import importlib
import os

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"


def test_library_exports_every_declared_symbol():
    _lib = importlib.import_module(PKG + "._lib")
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.exported_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.isac_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.isac_version()


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly (no CPU fallback)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(_lib.IsacError) as e:
        _lib.Context(0)
    assert e.value.status in (2, 3)


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), PKG)
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
