"""CPU: the C-ABI library builds/loads and exports every symbol include/oqupy_b200.h
declares (no compute calls without a GPU); the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from oqupy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "oqupy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == names
    _lib.load_library()
    assert _lib.load_library().b200_abi_version() == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    with pytest.raises(_lib.B200Error):
        _lib.CudaOps()
