"""The C-ABI library loads and exports every symbol include/pcb200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import PKG, ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(cuda_lib):
    lib = ctypes.CDLL(cuda_lib)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pcb200.h but not exported by libpcb200.so"
    lib.pcb_backend.restype = ctypes.c_char_p
    assert lib.pcb_backend() == b"cuda-sm_100a"


def test_binding_covers_header():
    import importlib
    L = importlib.import_module(PKG)._lib
    assert sorted(L.SIGNATURES) == _declared()


def test_no_cpu_fallback_without_device(cuda_lib):
    """On a box without a GPU the product must fail loudly (on the GPU box the context is created)."""
    import importlib
    pkg = importlib.import_module(PKG)
    pkg._lib.use_library(cuda_lib)
    n = ctypes.c_int(-1)
    rc = pkg._lib.lib().pcb_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(pkg.PcbError):
        pkg.devarray.Context(8, 0)


def test_supported_sizes(cuda_lib):
    lib = ctypes.CDLL(cuda_lib)
    buf = (ctypes.c_int * 64)()
    n = lib.pcb_supported_sizes(buf, 64)
    sizes = list(buf[:n])
    for need in (48, 100, 120, 150, 160, 256):
        assert need in sizes
