"""Test configuration.

Two builds of the SAME kernel sources are exercised through the SAME C ABI and Python host code:
  * backend "cuda": csrc/libpcb200.so on a real B200 -- tests marked ``gpu`` (the parity tests proper);
  * backend "emu":  tests/emu/_build/libpcb200_emu.so, the kernels compiled as host C++ against
    tests/emu/emu_cuda.h (fibers emulate the CUDA block/warp model).  CPU-only test infrastructure
    for index maps / staging / reductions / host logic at tiny N; never loaded by the product.
The oracle (oracle/pc_oracle.py) and the golden fixtures generated from the unmodified reference
(tests/golden, oracle/make_golden.py) are the checkers.
"""
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "linear-eigenvalue-problems-in-photonic-crystals_b200"
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("PCB200_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


def _package():
    return importlib.import_module(PKG)


def _build_module():
    spec = importlib.util.spec_from_file_location("pcb200_build", os.path.join(ROOT, PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


EMU_SIZES = [6, 8, 12, 16, 24, 32]


@pytest.fixture(scope="session")
def emu_lib():
    return _build_module().build(emu=True, sizes=EMU_SIZES, verbose=False)


@pytest.fixture(scope="session")
def cuda_lib():
    path = os.path.join(ROOT, PKG, "csrc", "libpcb200.so")
    if not os.path.exists(path):
        path = _build_module().build(verbose=False)
    return path


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", marks=pytest.mark.gpu, id="cuda")]


@pytest.fixture(params=BACKENDS)
def pcb(request):
    """The package bound to one backend."""
    pkg = _package()
    if request.param == "emu":
        pkg._lib.use_library(request.getfixturevalue("emu_lib"))
    else:
        pkg._lib.use_library(request.getfixturevalue("cuda_lib"))
        assert pkg.backend() == "cuda-sm_100a"
    pkg.backend_name = request.param
    yield pkg
    pkg.devarray.drop_contexts()
    pkg.lobpcg._helpers.clear()


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        man = json.load(f)
    return z, man


@pytest.fixture(scope="session")
def oracle():
    import pc_oracle
    return pc_oracle


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
