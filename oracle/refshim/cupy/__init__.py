"""NumPy-backed stand-in for the ``cupy`` module (TEST INFRASTRUCTURE ONLY).

The reference (`/root/reference/paper_2/*.py`) imports CuPy unconditionally
(e.g. ``pcfft.py:10-12``, ``lobpcg.py:13``) and CuPy is not installed in this
image.  This shim lets the *unmodified* reference modules run on the CPU so
that ``oracle/make_golden.py`` can generate golden vectors from the real
reference code.  It is never imported by the product package.

Surface covered: SURVEY.md Appendix B.
"""
import sys as _sys
import types as _types

import numpy as _np
from numpy import *  # noqa: F401,F403  (re-export the NumPy namespace)
from numpy import linalg, random  # noqa: F401
import scipy.sparse as sparse  # noqa: F401  (cupy.sparse.coo_matrix / csr_matrix)

from . import cublas  # noqa: F401  (importable submodule: ``from cupy.cublas import gemm``)

float32, float64 = _np.float32, _np.float64
complex64, complex128 = _np.complex64, _np.complex128
bool_ = _np.bool_
newaxis = _np.newaxis
inf = _np.inf


class ndarray(_np.ndarray):
    """numpy.ndarray plus the ``.get()`` device→host method CuPy arrays have."""

    def get(self):
        return _np.asarray(self)


def _as_dev(a):
    return a.view(ndarray) if isinstance(a, _np.ndarray) and not isinstance(a, ndarray) else a


def empty(*args, **kw):
    return _np.empty(*args, **kw).view(ndarray)


def zeros(*args, **kw):
    return _np.zeros(*args, **kw).view(ndarray)


def ones(*args, **kw):
    return _np.ones(*args, **kw).view(ndarray)


def asarray(a, *args, **kw):
    return _as_dev(_np.asarray(a, *args, **kw))


def array(a, *args, **kw):
    return _as_dev(_np.array(a, *args, **kw))


def fromfile(*args, **kw):
    return _as_dev(_np.fromfile(*args, **kw))


def asnumpy(a):
    return _np.asarray(a)


# ---------------------------------------------------------------------------
# cupy.ElementwiseKernel: the reference defines exactly two kernels
# (`_kernels.py:13-41` h_block_complex_kernel, `_kernels.py:43-71`
# a_block_complex_kernel).  ``raw`` arguments are indexed with the *logical*
# C-order flat index, the loop index ``i`` runs over the logical C-order
# elements of the output; rows are ``i / m``.  The bodies are restated with
# NumPy slicing on the logical (3nn, m) arrays.
# ---------------------------------------------------------------------------
class ElementwiseKernel:
    def __init__(self, in_params, out_params, operation, name="kernel", **kw):
        self.name = name

    def __call__(self, *args):
        if self.name == "h_block_complex_kernel":
            X, D0, D1, nn, m, Y = args
            x = _np.asarray(X).reshape(3 * nn, m)
            d0 = _np.asarray(D0).reshape(-1, 1)
            d1 = _np.asarray(D1).reshape(-1, 1)
            x1, x2, x3 = x[:nn], x[nn:2 * nn], x[2 * nn:]
            y = _np.empty((3 * nn, m), dtype=x.dtype)
            y[:nn] = d0[:nn] * x1 + d1[:nn] * x2 + d1[nn:2 * nn] * x3
            y[nn:2 * nn] = _np.conj(d1[:nn]) * x1 + d0[nn:2 * nn] * x2 + d1[2 * nn:] * x3
            y[2 * nn:] = _np.conj(d1[nn:2 * nn]) * x1 + d0[2 * nn:] * x3 + _np.conj(d1[2 * nn:]) * x2
            Y[...] = y.reshape(Y.shape)
            return Y
        if self.name == "a_block_complex_kernel":
            X, D, nn, m, Y = args
            x = _np.asarray(X).reshape(3 * nn, m)
            d = _np.asarray(D).reshape(-1, 1)
            x1, x2, x3 = x[:nn], x[nn:2 * nn], x[2 * nn:]
            y = _np.empty((3 * nn, m), dtype=x.dtype)
            y[:nn] = -d[2 * nn:] * x2 + d[nn:2 * nn] * x3
            y[nn:2 * nn] = d[2 * nn:] * x1 - d[:nn] * x3
            y[2 * nn:] = -d[nn:2 * nn] * x1 + d[:nn] * x2
            Y[...] = y.reshape(Y.shape)
            return Y
        raise NotImplementedError(f"refshim: unknown ElementwiseKernel {self.name!r}")


# ---------------------------------------------------------------------------
# Memory pool / device stubs.
# ---------------------------------------------------------------------------
class _Pool:
    def free_all_blocks(self):
        pass

    def used_bytes(self):
        return 0

    def total_bytes(self):
        return 0


_POOL = _Pool()


def get_default_memory_pool():
    return _POOL


class _Device:
    def __init__(self, *a):
        pass

    def synchronize(self):
        pass

    def use(self):
        pass


cuda = _types.ModuleType("cupy.cuda")
cuda.Device = _Device
cuda.set_allocator = lambda *a, **k: None
cuda.MemoryPool = _Pool
_sys.modules["cupy.cuda"] = cuda
_sys.modules["cupy.sparse"] = sparse
