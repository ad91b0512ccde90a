"""``cupy.cublas.gemm`` on NumPy (test infrastructure only; see package docstring).

Functional semantics: the product is computed before it is written to ``out``
(CuPy routes strided slices through temporaries), so ``out`` may alias ``a``.
Reference call sites: ``orthogonalization.py:15,143-144,183``,
``lobpcg.py:1254-1268``.
"""
import numpy as _np


def _op(t, a):
    a = _np.asarray(a)
    if t == "N":
        return a
    if t == "T":
        return a.T
    if t == "H":
        return a.conj().T
    raise ValueError(t)


def gemm(transa, transb, a, b, out=None, alpha=1.0, beta=0.0):
    c = _op(transa, a) @ _op(transb, b)
    if alpha != 1.0:
        c = alpha * c
    if out is None:
        import cupy as _cp
        return c.view(_cp.ndarray)
    if beta != 0.0:
        out[...] = c + beta * _np.asarray(out)
    else:
        out[...] = c
    return out
