"""Rayleigh-Ritz with Cholesky whitening (public surface of paper_2/orthogonalization.py used by the
default solver: hermitize :26-33, rayleigh_ritz_chol_sep :140-154, short_qr :36-46).

The two tall-skinny Gram products -- the only O(R) part -- run in one fused pass on the GPU
(pcb_gram2: S and HS are read once, Hermitian half only).  The remaining n_loc x n_loc dense algebra
(n_loc <= 96) is the library call the north star leaves on the host: LAPACK potrf / trtri / zheevd
through NumPy."""
import time

import numpy as np

from . import _lib as L
from .devarray import DeviceBlock


def hermitize(M):
    """(M + M^H)/2 (orthogonalization.py:26-33)."""
    return (M + M.T.conj()) / 2


def gram_pair(s, hs):
    """(hermitize(S^H S), hermitize(S^H HS)) for DeviceBlocks, as host arrays."""
    n = s.k
    G = np.empty((n, n), dtype=np.complex128)
    T = np.empty((n, n), dtype=np.complex128)
    L.check(L.lib().pcb_gram2(s.ctx.h, n, L.ptr_array(s.ptrs), L.ptr_array(hs.ptrs), G.ctypes.data, T.ctypes.data), "pcb_gram2")
    return G, T


RANK_TOL = 1e-14      # relative eigenvalue threshold of the rank-revealing fallback below


def gram_pair_top(s, hs, ntop):
    """Rows of the first `ntop` columns of the Gram pair (pcb_gram2_top): (G, T) with rows/columns < ntop valid."""
    n = s.k
    G = np.empty((n, n), dtype=np.complex128)
    T = np.empty((n, n), dtype=np.complex128)
    L.check(L.lib().pcb_gram2_top(s.ctx.h, n, int(ntop), L.ptr_array(s.ptrs), L.ptr_array(hs.ptrs), G.ctypes.data, T.ctypes.data),
            "pcb_gram2_top")
    return G, T


def rr_small(ss, shs, min_rank=None):
    """L = inv(chol(G)); eigh(L T L^H); E = L^H V (orthogonalization.py:148-151) on the small matrices.

    Fallback when G is numerically singular (Cholesky breaks down): the reference's CuPy path does not notice the
    breakdown (cuSOLVER potrf reports it through `info`, which cupy.linalg.cholesky ignores) and carries on with a partly
    unfactorised L; NumPy raises.  Typical trigger: the k-point after Gamma, warm-started with Gamma's zero modes (pure
    gradients), where K_P^-1 r is parallel to x up to ~1e-8 and cond(G) ~ 1e16.  Here the search space is instead whitened
    through the eigen-decomposition of G with the null directions (eigenvalue < RANK_TOL * max) removed -- the standard
    rank-revealing Rayleigh-Ritz; E then has fewer columns than rows.  Well-conditioned cases never take this branch."""
    try:
        Lm = np.linalg.inv(np.linalg.cholesky(ss))
    except np.linalg.LinAlgError:
        w, U = np.linalg.eigh(ss)
        keep = w > RANK_TOL * w[-1]
        if min_rank is not None and keep.sum() < min_rank:
            raise
        Wh = U[:, keep] / np.sqrt(w[keep])
        t = Wh.conj().T @ shs @ Wh
        lam, v = np.linalg.eigh(hermitize(t))
        return lam, Wh @ v
    t = (Lm @ shs) @ Lm.conj().T
    lam, v = np.linalg.eigh(t)
    return lam, Lm.conj().T @ v


def rayleigh_ritz_chol_sep(s, hs):
    """(lambda ascending, eigvec, seconds)  (orthogonalization.py:140-154).

    `s`, `hs`: DeviceBlocks (R x n_loc) -> fused device Gram pair; or small host matrices (the
    reference feeds the m x m Gram matrices themselves to RR for its initial lambda, lobpcg.py:379-381)."""
    t0 = time.time()
    if isinstance(s, DeviceBlock):
        ss, shs = gram_pair(s, hs)
    else:
        s, hs = np.asarray(s), np.asarray(hs)
        ss, shs = hermitize(s.conj().T @ s), hermitize(s.conj().T @ hs)
    lam, vec = rr_small(ss, shs)
    return lam, vec, time.time() - t0


def GEP_chol(T, G, herm=True, slice=None):
    """Small dense GEP T v = lambda G v reduced to an SEP by Cholesky (orthogonalization.py:99-115); host matrices.
    Returns (lambdas, eigvec, seconds), optionally the first `slice` pairs."""
    t0 = time.time()
    T, G = np.asarray(T), np.asarray(G)
    Lm = np.linalg.inv(np.linalg.cholesky(G))
    Tw = (Lm @ T) @ Lm.conj().T
    lam, vec = np.linalg.eigh(Tw) if herm else np.linalg.eig(Tw)
    vec = Lm.conj().T @ vec
    dt = time.time() - t0
    if slice is None:
        return lam, vec, dt
    return lam[:slice], vec[:, :slice], dt


def short_qr(x):
    """CholQR: x inv(chol(herm(x^H x)))^H (orthogonalization.py:36-46) on a DeviceBlock, in place; NumPy input is
    uploaded, orthonormalised on the device and returned as a host array."""
    if not isinstance(x, DeviceBlock):
        from .pcfft import _to_device
        xb = _to_device(x)
        short_qr(xb)
        return xb.get()
    G, _ = gram_pair(x, x)
    Linv = np.linalg.inv(np.linalg.cholesky(G))
    E = np.ascontiguousarray(Linv.conj().T)
    return block_times_small(x, E)


def block_times_small(x, E):
    """x <- x @ E for a DeviceBlock x (R x m) and host E (m x m), via the fused update kernel with an empty W/P part."""
    m = x.k
    hx, p_out, hp_out = x.copy(), DeviceBlock(x.ctx, m), DeviceBlock(x.ctx, m)     # HS twin and (zero) P outputs of the kernel
    E = np.ascontiguousarray(E, dtype=np.complex128)
    L.check(L.lib().pcb_update(x.ctx.h, m, m, L.ptr_array(x.ptrs), L.ptr_array(hx.ptrs), L.ptr_array(p_out.ptrs),
                               L.ptr_array(hp_out.ptrs), E.ctypes.data), "pcb_update")
    x.ctx.sync()          # the temporaries may be released once the kernel has run
    return x
