"""LOBPCG block kernels through the C ABI vs NumPy: Gram pair, fused update, residual+preconditioner, column dots,
layout round trip, device RNG."""
import numpy as np
import pytest

from conftest import relerr


def _blocks(pcb, N, k, seed):
    ctx = pcb.get_context(N)
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((ctx.R, k)) + 1j * rng.standard_normal((ctx.R, k))
    return ctx, a, ctx.from_host(a)


@pytest.mark.parametrize("N,n", [(6, 5), (8, 16), (8, 24), (6, 48), (6, 33), (6, 96)])
def test_gram_pair(pcb, N, n):
    ctx, s, S = _blocks(pcb, N, n, 1)
    # HS = D S with a real diagonal D: like H S in LOBPCG, S^H HS is Hermitian (pcb_gram2 computes one triangle)
    d = np.random.default_rng(2).standard_normal((ctx.R, 1))
    hs = d * s
    HS = ctx.from_host(hs)
    G, T = pcb.orthogonalization.gram_pair(S, HS)
    g = s.conj().T @ s
    t = s.conj().T @ hs
    assert relerr(G, (g + g.conj().T) / 2) < 1e-13
    assert relerr(T, (t + t.conj().T) / 2) < 1e-13


@pytest.mark.parametrize("N,m,n_act,first", [(6, 16, 16, False), (6, 16, 5, False), (8, 8, 3, False), (6, 16, 7, True), (6, 32, 20, False), (6, 20, 20, False)])
def test_update(pcb, N, m, n_act, first):
    """_sep_update_after_rr (lobpcg.py:1248-1270) incl. the soft-lock column views and in-place P."""
    import ctypes as C
    L = pcb._lib
    ctx, s, S = _blocks(pcb, N, 3 * m, 3)
    _, hs, HS = _blocks(pcb, N, 3 * m, 4)
    rng = np.random.default_rng(5)
    act = np.sort(rng.choice(m, n_act, replace=False))
    n_loc = m + (1 if first else 2) * n_act
    E = np.ascontiguousarray(rng.standard_normal((n_loc, m)) + 1j * rng.standard_normal((n_loc, m)))
    X, W, P = S[:, :m], S[:, m:2 * m], S[:, 2 * m:]
    HX, HW, HP = HS[:, :m], HS[:, m:2 * m], HS[:, 2 * m:]
    sl = X.ptrs + W.cols(act).ptrs + ([] if first else P.cols(act).ptrs)
    hl = HX.ptrs + HW.cols(act).ptrs + ([] if first else HP.cols(act).ptrs)
    L.check(L.lib().pcb_update(ctx.h, m, n_loc, L.ptr_array(sl), L.ptr_array(hl), L.ptr_array(P.ptrs), L.ptr_array(HP.ptrs),
                               E.ctypes.data), "pcb_update")
    for a, A in ((s, S), (hs, HS)):
        cols = [a[:, m + act]] + ([] if first else [a[:, 2 * m + act]])
        pn = np.concatenate(cols, axis=1) @ E[m:]
        xn = a[:, :m] @ E[:m] + pn
        got = A.get()
        assert relerr(got[:, :m], xn) < 1e-13
        assert relerr(got[:, 2 * m:], pn) < 1e-13
        assert np.array_equal(got[:, m:2 * m], a[:, m:2 * m])     # W untouched


def test_residual_precond_and_dots(pcb, oracle):
    N, k = 8, 11
    alpha = np.array([0.3, 0.0, 0.1])
    ne, mfd = pcb.numerical_experiments, pcb.discretization
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info("fcc", option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), None, inv_fft, relax[0])
    ctx, x, X = _blocks(pcb, N, k, 7)
    _, hx, HX = _blocks(pcb, N, k, 8)
    lam = np.linspace(0.5, 3.0, k)
    Wd = ctx.empty(k)
    nr = H.op.residual(X, HX, Wd, lam, precond=True)
    r = x * lam - hx
    assert np.allclose(nr, np.linalg.norm(r, axis=0), rtol=1e-13)
    ao, bo, io, sh, _ = oracle.assemble_symbols(N, "fcc", alpha)
    assert relerr(Wd.get(), oracle.h_block(r, io)) < 1e-12
    nr2 = H.op.residual(X, HX, Wd, lam, precond=False)
    assert relerr(Wd.get(), r) < 1e-15 and np.allclose(nr2, nr, rtol=1e-14)
    d = pcb.pcfft.column_dots(X, HX)
    assert relerr(d, np.einsum("ij,ij->j", x.conj(), hx)) < 1e-13
    assert np.allclose(pcb.environment.norms(X), np.linalg.norm(x, axis=0), rtol=1e-13)


def test_layout_roundtrip_views_and_rng(pcb):
    ctx, a, A = _blocks(pcb, 6, 7, 9)
    assert np.array_equal(A.get(), a)
    assert np.array_equal(A[:, 2:5].get(), a[:, 2:5])
    assert np.array_equal(A[:, [6, 0]].get(), a[:, [6, 0]])
    assert np.array_equal(A[:, 3].get(), a[:, 3])
    B = ctx.empty(7)
    B[:, :] = A
    B[:, 1:3] = a[:, 4:6]
    want = a.copy(); want[:, 1:3] = a[:, 4:6]
    assert np.array_equal(B.get(), want)
    r1, r2, r3 = ctx.random_block(3, 42).get(), ctx.random_block(3, 42).get(), ctx.random_block(3, 43).get()
    assert np.array_equal(r1, r2) and not np.array_equal(r1, r3)
    assert 0 <= r1.real.min() and r1.real.max() < 1 and 0 <= r1.imag.min() and r1.imag.max() < 1
    assert abs(r1.real.mean() - 0.5) < 0.05 and abs(np.corrcoef(r1.real[:, 0], r1.imag[:, 0])[0, 1]) < 0.1


def test_fft_roundtrip_and_scipy(pcb):
    import scipy.fft as sfft
    for N in ([6, 8, 12, 16] if pcb.backend_name == "emu" else [6, 8, 12, 16, 24, 32, 48, 64, 72, 80, 96, 100]):
        ctx, a, A = _blocks(pcb, N, 2, N)
        F = pcb.pcfft.fftn3(A)
        want = np.empty_like(a)
        for j in range(2):
            for c in range(3):
                v = a[c * N ** 3:(c + 1) * N ** 3, j].reshape(N, N, N)          # [i2, i1, i0]
                want[c * N ** 3:(c + 1) * N ** 3, j] = sfft.fftn(v).ravel()
        assert relerr(F.get(), want) < 1e-13, N
        back = pcb.pcfft.fftn3(F, inverse=True)
        assert relerr(back.get(), a) < 1e-13, N


def test_short_qr_and_gep_chol(pcb):
    """CholQR of a block on the device (orthogonalization.py:36-46) and the small GEP helper (:99-115)."""
    ctx, a, A = _blocks(pcb, 6, 9, 21)
    Q = pcb.orthogonalization.short_qr(A)
    q = Q.get()
    assert relerr(q.conj().T @ q, np.eye(9)) < 1e-12
    l = np.linalg.cholesky((a.conj().T @ a + (a.conj().T @ a).conj().T) / 2)
    assert relerr(q, a @ np.linalg.inv(l.conj().T)) < 1e-11
    rng = np.random.default_rng(3)
    B = rng.standard_normal((7, 7)) + 1j * rng.standard_normal((7, 7))
    G = B @ B.conj().T + 7 * np.eye(7)
    T = (B + B.conj().T) / 2
    lam, vec, _ = pcb.orthogonalization.GEP_chol(T, G)
    assert np.allclose(T @ vec, G @ vec * lam, atol=1e-10)
    import scipy.linalg
    assert np.allclose(lam, scipy.linalg.eigh(T, G, eigvals_only=True), atol=1e-11)


@pytest.mark.parametrize("N,n,ntop", [(6, 48, 16), (6, 48, 5), (8, 24, 8), (6, 33, 20), (6, 96, 32)])
def test_gram_pair_top_rows(pcb, N, n, ntop):
    """pcb_gram2_top: only the leading rows are accumulated; rows/columns < ntop equal the full Gram pair."""
    ctx, s, S = _blocks(pcb, N, n, 11)
    d = np.random.default_rng(12).standard_normal((ctx.R, 1))
    hs = d * s
    HS = ctx.from_host(hs)
    G, T = pcb.orthogonalization.gram_pair_top(S, HS, ntop)
    g, t = s.conj().T @ s, s.conj().T @ hs
    assert relerr(G[:ntop, :], g[:ntop, :]) < 1e-13 and relerr(G[:, :ntop], g[:, :ntop]) < 1e-13
    assert relerr(T[:ntop, :], t[:ntop, :]) < 1e-13 and relerr(T[:, :ntop], t[:, :ntop]) < 1e-13


def test_block_allocation_cache(pcb):
    """Context recycles freed blocks of the same size (a band-structure run drops and re-allocates GB blocks per k-point);
    trim() hands them back to the driver, and data written through a recycled block is what comes back."""
    import gc
    ctx = pcb.get_context(6)
    ctx.trim()
    a = ctx.from_host(np.full((ctx.R, 3), 1.0 + 2.0j))
    p0 = a.ptrs[0]
    nbytes = 16 * ctx.R * 3
    del a
    gc.collect()
    if nbytes >= (1 << 20):          # only blocks of at least 1 MiB are kept
        assert ctx._alloc_cached == nbytes
    b = ctx.from_host(np.full((ctx.R, 3), 3.0 - 1.0j))
    if nbytes >= (1 << 20):
        assert b.ptrs[0] == p0 and ctx._alloc_cached == 0
    assert np.all(b.get() == 3.0 - 1.0j)
    del b
    gc.collect()
    freed = ctx.trim()
    assert ctx._alloc_cached == 0 and (freed == nbytes or nbytes < (1 << 20))
    big = ctx.empty(400)               # 400 columns of N = 6: above 1 MiB, exercises the cached path on every backend
    p1 = big.ptrs[0]
    del big
    gc.collect()
    assert ctx._alloc_cached == 16 * ctx.R * 400
    again = ctx.empty(400)
    assert again.ptrs[0] == p1
    del again
    gc.collect()
    assert ctx.trim() == 16 * ctx.R * 400


def test_gram_top_with_the_real_operator(pcb, oracle):
    """The leading-rows Gram pair forms T as (H W)^H S instead of W^H (H S): with the actual Hermitian operator (FFT passes,
    dielectric, penalty term) both agree to rounding, which is what the incremental Gram update of the solver relies on."""
    N, d = 8, "sc_curv"
    mfd, ne = pcb.discretization, pcb.numerical_experiments
    alpha = np.array([np.pi, 0.4, -1.1])
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), mfd.chiral_handle(N, d), inv_fft, relax[0])
    ctx = H.op.ctx
    n, ntop = 24, 8
    s = oracle.random_x0(3 * N ** 3, n, 21)
    S = ctx.from_host(s)
    HS = ctx.empty(n)
    H.op.apply_into(pcb._lib.APPLY_H, S, HS)
    hs = HS.get()
    G, T = pcb.orthogonalization.gram_pair_top(S, HS, ntop)
    g, t = s.conj().T @ s, s.conj().T @ hs
    assert relerr(G[:ntop, :], g[:ntop, :]) < 1e-13
    assert relerr(T[:ntop, :], t[:ntop, :]) < 1e-12
    Gf, Tf = pcb.orthogonalization.gram_pair(S, HS)
    assert relerr(Tf, (t + t.conj().T) / 2) < 1e-12


@pytest.mark.parametrize("N,m,n_act,first", [(6, 16, 16, False), (8, 16, 5, False), (6, 10, 7, True), (8, 8, 3, False), (6, 32, 9, False)])
def test_update_fused_with_residual(pcb, oracle, N, m, n_act, first):
    """pcb_update_resid: the update (lobpcg.py:1248-1270) plus the NEXT residual / norms / K_P^-1 (:394-397,442) in one pass
    (k_update_res for m <= 16; two kernels for wider blocks) against NumPy + the oracle's preconditioner."""
    alpha = np.array([0.3, 0.0, 0.1])
    ne, mfd = pcb.numerical_experiments, pcb.discretization
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info("fcc", option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), None, inv_fft, relax[0])
    ctx, s, S = _blocks(pcb, N, 3 * m, 3)
    _, hs, HS = _blocks(pcb, N, 3 * m, 4)
    rng = np.random.default_rng(5)
    act = np.sort(rng.choice(m, n_act, replace=False))
    n_loc = m + (1 if first else 2) * n_act
    E = np.ascontiguousarray(rng.standard_normal((n_loc, m)) + 1j * rng.standard_normal((n_loc, m))) / np.sqrt(n_loc)
    lam = np.linspace(0.5, 3.0, m)
    X, W, Pb = S[:, :m], S[:, m:2 * m], S[:, 2 * m:]
    HX, HW, HP = HS[:, :m], HS[:, m:2 * m], HS[:, 2 * m:]
    DB = pcb.devarray.DeviceBlock
    s_loc = DB(ctx, _owners=S._owners, _ptrs=X.ptrs + W.cols(act).ptrs + ([] if first else Pb.cols(act).ptrs))
    hs_loc = DB(ctx, _owners=HS._owners, _ptrs=HX.ptrs + HW.cols(act).ptrs + ([] if first else HP.cols(act).ptrs))
    nr = H.op.update_resid(m, n_loc, s_loc, hs_loc, Pb, HP, E, lam, W)
    new = {}
    for key, a in (("s", s), ("hs", hs)):
        cols = [a[:, m + act]] + ([] if first else [a[:, 2 * m + act]])
        pn = np.concatenate(cols, axis=1) @ E[m:]
        new[key] = (a[:, :m] @ E[:m] + pn, pn)
    got_s, got_hs = S.get(), HS.get()
    assert relerr(got_s[:, :m], new["s"][0]) < 1e-13 and relerr(got_s[:, 2 * m:], new["s"][1]) < 1e-13
    assert relerr(got_hs[:, :m], new["hs"][0]) < 1e-13 and relerr(got_hs[:, 2 * m:], new["hs"][1]) < 1e-13
    assert np.array_equal(got_hs[:, m:2 * m], hs[:, m:2 * m])          # HW untouched
    r = new["s"][0] * lam - new["hs"][0]
    assert np.allclose(nr, np.linalg.norm(r, axis=0), rtol=1e-12)
    ao, bo, io, sh, _ = oracle.assemble_symbols(N, "fcc", alpha)
    assert relerr(got_s[:, m:2 * m], oracle.h_block(r, io)) < 1e-11      # W = K_P^-1 r
