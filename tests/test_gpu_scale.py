"""GPU-only parity at sizes the emulation cannot reach: oracle comparison at N = 24 ... 100, known answers from the
reference's shipped N = 120 band structures, and size-independent properties at the full BASELINE size."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture
def gpu(request):
    import importlib
    from conftest import PKG
    pkg = importlib.import_module(PKG)
    pkg._lib.use_library(request.getfixturevalue("cuda_lib"))
    assert pkg.backend() == "cuda-sm_100a"
    return pkg


def _ops(pcb, N, d_flag, alpha, typ, eps_opt=0):
    mfd, ne = pcb.discretization, pcb.numerical_experiments
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    Diels = None if typ is None else getattr(mfd, typ + "_handle")(N, d_flag, eps_opt=eps_opt)
    return ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0]), relax[0]


def _oracle_ops(oc, N, d_flag, alpha, typ, eps_opt=0):
    a, b, inv, shift, _ = oc.assemble_symbols(N, d_flag, alpha)
    diel = (lambda v: v) if typ is None else oc.HANDLES[typ](N, d_flag, eps_opt=eps_opt)
    return oc.pc_mfd_handle(a, b, diel, inv, shift)


def _axpby(pcb, ctx, x, y, a, b):
    L = pcb._lib
    L.check(L.lib().pcb_axpby(ctx.h, x.k, L.ptr_array(x.ptrs), L.ptr_array(y.ptrs), float(a), float(b)), "pcb_axpby")


@pytest.mark.parametrize("N,d_flag,typ,alpha", [
    (24, "sc_curv", "chiral", [np.pi, np.pi, np.pi]),
    (24, "bcc_sg", "pseudochiral_trivial", [0.3, 0.0, 2 * np.pi]),
    (24, "fcc", "pseudochiral_crossdof", [np.pi / 20, 0.0, 0.0]),
    (48, "sc_curv", "chiral", [np.pi, np.pi, np.pi]),
    (48, "bcc_dg", "pseudochiral_crossdof", [0.0, 0.0, 0.0]),
    (32, "fcc", "pseudochiral_trivial", [np.pi, 2 * np.pi, 0.0]),
    (64, "sc_curv", "chiral", [np.pi, 0.0, 0.0]),
    (72, "sc_curv", "pseudochiral_trivial", [np.pi, 0.0, 0.0]),
    (80, "sc_flat1", "chiral", [np.pi, np.pi, 0.0]),
    (96, "sc_flat2", "chiral", [np.pi, np.pi, 0.0]),
    (100, "sc_curv", "chiral", [np.pi, np.pi, np.pi]),
])
def test_operator_vs_oracle(gpu, oracle, N, d_flag, typ, alpha):
    alpha = np.array(alpha, dtype=float)
    (A, H, P), _ = _ops(gpu, N, d_flag, alpha, typ)
    Ao, Ho, Po = _oracle_ops(oracle, N, d_flag, alpha, typ)
    x = oracle.random_x0(3 * N ** 3, 3, N)
    assert relerr(H(x), Ho(x)) < 1e-12
    assert relerr(A(x), Ao(x)) < 1e-12
    assert relerr(P(x), Po(x)) < 1e-12


def test_index_sets_match_oracle(gpu, oracle):
    for d_flag in ("sc_flat1", "sc_flat2", "sc_curv", "bcc_sg", "bcc_dg", "fcc"):
        for N in (12, 24, 30):
            for dofs in ("edge", "volume"):
                assert np.array_equal(gpu.dielectric.compute_index(N, d_flag, dofs), oracle.diel_index(N, d_flag, dofs))


@pytest.mark.parametrize("N,d_flag,typ,alpha,seed", [
    (24, "sc_curv", "chiral", [np.pi, np.pi, np.pi], 1),
    (24, "bcc_sg", "pseudochiral_crossdof", [np.pi, np.pi, np.pi], 2),
])
def test_lobpcg_vs_oracle(gpu, oracle, N, d_flag, typ, alpha, seed):
    """Same x0 -> same iteration count, eigenvalues within 1e-10 relative (north star), residuals below tolerance."""
    alpha = np.array(alpha, dtype=float)
    (A, H, P), shift = _ops(gpu, N, d_flag, alpha, typ)
    Ao, Ho, Po = _oracle_ops(oracle, N, d_flag, alpha, typ)
    x0 = oracle.random_x0(3 * N ** 3, 16, seed)
    tr, tro = [], []
    lam, x, info = gpu.lobpcg.lobpcg_sep_softlock(H, P, x0, 10, trace=tr)
    lamo, xo, infoo = oracle.lobpcg_sep_softlock(Ho, Po, x0, 10, trace=tro)
    assert int(info[0]) == int(infoo[0])
    assert np.max(np.abs(lam[:10] - lamo[:10]) / np.abs(lamo[:10])) < 1e-10
    assert np.all(tr[-1]["res"][:10] < 1e-4)
    # eigenvectors: compare the invariant subspaces of the converged bands (phases/rotations are implementation defined)
    q1, _ = np.linalg.qr(x.get()[:, :10])
    q2, _ = np.linalg.qr(xo[:, :10])
    sv = np.linalg.svd(q1.conj().T @ q2, compute_uv=False)
    assert sv.min() > 1 - 1e-5


def test_shipped_band_structure_n120(gpu):
    """Rows of the reference's own published N = 120 band structures (paper_2/output/**/bandgap_*.json, RTX 4090 D).

    Finding (probed on B200, tools/probe_shipped.py): the shipped files were NOT produced by the paper_2 code as it stands --
    e.g. at R = (pi,pi,pi) of sc_curv they hold an exactly threefold-degenerate lowest band (0.38124755 x3) whereas the current
    reference code, run unmodified through oracle/refshim, and this implementation both give a 1 + 2 splitting
    (0.38224, 0.38300 x2), for stencil width k = 1 and k = 2 alike.  They therefore serve as a physics-level known answer
    (same lattice, same eps: bands agree to the discretisation error, a few 1e-3 in omega/2pi), not as a 1e-6 oracle."""
    rows = json.load(open(os.path.join(ROOT, "tests", "golden", "shipped_bands.json")))["rows"]
    ne = gpu.numerical_experiments
    keep = [r for r in rows if (r["type"], r["d_flag"], r["k_index"]) in
            (("chiral", "sc_curv", 59), ("chiral", "sc_curv", 19), ("pseudochiral_trivial", "sc_curv", 59), ("chiral", "bcc_sg", 59))]
    assert len(keep) == 4
    for row in keep:
        alpha = gpu.dielectric.kpath(row["d_flag"])[row["k_index"]]
        res = ne.eigen_1p(120, row["d_flag"], alpha, type=row["type"], nev=10, seed=7 + row["k_index"])
        assert res is not None, row
        want = np.array(row["frequencies"][:10])
        assert np.max(np.abs(res["omega_re"] - want)) < 8e-3, (row["type"], row["d_flag"], row["k_index"], res["omega_re"], want)
        assert np.all(res["residuals"] < 5e-3)


def test_full_size_properties_n120(gpu):
    """Size-independent properties at BASELINE's N = 120: Hermitian H and A; H = A + gamma K_B + shift; K_A^H annihilates the
    range of K_B (B A = 0); FFT round trip and Parseval; P0 H0 = I for the vacuum operator."""
    N = 120
    L = gpu._lib
    alpha = np.array([np.pi / 3, 2 * np.pi, 0.2])
    (A, H, P), shift = _ops(gpu, N, "fcc", alpha, "pseudochiral_trivial")
    ctx = gpu.get_context(N)
    X, Y = ctx.random_block(3, 11), ctx.random_block(3, 12)
    HX, HY, AX = H(X), H(Y), A(X)
    dots, norms = gpu.pcfft.column_dots, gpu.pcfft.column_norms
    assert np.allclose(dots(Y, HX), np.conj(dots(X, HY)), rtol=1e-11)
    assert np.allclose(dots(Y, AX), np.conj(dots(X, A(Y))), rtol=1e-11)
    KB = H.op.apply(L.APPLY_KB, X)
    _axpby(gpu, ctx, KB, AX, 1.0, 1.0)
    _axpby(gpu, ctx, X, AX, shift, 1.0)
    _axpby(gpu, ctx, HX, AX, -1.0, 1.0)
    assert np.all(norms(AX) < 1e-11 * norms(HX))
    Z = H.op.apply(L.APPLY_KAH, KB)
    assert np.all(norms(Z) < 1e-10 * norms(KB) * N)
    F = gpu.pcfft.fftn3(X)
    assert np.allclose(norms(F) ** 2, N ** 3 * norms(X) ** 2, rtol=1e-12)
    B = gpu.pcfft.fftn3(F, inverse=True)
    _axpby(gpu, ctx, X, B, -1.0, 1.0)
    assert np.all(norms(B) < 1e-13 * norms(X))
    (A0, H0, P0), _ = _ops(gpu, N, "fcc", alpha, None)
    PX = P0(H0(X))
    _axpby(gpu, ctx, X, PX, -1.0, 1.0)
    assert np.all(norms(PX) < 1e-9 * norms(X))


@pytest.mark.parametrize("N", [120, 128, 144, 150, 160, 192, 240, 256])
def test_large_sizes_fft_and_vacuum_inverse(gpu, N):
    """Every large plan: FFT round trip + Parseval and P0 H0 = I (exercises all five passes and the symbol tables)."""
    alpha = np.array([np.pi, 0.4, 0.0])
    (A0, H0, P0), _ = _ops(gpu, N, "sc_curv", alpha, None)
    ctx = gpu.get_context(N)
    norms = gpu.pcfft.column_norms
    X = ctx.random_block(1, N)
    F = gpu.pcfft.fftn3(X)
    assert np.allclose(norms(F) ** 2, N ** 3 * norms(X) ** 2, rtol=1e-12)
    B = gpu.pcfft.fftn3(F, inverse=True)
    _axpby(gpu, ctx, X, B, -1.0, 1.0)
    assert np.all(norms(B) < 1e-13 * norms(X))
    del F, B
    PX = P0(H0(X))
    _axpby(gpu, ctx, X, PX, -1.0, 1.0)
    assert np.all(norms(PX) < 1e-9 * norms(X))
