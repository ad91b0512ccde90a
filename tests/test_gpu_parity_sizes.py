"""GPU parity at the sizes the benchmark and the BASELINE configs run at (VERDICT r01 "pin the configuration you benchmark").

  * H / A / P at N = 120 for every dielectric type vs the oracle at 1e-12 -- the kernel instantiations bench.py times
    (plane-mode passes for N = 120, the Paper-2 dielectric variants) on the BASELINE lattices;
  * one column at N = 160 and N = 256: 3-D FFT vs scipy.fft and the vacuum / isotropic operator vs the oracle;
  * every large FFT plan vs scipy.fft directly (a round trip cannot see a permutation shared by forward and inverse);
  * BASELINE configs[0] (C1: sc_curv, N = 48, 10 bands, rng(0) x0) against the golden generated from the unmodified
    reference (oracle/make_golden_c1.py): identical iteration count, eigenvalues within 1e-10;
  * Gram pair / leading-rows Gram / fused update vs NumPy at N = 48 (R = 331 776: thousands of row tiles per CTA, so the
    cp.async multi-stage steady state of the block kernels is what is checked), m = 16 and m = 32.
"""
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
    yield pkg
    pkg.devarray.drop_contexts()
    pkg.lobpcg._helpers.clear()


def _ops(pcb, N, d_flag, alpha, typ, eps_opt=0):
    mfd, ne = pcb.discretization, pcb.numerical_experiments
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    Diels = None if typ is None else getattr(mfd, typ + "_handle")(N, d_flag, eps_opt=eps_opt)
    return ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0]), relax[0]


def _oracle_ops(pcb, oc, N, d_flag, alpha, typ, eps_opt=0):
    """Oracle operator; the index sets come from the product's chunked geometry code (bit-identical to oracle.diel_index,
    test_index_sets_match_oracle / test_index_sets_n120) because the oracle's dense evaluation needs minutes at N = 120."""
    a, b, inv, shift, _ = oc.assemble_symbols(N, d_flag, alpha)
    if typ is None:
        diel = lambda v: v      # noqa: E731
    else:
        kw = {"ind_e": pcb.dielectric.compute_index(N, d_flag, "edge")}
        if typ == "pseudochiral_trivial":
            kw["ind_v"] = pcb.dielectric.compute_index(N, d_flag, "volume")
        diel = oc.HANDLES[typ](N, d_flag, eps_opt=eps_opt, **kw)
    return oc.pc_mfd_handle(a, b, diel, inv, shift)


@pytest.mark.parametrize("N,d_flag,typ,alpha,cols", [
    (120, "fcc", "chiral", [0.3 * np.pi, 2 * np.pi, 0.0], 3),              # bench.py's workload (BASELINE configs[1])
    (120, "sc_curv", "chiral", [np.pi, np.pi, np.pi], 2),
    (120, "bcc_sg", "pseudochiral_trivial", [np.pi / 20, 0.0, 0.0], 2),    # BASELINE configs[2]
    (120, "bcc_sg", "pseudochiral_crossdof", [np.pi, 0.5 * np.pi, 0.0], 2),
    (120, "fcc", None, [np.pi, 0.4, 0.0], 1),
    (96, "bcc_dg", "pseudochiral_trivial", [np.pi, np.pi, 0.0], 2),
    (64, "bcc_sg", "pseudochiral_crossdof", [0.0, 0.0, 0.0], 2),
    (160, "bcc_dg", "pseudochiral_trivial", [np.pi, 0.0, np.pi], 1),       # BASELINE configs[3]
    (160, "bcc_dg", "pseudochiral_crossdof", [np.pi, 0.0, np.pi], 1),
    (256, "sc_curv", "chiral", [np.pi, np.pi, np.pi], 1),                  # BASELINE configs[4]
    (160, "fcc", "chiral", [0.3 * np.pi, 2 * np.pi, 0.0], 2),              # z-split plane mode (N = 128, 144, 160): isotropic M ...
    (160, "sc_curv", None, [np.pi, 0.4, 0.0], 1),
    (144, "fcc", "chiral", [np.pi, np.pi, 0.0], 1),
    (144, "bcc_sg", "pseudochiral_crossdof", [np.pi, 0.5 * np.pi, 0.0], 1),   # ... and the cross-DoF halves on half planes
    (128, "sc_curv", "chiral", [np.pi, np.pi, np.pi], 2),
    (128, "bcc_dg", "pseudochiral_crossdof", [0.0, 0.0, 2 * np.pi], 1),
    (144, "bcc_sg", "pseudochiral_trivial", [np.pi / 20, 0.0, 0.0], 1),      # ... and the coupled 3x3 M on clusters of three half-plane CTAs
    (128, "bcc_dg", "pseudochiral_trivial", [np.pi, np.pi, 0.0], 2),
])
def test_operator_vs_oracle_at_baseline_sizes(gpu, oracle, N, d_flag, typ, alpha, cols):
    alpha = np.array(alpha, dtype=float)
    oracle.FFT_WORKERS = os.cpu_count() or 1
    (A, H, P), _ = _ops(gpu, N, d_flag, alpha, typ)
    Ao, Ho, Po = _oracle_ops(gpu, oracle, N, d_flag, alpha, typ)
    x = oracle.random_x0(3 * N ** 3, cols, N + cols)
    assert relerr(H(x), Ho(x)) < 1e-12
    if N <= 160:
        assert relerr(A(x), Ao(x)) < 1e-12
    if N <= 120:
        assert relerr(P(x), Po(x)) < 1e-12


@pytest.mark.parametrize("N,d_flag,typ", [(64, "fcc", "chiral"), (96, "bcc_sg", "pseudochiral_crossdof"), (96, "sc_curv", None),
                                         (96, "bcc_sg", "pseudochiral_trivial"), (64, "bcc_dg", "pseudochiral_trivial")])
def test_z_split_plane_mode_at_plane_sizes(gpu, oracle, N, d_flag, typ):
    """The z-split form of the plane mode forced on where the whole-plane form is the default: both match the oracle."""
    alpha = np.array([0.3 * np.pi, 2 * np.pi, 0.0])
    oracle.FFT_WORKERS = os.cpu_count() or 1
    ctx = gpu.get_context(N)
    Ao, Ho, Po = _oracle_ops(gpu, oracle, N, d_flag, alpha, typ)
    x = oracle.random_x0(3 * N ** 3, 2, N)
    want_a, want_h = Ao(x), Ho(x)
    out = {}
    try:
        for split in (1, 0):
            ctx.option("plane_split", split)
            (A, H, P), _ = _ops(gpu, N, d_flag, alpha, typ)
            out[split] = A(x)
            assert relerr(out[split], want_a) < 1e-12
            assert relerr(H(x), want_h) < 1e-12
    finally:
        ctx.option("plane_split", 0)
    assert not np.array_equal(out[0], out[1])


def test_operator_16_columns_n120_matches_column_by_column(gpu, oracle):
    """The 16-column block apply bench.py times equals 16 single-column applies bit for bit (every column of a launch takes
    the same code path), and the first column matches the oracle."""
    N = 120
    alpha = np.array([0.3 * np.pi, 2 * np.pi, 0.0])
    (A, H, P), _ = _ops(gpu, N, "fcc", alpha, "chiral")
    ctx = gpu.get_context(N)
    X = ctx.random_block(16, 5)
    Y = H(X)
    y = Y.get()
    for j in (0, 7, 15):
        yj = H(X[:, j:j + 1]).get()
        assert np.array_equal(yj[:, 0], y[:, j])
    oracle.FFT_WORKERS = os.cpu_count() or 1
    Ao, Ho, Po = _oracle_ops(gpu, oracle, N, "fcc", alpha, "chiral")
    x0 = X[:, 0:1].get()
    assert relerr(y[:, 0:1], Ho(x0)) < 1e-12


def test_index_sets_n120(gpu, oracle):
    """Geometry at the benchmark size: the product's index sets equal the oracle's dense evaluation (dielectric.py:201-261)."""
    for d_flag, dofs in (("fcc", "edge"), ("bcc_sg", "volume")):
        assert np.array_equal(gpu.dielectric.compute_index(120, d_flag, dofs), oracle.diel_index(120, d_flag, dofs))


@pytest.mark.parametrize("N", [120, 128, 144, 150, 160, 192, 240, 256])
def test_large_fft_vs_scipy(gpu, N):
    """3-D DFT of the three component grids of one column vs scipy.fft.fftn / ifftn (pcfft.py:149,151)."""
    import scipy.fft as sf
    ctx = gpu.get_context(N)
    rng = np.random.default_rng(N)
    x = rng.standard_normal((3 * N ** 3, 1)) + 1j * rng.standard_normal((3 * N ** 3, 1))
    w = os.cpu_count() or 1
    f = gpu.pcfft.fftn3(x, n=N)
    want = sf.fftn(x.reshape(3, N, N, N), axes=(1, 2, 3), workers=w).reshape(-1, 1)
    assert relerr(f, want) < 1e-13
    b = gpu.pcfft.fftn3(x, n=N, inverse=True)
    want = sf.ifftn(x.reshape(3, N, N, N), axes=(1, 2, 3), workers=w).reshape(-1, 1)
    assert relerr(b, want) < 1e-13


def test_c1_solve_vs_reference_golden(gpu, oracle):
    """BASELINE configs[0] / SURVEY 8(d) C1 from the same x0 as the unmodified reference (tests/golden/c1_golden.json)."""
    with open(os.path.join(ROOT, "tests", "golden", "c1_golden.json")) as f:
        g = json.load(f)
    N, nev, m = g["N"], g["nev"], g["m"]
    alpha = np.array(g["alpha"])
    (A, H, P), shift = _ops(gpu, N, g["d_flag"], alpha, g["type"])
    assert shift == g["shift"]
    x0 = oracle.random_x0(3 * N ** 3, m, g["seed"])
    tr = []
    lam, x, info = gpu.lobpcg.lobpcg_sep_softlock(H, P, x0, nev, tol=g["tol"], history=True, trace=tr)
    assert int(info[0]) == g["iters"]
    want = np.array(g["lambdas"])
    assert np.max(np.abs(lam[:nev] - want[:nev]) / np.abs(want[:nev])) < 1e-10
    assert np.all(tr[-1]["res"][:nev] < g["tol"])
    hist = np.array(g["res_history"])
    assert np.allclose(info[2:], hist, rtol=5e-3)
    wpnt, wre = gpu.numerical_experiments.recompute_normalize_print(lam[:nev].copy(), x[:, :nev], A, shift)
    assert np.max(np.abs(np.asarray(wre) - np.array(g["omega_re"])) / np.array(g["omega_re"])) < 1e-9


@pytest.mark.parametrize("m,n_act", [(16, 16), (16, 5), (32, 32), (32, 11)])
def test_block_kernels_n48_vs_numpy(gpu, m, n_act):
    """pcb_gram2, pcb_gram2_top and pcb_update at R = 331 776 rows against NumPy (orthogonalization.py:143-144, lobpcg.py:1248-1270)."""
    N = 48
    L = gpu._lib
    ctx = gpu.get_context(N)
    rng = np.random.default_rng(100 * m + n_act)
    R = ctx.R
    s = rng.standard_normal((R, 3 * m)) + 1j * rng.standard_normal((R, 3 * m))
    d = rng.standard_normal((R, 1))
    hs = d * s                      # HS = D S with a real diagonal D: S^H HS is Hermitian like S^H (H S)
    S, HS = ctx.from_host(s), ctx.from_host(hs)
    act = np.sort(rng.choice(m, n_act, replace=False))
    X, W, P = S[:, :m], S[:, m:2 * m], S[:, 2 * m:]
    HX, HW, HP = HS[:, :m], HS[:, m:2 * m], HS[:, 2 * m:]
    cols = np.concatenate((np.arange(m), m + act, 2 * m + act))
    n_loc = len(cols)
    DB = gpu.devarray.DeviceBlock
    s_loc = DB(ctx, _owners=S._owners, _ptrs=X.ptrs + W.cols(act).ptrs + P.cols(act).ptrs)
    hs_loc = DB(ctx, _owners=HS._owners, _ptrs=HX.ptrs + HW.cols(act).ptrs + HP.cols(act).ptrs)
    G, T = gpu.orthogonalization.gram_pair(s_loc, hs_loc)
    sl, hl = s[:, cols], hs[:, cols]
    g = sl.conj().T @ sl
    t = sl.conj().T @ hl
    assert relerr(G, (g + g.conj().T) / 2) < 1e-13
    assert relerr(T, (t + t.conj().T) / 2) < 1e-13
    # leading rows only (incremental Gram): column order [W_act | X | P_act]
    s_k = DB(ctx, _owners=S._owners, _ptrs=W.cols(act).ptrs + X.ptrs + P.cols(act).ptrs)
    hs_k = DB(ctx, _owners=HS._owners, _ptrs=HW.cols(act).ptrs + HX.ptrs + HP.cols(act).ptrs)
    Gk, Tk = gpu.orthogonalization.gram_pair_top(s_k, hs_k, n_act)
    ck = np.concatenate((m + act, np.arange(m), 2 * m + act))
    gk = s[:, ck[:n_act]].conj().T @ s[:, ck]
    tk = s[:, ck[:n_act]].conj().T @ hs[:, ck]
    assert relerr(Gk[:n_act], gk) < 1e-13
    assert relerr(Tk[:n_act], tk) < 1e-13
    # fused update
    E = np.ascontiguousarray(rng.standard_normal((n_loc, m)) + 1j * rng.standard_normal((n_loc, m))) / np.sqrt(n_loc)
    L.check(L.lib().pcb_update(ctx.h, m, n_loc, L.ptr_array(s_loc.ptrs), L.ptr_array(hs_loc.ptrs), L.ptr_array(P.ptrs),
                               L.ptr_array(HP.ptrs), E.ctypes.data), "pcb_update")
    for a, Ab in ((s, S), (hs, HS)):
        pn = np.concatenate((a[:, m + act], a[:, 2 * m + act]), axis=1) @ E[m:]
        xn = a[:, :m] @ E[:m] + pn
        got = Ab.get()
        assert relerr(got[:, :m], xn) < 1e-13
        assert relerr(got[:, 2 * m:], pn) < 1e-13
        assert np.array_equal(got[:, m:2 * m], a[:, m:2 * m])


def test_failed_allocation_does_not_poison_later_launches(gpu):
    """a failed cudaMalloc (the case Context.malloc retries after trimming its cache) must leave no sticky error
    behind -- the next launch + sync has to succeed."""
    import ctypes as C
    L = gpu._lib
    ctx = gpu.get_context(24)
    p = C.c_void_p()
    rc = L.lib().pcb_malloc(ctx.h, C.c_size_t(1 << 46), C.byref(p))      # 64 TiB: fails
    assert rc != 0
    X = ctx.random_block(2, 3)       # launches k_fill_uniform
    ctx.sync()                       # raises if the old cudaErrorMemoryAllocation is still pending
    assert np.isfinite(gpu.pcfft.column_norms(X)).all()


@pytest.mark.parametrize("d_flag", ["fcc", "bcc_dg", "sc_curv"])
def test_device_geometry_n120(gpu, d_flag):
    """SURVEY 8f N4 at the benchmark size: Omega_1 evaluated on the GPU equals the host NumPy evaluation (dielectric.py:201-261)
    bit for bit and takes well under half a second (host: 5-8 s for FCC)."""
    N = 120
    gpu.get_context(N)
    ind_e, ind_v = gpu.dielectric.device_index_sets(N, d_flag)
    assert gpu.dielectric.geometry_stats["seconds"] < 0.5
    assert np.array_equal(ind_e, gpu.dielectric.compute_index(N, d_flag, "edge"))
    assert np.array_equal(ind_v, gpu.dielectric.compute_index(N, d_flag, "volume"))


@pytest.mark.parametrize("m,n_act", [(16, 16), (16, 6)])
def test_fused_update_residual_n48(gpu, oracle, m, n_act):
    """k_update_res at R = 331 776 rows (thousands of 3 x 16-cell tiles per CTA: the double-buffered steady state)."""
    import test_block_kernels as tb
    gpu.backend_name = "cuda"
    tb.test_update_fused_with_residual(gpu, oracle, 48, m, n_act, False)


def test_solver_on_z_split_plane_mode_matches_five_pass(gpu, oracle):
    """A whole LOBPCG solve at N = 128 (z-split plane mode: the operator structure BASELINE configs[3] runs at N = 160) from the
    same x0 on the three-pass and on the five-pass operator: identical iteration count, eigenvalues within 1e-10 (north star),
    Hermiticity of the z-split operator to rounding."""
    N, d_flag, typ, nev = 128, "bcc_dg", "pseudochiral_crossdof", 10
    alpha = np.array([np.pi, 0.0, np.pi])
    ctx = gpu.get_context(N)
    x0 = oracle.random_x0(3 * N ** 3, 16, 5)
    out = {}
    try:
        for plane in (1, 0):
            ctx.option("plane", plane)
            (A, H, P), _ = _ops(gpu, N, d_flag, alpha, typ)
            lam, x, info = gpu.lobpcg.lobpcg_sep_softlock(H, P, x0, nev)
            assert lam is not None
            out[plane] = (np.array(lam[:nev]), int(info[0]))
            if plane == 1:
                X, Y = ctx.random_block(2, 3), ctx.random_block(2, 4)
                dots = gpu.pcfft.column_dots
                assert np.allclose(dots(Y, H(X)), np.conj(dots(X, H(Y))), rtol=1e-11)
    finally:
        ctx.option("plane", 1)
    assert out[1][1] == out[0][1]
    assert np.max(np.abs(out[1][0] - out[0][0]) / np.abs(out[0][0])) < 1e-10
