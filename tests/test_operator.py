"""Operator parity: CUDA path (through the C ABI) vs the golden outputs of the unmodified reference
and vs the oracle on seeded inputs."""
import numpy as np
import pytest

from conftest import relerr

TOL = 1e-12   # relative, complex128: FFT + symbol rounding differences only


def _setup(pcb, oracle, case):
    N, d_flag, alpha, typ = case["N"], case["d_flag"], np.array(case["alpha"]), case["type"]
    ne, mfd = pcb.numerical_experiments, pcb.discretization
    relax, pnt = mfd.set_relaxation(alpha)
    ct = pcb.dielectric.diel_info(d_flag, option="ct")
    a_fft, b_fft = mfd.fft_blocks(N, 1, ct, alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    b_fft = (pnt * b_fft[0], pnt * b_fft[1])
    Diels = None if typ is None else getattr(mfd, typ + "_handle")(N, d_flag, eps_opt=case["eps_opt"])
    A, H, P = ne.pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, relax[0])
    x = oracle.random_x0(3 * N ** 3, case["m"], case["seed"])
    return A, H, P, Diels, x


def test_operator_vs_reference_golden(pcb, oracle, golden):
    z, man = golden
    for case in man["operator"]:
        if pcb.backend_name == "emu" and case["N"] > 8:
            continue
        A, H, P, Diels, x = _setup(pcb, oracle, case)
        key = case["key"]
        assert relerr(H(x), z[key + "_H"]) < TOL, key
        assert relerr(A(x), z[key + "_A"]) < TOL, key
        assert relerr(P(x), z[key + "_P"]) < TOL, key
        if Diels is not None:
            assert relerr(Diels(x), z[key + "_M"]) < TOL, key


def test_symbols_vs_reference_golden(pcb, golden):
    z, man = golden
    mfd = pcb.discretization
    for case in man["symbols"]:
        N, alpha = case["N"], np.array(case["alpha"])
        ct = pcb.dielectric.diel_info(case["d_flag"], option="ct")
        relax, pnt = mfd.set_relaxation(alpha)
        assert relax[0] == pytest.approx(case["shift"], rel=0, abs=0)
        assert pnt == pytest.approx(case["gamma"], rel=1e-15)
        a_fft, b_fft = mfd.fft_blocks(N, 1, ct, alpha=alpha)
        k = case["key"]
        assert relerr(a_fft.toarray(), z[k + "_a"]) < 1e-14
        b0, b1 = mfd.PenaltySymbols(a_fft, pnt).toarray()
        assert relerr(b0, z[k + "_b0"]) < 1e-14 and relerr(b1, z[k + "_b1"]) < 1e-14


@pytest.mark.parametrize("N,d_flag,typ,alpha", [
    (8, "sc_curv", "chiral", [np.pi, np.pi, np.pi]),          # N % 8 == 0: three-pass plane mode (fused y/z pass)
    (16, "fcc", "chiral", [0.3, 2 * np.pi, 0.0]),
    (16, "bcc_sg", None, [0.0, 0.0, 0.0]),
    (16, "sc_curv", "pseudochiral_trivial", [np.pi, 0.0, 0.0]),   # coupled 3x3 M: five-pass path
    (12, "fcc", "chiral", [np.pi, np.pi, 0.0]),                    # N % 8 != 0: five-pass path
    (8, "bcc_sg", "pseudochiral_crossdof", [0.3, 0.0, 2 * np.pi]),  # plane halves around the stencil on the slot layout (Cooley-Tukey slots)
    (16, "fcc", "pseudochiral_crossdof", [np.pi, 0.0, 0.0]),
    (24, "bcc_dg", "pseudochiral_crossdof", [np.pi, np.pi, 0.0]),   # ... Good-Thomas slots (24 = 8 x 3, like 120 = 8 x 15)
    (24, "sc_curv", "chiral", [np.pi, np.pi, np.pi]),
    (24, "bcc_sg", "pseudochiral_trivial", [0.0, 0.0, 0.0]),        # CUDA: clusters of three CTAs (24 % 3 == 0)
])
def test_operator_vs_oracle_small(pcb, oracle, N, d_flag, typ, alpha):
    alpha = np.array(alpha, dtype=float)
    case = {"N": N, "d_flag": d_flag, "alpha": alpha, "type": typ, "eps_opt": 0, "m": 2, "seed": 31 + N}
    A, H, P, Diels, x = _setup(pcb, oracle, case)
    a, b, inv, shift, _ = oracle.assemble_symbols(N, d_flag, alpha)
    diel = (lambda v: v) if typ is None else oracle.HANDLES[typ](N, d_flag)
    Ao, Ho, Po = oracle.pc_mfd_handle(a, b, diel, inv, shift)
    assert relerr(H(x), Ho(x)) < TOL
    assert relerr(A(x), Ao(x)) < TOL


def test_dropin_symbol_multiplies_and_fft(pcb, oracle):
    """The reference-named point-wise functions: A_block (K_A, K_A^H), H_block (gamma K_B, K_P^-1) and the batched FFT."""
    N, d_flag = 8, "fcc"
    alpha = np.array([0.4, 2 * np.pi, 0.1])
    mfd, pcfft = pcb.discretization, pcb.pcfft
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    ao, bo, io, shift, gamma = oracle.assemble_symbols(N, d_flag, alpha)
    x = oracle.random_x0(3 * N ** 3, 3, 77)
    assert relerr(pcfft.A_block(x, a_fft), oracle.a_block(x, ao)) < TOL
    assert relerr(pcfft.A_block_kernel(x, -a_fft.conj()), oracle.a_block(x, -ao.conj())) < TOL
    assert relerr(pcfft.H_block_kernel(x, (pnt * b_fft[0], pnt * b_fft[1])), oracle.h_block(x, bo)) < TOL
    assert relerr(pcfft.H_block(x, inv_fft), oracle.h_block(x, io)) < TOL
    v = x[:, 0]                                              # 1-D vectors keep their shape
    assert pcfft.H_block(v, inv_fft).shape == v.shape
    assert relerr(pcfft.AMA_BB(v, a_fft, (pnt * b_fft[0], pnt * b_fft[1]), None, relax[0]),
                  oracle.AMA_BB(v, ao, bo, lambda u: u, shift)) < TOL
    import scipy.fft as sfft
    F = pcfft.fftn3(x)
    want = np.concatenate([sfft.fftn(x[c * N ** 3:(c + 1) * N ** 3].reshape(N, N, N, 3), axes=(0, 1, 2)).reshape(N ** 3, 3) for c in range(3)])
    assert relerr(F, want) < TOL


@pytest.mark.parametrize("typ", ["chiral", "pseudochiral_crossdof"])
def test_fourth_order_stencils_k2(pcb, oracle, typ):
    """Stencil half-width k = 2 (4-point mimetic stencils, discretization.py:152-193; cross-DoF averaging with 4 taps)."""
    N, d_flag, k = 12, "sc_curv", 2
    alpha = np.array([np.pi, 0.3, 0.0])
    mfd, ne = pcb.discretization, pcb.numerical_experiments
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, k, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    Diels = getattr(mfd, typ + "_handle")(N, d_flag, k=k)
    A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
    ao, bo, io, shift, _ = oracle.assemble_symbols(N, d_flag, alpha, k=k)
    diel = oracle.HANDLES[typ](N, d_flag, k=k) if typ == "pseudochiral_crossdof" else oracle.HANDLES[typ](N, d_flag)
    Ao, Ho, Po = oracle.pc_mfd_handle(ao, bo, diel, io, shift)
    x = oracle.random_x0(3 * N ** 3, 2, 5)
    assert relerr(H(x), Ho(x)) < TOL
    assert relerr(P(x), Po(x)) < TOL
    assert relerr(Diels(x), diel(x)) < TOL


@pytest.mark.parametrize("d_flag", ["sc_flat1", "sc_flat2", "sc_curv", "bcc_sg", "bcc_dg", "fcc"])
def test_device_geometry_matches_host(pcb, oracle, d_flag):
    """Omega_1 classified on the device (pcb_geometry_mask + host arbitration of the boundary points) == the NumPy evaluation
    of dielectric.py:104-261 bit for bit, for edge and volume DoFs."""
    sizes = (8, 12, 16) if pcb.backend_name == "emu" else (16, 24, 48, 100)
    for N in sizes:
        ind_e, ind_v = pcb.dielectric.device_index_sets(N, d_flag)
        assert np.array_equal(ind_e, pcb.dielectric.compute_index(N, d_flag, "edge")), (d_flag, N)
        assert np.array_equal(ind_v, pcb.dielectric.compute_index(N, d_flag, "volume")), (d_flag, N)
        if N <= 24:
            assert np.array_equal(ind_e, oracle.diel_index(N, d_flag, "edge"))


@pytest.mark.parametrize("typ", ["chiral", None, "pseudochiral_trivial"])
def test_five_sweep_plane_pass(pcb, oracle, typ):
    """k_mid2 (2-D register tiles, five shared-memory sweeps; the default at N = 120) forced on at N = 24 = 8 x 3, where the
    emulation can run it, against the oracle; the seven-sweep pass gives the same result to rounding."""
    N, d_flag = 24, "fcc"
    alpha = np.array([0.3 * np.pi, 2 * np.pi, 0.0])
    ctx = pcb.get_context(N)
    case = {"N": N, "d_flag": d_flag, "alpha": alpha, "type": typ, "eps_opt": 3 if typ == "pseudochiral_trivial" else 0, "m": 2, "seed": 9}
    a, b, inv, shift, _ = oracle.assemble_symbols(N, d_flag, alpha)
    diel = (lambda v: v) if typ is None else oracle.HANDLES[typ](N, d_flag, eps_opt=case["eps_opt"])
    Ao, Ho, Po = oracle.pc_mfd_handle(a, b, diel, inv, shift)
    out = {}
    try:
        for five in (1, 0):
            ctx.option("mid_five", five)
            A, H, P, Diels, x = _setup(pcb, oracle, case)
            out[five] = H(x)
            assert relerr(out[five], Ho(x)) < TOL
            assert relerr(A(x), Ao(x)) < TOL
    finally:
        ctx.option("mid_five", -1)
    assert relerr(out[1], out[0]) < 1e-13
    if not (typ == "pseudochiral_trivial" and pcb.backend_name == "emu"):      # (no clusters in the emulation: five-pass path both times)
        assert not np.array_equal(out[1], out[0])      # two different kernels did run


@pytest.mark.parametrize("structure", [1, 2, 0])
def test_crossdof_pass_structures(pcb, oracle, structure):
    """The three pass structures of the cross-DoF dielectric give the oracle's H: 1 = plane halves with the stencil fused into
    the inverse half (default), 2 = plane halves around the stencil kernel on the slot layout, 0 = split five-pass path.
    eps_opt = 3 couples all three component pairs (in-plane and across i0 planes)."""
    N, d_flag = 16, "bcc_dg"
    alpha = np.array([np.pi, 0.0, np.pi])
    ctx = pcb.get_context(N)
    case = {"N": N, "d_flag": d_flag, "alpha": alpha, "type": "pseudochiral_crossdof", "eps_opt": 3, "m": 2, "seed": 21}
    a, b, inv, shift, _ = oracle.assemble_symbols(N, d_flag, alpha)
    Ao, Ho, Po = oracle.pc_mfd_handle(a, b, oracle.HANDLES["pseudochiral_crossdof"](N, d_flag, eps_opt=3), inv, shift)
    try:
        ctx.option("plane_cross", structure)
        A, H, P, Diels, x = _setup(pcb, oracle, case)
        assert relerr(H(x), Ho(x)) < TOL
        assert relerr(A(x), Ao(x)) < TOL
    finally:
        ctx.option("plane_cross", 1)


@pytest.mark.parametrize("N,typ,plane_cross,k", [
    (16, None, 1, 1), (16, "chiral", 1, 1), (16, "pseudochiral_crossdof", 1, 1), (16, "pseudochiral_crossdof", 2, 2),
    (16, "pseudochiral_trivial", 1, 1),
    (32, "chiral", 1, 1), (32, "pseudochiral_crossdof", 1, 2),
])
def test_z_split_plane_mode(pcb, oracle, N, typ, plane_cross, k):
    """Z-split plane mode (the plane mode of N = 128, 144, 160: half planes of N/2 x N, the radix-2 step of the z transform inside
    the x passes; ZSplit in pcb_operator.cuh) forced on at sizes where both forms exist: equal to the oracle, and to the whole-plane
    form up to rounding while not bit-identical (two different sets of kernels did run).  The coupled 3x3 M has no z-split
    form and must fall back to the five-pass path."""
    d_flag = "bcc_dg" if typ and typ.startswith("pseudo") else "fcc"
    alpha = np.array([0.3 * np.pi, 2 * np.pi, 0.0])
    ctx = pcb.get_context(N)
    eps_opt = 3 if typ and typ.startswith("pseudo") else 0
    a, b, inv, shift, _ = oracle.assemble_symbols(N, d_flag, alpha, k=k)
    okw = {"k": k} if typ == "pseudochiral_crossdof" else {}
    diel = (lambda v: v) if typ is None else oracle.HANDLES[typ](N, d_flag, eps_opt=eps_opt, **okw)
    Ao, Ho, Po = oracle.pc_mfd_handle(a, b, diel, inv, shift)
    mfd, ne = pcb.discretization, pcb.numerical_experiments
    x = oracle.random_x0(3 * N ** 3, 2, 77 + N)
    out = {}
    try:
        ctx.option("plane_cross", plane_cross)
        for split in (1, 0):
            ctx.option("plane_split", split)
            relax, pnt = mfd.set_relaxation(alpha)
            a_fft, b_fft = mfd.fft_blocks(N, k, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
            inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
            Diels = None if typ is None else getattr(mfd, typ + "_handle")(N, d_flag, eps_opt=eps_opt, **okw)
            A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
            out[split] = A(x)
            assert relerr(out[split], Ao(x)) < TOL
            assert relerr(H(x), Ho(x)) < TOL
    finally:
        ctx.option("plane_split", 0)
        ctx.option("plane_cross", 1)
    assert relerr(out[1], out[0]) < 1e-13
    if typ != "pseudochiral_trivial":      # (the coupled M takes the five-pass path in both forms at N = 16)
        assert not np.array_equal(out[1], out[0])


def test_z_split_dielectric_is_bound_to_its_form(pcb, oracle):
    """A dielectric carries masks in the slot order of the plane-mode form it was created under: using it after the form of the
    context changed is refused instead of silently applying a permuted M."""
    N, d_flag = 16, "fcc"
    ctx = pcb.get_context(N)
    mfd, ne = pcb.discretization, pcb.numerical_experiments
    alpha = np.array([np.pi, 0.0, 0.0])
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    Diels = mfd.chiral_handle(N, d_flag)
    try:
        ctx.option("plane_split", 1)
        with pytest.raises(Exception):
            ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
    finally:
        ctx.option("plane_split", 0)
