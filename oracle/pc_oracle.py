"""CPU oracle for the photonic-crystal hot path  --  TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the reference algorithm for the path named in
BASELINE.json (paper_2: matrix-free operator ``(A M A^H + gamma B^H B) X`` and
the soft-locking LOBPCG).  It exists so that the CUDA path can be checked; it is
never imported by the product package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.

Pinning: every function below is checked against the UNMODIFIED reference
modules, executed in this container through ``oracle/refshim`` (NumPy-backed
``cupy``), by ``oracle/make_golden.py``, ``oracle/make_golden_mixed.py`` and
``oracle/make_golden_c1.py`` (which assert the agreement and write the fixtures
of ``tests/golden/``); ``tests/test_operator.py`` / ``tests/test_lobpcg.py`` then
hold the oracle and the CUDA path against those fixtures on every run.  The third-party arithmetic (FFT, GEMM, Cholesky, eigh) is CuPy ->
cuFFT/cuBLAS/cuSOLVER in the reference (no pinned version) and
pocketfft/OpenBLAS/LAPACK here: standard DFT/GEMM/zheevd semantics.

Conventions (reference file:line, relative to /root/reference/paper_2):
  * row r = c*nn + i0 + N*i1 + N^2*i2, c = field component, i0 fastest
    (discretization.py:326-328, dielectric.py:111); blocks are (3nn, m)
    row-major complex128.
"""
from __future__ import annotations

import time
from fractions import Fraction

import numpy as np
import scipy.fft as sfft

pi = np.pi

# ---------------------------------------------------------------------------
# Constants  (environment.py:19-56, 72-82)
# ---------------------------------------------------------------------------
K = 1
NEV = 10
SCAL = 1
TOL = 1e-4
GAP = 20
MAXITER = 500

CHIRAL_EPS_EG = {"sc_flat1": 13.0, "sc_flat2": 13.0, "sc_curv": 13.0, "bcc_sg": 16.0, "bcc_dg": 16.0,
                 "fcc": 13.0}
_g = (1 + 0.875 ** 2) ** 0.5
PSEUDOCHIRAL_EPS_LOC = [
    np.array([_g, _g, 1.0, -1j * 0.875, 0.0, 0.0]),
    np.array([_g, 1.0, _g, 0.0, 1j * 0.875, 0.0]),
    np.array([1.0346, 0.5059, 0.2595, -0.0163 - 0.2319j, 0.027 + 0.0827j, -0.2743 - 0.0076j]),
    np.array([3.0, 3.0, 3.0, np.sqrt(3) + 1j, 1j, np.sqrt(2) * (1 + 1j)]) / 5.0,
]
CT = {"sc": [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
      "bcc": [[0, 1, 1], [1, 0, 1], [1, 1, 0]],
      "fcc": [[-1, 1, 1], [1, -1, 1], [1, 1, -1]]}
SYM = {"sc": [[0, 0, 0], [pi, 0, 0], [pi, pi, 0], [pi, pi, pi], [0, 0, 0]],
       "bcc": [[0, 0, 2 * pi], [0, 0, 0], [pi, pi, pi], [0, 0, 2 * pi], [pi, 0, pi], [0, 0, 0],
               [0, 2 * pi, 0], [pi, pi, pi], [pi, 0, pi]],
       "fcc": [[0, 2 * pi, 0], [pi / 2, 2 * pi, pi / 2], [pi, pi, pi], [0, 0, 0], [0, 2 * pi, 0],
               [pi, 2 * pi, 0], [3 * pi / 2, 3 * pi / 2, 0]]}

FFT_WORKERS = 1  # scipy.fft workers; bench.py raises it to the host core count


def lattice_ct(d_flag):
    """Coordinate-transform matrix of a lattice flag (dielectric.py:20-35)."""
    return np.array(CT[d_flag.split("_")[0]])


def lattice_sym(d_flag):
    return np.array(SYM[d_flag.split("_")[0]], dtype=float)


def kpath(d_flag, gap=GAP):
    """All k-points of the band path (numerical_experiments.py:342-346):
    GAP points per segment, end point included, start point excluded."""
    sym = lattice_sym(d_flag)
    n_pt = sym.shape[0] - 1
    alphas = np.zeros((n_pt * gap, 3))
    for i in range(n_pt):
        alphas[(i + 1) * gap - 1] = sym[i + 1]
        for j in range(gap - 1):
            alphas[i * gap + j] = ((j + 1) * sym[i + 1] + (gap - j - 1) * sym[i]) / gap
    return alphas


def block_width(nev, ratio=0.6):
    """m = nev + round(0.6 nev) (numerical_experiments.py:64,423)."""
    return nev + round(nev * ratio)


# ---------------------------------------------------------------------------
# Relaxation, stencils, Fourier symbols  (discretization.py:31-49,152-346)
# ---------------------------------------------------------------------------
def set_relaxation(alpha, scal=SCAL):
    """(shift, 0.6), gamma  (discretization.py:31-49)."""
    nu = np.linalg.norm(np.asarray(alpha, dtype=float) / scal)
    if nu > 1:
        return (0, 0.6), 4 * pi * pi
    if nu == 0:
        return (1.0 / pi, 0.6), 4 * pi * pi
    return (nu, 0.6), (2 * pi / nu) ** 2


def mfd_stencil(k, deriv_order):
    """2k-point symmetric stencil on the points {2(j-k)+1} (discretization.py:152-193).

    The reference solves the moment system symbolically with sympy; here it is
    solved exactly over the rationals (same numbers after conversion to float):
    sum_j c_j p_j^i = (deriv_order+1) if i == deriv_order else 0, i = 0..2k-1.
    """
    n = 2 * k
    pts = [Fraction(2 * (j - k) + 1) for j in range(n)]
    a = [[p ** i for p in pts] + [Fraction(deriv_order + 1 if i == deriv_order else 0)] for i in range(n)]
    for col in range(n):  # Gauss-Jordan over Fractions
        piv = next(r for r in range(col, n) if a[r][col] != 0)
        a[col], a[piv] = a[piv], a[col]
        a[col] = [v / a[col][col] for v in a[col]]
        for r in range(n):
            if r != col and a[r][col] != 0:
                f = a[r][col]
                a[r] = [vr - f * vc for vr, vc in zip(a[r], a[col])]
    return np.array([float(a[r][n]) for r in range(n)])


def circulant_symbol(sten, k, N):
    """Eigenvalues of the periodic stencil operator: N*ifft(first row)
    (discretization.py:195-222 called with L=2k, ind=k)."""
    c = np.zeros(N, dtype=complex)
    c[0:k + 1] = sten[k - 1:2 * k]
    if k > 1:
        c[N - k + 1:N] = sten[0:k - 1]
    return np.fft.ifft(c) * N


def symbol_tables(N, k=K, scal=SCAL):
    """1-D tables D1 (derivative /h, h = scal/N) and D0 (averaging)
    (discretization.py:317-324)."""
    h = scal / N
    D1 = circulant_symbol(mfd_stencil(k, 1) / h, k, N)
    D0 = circulant_symbol(mfd_stencil(k, 0), k, N)
    return D1, D0


def fft_blocks(N, k, ct, alpha=None, scal=SCAL):
    """Fourier symbols of the curl block (discretization.py:301-346).

    alpha None -> (D, Di), the alpha-independent parts; else (A, (B_diag, B_sdiag))."""
    n = N ** 3
    D1, D0 = symbol_tables(N, k, scal)
    d01 = np.tile(D1, N * N)
    d02 = np.tile(np.repeat(D1, N), N)
    d03 = np.repeat(D1, N * N)
    e = [np.tile(D0, N * N), np.tile(np.repeat(D0, N), N), np.repeat(D0, N * N)]
    rows = [ct[c][0] * d01 + ct[c][1] * d02 + ct[c][2] * d03 for c in range(3)]
    if alpha is None:
        return np.concatenate(rows), np.concatenate(e)
    al = np.asarray(alpha, dtype=float) / scal
    A = np.concatenate([rows[c] + 1j * al[c] * e[c] for c in range(3)])
    return A, b_from_a(A, n)


def b_from_a(A, n):
    """(|A_c|^2, (conj(A0)A1, conj(A0)A2, conj(A1)A2)) (discretization.py:343-344)."""
    a0, a1, a2 = A[:n], A[n:2 * n], A[2 * n:]
    diag = np.concatenate(((a0 * a0.conj()).real, (a1 * a1.conj()).real, (a2 * a2.conj()).real)).astype(float)
    sdiag = np.concatenate((a0.conj() * a1, a0.conj() * a2, a1.conj() * a2))
    return diag, sdiag


def inverse_3x3_block(D11, D22, D33, D12, D13, D23, shift=0.0):
    """Point-wise inverse of a Hermitian 3x3 block matrix with diagonal blocks
    (cofactor formulas, discretization.py:224-270)."""
    if shift != 0.0:
        D11, D22, D33 = D11 + shift, D22 + shift, D33 + shift
    det = (D11 * D22 * D33 - (D11 * (D23 * D23.conj()) + D22 * (D13 * D13.conj()) + D33 * (D12 * D12.conj()))) \
        + 2 * (D12 * D23 * D13.conj()).real
    fd = np.concatenate(((D22 * D33 - D23 * D23.conj()) / det,
                         (D11 * D33 - D13 * D13.conj()) / det,
                         (D11 * D22 - D12 * D12.conj()) / det))
    det, fd = det.real, fd.real
    fs = np.concatenate(((D13 * D23.conj() - D12 * D33) / det,
                         (D12 * D23 - D13 * D22) / det,
                         (D13 * D12.conj() - D11 * D23) / det))
    return fd, fs


def inverse_3x3_B(B, pnt, shift=0.0):
    """inv(K_A K_A^H + gamma K_B + shift I) symbols (discretization.py:284-295)."""
    b0, b1 = B
    n = len(b0) // 3
    return inverse_3x3_block(pnt * b0[:n] + b0[n:2 * n] + b0[2 * n:],
                             b0[:n] + pnt * b0[n:2 * n] + b0[2 * n:],
                             b0[:n] + b0[n:2 * n] + pnt * b0[2 * n:],
                             (pnt - 1) * b1[:n], (pnt - 1) * b1[n:2 * n], (pnt - 1) * b1[2 * n:], shift=shift)


def assemble_symbols(N, d_flag, alpha, k=K, scal=SCAL):
    """Per-k-point assembly (numerical_experiments.py:33-71 and :421-446).

    Returns a_fft, b_fft=(diag,sdiag) (already times gamma), inv_fft, shift, gamma."""
    (shift, _), pnt = set_relaxation(alpha, scal)
    ct = lattice_ct(d_flag)
    a_fft, b_fft = fft_blocks(N, k, ct, alpha=alpha, scal=scal)
    inv_fft = inverse_3x3_B(b_fft, pnt, shift)
    a_fft = a_fft / scal
    b_fft = (pnt * b_fft[0] / scal / scal, pnt * b_fft[1] / scal / scal)
    inv_fft = (inv_fft[0] * scal * scal, inv_fft[1] * scal * scal)
    return a_fft, b_fft, inv_fft, shift, pnt


# ---------------------------------------------------------------------------
# Geometry  (dielectric.py:104-261)
# ---------------------------------------------------------------------------
def _ijk(N):
    a = np.arange(N)
    return np.tile(a, N * N), np.tile(np.repeat(a, N), N), np.repeat(a, N * N)


def mesh3d_edge_dofs(N):
    """(3N^3, 3) edge-DoF coordinates (dielectric.py:104-117)."""
    I, J, Kk = _ijk(N)
    return np.vstack((np.column_stack(((I + 0.5) / N, J / N, Kk / N)),
                      np.column_stack((I / N, (J + 0.5) / N, Kk / N)),
                      np.column_stack((I / N, J / N, (Kk + 0.5) / N))))


def mesh3d_volume_dofs(N):
    """(N^3, 3) cell-centre coordinates (dielectric.py:119-130)."""
    I, J, Kk = _ijk(N)
    return np.column_stack(((I + 0.5) / N, (J + 0.5) / N, (Kk + 0.5) / N))


def flag_sc_flat1(c):
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    return np.where((x <= 0.25) & (y <= 0.25) | (x <= 0.25) & (z <= 0.25) | (y <= 0.25) & (z <= 0.25))[0]


def flag_sc_flat2(c):
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    return np.where((x <= 0.25) & (y <= 0.25) | (x <= 0.25) & (z >= 0.25) & (z <= 0.5)
                    | (y >= 0.5) & (y <= 0.75) & (z >= 0.5) & (z <= 0.75) | (x >= 0.5) & (x <= 0.75) & (z >= 0.75))[0]


def flag_sc_curv(c_in):
    """Sphere R=0.345 + three axis cylinders r=0.11 about the cell centre (dielectric.py:173-181)."""
    r1, R1 = 0.11, 0.345
    c = c_in - 0.5
    x2, y2, z2 = c[:, 0] ** 2, c[:, 1] ** 2, c[:, 2] ** 2
    return np.where((x2 + y2 + z2 <= R1 ** 2) | (x2 + y2 <= r1 ** 2) | (x2 + z2 <= r1 ** 2) | (y2 + z2 <= r1 ** 2))[0]


def _gyroid(c):
    r = c.T
    return np.sin(2 * pi * r[0]) * np.cos(2 * pi * r[1]) + np.sin(2 * pi * r[1]) * np.cos(2 * pi * r[2]) + \
        np.sin(2 * pi * r[2]) * np.cos(2 * pi * r[0])


def flag_bcc_sg(c):
    return np.where(_gyroid(c) > 1.1)[0]


def flag_bcc_dg(c):
    return np.where(np.abs(_gyroid(c)) > 1.1)[0]


def flag_fcc(c_in):
    """Diamond network: 18 spheres r=0.12 + 16 prolate spheroids (semi-minor 0.11)
    along the four bond directions (dielectric.py:201-261)."""
    r, b = 0.12, 0.11
    x = np.array(c_in, dtype=float).T  # (3, P)
    a = np.array([[0, 0, 0.5, 0.5], [0, 0.5, 0, 0.5], [0, 0.5, 0.5, 0]], dtype=float)
    cnt = np.ones(3) * 0.25
    corners = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 1, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1],
                        [0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0], [1, 0.5, 0.5], [0.5, 1, 0.5],
                        [0.5, 0.5, 1]], dtype=float).T
    centres = np.hstack((corners, cnt[:, None] + a))  # (3, 18)
    sphere = np.any(np.sum((x[:, :, None] - centres[:, None, :]) ** 2, axis=0) < r * r, axis=1)
    inside = np.zeros((x.shape[1], 4), dtype=bool)
    for i in range(4):
        o = (a[:, i] + cnt) / 2
        d = (a[:, i] - cnt) / 2
        c = np.linalg.norm(d)
        d = d / c
        X = x[:, None, :] - (o[:, None] + a)[:, :, None]  # (3, 4, P)
        a_val = np.hypot(b, c)
        L1 = np.tensordot(d, X, axes=([0], [0])) ** 2
        L2 = np.sum(X ** 2, axis=0) - L1
        inside[:, i] = np.any((L1 / a_val ** 2) + (L2 / b ** 2) < 1, axis=0)
    return np.where(sphere | np.any(inside, axis=1))[0]


FLAGS = {"sc_flat1": flag_sc_flat1, "sc_flat2": flag_sc_flat2, "sc_curv": flag_sc_curv,
         "bcc_sg": flag_bcc_sg, "bcc_dg": flag_bcc_dg, "fcc": flag_fcc}


def diel_index(N, d_flag, dofs="edge"):
    """int64 indices of the DoFs inside the dielectric (dielectric.py:58-97, the compute branch)."""
    mesh = mesh3d_edge_dofs(N) if dofs == "edge" else mesh3d_volume_dofs(N)
    return FLAGS[d_flag](mesh @ np.linalg.inv(lattice_ct(d_flag).T)).astype(np.int64)


# ---------------------------------------------------------------------------
# Dielectric operators M  (discretization.py:352-453)
# ---------------------------------------------------------------------------
def _eps_loc(d_flag, eps_opt, eps_mat):
    if eps_mat is not None:
        return np.asarray(eps_mat, dtype=complex)
    return PSEUDOCHIRAL_EPS_LOC[eps_opt] / CHIRAL_EPS_EG[d_flag]


def _eps_diag(N, ind_e, eps_loc):
    nn = N ** 3
    d = np.ones(3 * nn)
    d[ind_e[ind_e < nn]] = eps_loc[0].real
    d[ind_e[(ind_e >= nn) & (ind_e < 2 * nn)]] = eps_loc[1].real
    d[ind_e[ind_e >= 2 * nn]] = eps_loc[2].real
    return d


def chiral_handle(N, d_flag, eps_opt=0, ind_e=None):
    """Isotropic M: rows in Omega_1 divided by eps (discretization.py:352-366)."""
    ind_e = diel_index(N, d_flag, "edge") if ind_e is None else ind_e
    eps1 = CHIRAL_EPS_EG[d_flag] if (eps_opt is None or eps_opt == 0) else eps_opt

    def M(x_in):
        x = x_in.copy()
        x[ind_e] /= eps1
        return x
    return M


def pseudochiral_trivial_handle(N, d_flag, eps_opt=0, eps_mat=None, ind_e=None, ind_v=None):
    """M_Trivial (discretization.py:368-401): diagonal eps_cc on edge DoFs in Omega_1,
    off-diagonal eps_ab coupling the three components at the SAME grid point for
    cells whose volume DoF is in Omega_1."""
    nn = N ** 3
    eps = _eps_loc(d_flag, eps_opt, eps_mat)
    ind_e = diel_index(N, d_flag, "edge") if ind_e is None else ind_e
    ind_v = diel_index(N, d_flag, "volume") if ind_v is None else ind_v
    diag = _eps_diag(N, ind_e, eps)[:, None]
    vol = np.zeros(nn)
    vol[ind_v] = 1.0
    vol = vol[:, None]

    def M(x):
        one = x.ndim == 1
        x = x.reshape(3 * nn, -1)
        x0, x1, x2 = x[:nn], x[nn:2 * nn], x[2 * nn:]
        y = diag * x
        y[:nn] += vol * (eps[3] * x1 + eps[4] * x2)
        y[nn:2 * nn] += vol * (np.conj(eps[3]) * x0 + eps[5] * x2)
        y[2 * nn:] += vol * (np.conj(eps[4]) * x0 + np.conj(eps[5]) * x1)
        return y.ravel() if one else y
    return M


def pseudochiral_crossdof_handle(N, d_flag, eps_opt=0, eps_mat=None, k=1, ind_e=None):
    """M_CrossDoF (discretization.py:403-453), restated as an explicit stencil.

    Off-diagonal block (a,b), a<b:  eps_ab * S_ab,  S_ab = (I_a T_ab + T_ab I_b)/2 with I_c the
    indicator of component-c edge DoFs in Omega_1 and
        T_01 = c(i2) x c^T(i1) x I(i0),  T_02 = c(i2) x I(i1) x c^T(i0),  T_12 = I(i2) x c(i1) x c^T(i0)
    (Kronecker factors slowest index first), c = periodic 2k-point averaging stencil: row r has
    weights sten[j] at columns r + (1-k+j); c^T is its transpose.  The lower blocks are the
    Hermitian transposes."""
    nn = N ** 3
    eps = _eps_loc(d_flag, eps_opt, eps_mat)
    ind_e = diel_index(N, d_flag, "edge") if ind_e is None else ind_e
    diag = _eps_diag(N, ind_e, eps)
    ind = np.zeros(3 * nn)
    ind[ind_e] = 1.0
    I = [ind[c * nn:(c + 1) * nn].reshape(N, N, N) for c in range(3)]  # [i2, i1, i0]
    sten = mfd_stencil(k, 0)
    offs = [1 - k + j for j in range(2 * k)]

    def c_apply(v, axis):      # (c v)[r] = sum_j sten[j] v[r + offs[j]]
        return sum(s * np.roll(v, -o, axis=axis) for s, o in zip(sten, offs))

    def ct_apply(v, axis):     # (c^T v)[r] = sum_j sten[j] v[r - offs[j]]
        return sum(s * np.roll(v, o, axis=axis) for s, o in zip(sten, offs))

    # axes of a (N,N,N,cols) array indexed [i2, i1, i0, col]
    def T(a, b, v):
        if (a, b) == (0, 1):
            return c_apply(ct_apply(v, 1), 0)
        if (a, b) == (0, 2):
            return c_apply(ct_apply(v, 2), 0)
        return c_apply(ct_apply(v, 2), 1)

    def Tt(a, b, v):           # transpose of T_ab
        if (a, b) == (0, 1):
            return ct_apply(c_apply(v, 1), 0)
        if (a, b) == (0, 2):
            return ct_apply(c_apply(v, 2), 0)
        return ct_apply(c_apply(v, 2), 1)

    pairs = [((0, 1), eps[3]), ((0, 2), eps[4]), ((1, 2), eps[5])]

    def M(x):
        one = x.ndim == 1
        x = x.reshape(3 * nn, -1)
        m = x.shape[1]
        xs = [x[c * nn:(c + 1) * nn].reshape(N, N, N, m) for c in range(3)]
        ys = [diag[c * nn:(c + 1) * nn].reshape(N, N, N, 1) * xs[c] for c in range(3)]
        for (a, b), e in pairs:
            if e == 0:
                continue
            Ia, Ib = I[a][..., None], I[b][..., None]
            # y_a += eps_ab * S_ab x_b ;  y_b += conj(eps_ab) * S_ab^T x_a
            ys[a] = ys[a] + e * 0.5 * (Ia * T(a, b, xs[b]) + T(a, b, Ib * xs[b]))
            ys[b] = ys[b] + np.conj(e) * 0.5 * (Tt(a, b, Ia * xs[a]) + Ib * Tt(a, b, xs[a]))
        y = np.concatenate([v.reshape(nn, m) for v in ys], axis=0)
        return y.ravel() if one else y
    return M


HANDLES = {"chiral": chiral_handle, "pseudochiral_trivial": pseudochiral_trivial_handle,
           "pseudochiral_crossdof": pseudochiral_crossdof_handle}


# ---------------------------------------------------------------------------
# Operator  (pcfft.py:18-181, _kernels.py:13-71)
# ---------------------------------------------------------------------------
def a_block(X, D):
    """Cross product with the symbol vector D (pcfft.py:91-108; _kernels.py:43-71)."""
    n = X.shape[0] // 3
    X2 = X.reshape(3 * n, -1)
    d0, d1, d2 = D[:n, None], D[n:2 * n, None], D[2 * n:, None]
    x0, x1, x2 = X2[:n], X2[n:2 * n], X2[2 * n:]
    Y = np.concatenate((-d2 * x1 + d1 * x2, d2 * x0 - d0 * x2, -d1 * x0 + d0 * x1), axis=0)
    return Y.reshape(X.shape)


def h_block(X, DIAG):
    """Hermitian 3x3 block-diagonal multiply, DIAG=(real diag, complex (12,13,23))
    (pcfft.py:50-70; _kernels.py:13-41)."""
    n = X.shape[0] // 3
    X2 = X.reshape(3 * n, -1)
    dg, sd = DIAG
    x0, x1, x2 = X2[:n], X2[n:2 * n], X2[2 * n:]
    s01, s02, s12 = sd[:n, None], sd[n:2 * n, None], sd[2 * n:, None]
    Y = np.concatenate((dg[:n, None] * x0 + s01 * x1 + s02 * x2,
                        s01.conj() * x0 + dg[n:2 * n, None] * x1 + s12 * x2,
                        s02.conj() * x0 + dg[2 * n:, None] * x2 + s12.conj() * x1), axis=0)
    return Y.reshape(X.shape)


def AMA(X, D_A, diel):
    """K_A IFFT3( M FFT3( K_A^H X ) )  (pcfft.py:130-158); same F-order views as the reference."""
    one = X.ndim == 1
    nn3 = X.shape[0]
    m = 1 if one else X.shape[1]
    nn = nn3 // 3
    n = round(nn ** (1 / 3))
    AX = a_block(X.reshape(nn3, m), -D_A.conj()).reshape(n, n, n, 3 * m, order="F")
    AX = sfft.fftn(AX, axes=(0, 1, 2), workers=FFT_WORKERS)
    AX = diel(AX.reshape(nn3, m, order="F")).reshape(n, n, n, 3 * m, order="F")
    AX = sfft.ifftn(AX, axes=(0, 1, 2), workers=FFT_WORKERS)
    AX = a_block(AX.reshape(nn3, m, order="F"), D_A)
    return AX.ravel() if one else AX


def AMA_BB(X, D_A, D_B, diel, shift=0):
    """(A M A^H + gamma B^H B + shift) X  (pcfft.py:160-181)."""
    HX = AMA(X, D_A, diel)
    HX = HX + h_block(X, D_B)
    if shift != 0:
        HX = HX + shift * X
    return HX


def pc_mfd_handle(a_fft, b_fft, diel, inv_fft, shift=0.0):
    """(A_func, H_func, P_func)  (numerical_experiments.py:73-85)."""
    return (lambda x: AMA(x, a_fft, diel),
            lambda x: AMA_BB(x, a_fft, b_fft, diel, shift),
            lambda x: h_block(x, inv_fft))


# ---------------------------------------------------------------------------
# Rayleigh-Ritz and LOBPCG  (orthogonalization.py:26-33,140-154; lobpcg.py:325-492,1248-1270)
# ---------------------------------------------------------------------------
def hermitize(M):
    return (M + M.T.conj()) / 2


def rayleigh_ritz_chol_sep(s, hs):
    """Cholesky-whitened Rayleigh-Ritz (orthogonalization.py:140-154)."""
    ss = hermitize(s.conj().T @ s)
    shs = hermitize(s.conj().T @ hs)
    return rr_small(ss, shs)


def rr_small(ss, shs):
    """The n_loc x n_loc part of RR: L = inv(chol(G)); eigh(L T L^H); E = L^H V
    (orthogonalization.py:148-151)."""
    L = np.linalg.inv(np.linalg.cholesky(ss))
    t = (L @ shs) @ L.conj().T
    lam, v = np.linalg.eigh(t)
    return lam, L.conj().T @ v


def column_norms(X):
    """sqrt(diag(X^H X))  (environment.py:131-143)."""
    return np.sqrt(np.einsum("ij,ij->j", X.conj(), X).real)


def lobpcg_sep_softlock(h_func_in, p_func, x0, nev, shift=0.0, tol=TOL, maxiter=MAXITER, history=False,
                        maxstagniter=50, verbose=False, trace=None):
    """Soft-locking LOBPCG, [X W P] stored in one (R, 3m) block (lobpcg.py:325-492).

    ``trace`` (optional list) receives per-iteration dicts (res_nrms, n_act, lambdas) for tests."""
    m = x0.shape[1]
    R = x0.shape[0]
    h_func = h_func_in if shift == 0.0 else (lambda x: h_func_in(x) + shift * x)
    res_his = np.empty(maxiter)
    s = np.empty((R, 3 * m), dtype=np.complex128)
    hs = np.empty((R, 3 * m), dtype=np.complex128)
    s[:, :m] = x0
    hs[:, :m] = h_func(s[:, :m])
    # Initial lambda (lobpcg.py:378-381): the m x m Gram matrices are themselves fed to RR.
    ss = hermitize(s[:, :m].conj().T @ s[:, :m])
    shs = hermitize(s[:, :m].conj().T @ hs[:, :m])
    lambdas, _ = rayleigh_ritz_chol_sep(ss, shs)
    t0 = time.time()
    it = 0
    for it in range(maxiter):
        s[:, m:2 * m] = s[:, :m] * lambdas - hs[:, :m]
        res = column_norms(s[:, m:2 * m])
        res_his[it] = np.linalg.norm(res[:nev])
        act = np.where(res > tol)[0]
        n_act = len(act)
        if trace is not None:
            trace.append({"res": res.copy(), "n_act": n_act, "lambdas": np.array(lambdas[:m])})
        if verbose:
            print(f"Iter = {it:<4d}, res_nrm = {np.linalg.norm(res):<6.2e}, n_act = {n_act:<3d}.")
        if np.isnan(res).any():
            return None, None, None
        if (it > maxstagniter and (res[0] > 1000 or res[0] > res_his[1])) or (it > 2 * maxstagniter and res[0] > 50):
            if not np.linalg.norm(res[:nev]) < res_his[maxstagniter // 2] * 0.1:
                return None, None, None
        if max(res[:nev]) < tol:
            break
        n_loc = m + 2 * n_act if it > 0 else m + n_act
        if n_act < m:   # in-place, ascending compaction (lobpcg.py:431-436)
            for i0 in range(n_act):
                s[:, m + i0] = s[:, m + act[i0]]
            for i0 in range(n_act):
                s[:, m + n_act + i0] = s[:, 2 * m + act[i0]]
                hs[:, m + n_act + i0] = hs[:, 2 * m + act[i0]]
        s[:, m:m + n_act] = p_func(s[:, m:m + n_act])
        hs[:, m:m + n_act] = h_func(s[:, m:m + n_act])
        try:
            lambdas, E = rayleigh_ritz_chol_sep(s[:, :n_loc], hs[:, :n_loc])
        except np.linalg.LinAlgError:
            return None, None, None
        if np.isnan(lambdas).any() or np.isnan(E).any():
            return None, None, None
        lambdas, E = lambdas[:m], E[:, :m]
        # _sep_update_after_rr (lobpcg.py:1248-1270)
        if it > 0:
            pn = s[:, m + n_act:n_loc] @ E[m + n_act:] + s[:, m:m + n_act] @ E[m:m + n_act]
            hpn = hs[:, m + n_act:n_loc] @ E[m + n_act:] + hs[:, m:m + n_act] @ E[m:m + n_act]
        else:
            pn = s[:, m:m + n_act] @ E[m:]
            hpn = hs[:, m:m + n_act] @ E[m:]
        s[:, 2 * m:] = pn
        hs[:, 2 * m:] = hpn
        s[:, :m] = s[:, :m] @ E[:m] + pn
        hs[:, :m] = hs[:, :m] @ E[:m] + hpn
    t_tot = time.time() - t0
    info = np.array([it, t_tot])
    if history:
        info = np.append(info, res_his[1:it])
    return lambdas[:m] - shift, s[:, :m], info


def lobpcg_sep_softlock_mixedprecision(h_func, p_func, x0, nev, tol=TOL, maxiter=MAXITER, history=False, trace=None):
    """Soft-locking LOBPCG with the preconditioner input rounded to complex64 (lobpcg.py:494-629).

    Same iteration as ``lobpcg_sep_softlock`` except: initial lambda = eigvalsh(herm(X^H HX)) (:524), no stagnation
    rules, NaN residuals raise (:549), ``p_func`` receives ``W.astype(complex64)`` and its result is widened again
    (:574-577), and the return value is ``(lambdas[:nev], s[:, :nev], info)`` (:629)."""
    m = x0.shape[1]
    R = x0.shape[0]
    res_his = np.empty(maxiter)
    s = np.empty((R, 3 * m), dtype=np.complex128)
    hs = np.empty((R, 3 * m), dtype=np.complex128)
    s[:, :m] = x0
    hs[:, :m] = h_func(s[:, :m])
    lambdas = np.linalg.eigvalsh(hermitize(s[:, :m].conj().T @ hs[:, :m]))
    t0 = time.time()
    it = 0
    for it in range(maxiter):
        s[:, m:2 * m] = s[:, :m] * lambdas - hs[:, :m]
        res = column_norms(s[:, m:2 * m])
        res_his[it] = np.linalg.norm(res[:nev])
        if np.isnan(res).any():
            raise ValueError("Nan occurs in residuals.")
        act = np.where(res > tol)[0]
        n_act = len(act)
        if trace is not None:
            trace.append({"res": res.copy(), "n_act": n_act, "lambdas": np.array(lambdas[:m])})
        n_loc = m + 2 * n_act if it > 0 else m + n_act
        if max(res[:nev]) < tol:
            break
        if n_act < m:
            for i0 in range(n_act):
                s[:, m + i0] = s[:, m + act[i0]]
            for i0 in range(n_act):
                s[:, m + n_act + i0] = s[:, 2 * m + act[i0]]
                hs[:, m + n_act + i0] = hs[:, 2 * m + act[i0]]
        s[:, m:m + n_act] = np.asarray(p_func(s[:, m:m + n_act].astype(np.complex64))).astype(np.complex128)
        hs[:, m:m + n_act] = h_func(s[:, m:m + n_act])
        lambdas, E = rayleigh_ritz_chol_sep(s[:, :n_loc], hs[:, :n_loc])
        lambdas, E = lambdas[:m], E[:, :m]
        if it > 0:
            pn = s[:, m + n_act:n_loc] @ E[m + n_act:] + s[:, m:m + n_act] @ E[m:m + n_act]
            hpn = hs[:, m + n_act:n_loc] @ E[m + n_act:] + hs[:, m:m + n_act] @ E[m:m + n_act]
        else:
            pn = s[:, m:m + n_act] @ E[m:]
            hpn = hs[:, m:m + n_act] @ E[m:]
        s[:, 2 * m:] = pn
        hs[:, 2 * m:] = hpn
        s[:, :m] = s[:, :m] @ E[:m] + pn
        hs[:, :m] = hs[:, :m] @ E[:m] + hpn
    info = np.array([it, time.time() - t0])
    if history:
        info = np.append(info, res_his[1:it])
    return lambdas[:nev], s[:, :nev], info


# ---------------------------------------------------------------------------
# Post-processing  (numerical_experiments.py:87-158)
# ---------------------------------------------------------------------------
def _sqrt_robust(a):
    return 0.0 if (a <= 0) and (a > -1e-8) else a ** 0.5


def recompute_normalize(lambdas_in, x, A_func, shift=0.0, scal=SCAL):
    """Returns (omega_pnt/2pi, omega_re/2pi, residual norms); raises on spurious modes."""
    adax = A_func(x)
    lam_pnt = lambdas_in - shift if shift > 0.0 else np.array(lambdas_in, dtype=float)
    Rm = adax - x * lam_pnt
    lam_re = (np.einsum("ij,ij->j", x.conj(), adax) / np.einsum("ij,ij->j", x.conj(), x)).real
    w_pnt = np.array([_sqrt_robust(v) * scal / (2 * pi) for v in lam_pnt])
    w_re = np.array([_sqrt_robust(v) * scal / (2 * pi) for v in lam_re])
    if np.any(w_pnt - w_re > 1e-3):
        raise ValueError("Spurious eigenvalues occur.")
    return w_pnt, w_re, column_norms(Rm)


# ---------------------------------------------------------------------------
# Convenience: one k-point end to end (numerical_experiments.py:209-247)
# ---------------------------------------------------------------------------
def random_x0(R, m, seed):
    """x0 = U[0,1) + i U[0,1); real part drawn first."""
    rng = np.random.default_rng(seed)
    re = rng.random((R, m))
    im = rng.random((R, m))
    return re + 1j * im


def eigen_1p(N, d_flag, alpha, type="chiral", nev=NEV, x0=None, seed=0, tol=TOL, maxiter=MAXITER,
             eps_opt=0, k=K, history=False, trace=None):
    a_fft, b_fft, inv_fft, shift, _ = assemble_symbols(N, d_flag, np.asarray(alpha, dtype=float), k=k)
    m = block_width(nev)
    if x0 is None:
        x0 = random_x0(3 * N ** 3, m, seed)
    diel = (lambda v: v) if type is None else HANDLES[type](N, d_flag, eps_opt=eps_opt)
    A_func, H_func, P_func = pc_mfd_handle(a_fft, b_fft, diel, inv_fft, shift)
    lam, x, info = lobpcg_sep_softlock(H_func, P_func, x0, nev, tol=tol, maxiter=maxiter, history=history,
                                       trace=trace)
    if lam is None:
        return None
    w_pnt, w_re, res = recompute_normalize(lam[:nev], x[:, :nev], A_func, shift)
    return {"lambdas": lam, "x": x, "info": info, "omega_pnt": w_pnt, "omega_re": w_re, "residuals": res,
            "shift": shift}
