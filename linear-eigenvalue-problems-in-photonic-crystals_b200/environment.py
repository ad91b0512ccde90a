"""Global constants of the hot path (mirrors paper_2/environment.py:19-82 -- same names and values --
so that code written against the reference reads them unchanged)."""
import os
import time

import numpy as np
from numpy import pi

# Paths (environment.py:19-20): relative to the working directory, as in the reference.
OUTPUT_PATH = "output/"
DIEL_PATH = "dielectric_examples/"

# Global parameters (environment.py:23-27)
K = 1          # stencil half-width (accuracy 2k)
NEV = 10       # number of desired eigenpairs
SCAL = 1       # lattice scaling constant
TOL = 1e-4     # residual tolerance
GAP = 20       # k-points per segment of the Brillouin-zone path

# Solver settings (environment.py:30-32)
MAXITER = 500
RESTART_MAX = 100
N_SUBSPACE = 40

# Lattice flags (environment.py:35-40)
SC_F1, SC_F2, SC_C = "sc_flat1", "sc_flat2", "sc_curv"
BCC_SG, BCC_DG, FCC = "bcc_sg", "bcc_dg", "fcc"

# Dielectric types (environment.py:43-46)
TYPE0, TYPE1, TYPE2 = "chiral", "pseudochiral_trivial", "pseudochiral_crossdof"

CHIRAL_EPS_EG = {SC_F1: 13.0, SC_F2: 13.0, SC_C: 13.0, BCC_SG: 16.0, BCC_DG: 16.0, FCC: 13.0}

_g = (1 + 0.875 ** 2) ** 0.5
PSEUDOCHIRAL_EPS_LOC = [
    np.array([_g, _g, 1.0, -1j * 0.875, 0.0, 0.0]),
    np.array([_g, 1.0, _g, 0.0, 1j * 0.875, 0.0]),
    np.array([1.0346, 0.5059, 0.2595, -0.0163 - 0.2319j, 0.027 + 0.0827j, -0.2743 - 0.0076j]),
    np.array([3.0, 3.0, 3.0, np.sqrt(3) + 1j, 1j, np.sqrt(2) * (1 + 1j)]) / 5.0,
]

RED, GREEN, YELLOW, BLUE = "\033[31m", "\033[32m", "\033[33m", "\033[34m"
MAGENTA, CYAN, WHITE, RESET = "\033[35m", "\033[36m", "\033[37m", "\033[0m"

# Coordinate transforms and Brillouin-zone symmetry points (environment.py:72-82)
DIEL_LIB = {
    "CT_sc": [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
    "CT_bcc": [[0, 1, 1], [1, 0, 1], [1, 1, 0]],
    "CT_fcc": [[-1, 1, 1], [1, -1, 1], [1, 1, -1]],
    "sym_sc": [[0, 0, 0], [pi, 0, 0], [pi, pi, 0], [pi, pi, pi], [0, 0, 0]],
    "sym_bcc": [[0, 0, 2 * pi], [0, 0, 0], [pi, pi, pi], [0, 0, 2 * pi], [pi, 0, pi], [0, 0, 0],
                [0, 2 * pi, 0], [pi, pi, pi], [pi, 0, pi]],
    "sym_fcc": [[0, 2 * pi, 0], [pi / 2, 2 * pi, pi / 2], [pi, pi, pi], [0, 0, 0], [0, 2 * pi, 0],
                [pi, 2 * pi, 0], [3 * pi / 2, 3 * pi / 2, 0]],
}

# Verbosity of the ported runners: the reference prints unconditionally; PCB200_QUIET=1 silences the tables.
QUIET = os.environ.get("PCB200_QUIET", "0") == "1"


def say(*a, **k):
    if not QUIET:
        print(*a, **k)


def owari_cuda():
    """Device-synchronised wall clock (environment.py:179-180)."""
    from . import devarray
    for ctx in list(devarray._contexts.values()):
        ctx.sync()
    return time.time()


def norm(X):
    """Frobenius norm of a block / 2-norm of a vector (environment.py:117-129), reduced on the device."""
    from . import pcfft
    import numpy as _np
    return float(_np.sqrt(_np.sum(pcfft.column_dots(X, X).real)))


def norms(X):
    """Column 2-norms (environment.py:131-143) of a DeviceBlock or NumPy array."""
    from . import pcfft
    return pcfft.column_norms(X)


def dots(X, Y):
    """diag(X^H Y) (environment.py:145-157)."""
    from . import pcfft
    return pcfft.column_dots(X, Y)
