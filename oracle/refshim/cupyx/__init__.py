"""NumPy/SciPy-backed stand-in for ``cupyx`` (test infrastructure only)."""
from . import scipy  # noqa: F401
