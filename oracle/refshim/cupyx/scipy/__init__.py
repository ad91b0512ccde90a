from . import fft  # noqa: F401
