"""``cupyx.scipy.fft.{fftn,ifftn}`` -> ``scipy.fft`` (reference: ``pcfft.py:149,151``)."""
import os as _os
import scipy.fft as _sf

_WORKERS = int(_os.environ.get("REFSHIM_FFT_WORKERS", "1"))


def fftn(x, s=None, axes=None, norm=None, overwrite_x=False, **kw):
    return _sf.fftn(x, s=s, axes=axes, norm=norm, workers=_WORKERS)


def ifftn(x, s=None, axes=None, norm=None, overwrite_x=False, **kw):
    return _sf.ifftn(x, s=s, axes=axes, norm=norm, workers=_WORKERS)
