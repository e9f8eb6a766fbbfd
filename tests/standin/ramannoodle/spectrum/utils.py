"""ramannoodle/spectrum/utils.py:12-73."""
from oracle import numpy_port as ora


def convolve_spectrum(wavenumbers, intensities, function="gaussian", width=5, out_wavenumbers=None):
    return ora.convolve_spectrum(wavenumbers, intensities, function, width, out_wavenumbers)
