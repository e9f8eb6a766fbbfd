from ramannoodle.spectrum import utils  # noqa: F401
from ramannoodle.spectrum._raman import MDRamanSpectrum  # noqa: F401
