"""ramannoodle/spectrum/_raman.py:197-309."""
from oracle import numpy_port as ora


class MDRamanSpectrum:
    def __init__(self, polarizability_ts, timestep):
        self._polarizability_ts = polarizability_ts
        self._timestep = timestep

    @property
    def polarizability_ts(self):
        return self._polarizability_ts

    # pylint: disable=too-many-arguments,too-many-positional-arguments
    def measure(self, orientation="polycrystalline", laser_correction=False, laser_wavelength=522,
                bose_einstein_correction=False, temperature=300):
        if orientation != "polycrystalline":
            raise NotImplementedError("only polycrystalline spectra are supported for now")
        return ora.md_measure(self._polarizability_ts, self._timestep, laser_correction, laser_wavelength,
                              bose_einstein_correction, temperature)
