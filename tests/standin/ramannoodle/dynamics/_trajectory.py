"""ramannoodle/dynamics/_trajectory.py:16-109 (the parts on the hot path)."""
from oracle import numpy_port as ora
from ramannoodle.spectrum._raman import MDRamanSpectrum


class Trajectory:
    def __init__(self, positions_ts, timestep):
        self._positions_ts = ora.apply_pbc(positions_ts)
        self._timestep = timestep

    def get_raman_spectrum(self, polarizability_model):
        try:
            polarizability_ts = polarizability_model.calc_polarizabilities(self._positions_ts)
        except ValueError as exc:
            raise ValueError("polarizability_model and trajectory are incompatible") from exc
        return MDRamanSpectrum(polarizability_ts, self._timestep)
