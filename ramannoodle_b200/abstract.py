"""The three-method plugin contract of the reference (``ramannoodle/abstract.py:10-83``)."""
from __future__ import annotations

from abc import ABC, abstractmethod


class PolarizabilityModel(ABC):  # pylint: disable=too-few-public-methods
    """``calc_polarizabilities(positions_batch (S,N,3) fractional) -> (S,3,3)``."""

    @abstractmethod
    def calc_polarizabilities(self, positions_batch):
        """Return polarizabilities for a batch of fractional positions."""


class RamanSpectrum(ABC):  # pylint: disable=too-few-public-methods
    """``measure(...) -> (wavenumbers, intensities)``."""

    @abstractmethod
    def measure(self, orientation="polycrystalline", laser_correction=False,
                laser_wavelength=522, bose_einstein_correction=False, temperature=300):
        """Calculate and return a raw Raman spectrum."""


class Dynamics(ABC):  # pylint: disable=too-few-public-methods
    """``get_raman_spectrum(polarizability_model) -> RamanSpectrum``."""

    @abstractmethod
    def get_raman_spectrum(self, polarizability_model):
        """Calculate a Raman spectrum using a polarizability model."""
