"""CPU oracle: numpy/scipy restatement of ramannoodle's MD-Raman hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and
only as the checker / the timed CPU baseline.  The product package ``ramannoodle_b200``
never imports anything under ``oracle/``.

Every function follows the reference statement by statement (same numpy calls, same order
of operations) so that on identical inputs it is bit-identical to the reference; this is
pinned in ``tests/test_oracle_vs_reference.py`` against the unmodified reference imported
from ``/root/reference`` (when present) and in ``tests/test_oracle_golden.py`` against the
committed golden vectors generated from the reference by ``oracle/make_golden.py``.

Third-party arithmetic on the path (not vendored in the reference; only lower-bounded in
its ``pyproject.toml:16-27``, installed in this image: numpy 2.3.5, scipy 1.18.1):
``scipy.interpolate.BSpline.__call__`` (de Boor, restated independently in
``oracle/oracle.c`` and checked against scipy), ``scipy.signal.correlate`` and
``scipy.fftpack.fft/fftfreq``.  Like the reference, this port calls scipy for those.

Citations are ``file:line`` relative to ``/root/reference``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy
import scipy.fftpack
import scipy.signal
from scipy.interpolate import BSpline

BOLTZMANN_CONSTANT = 8.617333262e-5  # eV/K, ramannoodle/constants.py:249


# --------------------------------------------------------------------------------------
# structure/utils.py
# --------------------------------------------------------------------------------------
def apply_pbc(positions):
    """ramannoodle/structure/utils.py:13-29 — ``positions - positions // 1``."""
    return positions - positions // 1


def apply_pbc_displacement(displacement):
    """ramannoodle/structure/utils.py:32-48 — wrap into (-0.5, 0.5]."""
    return np.where(displacement % 1 > 0.5, displacement % 1 - 1, displacement % 1)


def calc_displacement(positions_1, positions_2):
    """ramannoodle/structure/utils.py:110-135."""
    positions_1 = apply_pbc(positions_1)
    positions_2 = apply_pbc(positions_2)
    return apply_pbc_displacement(positions_2 - positions_1)


def get_cart_displacement(lattice, displacement):
    """ramannoodle/structure/_reference.py:268-285 — wrap again, then ``@ lattice``."""
    displacement = apply_pbc_displacement(displacement)
    return displacement @ lattice


# --------------------------------------------------------------------------------------
# pmodel/_interpolation.py  (state container + calc_polarizabilities)
# --------------------------------------------------------------------------------------
@dataclass
class OracleModel:
    """The state ``InterpolationModel``/``ARTModel`` hold after construction.

    Mirrors ``ramannoodle/pmodel/_interpolation.py:110-115``: ``ref_positions`` and
    ``lattice`` come from ``_ref_structure``; ``basis_vectors`` is ``_cart_basis_vectors``
    (J arrays (N,3)); ``splines`` is ``_interpolations`` given as (t, c, k) triples with
    ``c.shape == (n,3,3)``; ``mask`` is ``_mask`` (J,) bool.
    """

    ref_positions: np.ndarray
    lattice: np.ndarray
    ref_polarizability: np.ndarray
    basis_vectors: list = field(default_factory=list)
    splines: list = field(default_factory=list)  # (t, c, k)
    mask: np.ndarray = field(default_factory=lambda: np.array([], dtype=bool))

    @property
    def num_atoms(self) -> int:
        return self.ref_positions.shape[0]

    @classmethod
    def from_reference_model(cls, model) -> "OracleModel":
        """Read the private state of a reference ``InterpolationModel``/``ARTModel``."""
        return cls(
            ref_positions=np.array(model._ref_structure.positions, dtype=np.float64),
            lattice=np.array(model._ref_structure.lattice, dtype=np.float64),
            ref_polarizability=np.array(model._ref_polarizability, dtype=np.float64),
            basis_vectors=[np.array(v, dtype=np.float64) for v in model._cart_basis_vectors],
            splines=[(np.array(s.t), np.array(s.c), int(s.k)) for s in model._interpolations],
            mask=np.array(model._mask, dtype=bool),
        )


def calc_polarizabilities(model: OracleModel, positions_batch):
    """ramannoodle/pmodel/_interpolation.py:191-252 (error wrapping omitted).

    ``alpha_s = alpha_ref + sum_j (1 - mask_j) * B_j(v_j . vec(cart_disp_s))``.
    """
    delta = np.zeros((positions_batch.shape[0], 3, 3))  # :214-216
    cart = get_cart_displacement(  # :217-219
        model.lattice, calc_displacement(model.ref_positions, positions_batch)
    )
    cart = cart.reshape(cart.shape[0], cart.shape[1] * cart.shape[2])  # :220-223
    for basis_vector, (t, c, k), mask in zip(  # :233-244
        model.basis_vectors, model.splines, model.mask, strict=True
    ):
        amplitudes = np.einsum("i,ji", basis_vector.flatten(), cart)
        interpolation = BSpline(t, c, k, extrapolate=True)
        delta += (1 - mask) * np.array(interpolation(amplitudes), dtype="float64")
    return delta + model.ref_polarizability  # :252


def calc_cart_displacements(model: OracleModel, positions_batch):
    """The (S, 3N) Cartesian displacement matrix of ``_interpolation.py:217-223``."""
    cart = get_cart_displacement(
        model.lattice, calc_displacement(model.ref_positions, positions_batch)
    )
    return cart.reshape(cart.shape[0], cart.shape[1] * cart.shape[2])


def get_polarizability(model: OracleModel, cart_displacements):
    """``_interpolation.py:233-252`` starting from precomputed (S,3N) displacements."""
    delta = np.zeros((cart_displacements.shape[0], 3, 3))
    for basis_vector, (t, c, k), mask in zip(
        model.basis_vectors, model.splines, model.mask, strict=True
    ):
        amplitudes = np.einsum("i,ji", basis_vector.flatten(), cart_displacements)
        delta += (1 - mask) * np.array(
            BSpline(t, c, k, extrapolate=True)(amplitudes), dtype="float64"
        )
    return delta + model.ref_polarizability


# --------------------------------------------------------------------------------------
# spectrum/utils.py
# --------------------------------------------------------------------------------------
def _calc_autocorrelation(signal):
    """ramannoodle/spectrum/utils.py:76-92."""
    autocorrelation = scipy.signal.correlate(signal, signal, "full")
    autocorrelation = autocorrelation[(len(autocorrelation) - 1) // 2:]
    return autocorrelation


def calc_signal_spectrum(signal, sampling_rate):
    """ramannoodle/spectrum/utils.py:95-124."""
    autocorrelation = _calc_autocorrelation(signal)
    wavenumbers = (
        scipy.fftpack.fftfreq(autocorrelation.size, sampling_rate)
        * 33.35640951981521
        * 1e3
    )
    intensities = np.real(scipy.fftpack.fft(autocorrelation))
    return wavenumbers[wavenumbers >= 0], intensities[wavenumbers >= 0]


def convolve_spectrum(wavenumbers, intensities, function="gaussian", width=5,
                      out_wavenumbers=None):
    """ramannoodle/spectrum/utils.py:12-73 (argument verification omitted)."""
    if out_wavenumbers is None:  # :42-46
        min_wavenumber = np.min(wavenumbers) - 100
        max_wavenumber = np.max(wavenumbers) + 100
        num_samples = int(np.rint(max_wavenumber - min_wavenumber))
        out_wavenumbers = np.linspace(min_wavenumber, max_wavenumber, num_samples)
    if width <= 0:
        raise ValueError(f"invalid width: {width} <= 0")
    convolved_intensities = out_wavenumbers * 0
    for wavenumber, intensity in zip(wavenumbers, intensities):  # :58-72
        factor = 0
        if function == "gaussian":
            factor = (
                (1 / width)
                * (1 / np.sqrt(2 * np.pi))
                * np.exp(-((wavenumber - out_wavenumbers) ** 2) / (2 * width**2))
            )
        elif function == "lorentzian":
            factor = (1 / np.pi) * (
                0.5 * width / ((wavenumber - out_wavenumbers) ** 2 + (0.5 * width) ** 2)
            )
        else:
            raise ValueError(f"unsupported convolution type: {function}")
        convolved_intensities += factor * intensity
    return (out_wavenumbers, convolved_intensities)


# --------------------------------------------------------------------------------------
# spectrum/_raman.py
# --------------------------------------------------------------------------------------
def get_bose_einstein_correction(wavenumbers, temperature):
    """ramannoodle/spectrum/_raman.py:13-40."""
    if temperature <= 0:
        raise ValueError(f"invalid temperature: {temperature} <= 0")
    energy = wavenumbers * 29979245800.0 * 4.1357e-15  # in eV
    return 1 / (1 - np.exp(-energy / (BOLTZMANN_CONSTANT * temperature)))


def get_laser_correction(wavenumbers, laser_wavenumber):
    """ramannoodle/spectrum/_raman.py:43-69."""
    if laser_wavenumber <= 0:
        raise ValueError(f"invalid laser_wavenumber: {laser_wavenumber} <= 0")
    return ((wavenumbers - laser_wavenumber) / 10000) ** 4 / wavenumbers


def md_measure(polarizability_ts, timestep, laser_correction=False, laser_wavelength=522,
               bose_einstein_correction=False, temperature=300):
    """``MDRamanSpectrum.measure`` — ramannoodle/spectrum/_raman.py:241-309."""
    ad = np.diff(polarizability_ts, axis=0)  # :282
    wavenumbers, _ = calc_signal_spectrum(ad[:, 0, 0], timestep)
    alpha2 = (1 / 9) * calc_signal_spectrum(
        ad[:, 0, 0] + ad[:, 1, 1] + ad[:, 2, 2], timestep
    )[1]
    gamma2 = (
        (1 / 2) * calc_signal_spectrum(ad[:, 0, 0] - ad[:, 1, 1], timestep)[1]
        + (1 / 2) * calc_signal_spectrum(ad[:, 1, 1] - ad[:, 2, 2], timestep)[1]
        + (1 / 2) * calc_signal_spectrum(ad[:, 2, 2] - ad[:, 0, 0], timestep)[1]
        + 3 * calc_signal_spectrum(ad[:, 0, 1], timestep)[1]
        + 3 * calc_signal_spectrum(ad[:, 1, 2], timestep)[1]
        + 3 * calc_signal_spectrum(ad[:, 0, 2], timestep)[1]
    )
    intensities = 45.0 * alpha2 + 7.0 * gamma2
    intensities = intensities[1:]  # :299-301
    wavenumbers = wavenumbers[1:]
    if laser_correction:
        laser_wavenumber = 10000000 / laser_wavelength
        intensities *= get_laser_correction(wavenumbers, laser_wavenumber)
    if bose_einstein_correction:
        intensities *= get_bose_einstein_correction(wavenumbers, temperature)
    return wavenumbers, intensities


def trajectory_positions(positions_ts):
    """``Trajectory.__init__`` stores ``apply_pbc(positions_ts)`` — dynamics/_trajectory.py:45."""
    return apply_pbc(positions_ts)


def get_raman_spectrum(model: OracleModel, positions_ts, timestep, **measure_kwargs):
    """Trajectory.get_raman_spectrum + measure — dynamics/_trajectory.py:71-90."""
    alpha = calc_polarizabilities(model, trajectory_positions(positions_ts))
    return md_measure(alpha, timestep, **measure_kwargs)


# --------------------------------------------------------------------------------------
# dynamics/_phonon.py + PhononRamanSpectrum
# --------------------------------------------------------------------------------------
RAMAN_TENSOR_CENTRAL_DIFFERENCE = 0.001  # ramannoodle/constants.py:248


def phonon_raman_tensors(model: OracleModel, ref_positions, displacements):
    """``Phonons.get_raman_spectrum`` — ramannoodle/dynamics/_phonon.py:82-108: central
    differences, two S=1 ``calc_polarizabilities`` calls per mode."""
    raman_tensors = []
    for displacement in displacements:
        epsilon = displacement * RAMAN_TENSOR_CENTRAL_DIFFERENCE
        plus = calc_polarizabilities(model, np.array([ref_positions + epsilon]))[0]
        minus = calc_polarizabilities(model, np.array([ref_positions - epsilon]))[0]
        raman_tensors.append((plus - minus) / RAMAN_TENSOR_CENTRAL_DIFFERENCE)
    return np.array(raman_tensors)


def phonon_measure(phonon_wavenumbers, raman_tensors, laser_correction=False, laser_wavelength=522,
                   bose_einstein_correction=False, temperature=300):
    """``PhononRamanSpectrum.measure`` — ramannoodle/spectrum/_raman.py:128-194."""
    alpha_squared = ((raman_tensors[:, 0, 0] + raman_tensors[:, 1, 1] + raman_tensors[:, 2, 2]) / 3.0) ** 2
    gamma_squared = (
        (raman_tensors[:, 0, 0] - raman_tensors[:, 1, 1]) ** 2
        + (raman_tensors[:, 0, 0] - raman_tensors[:, 2, 2]) ** 2
        + (raman_tensors[:, 1, 1] - raman_tensors[:, 2, 2]) ** 2
        + 6.0 * (raman_tensors[:, 0, 1] ** 2 + raman_tensors[:, 0, 2] ** 2 + raman_tensors[:, 1, 2] ** 2)
    ) / 2.0
    intensities = 45.0 * alpha_squared + 7.0 * gamma_squared
    if laser_correction:
        laser_wavenumber = 10000000 / laser_wavelength
        intensities *= get_laser_correction(phonon_wavenumbers, laser_wavenumber)
    if bose_einstein_correction:
        intensities *= get_bose_einstein_correction(phonon_wavenumbers, temperature)
    return phonon_wavenumbers, intensities
