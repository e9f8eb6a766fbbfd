"""``install()``: run an importable ``ramannoodle`` on the GPU without changing user code.

north_star: "the public Python API stays as it is".  ``install()`` patches the four entry points of
the hot path inside the reference package, so existing scripts —

    model = ramannoodle.pmodel.ARTModel(...); model.add_art_from_files(...)
    trajectory = ramannoodle.io.generic.read_trajectory(...)
    spectrum = trajectory.get_raman_spectrum(model); spectrum.measure(...)

— keep their objects and calls and get the CUDA path:

* ``InterpolationModel.calc_polarizabilities`` (``pmodel/_interpolation.py:191``; inherited by
  ``ARTModel``) evaluates through a ``ramannoodle_b200`` model that is (re)packed whenever a fingerprint
  of the LIVE reference object changes (DOFs added, ``mask`` reassigned or edited, lists replaced),
  so the reference's mutable models (``add_dof``, ``mask`` setter, ``unmask``, ``get_masked_model``'s
  deep copies) keep working;
* ``Trajectory.get_raman_spectrum`` (``dynamics/_trajectory.py:71``) additionally keeps the series on
  the device for ``measure`` and can page-lock the trajectory once for full PCIe bandwidth;
* ``MDRamanSpectrum.measure`` (``spectrum/_raman.py:241``) and ``convolve_spectrum``
  (``spectrum/utils.py:12``) run the CUDA kernels.

Exceptions keep the reference's types and messages (``UserError`` is re-raised as the reference's
class).  ``uninstall()`` restores the original functions.  There is still no CPU fallback: on a
machine without an sm_100 GPU the patched calls raise ``NativeLibraryError``.
"""
from __future__ import annotations

import importlib
import weakref

import numpy as np

from . import _lib
from .exceptions import UserError
from .pmodel import ARTModel, InterpolationModel
from .spectrum import MDRamanSpectrum, convolve_spectrum
from .state import ModelState

_ORIGINALS: dict = {}
_CACHE_ATTR = "_rn_b200_accel"
_SERIES_ATTR = "_rn_b200_device_series"
_PINNED_ATTR = "_rn_b200_pinned"


def live_fingerprint(model) -> tuple:
    """Changes whenever the state ``calc_polarizabilities`` reads from a reference model changes:
    list identities and lengths, the last elements' identities, the mask's bytes, the reference
    polarizability's bytes, the reference structure's identity."""
    vectors = model._cart_basis_vectors  # pylint: disable=protected-access
    splines = model._interpolations  # pylint: disable=protected-access
    mask = np.asarray(model._mask)  # pylint: disable=protected-access
    return (id(vectors), len(vectors), id(vectors[-1]) if vectors else 0,
            id(splines), len(splines), id(splines[-1]) if splines else 0,
            mask.tobytes(), np.asarray(model._ref_polarizability).tobytes(),  # pylint: disable=protected-access
            id(model._ref_structure))  # pylint: disable=protected-access


def accelerated(model, device=None):
    """The ``ramannoodle_b200`` evaluator of a reference model, re-packed when the model changed."""
    fingerprint = live_fingerprint(model)
    cached = model.__dict__.get(_CACHE_ATTR)
    if cached is not None and cached[0] == fingerprint:
        return cached[1]
    cls = ARTModel if type(model).__name__ == "ARTModel" else InterpolationModel
    accel = cls(ModelState.from_reference(model), device=device)
    model.__dict__[_CACHE_ATTR] = (fingerprint, accel)
    return accel


def _reference_user_error():
    return importlib.import_module("ramannoodle.exceptions").UserError


def _patched_calc_polarizabilities(self, positions_batch):
    try:
        return accelerated(self).calc_polarizabilities(positions_batch)
    except UserError as exc:
        raise _reference_user_error()(str(exc)) from exc


def _patched_get_raman_spectrum(self, polarizability_model):
    spectrum_cls = importlib.import_module("ramannoodle.spectrum._raman").MDRamanSpectrum
    interpolation_cls = importlib.import_module("ramannoodle.pmodel._interpolation").InterpolationModel
    if not isinstance(polarizability_model, interpolation_cls):  # e.g. PotGNN: the reference's own path
        return _ORIGINALS[("ramannoodle.dynamics._trajectory", "Trajectory", "get_raman_spectrum")](
            self, polarizability_model)
    positions = self._positions_ts  # pylint: disable=protected-access
    if _OPTIONS["pin_trajectories"] and not self.__dict__.get(_PINNED_ATTR) and positions.flags["C_CONTIGUOUS"]:
        # page-lock the trajectory once: later evaluations stream it at full PCIe bandwidth
        status = _lib.lib().rn_host_register(positions.ctypes.data, positions.nbytes)
        self.__dict__[_PINNED_ATTR] = "registered" if status == 0 else "failed"
        if status == 0:  # unlock the pages before numpy frees them
            weakref.finalize(self, _lib.lib().rn_host_unregister, positions.ctypes.data)
    try:
        accel = accelerated(polarizability_model)
        series = accel.calc_polarizabilities_to_device(positions)
    except UserError as exc:
        raise _reference_user_error()(str(exc)) from exc
    except ValueError as exc:
        raise ValueError("polarizability_model and trajectory are incompatible") from exc
    spectrum = spectrum_cls(series.cpu().numpy(), self._timestep)  # pylint: disable=protected-access
    spectrum.__dict__[_SERIES_ATTR] = series  # measure() starts from the device copy
    return spectrum


# pylint: disable=too-many-arguments,too-many-positional-arguments
def _patched_measure(self, orientation="polycrystalline", laser_correction=False, laser_wavelength=522,
                     bose_einstein_correction=False, temperature=300):
    series = self.__dict__.get(_SERIES_ATTR)
    if series is None:
        series = self._polarizability_ts  # pylint: disable=protected-access
    return MDRamanSpectrum(series, self._timestep).measure(  # pylint: disable=protected-access
        orientation, laser_correction, laser_wavelength, bose_einstein_correction, temperature)


_OPTIONS = {"pin_trajectories": False}
_TARGETS = (
    ("ramannoodle.pmodel._interpolation", "InterpolationModel", "calc_polarizabilities", _patched_calc_polarizabilities),
    ("ramannoodle.dynamics._trajectory", "Trajectory", "get_raman_spectrum", _patched_get_raman_spectrum),
    ("ramannoodle.spectrum._raman", "MDRamanSpectrum", "measure", _patched_measure),
    ("ramannoodle.spectrum.utils", None, "convolve_spectrum", convolve_spectrum),
)


def install(pin_trajectories: bool = False) -> list:
    """Patch the hot path of an importable ``ramannoodle`` (see the module docstring).  Returns the
    patched names.  ``pin_trajectories=True`` page-locks a ``Trajectory``'s positions the first time it
    is evaluated (worth it when the trajectory is evaluated more than once, e.g. mask studies)."""
    _OPTIONS["pin_trajectories"] = bool(pin_trajectories)
    patched = []
    for module_name, owner_name, attribute, replacement in _TARGETS:
        module = importlib.import_module(module_name)
        owner = getattr(module, owner_name) if owner_name else module
        key = (module_name, owner_name, attribute)
        _ORIGINALS.setdefault(key, getattr(owner, attribute))
        setattr(owner, attribute, replacement)
        patched.append(".".join(part for part in key if part))
    return patched


def uninstall() -> None:
    """Undo ``install()``."""
    for (module_name, owner_name, attribute), original in list(_ORIGINALS.items()):
        module = importlib.import_module(module_name)
        owner = getattr(module, owner_name) if owner_name else module
        setattr(owner, attribute, original)
    _ORIGINALS.clear()
