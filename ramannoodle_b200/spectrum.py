"""MD Raman spectrum and smearing — drop-ins for ``ramannoodle.spectrum``.

``MDRamanSpectrum.measure`` (``ramannoodle/spectrum/_raman.py:241-309``),
``calc_signal_spectrum`` (``ramannoodle/spectrum/utils.py:95-124``) and
``convolve_spectrum`` (``ramannoodle/spectrum/utils.py:12-73``) keep their signatures, return
types and error messages; the arithmetic runs in the CUDA library.
"""
from __future__ import annotations

import ctypes
import threading
from collections import OrderedDict

import numpy as np

from . import _lib
from .abstract import RamanSpectrum
from .exceptions import get_type_error, verify_ndarray, verify_ndarray_shape

BOLTZMANN_CONSTANT = 8.617333262e-5  # eV/K (ramannoodle/constants.py:249)


def _torch():
    import torch  # pylint: disable=import-outside-toplevel

    return torch


def _is_torch_tensor(obj) -> bool:
    return type(obj).__module__.startswith("torch") and hasattr(obj, "data_ptr")


def _current_device() -> int:
    torch = _torch()
    if not torch.cuda.is_available():
        _lib.require_device(0)  # raises NativeLibraryError with the reason
    return int(torch.cuda.current_device())


def _stream(device: int) -> ctypes.c_void_p:
    return ctypes.c_void_p(_torch().cuda.current_stream(device).cuda_stream)


class _Plan:
    """``rn_spectrum_plan`` for one series length on one device."""

    def __init__(self, num_frames: int, device: int) -> None:
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().rn_spectrum_plan_create(num_frames, device, ctypes.byref(handle)),
                   "rn_spectrum_plan_create")
        self.handle = handle

    def close(self) -> None:
        if getattr(self, "handle", None):
            _lib.lib().rn_spectrum_plan_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


_PLAN_CACHE: "OrderedDict[tuple, _Plan]" = OrderedDict()
_PLAN_CACHE_SIZE = 4
_PLAN_CACHE_LOCK = threading.Lock()


def _get_plan(num_frames: int, device: int) -> _Plan:
    """A plan owns its work buffers, so it is shared only by calls that are ordered anyway: the same
    Python thread AND the same CUDA stream (two streams or two threads measuring series of the same length
    would otherwise race on the buffers)."""
    stream_id = 0
    try:
        import torch  # pylint: disable=import-outside-toplevel

        if torch.cuda.is_available():
            stream_id = int(torch.cuda.current_stream(device).cuda_stream)
    except (ImportError, RuntimeError):
        stream_id = 0
    key = (int(num_frames), int(device), stream_id, threading.get_ident())
    with _PLAN_CACHE_LOCK:
        return _get_plan_locked(key, num_frames, device)


def _get_plan_locked(key, num_frames: int, device: int) -> _Plan:
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        _lib.require_device(device)
        while len(_PLAN_CACHE) >= _PLAN_CACHE_SIZE:
            _, old = _PLAN_CACHE.popitem(last=False)
            old.close()
        plan = _Plan(num_frames, device)
        _PLAN_CACHE[key] = plan
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


def clear_plan_cache() -> None:
    with _PLAN_CACHE_LOCK:
        while _PLAN_CACHE:
            _, plan = _PLAN_CACHE.popitem()
            plan.close()


def get_bose_einstein_correction(wavenumbers, temperature):
    """``ramannoodle/spectrum/_raman.py:13-40`` (elementwise; evaluated on the host when
    called directly — inside ``measure`` the correction is fused into the GPU kernel)."""
    try:
        if temperature <= 0:
            raise ValueError(f"invalid temperature: {temperature} <= 0")
    except TypeError as exc:
        raise get_type_error("temperature", temperature, "float") from exc
    try:
        energy = wavenumbers * 29979245800.0 * 4.1357e-15  # in eV
        return 1 / (1 - np.exp(-energy / (BOLTZMANN_CONSTANT * temperature)))
    except TypeError as exc:
        raise get_type_error("wavenumbers", wavenumbers, "ndarray") from exc


def get_laser_correction(wavenumbers, laser_wavenumber):
    """``ramannoodle/spectrum/_raman.py:43-69``."""
    try:
        if laser_wavenumber <= 0:
            raise ValueError(f"invalid laser_wavenumber: {laser_wavenumber} <= 0")
    except TypeError as exc:
        raise get_type_error("laser_wavenumber", laser_wavenumber, "float") from exc
    try:
        return ((wavenumbers - laser_wavenumber) / 10000) ** 4 / wavenumbers
    except TypeError as exc:
        raise get_type_error("wavenumbers", wavenumbers, "ndarray") from exc


class MDRamanSpectrum(RamanSpectrum):
    """Molecular-dynamics Raman spectrum (``ramannoodle/spectrum/_raman.py:197-309``).

    ``polarizability_ts`` is an array with shape (S,3,3) — numpy, or a CUDA tensor left on the
    device by ``Trajectory.get_raman_spectrum``.
    """

    def __init__(self, polarizability_ts, timestep: float):
        verify_ndarray_shape("polarizability_ts", polarizability_ts, (None, 3, 3))
        self._polarizability_ts = polarizability_ts
        self._timestep = timestep

    @property
    def polarizability_ts(self) -> np.ndarray:
        """The (S,3,3) series as numpy (copied from the GPU on first access if needed)."""
        if _is_torch_tensor(self._polarizability_ts):
            return self._polarizability_ts.detach().cpu().numpy()
        return self._polarizability_ts

    @property
    def timestep(self) -> float:
        return self._timestep

    def _device_series(self):
        torch = _torch()
        series = self._polarizability_ts
        if _is_torch_tensor(series):
            if not series.is_cuda:
                series = series.to(f"cuda:{_current_device()}")
            return series.to(torch.float64).contiguous()
        device = _current_device()
        return torch.from_numpy(np.ascontiguousarray(series, dtype=np.float64)).to(f"cuda:{device}")

    # pylint: disable=too-many-arguments,too-many-positional-arguments
    def measure_device(self, orientation="polycrystalline", laser_correction=False, laser_wavelength=522,
                       bose_einstein_correction=False, temperature=300):
        """``measure`` with the result left on the GPU (two CUDA tensors)."""
        if orientation != "polycrystalline":
            raise NotImplementedError("only polycrystalline spectra are supported for now")
        laser_wavenumber = None
        if laser_correction:
            laser_wavenumber = 10000000 / laser_wavelength
            try:
                if laser_wavenumber <= 0:
                    raise ValueError(f"invalid laser_wavenumber: {laser_wavenumber} <= 0")
            except TypeError as exc:
                raise get_type_error("laser_wavenumber", laser_wavenumber, "float") from exc
        if bose_einstein_correction:
            try:
                if temperature <= 0:
                    raise ValueError(f"invalid temperature: {temperature} <= 0")
            except TypeError as exc:
                raise get_type_error("temperature", temperature, "float") from exc
        torch = _torch()
        series = self._device_series()
        num_frames = int(series.shape[0])
        if num_frames < 2:
            raise ValueError("polarizability_ts must contain at least 2 configurations")
        device = int(series.device.index or 0)
        points = int(_lib.lib().rn_spectrum_num_points(num_frames))
        with torch.cuda.device(device):
            wavenumbers = torch.empty(points, dtype=torch.float64, device=series.device)
            intensities = torch.empty(points, dtype=torch.float64, device=series.device)
            if points > 0:
                plan = _get_plan(num_frames, device)
                status = _lib.lib().rn_md_spectrum(
                    plan.handle, ctypes.c_void_p(series.data_ptr()), float(self._timestep),
                    1 if laser_correction else 0, float(laser_wavelength) if laser_correction else 0.0,
                    1 if bose_einstein_correction else 0, float(temperature) if bose_einstein_correction else 0.0,
                    ctypes.c_void_p(wavenumbers.data_ptr()), ctypes.c_void_p(intensities.data_ptr()), _stream(device))
                _lib.check(status, "rn_md_spectrum")
        return wavenumbers, intensities

    def measure(self, orientation="polycrystalline", laser_correction=False, laser_wavelength=522,
                bose_einstein_correction=False, temperature=300):
        """Raw polycrystalline Raman spectrum -> (wavenumbers, intensities), numpy arrays of
        shape (ceil((S-1)/2) - 1,)."""
        wavenumbers, intensities = self.measure_device(orientation, laser_correction, laser_wavelength,
                                                       bose_einstein_correction, temperature)
        return wavenumbers.cpu().numpy(), intensities.cpu().numpy()


class PhononRamanSpectrum(RamanSpectrum):
    """Phonon-based first-order Raman spectrum (``ramannoodle/spectrum/_raman.py:72-194``).

    ``measure`` is O(M) elementwise work on M = 3N Raman tensors (a few hundred numbers), so it
    stays on the host with the reference's exact formulas; the data-parallel part of the phonon
    path — the 2·M central-difference polarizabilities — is batched onto the GPU by
    ``ramannoodle_b200.dynamics.Phonons.get_raman_spectrum``."""

    def __init__(self, phonon_wavenumbers, raman_tensors) -> None:
        verify_ndarray_shape("phonon_wavenumbers", phonon_wavenumbers, (None,))
        verify_ndarray_shape("raman_tensors", raman_tensors, (len(phonon_wavenumbers), 3, 3))
        self._phonon_wavenumbers = phonon_wavenumbers
        self._raman_tensors = raman_tensors

    @property
    def phonon_wavenumbers(self) -> np.ndarray:
        return self._phonon_wavenumbers.copy()

    @property
    def raman_tensors(self) -> np.ndarray:
        return self._raman_tensors.copy()

    # pylint: disable=too-many-arguments,too-many-positional-arguments
    def measure(self, orientation="polycrystalline", laser_correction=False, laser_wavelength=522,
                bose_einstein_correction=False, temperature=300):
        if orientation != "polycrystalline":
            raise NotImplementedError("only polycrystalline spectra are supported for now")
        tensors = self._raman_tensors
        alpha_squared = ((tensors[:, 0, 0] + tensors[:, 1, 1] + tensors[:, 2, 2]) / 3.0) ** 2
        gamma_squared = (
            (tensors[:, 0, 0] - tensors[:, 1, 1]) ** 2
            + (tensors[:, 0, 0] - tensors[:, 2, 2]) ** 2
            + (tensors[:, 1, 1] - tensors[:, 2, 2]) ** 2
            + 6.0 * (tensors[:, 0, 1] ** 2 + tensors[:, 0, 2] ** 2 + tensors[:, 1, 2] ** 2)
        ) / 2.0
        intensities = 45.0 * alpha_squared + 7.0 * gamma_squared
        if laser_correction:
            laser_wavenumber = 10000000 / laser_wavelength
            intensities *= get_laser_correction(self._phonon_wavenumbers, laser_wavenumber)
        if bose_einstein_correction:
            intensities *= get_bose_einstein_correction(self._phonon_wavenumbers, temperature)
        return self._phonon_wavenumbers, intensities


def calc_signal_spectrum(signal, sampling_rate: float):
    """Spectrum of one real signal (``ramannoodle/spectrum/utils.py:95-124``): the
    positive-frequency Fourier transform of its autocorrelation; ceil(S/2) points."""
    verify_ndarray_shape("signal", signal, (None,))
    torch = _torch()
    device = _current_device()
    data = torch.from_numpy(np.ascontiguousarray(signal, dtype=np.float64)).to(f"cuda:{device}")
    length = int(data.shape[0])
    if length < 1:
        raise ValueError("signal is empty")
    points = (length + 1) // 2
    with torch.cuda.device(device):
        wavenumbers = torch.empty(points, dtype=torch.float64, device=data.device)
        intensities = torch.empty(points, dtype=torch.float64, device=data.device)
        plan = _get_plan(length + 1, device)
        status = _lib.lib().rn_signal_spectrum(plan.handle, ctypes.c_void_p(data.data_ptr()), float(sampling_rate),
                                               ctypes.c_void_p(wavenumbers.data_ptr()),
                                               ctypes.c_void_p(intensities.data_ptr()), _stream(device))
        _lib.check(status, "rn_signal_spectrum")
    return wavenumbers.cpu().numpy(), intensities.cpu().numpy()


def convolve_spectrum(wavenumbers, intensities, function: str = "gaussian", width: float = 5,
                      out_wavenumbers=None):
    """Smear a spectrum (``ramannoodle/spectrum/utils.py:12-73``); same defaults, output grid
    and error messages.  Accepts numpy arrays (or CUDA tensors) and returns numpy arrays."""
    torch = _torch()
    if out_wavenumbers is None:
        if _is_torch_tensor(wavenumbers) and wavenumbers.is_cuda and wavenumbers.numel() > 0:
            # the default grid needs two numbers, not a copy of the spectrum
            low, high = (float(v) for v in torch.aminmax(wavenumbers))
        else:
            wn_host = wavenumbers.detach().cpu().numpy() if _is_torch_tensor(wavenumbers) else wavenumbers
            low, high = np.min(wn_host), np.max(wn_host)
        min_wavenumber = low - 100
        max_wavenumber = high + 100
        num_samples = int(np.rint(max_wavenumber - min_wavenumber))
        out_wavenumbers = np.linspace(min_wavenumber, max_wavenumber, num_samples)
    verify_ndarray_shape("out_wavenumbers", out_wavenumbers, (None,))
    verify_ndarray_shape("wavenumbers", wavenumbers, (None,))
    verify_ndarray_shape("intensities", intensities, (len(wavenumbers),))
    try:
        if width <= 0:
            raise ValueError(f"invalid width: {width} <= 0")
    except TypeError as exc:
        raise get_type_error("width", width, "float") from exc
    verify_ndarray("out_wavenumbers", out_wavenumbers)
    if function not in ("gaussian", "lorentzian"):
        raise ValueError(f"unsupported convolution type: {function}")
    kind = 0 if function == "gaussian" else 1

    device = _current_device()
    _lib.require_device(device)

    def to_device(array):
        if _is_torch_tensor(array):
            return array.to(device=f"cuda:{device}", dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64)).to(f"cuda:{device}")

    d_wn, d_in, d_out_wn = to_device(wavenumbers), to_device(intensities), to_device(out_wavenumbers)
    num_in, num_out = int(d_wn.shape[0]), int(d_out_wn.shape[0])
    with torch.cuda.device(device):
        d_out = torch.empty(num_out, dtype=torch.float64, device=d_wn.device)
        ws_bytes = int(_lib.lib().rn_convolve_workspace_size(num_in, num_out))
        workspace = torch.empty(max(ws_bytes // 8, 1), dtype=torch.float64, device=d_wn.device)
        status = _lib.lib().rn_convolve_spectrum(
            ctypes.c_void_p(d_wn.data_ptr()), ctypes.c_void_p(d_in.data_ptr()), num_in, kind, float(width),
            ctypes.c_void_p(d_out_wn.data_ptr()), num_out, ctypes.c_void_p(d_out.data_ptr()),
            ctypes.c_void_p(workspace.data_ptr()), _stream(device))
        _lib.check(status, "rn_convolve_spectrum")
    out_host = out_wavenumbers.detach().cpu().numpy() if _is_torch_tensor(out_wavenumbers) else out_wavenumbers
    return (out_host, d_out.cpu().numpy())
