"""Trajectory ingest — fast replacements for the reference's trajectory readers
(``ramannoodle/io/generic.py: read_trajectory``):

* XDATCAR: ``ramannoodle/io/vasp/xdatcar.py:21-81`` (``read_positions_ts`` / ``read_trajectory``),
* OUTCAR molecular dynamics: ``ramannoodle/io/vasp/outcar.py:497-538``.

The text is parsed by the native library (mmap + a pool of threads running a correctly rounded
decimal parser, so every value equals Python's ``float(token)``), straight into the caller's buffer
(page-locked when ``read_trajectory`` runs on a GPU box, so the ``Trajectory`` streams to the device
at full PCIe bandwidth).  Only direct-coordinate XDATCAR frames are handled; vasprun.xml
trajectories are not covered — use the reference's reader for those.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from .dynamics import Trajectory


class InvalidFileException(Exception):
    """File cannot be read, likely due to an invalid or unexpected format
    (``ramannoodle/exceptions.py:12-13``)."""


def _checked_path(filepath) -> str:
    path = os.fspath(filepath)
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found")
    return path


def _scan(path: str):
    frames = ctypes.c_int64()
    atoms = ctypes.c_int64()
    lattice = np.zeros((3, 3))
    status = _lib.lib().rn_xdatcar_scan(path.encode(), ctypes.byref(frames), ctypes.byref(atoms),
                                        ctypes.c_void_p(lattice.ctypes.data))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return int(frames.value), int(atoms.value), lattice


def _scan_outcar(path: str):
    frames = ctypes.c_int64()
    atoms = ctypes.c_int64()
    lattice = np.zeros((3, 3))
    timestep = ctypes.c_double()
    status = _lib.lib().rn_outcar_scan(path.encode(), ctypes.byref(frames), ctypes.byref(atoms),
                                       ctypes.c_void_p(lattice.ctypes.data), ctypes.byref(timestep))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return int(frames.value), int(atoms.value), lattice, float(timestep.value)


def _output(out, frames: int, atoms: int) -> np.ndarray:
    if out is None:
        return np.empty((frames, atoms, 3), dtype=np.float64)
    if out.shape != (frames, atoms, 3) or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float64 array of shape {(frames, atoms, 3)}")
    return out


def _pinned(frames: int, atoms: int):
    """A page-locked (frames, atoms, 3) buffer when a GPU is present (owner tensor, numpy view)."""
    try:
        import torch  # pylint: disable=import-outside-toplevel

        if torch.cuda.is_available() and frames > 0:
            owner = torch.empty((frames, atoms, 3), dtype=torch.float64, pin_memory=True)
            return owner, owner.numpy()
    except (ImportError, RuntimeError):
        pass
    return None, None


def read_positions_ts(filepath, num_threads: int = 0, out: np.ndarray | None = None,
                      wrap: bool = False) -> np.ndarray:
    """Fractional positions time series (S,N,3) from a VASP XDATCAR file; ``wrap`` applies the
    periodic wrap ``x - x // 1`` while parsing (what ``Trajectory`` does to its input)."""
    path = _checked_path(filepath)
    frames, atoms, _ = _scan(path)
    out = _output(out, frames, atoms)
    status = _lib.lib().rn_xdatcar_read(path.encode(), ctypes.c_void_p(out.ctypes.data), frames, atoms, num_threads,
                                        int(bool(wrap)))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return out


def read_lattice(filepath) -> np.ndarray:
    """Scaled lattice (3,3) of an XDATCAR file (rows are lattice vectors, Å)."""
    return _scan(_checked_path(filepath))[2]


def read_outcar_positions_ts(filepath, num_threads: int = 0, out: np.ndarray | None = None, wrap: bool = False,
                             cartesian: bool = False):
    """``(positions_ts, lattice, timestep)`` of an OUTCAR molecular-dynamics run: fractional
    positions ``cart @ inv(lattice)`` (``outcar.py:529``), or the Cartesian coordinates as written
    (Å) with ``cartesian=True``.  Machine-learned / ab-initio duplicate steps are dropped as the
    reference does (``outcar.py:520-527``)."""
    path = _checked_path(filepath)
    frames, atoms, lattice, timestep = _scan_outcar(path)
    out = _output(out, frames, atoms)
    inverse = None if cartesian else np.ascontiguousarray(np.linalg.inv(lattice))
    status = _lib.lib().rn_outcar_read(path.encode(), ctypes.c_void_p(out.ctypes.data), frames, atoms,
                                       None if inverse is None else ctypes.c_void_p(inverse.ctypes.data), num_threads,
                                       int(bool(wrap)))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return out, lattice, timestep


def read_trajectory(filepath, timestep: float | None = None, file_format: str = "xdatcar",
                    num_threads: int = 0) -> Trajectory:
    """``Trajectory`` from a trajectory file (``ramannoodle/io/generic.py: read_trajectory``).

    ``file_format="xdatcar"`` needs ``timestep`` (fs; XDATCAR files do not store it,
    ``xdatcar.py:59-81``); ``"outcar"`` reads it from the file (``outcar.py:481-494``) unless given.
    The text is parsed straight into the (pinned) buffer the ``Trajectory`` owns."""
    path = _checked_path(filepath)
    if file_format == "xdatcar":
        if timestep is None:
            raise ValueError("timestep is required for xdatcar trajectories")
        frames, atoms, _ = _scan(path)
        owner, out = _pinned(frames, atoms)
        positions = read_positions_ts(path, num_threads=num_threads, out=out, wrap=True)
    elif file_format == "outcar":
        frames, atoms, _, file_timestep = _scan_outcar(path)
        owner, out = _pinned(frames, atoms)
        positions, _, _ = read_outcar_positions_ts(path, num_threads=num_threads, out=out, wrap=True)
        if timestep is None:
            timestep = file_timestep
    else:
        raise ValueError(f"unsupported format: {file_format}")
    return Trajectory._from_wrapped(positions, timestep, owner)  # pylint: disable=protected-access
