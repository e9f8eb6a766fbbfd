"""Trajectory ingest — fast replacement for ``ramannoodle.io.vasp.xdatcar.read_positions_ts`` /
``read_trajectory`` (``ramannoodle/io/vasp/xdatcar.py:21-81``).

The text is parsed by the native library (mmap + a pool of threads running a correctly rounded
decimal parser, so every value equals Python's ``float(token)``), straight into the caller's buffer
(page-locked when ``read_trajectory`` runs on a GPU box, so the ``Trajectory`` streams to the device
at full PCIe bandwidth).  Only direct-coordinate
XDATCAR frames are handled; for anything else use the reference's readers.
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


def _scan(path: str):
    frames = ctypes.c_int64()
    atoms = ctypes.c_int64()
    lattice = np.zeros((3, 3))
    status = _lib.lib().rn_xdatcar_scan(path.encode(), ctypes.byref(frames), ctypes.byref(atoms),
                                        ctypes.c_void_p(lattice.ctypes.data))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return int(frames.value), int(atoms.value), lattice


def read_positions_ts(filepath, num_threads: int = 0, out: np.ndarray | None = None,
                      wrap: bool = False) -> np.ndarray:
    """Fractional positions time series (S,N,3) from a VASP XDATCAR file; ``wrap`` applies the
    periodic wrap ``x - x // 1`` while parsing (what ``Trajectory`` does to its input)."""
    path = os.fspath(filepath)
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found")
    frames, atoms, _ = _scan(path)
    if out is None:
        out = np.empty((frames, atoms, 3), dtype=np.float64)
    elif out.shape != (frames, atoms, 3) or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float64 array of shape {(frames, atoms, 3)}")
    status = _lib.lib().rn_xdatcar_read(path.encode(), ctypes.c_void_p(out.ctypes.data), frames, atoms, num_threads,
                                        int(bool(wrap)))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return out


def read_lattice(filepath) -> np.ndarray:
    """Scaled lattice (3,3) of an XDATCAR file (rows are lattice vectors, Å)."""
    path = os.fspath(filepath)
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found")
    return _scan(path)[2]


def read_trajectory(filepath, timestep: float, file_format: str = "xdatcar", num_threads: int = 0) -> Trajectory:
    """``Trajectory`` from a trajectory file (``ramannoodle/io/generic.py: read_trajectory``);
    ``file_format`` must be ``"xdatcar"`` (the timestep is not stored in XDATCAR files)."""
    if file_format != "xdatcar":
        raise ValueError(f"unsupported format: {file_format}")
    path = os.fspath(filepath)
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found")
    frames, atoms, _ = _scan(path)
    owner = None
    try:
        import torch  # pylint: disable=import-outside-toplevel

        if torch.cuda.is_available() and frames > 0:
            owner = torch.empty((frames, atoms, 3), dtype=torch.float64, pin_memory=True)
    except (ImportError, RuntimeError):
        owner = None
    out = owner.numpy() if owner is not None else None
    positions = read_positions_ts(path, num_threads=num_threads, out=out, wrap=True)
    return Trajectory._from_wrapped(positions, timestep, owner)  # pylint: disable=protected-access
