"""Trajectory ingest — fast replacements for the reference's trajectory readers
(``ramannoodle/io/generic.py: read_trajectory``):

* XDATCAR: ``ramannoodle/io/vasp/xdatcar.py:21-81`` (``read_positions_ts`` / ``read_trajectory``),
* OUTCAR molecular dynamics: ``ramannoodle/io/vasp/outcar.py:497-538``,
* vasprun.xml molecular dynamics: ``ramannoodle/io/vasp/vasprun.py:298-330``.

The text is parsed by the native library (mmap + a pool of threads running a correctly rounded
decimal parser, so every value equals Python's ``float(token)``), straight into the caller's buffer
(page-locked when ``read_trajectory`` runs on a GPU box, so the ``Trajectory`` streams to the device
at full PCIe bandwidth).  XDATCAR files with frames in Cartesian coordinates (VASP writes Direct) take a
Python walk with the reference's own numpy calls (``poscar.py:118-119``).  vasprun.xml files the
native tokenizer does not understand are re-read with the standard library's ElementTree under the
reference's rules.
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


class _CartesianFrames(Exception):
    """The native scan met a frame in Cartesian coordinates (the Python walk below converts those)."""


def _scan(path: str):
    frames = ctypes.c_int64()
    atoms = ctypes.c_int64()
    lattice = np.zeros((3, 3))
    status = _lib.lib().rn_xdatcar_scan(path.encode(), ctypes.byref(frames), ctypes.byref(atoms),
                                        ctypes.c_void_p(lattice.ctypes.data))
    if status == -3:  # RN_ERR_UNSUPPORTED
        raise _CartesianFrames(_lib.last_error())
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return int(frames.value), int(atoms.value), lattice


def _xdatcar_header_python(file):
    """``(lattice, num_atoms)`` from the header lines of an XDATCAR file (``poscar.py:13-77``)."""
    file.readline()
    line = file.readline()
    try:
        scale_factor = float(line)
    except ValueError as exc:
        raise InvalidFileException(f"scale factor could not be parsed: {line}") from exc
    rows = []
    for _ in range(3):
        line = file.readline()
        try:
            vector = np.array([float(item) for item in line.split()[0:3]])
        except ValueError as exc:
            raise InvalidFileException(f"lattice could not be parsed: {line}") from exc
        if vector.shape != (3,):
            raise InvalidFileException(f"lattice could not be parsed: {line}")
        rows.append(vector)
    symbols = file.readline().split()
    if len(symbols) == 0:
        raise InvalidFileException("no atom symbols found")
    line = file.readline()
    counts = line.split()
    if len(counts) != len(symbols):
        raise InvalidFileException(f"wrong number of ion counts: {len(counts)} != {len(symbols)}")
    try:
        num_atoms = sum(int(count) for count in counts)
    except ValueError as exc:
        raise InvalidFileException(f"could not parse counts: {line}") from exc
    return np.array(rows) * scale_factor, num_atoms


def _xdatcar_positions_python(path: str) -> np.ndarray:
    """The reference's XDATCAR walk (``xdatcar.py:21-56``, ``poscar.py:13-119``), used for files whose frames
    are in Cartesian coordinates — VASP writes Direct, so this is the rare path: each such frame becomes
    ``positions @ np.linalg.inv(lattice)`` (``poscar.py:118-119``), the same numpy calls as the reference."""
    positions_ts = []
    with open(path, "r", encoding="utf-8") as file:
        lattice, num_atoms = _xdatcar_header_python(file)
        while True:
            label = file.readline()
            if len(label.strip()) == 0 or label[0].strip() == "":
                break  # "missing first character in coordinate format": the end of the series
            if label[0].lower() == "s":  # selective dynamics
                label = file.readline()
            cart_mode = label[:1].lower() == "c"
            if not cart_mode and label[:1].lower() != "d":
                raise InvalidFileException(f"unrecognized coordinate format: {label}")
            positions = []
            for _ in range(num_atoms):
                line = file.readline()
                try:
                    position = [float(item) for item in line.split()[0:3]]
                except ValueError as exc:
                    raise InvalidFileException(f"positions could not be parsed: {line}") from exc
                if len(position) != 3:
                    raise InvalidFileException(f"positions could not be parsed: {line}")
                positions.append(position)
            frame = np.array(positions)
            positions_ts.append(frame @ np.linalg.inv(lattice) if cart_mode else frame)
    return np.array(positions_ts)


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
    try:
        frames, atoms, _ = _scan(path)
    except _CartesianFrames:
        positions = _xdatcar_positions_python(path)
        if wrap:
            positions = positions - positions // 1
        if out is not None and positions.shape == out.shape:
            out[...] = positions
            return out
        return positions
    out = _output(out, frames, atoms)
    status = _lib.lib().rn_xdatcar_read(path.encode(), ctypes.c_void_p(out.ctypes.data), frames, atoms, num_threads,
                                        int(bool(wrap)))
    if status != 0:
        raise InvalidFileException(_lib.last_error())
    return out


def read_lattice(filepath) -> np.ndarray:
    """Scaled lattice (3,3) of an XDATCAR file (rows are lattice vectors, Å)."""
    path = _checked_path(filepath)
    try:
        return _scan(path)[2]
    except _CartesianFrames:
        with open(path, "r", encoding="utf-8") as file:
            return _xdatcar_header_python(file)[0]


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


def _vasprun_positions_etree(path: str):
    """The reference's rules (``vasprun.py:53-70,281-330``) on a standard-library ElementTree: the
    fallback for files the native tokenizer leaves alone, and its cross-check in the tests."""
    import xml.etree.ElementTree as ET  # pylint: disable=import-outside-toplevel

    try:
        root = ET.parse(path).getroot()
    except ET.ParseError as exc:
        raise InvalidFileException("root xml element could not be found") from exc
    positions_ts = []
    for structure in root.iterfind("structure"):
        if "name" in structure.attrib:  # skip named structures
            continue
        varray = structure.find("varray")
        if varray is None:
            raise InvalidFileException("structure varray not found")
        rows = []
        for child in varray:
            if child.text is None:
                raise InvalidFileException("varray child text not found")
            rows.append([float(token) for token in child.text.split()])
        positions_ts.append(np.array(rows))
    if len(positions_ts) == 0:
        raise InvalidFileException("no trajectory found")
    element = root.find("./parameters/separator[@name='ionic']/i/[@name='POTIM']")
    if element is None:
        raise InvalidFileException("timestep not found")
    if element.text is None:
        raise InvalidFileException("potim element has no text")
    return np.array(positions_ts), float(element.text.strip())


def read_vasprun_positions_ts(filepath, num_threads: int = 0, out: np.ndarray | None = None, wrap: bool = False):
    """``(positions_ts, timestep)`` of a vasprun.xml molecular-dynamics run: the unnamed root-level
    ``structure`` elements are the frames, POTIM is the timestep (``vasprun.py:298-330``)."""
    path = _checked_path(filepath)
    frames, atoms, timestep = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    status = _lib.lib().rn_vasprun_scan(path.encode(), ctypes.byref(frames), ctypes.byref(atoms), ctypes.byref(timestep))
    if status == 0:
        out = _output(out, int(frames.value), int(atoms.value))
        status = _lib.lib().rn_vasprun_read(path.encode(), ctypes.c_void_p(out.ctypes.data), int(frames.value),
                                            int(atoms.value), num_threads, int(bool(wrap)))
        if status == 0:
            return out, float(timestep.value)
    if status == -1:
        raise InvalidFileException(_lib.last_error())
    # RN_ERR_UNSUPPORTED (irregular markup / rows) or anything else: the ElementTree walk decides
    positions, file_timestep = _vasprun_positions_etree(path)
    if wrap:
        positions = positions - positions // 1
    return positions, file_timestep


def read_trajectory(filepath, timestep: float | None = None, file_format: str = "xdatcar",
                    num_threads: int = 0) -> Trajectory:
    """``Trajectory`` from a trajectory file (``ramannoodle/io/generic.py: read_trajectory``).

    ``file_format="xdatcar"`` needs ``timestep`` (fs; XDATCAR files do not store it,
    ``xdatcar.py:59-81``); ``"outcar"`` and ``"vasprun.xml"`` read it from the file (``outcar.py:481-494``,
    ``vasprun.py:281-295``) unless given.
    The text is parsed straight into the (pinned) buffer the ``Trajectory`` owns."""
    path = _checked_path(filepath)
    if file_format == "xdatcar":
        if timestep is None:
            raise ValueError("timestep is required for xdatcar trajectories")
        try:
            frames, atoms, _ = _scan(path)
        except _CartesianFrames:  # rare: Cartesian frames go through the Python walk and the usual constructor
            return Trajectory(read_positions_ts(path), timestep)
        owner, out = _pinned(frames, atoms)
        positions = read_positions_ts(path, num_threads=num_threads, out=out, wrap=True)
    elif file_format == "outcar":
        frames, atoms, _, file_timestep = _scan_outcar(path)
        owner, out = _pinned(frames, atoms)
        positions, _, _ = read_outcar_positions_ts(path, num_threads=num_threads, out=out, wrap=True)
        if timestep is None:
            timestep = file_timestep
    elif file_format in ("vasprun.xml", "vasprun"):
        frames, atoms, file_timestep = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
        owner = out = None
        if _lib.lib().rn_vasprun_scan(path.encode(), ctypes.byref(frames), ctypes.byref(atoms),
                                      ctypes.byref(file_timestep)) == 0:
            owner, out = _pinned(int(frames.value), int(atoms.value))
        positions, parsed_timestep = read_vasprun_positions_ts(path, num_threads=num_threads, out=out, wrap=True)
        if positions is not out:
            return Trajectory(positions, parsed_timestep if timestep is None else timestep)
        if timestep is None:
            timestep = parsed_timestep
    else:
        raise ValueError(f"unsupported format: {file_format}")
    return Trajectory._from_wrapped(positions, timestep, owner)  # pylint: disable=protected-access
