"""Error types and message formats at the drop-in boundary.

The reference pins exact substrings of these messages in its tests
(``test/tests/test_trajectory_spectrum.py:96-139``, ``test_art.py:358-369``,
``test_phonon_spectrum.py:452-609``); the formats below reproduce
``ramannoodle/exceptions.py:57-105`` so the parity tests read like the reference's own.
"""
from __future__ import annotations

from typing import Any, Sequence


class UserError(Exception):
    """The user has done something they shouldn't (``ramannoodle/exceptions.py:33-39``)."""


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing, failed to load, or a native call returned an error."""


def shape_string(shape: Sequence[int | None]) -> str:
    """``(3,_,3)``-style rendering; ``None`` prints as ``_`` (``exceptions.py:42-54``)."""
    body = ",".join("_" if dim is None else str(dim) for dim in shape)
    if len(shape) == 1:
        body += ","
    return f"({body})"


def get_type_error(name: str, value: Any, correct_type: str) -> TypeError:
    """``exceptions.py:57-63``."""
    return TypeError(f"{name} should have type {correct_type}, not {type(value).__name__}")


def get_shape_error(name: str, array: Any, desired_shape: str) -> ValueError:
    """``exceptions.py:66-72``."""
    return ValueError(f"{name} has wrong shape: {shape_string(array.shape)} != {desired_shape}")


def verify_ndarray(name: str, array: Any) -> None:
    """``exceptions.py:75-83``."""
    if not hasattr(array, "shape"):
        raise get_type_error(name, array, "ndarray")


def verify_ndarray_shape(name: str, array: Any, shape: Sequence[int | None]) -> None:
    """``exceptions.py:86-105``: int entries are checked, ``None`` entries are free."""
    if not hasattr(array, "shape") or not hasattr(array, "ndim"):
        raise get_type_error(name, array, "ndarray")
    if len(shape) != array.ndim:
        raise get_shape_error(name, array, shape_string(shape))
    for have, want in zip(array.shape, shape):
        if want is not None and have != want:
            raise get_shape_error(name, array, shape_string(shape))
