"""Vectorised stand-ins for the reference's basis-vector scans (host-side).

``InterpolationModel.add_dof`` / ``add_art`` check every new degree of freedom against all
existing basis vectors with Python loops (``ramannoodle/structure/_symmetry_utils.py:42-133``,
called from ``pmodel/_interpolation.py:292,568,678`` and ``structure/_reference.py:226,242``):
one ``np.linalg.norm``, one ``np.dot`` and one ``np.allclose`` per pair, O(J²) interpreter
round trips over a model's construction — minutes for J = 576, hours for the 4 600-DOF supercell.
The functions below return the same indices from one matrix-vector product per call; the matrix of
normalised basis vectors is kept between calls while the caller keeps appending to the same list
(which is what ``add_dof`` does with ``self._cart_basis_vectors``).

``accelerate_construction()`` swaps them into an importable ``ramannoodle``; nothing else of model
construction (symmetry operations, spline fitting) is touched, and nothing here runs on the GPU.
"""
from __future__ import annotations

import numpy as np

from .exceptions import get_type_error

# np.allclose / np.isclose defaults, as the reference uses them (b = +-1): |a - b| <= atol + rtol |b|
_TOL = 1e-8 + 1e-5 * 1.0


class _RowCache:
    """Normalised, flattened rows of one caller-owned list that only ever grows at its end."""

    def __init__(self) -> None:
        self.key = None
        self.rows = np.empty((0, 0))
        self.count = 0
        self.last = None

    def matrix(self, vectors) -> np.ndarray:
        if not isinstance(vectors, list):
            return _normalised_rows(list(vectors), 0, None)[0]
        count = len(vectors)
        same = (self.key == id(vectors) and 0 < self.count <= count and vectors[self.count - 1] is self.last)
        start = self.count if same else 0
        if count == start:
            return self.rows[:count]
        fresh, width = _normalised_rows(vectors, start, self.rows.shape[1] if same else None)
        if not same or self.rows.shape[0] < count:
            grown = np.empty((max(count, 2 * self.rows.shape[0] if same else count), width))
            if same:
                grown[:start] = self.rows[:start]
            self.rows = grown
        self.rows[start:count] = fresh
        self.key, self.count, self.last = id(vectors), count, vectors[count - 1]
        return self.rows[:count]


def _normalised_rows(vectors, start: int, width):
    rows = []
    for index in range(start, len(vectors)):
        vector = vectors[index]
        try:
            flat = np.asarray(vector, dtype=np.float64).reshape(-1)
            rows.append(flat / np.linalg.norm(flat))
        except (TypeError, ValueError) as exc:
            raise get_type_error(f"vectors[{index}]", vector, "ndarray") from exc
    if not rows:
        return np.empty((0, width or 0)), width or 0
    size = rows[0].shape[0] if width is None else width
    for index, row in enumerate(rows):
        if row.shape[0] != size:
            raise ValueError(f"vectors[{start + index}] has {row.shape[0]} components, expected {size}")
    return np.stack(rows), size


_ORTHOGONAL_CACHE = _RowCache()


def _unit(vector, name: str) -> np.ndarray:
    try:
        flat = np.asarray(vector, dtype=np.float64).reshape(-1)
        return flat / float(np.linalg.norm(flat))
    except (TypeError, ValueError) as exc:
        raise get_type_error(name, vector, "ndarray") from exc


def _first(mask: np.ndarray) -> int:
    hits = np.flatnonzero(mask)
    return int(hits[0]) if hits.size else -1


def is_orthogonal_to_all(vector_1, vectors) -> int:
    """First index of a vector that is not orthogonal to ``vector_1``, else -1
    (``_symmetry_utils.py:42-75``: ``not np.allclose(dot + 1, 1)`` on normalised vectors)."""
    unit = _unit(vector_1, "vector_1")
    rows = _ORTHOGONAL_CACHE.matrix(vectors)
    if rows.shape[0] == 0:
        return -1
    dots = rows @ unit
    return _first(~(np.abs((dots + 1.0) - 1.0) <= _TOL))


def _collinear(vector_1, vectors) -> np.ndarray:
    unit = _unit(vector_1, "vector_1")
    rows, _ = _normalised_rows(list(vectors), 0, None)
    if rows.shape[0] == 0:
        return np.zeros(0, dtype=bool)
    dots = rows @ unit
    # are_collinear (:13-39): np.allclose(dot, 1) or np.isclose(dot, -1)
    return (np.abs(dots - 1.0) <= _TOL) | (np.abs(dots + 1.0) <= _TOL)


def is_collinear_with_all(vector_1, vectors) -> int:
    """First index of a vector that is not collinear with ``vector_1``, else -1 (``:78-104``)."""
    return _first(~_collinear(vector_1, vectors))


def is_non_collinear_with_all(vector_1, vectors) -> int:
    """First index of a vector that is collinear with ``vector_1``, else -1 (``:107-133``)."""
    return _first(_collinear(vector_1, vectors))


_PATCHED: dict = {}


def accelerate_construction() -> list:
    """Swap the vectorised scans into an importable ``ramannoodle`` (its ``_symmetry_utils`` module and
    the names ``pmodel/_interpolation.py`` imported from it).  Returns the patched attribute names;
    ``restore_construction()`` undoes it."""
    import importlib  # pylint: disable=import-outside-toplevel

    utils = importlib.import_module("ramannoodle.structure._symmetry_utils")
    interpolation = importlib.import_module("ramannoodle.pmodel._interpolation")
    replacements = {"is_orthogonal_to_all": is_orthogonal_to_all, "is_collinear_with_all": is_collinear_with_all,
                    "is_non_collinear_with_all": is_non_collinear_with_all}
    patched = []
    for module in (utils, interpolation):
        for name, function in replacements.items():
            if hasattr(module, name):
                key = (module.__name__, name)
                _PATCHED.setdefault(key, getattr(module, name))
                setattr(module, name, function)
                patched.append(f"{module.__name__}.{name}")
    return patched


def restore_construction() -> None:
    """Undo ``accelerate_construction()``."""
    import importlib  # pylint: disable=import-outside-toplevel

    for (module_name, name), original in list(_PATCHED.items()):
        setattr(importlib.import_module(module_name), name, original)
    _PATCHED.clear()
