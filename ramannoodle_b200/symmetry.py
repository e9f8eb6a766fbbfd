"""Space-group search for model construction (host-side, nothing here runs on the GPU).

``ReferenceStructure.__init__`` asks ``spglib.get_symmetry(cell, symprec, angle_tolerance)`` for the
rotations, translations and equivalent atoms of the cell (``ramannoodle/structure/_reference.py:114-122``)
and everything in model construction that uses symmetry is derived from those three arrays
(``_reference.py:26-47`` permutation matrices, ``:183-266`` ``get_equivalent_displacements``,
``pmodel/_interpolation.py:254-342``).  ``get_symmetry`` below returns the same dictionary from a direct
search, so that the models of this package's workloads — up to the 1536-atom, 4608-DOF supercell — can be
built through the reference's own ``add_dof`` / ``add_art`` where spglib is not installed:

* rotations: integer matrices ``W`` with entries in {-1, 0, 1} and ``|det W| = 1`` that keep the metric
  ``G = L L^T`` (``W^T G W = G`` within the tolerances) — complete for reduced cells and their supercells;
* translations: ``t = p_j - W p_a`` for the atoms j of the rarest species (a: its first atom); ``(W, t)`` is
  kept when every atom lands on an atom of its own species within ``symprec`` (Cartesian, minimum image);
  candidates are screened on a few atoms first and confirmed with a periodic k-d tree;
* equivalent atoms: orbits under the accepted operations, labelled by their smallest index (the convention
  ``get_equivalent_atom_dict`` relies on, ``_reference.py:166-181``).

Conventions are spglib's: ``x' = W x + t`` on fractional column vectors, identity first.
``install_spglib_stand_in()`` registers this module as ``spglib`` when the real one cannot be imported.
"""
from __future__ import annotations

import itertools
import sys
import types

import numpy as np

__all__ = ["get_symmetry", "install_spglib_stand_in"]


def _lattice_rotations(lattice: np.ndarray, symprec: float, angle_tolerance: float) -> np.ndarray:
    """Integer matrices W (n,3,3), entries in {-1,0,1}, |det| = 1, that map the lattice onto itself:
    the lengths of the transformed basis vectors agree within ``symprec`` and the angles between them
    within ``angle_tolerance`` degrees (< 0: within the arc ``symprec`` subtends at the vectors' ends)."""
    metric = lattice @ lattice.T
    entries = np.array(list(itertools.product((-1, 0, 1), repeat=9)), dtype=np.int64).reshape(-1, 3, 3)
    det = np.rint(np.linalg.det(entries.astype(np.float64))).astype(np.int64)
    entries = entries[np.abs(det) == 1]
    w = entries.astype(np.float64)
    new = np.transpose(w, (0, 2, 1)) @ metric @ w
    lengths = np.sqrt(np.diagonal(metric))
    new_lengths = np.sqrt(np.abs(np.diagonal(new, axis1=1, axis2=2)))
    ok = np.all(np.abs(new_lengths - lengths) <= symprec, axis=1)
    for i, j in ((0, 1), (0, 2), (1, 2)):
        cos_old = metric[i, j] / (lengths[i] * lengths[j])
        cos_new = new[:, i, j] / np.maximum(new_lengths[:, i] * new_lengths[:, j], 1e-300)
        if angle_tolerance > 0:
            diff = np.abs(np.degrees(np.arccos(np.clip(cos_new, -1, 1))) - np.degrees(np.arccos(np.clip(cos_old, -1, 1))))
            ok &= diff <= angle_tolerance
        else:
            # spglib's length-based criterion: sin(dtheta) * mean length <= symprec
            sin_diff = np.abs(np.sqrt(np.clip(1 - cos_new ** 2, 0, 1)) * cos_old - cos_new * np.sqrt(max(0.0, 1 - cos_old ** 2)))
            ok &= sin_diff * 0.5 * (lengths[i] + lengths[j]) <= symprec
    return entries[ok]


def _min_image(delta: np.ndarray) -> np.ndarray:
    return delta - np.rint(delta)


class _SpeciesIndex:
    """Periodic nearest-neighbour lookup of the atoms of one species (fractional coordinates)."""

    def __init__(self, positions: np.ndarray, members: np.ndarray) -> None:
        from scipy.spatial import cKDTree  # pylint: disable=import-outside-toplevel

        self.members = members
        self.points = positions[members] % 1.0
        self.points[self.points >= 1.0] = 0.0
        self.tree = cKDTree(self.points, boxsize=1.0)

    def nearest(self, query: np.ndarray) -> np.ndarray:
        wrapped = query % 1.0
        wrapped[wrapped >= 1.0] = 0.0
        _, index = self.tree.query(wrapped, k=1)
        return index


def _maps_onto_itself(rotation, translation, positions, lattice, species_sets, indexes, symprec):
    """The atom permutation of (W, t): image of atom i is atom perm[i], or None if some atom has no partner
    of its species within ``symprec``."""
    image = positions @ rotation.T + translation
    perm = np.empty(len(positions), dtype=np.int64)
    for members, index in zip(species_sets, indexes):
        nearest = index.nearest(image[members])
        delta = _min_image(image[members] - index.points[nearest]) @ lattice
        if np.any(np.einsum("ij,ij->i", delta, delta) > symprec * symprec):
            return None
        perm[members] = members[nearest]
    return perm


def get_symmetry(cell, symprec: float = 1e-5, angle_tolerance: float = -1.0, mag_symprec: float = -1.0,  # pylint: disable=unused-argument
                 is_magnetic: bool = True) -> dict | None:  # pylint: disable=unused-argument
    """``spglib.get_symmetry`` for ``cell = (lattice, positions, numbers)`` (rows of ``lattice`` are the
    basis vectors in Å, positions fractional): ``{"rotations": (n,3,3) int32, "translations": (n,3),
    "equivalent_atoms": (N,) int32}``; None when the cell is malformed (spglib's failure value, which
    ``ReferenceStructure`` turns into ``SymmetryException``)."""
    try:
        lattice = np.array(cell[0], dtype=np.float64)
        positions = np.array(cell[1], dtype=np.float64)
        numbers = np.array(cell[2])
    except (TypeError, ValueError, IndexError):
        return None
    if lattice.shape != (3, 3) or positions.ndim != 2 or positions.shape[1] != 3 or numbers.shape != (len(positions),) \
            or len(positions) == 0 or abs(np.linalg.det(lattice)) < 1e-12:
        return None
    num_atoms = len(positions)
    species = sorted(set(numbers.tolist()), key=lambda z: (int(np.sum(numbers == z)), z))
    species_sets = [np.flatnonzero(numbers == z) for z in species]
    indexes = [_SpeciesIndex(positions, members) for members in species_sets]
    anchor_set = species_sets[0]
    anchor = positions[anchor_set[0]]
    # a few atoms of every species screen the candidates before the full check
    probe = np.concatenate([members[:3] for members in species_sets])
    probe_species = np.concatenate([np.full(min(3, len(m)), k) for k, m in enumerate(species_sets)])

    rotations, translations, perms = [], [], []
    for rotation in _lattice_rotations(lattice, symprec, angle_tolerance):
        w = rotation.astype(np.float64)
        candidates = positions[anchor_set] - w @ anchor
        candidates -= np.floor(candidates)
        candidates[candidates >= 1.0] = 0.0
        # screening: image of the probe atoms under every candidate translation at once
        rotated = positions[probe] @ w.T
        alive = np.ones(len(candidates), dtype=bool)
        for k, index in enumerate(indexes):
            rows = np.flatnonzero(probe_species == k)
            if rows.size == 0:
                continue
            image = (rotated[rows][None, :, :] + candidates[:, None, :]).reshape(-1, 3)
            nearest = index.nearest(image)
            delta = _min_image(image - index.points[nearest]) @ lattice
            far = (np.einsum("ij,ij->i", delta, delta) > symprec * symprec).reshape(len(candidates), rows.size)
            alive &= ~far.any(axis=1)
        kept: list[np.ndarray] = []
        for translation in candidates[alive]:
            if any(np.all(np.abs(_min_image(translation - other)) @ np.abs(lattice) <= symprec) for other in kept):
                continue  # the same translation reached from two anchor images
            perm = _maps_onto_itself(w, translation, positions, lattice, species_sets, indexes, symprec)
            if perm is None:
                continue
            kept.append(translation)
            rotations.append(rotation)
            translations.append(translation)
            perms.append(perm)
    if not rotations:
        return None
    # identity first, then by rotation / translation (a deterministic order; spglib's own differs)
    identity = np.eye(3, dtype=np.int64)

    def order(k: int):
        is_identity = np.array_equal(rotations[k], identity) and np.allclose(_min_image(translations[k]), 0.0, atol=1e-9)
        return (0 if is_identity else 1, 0 if np.array_equal(rotations[k], identity) else 1,
                tuple(-rotations[k].reshape(-1)), tuple(np.round(translations[k], 9)))

    ranks = sorted(range(len(rotations)), key=order)
    # orbits: union-find over the permutations, labelled by the smallest member
    parent = np.arange(num_atoms)

    def find(i: int) -> int:
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for perm in perms:
        for i in range(num_atoms):
            a, b = find(i), find(int(perm[i]))
            if a != b:
                parent[max(a, b)] = min(a, b)
    equivalent = np.array([find(i) for i in range(num_atoms)], dtype=np.int32)
    return {"rotations": np.array([rotations[k] for k in ranks], dtype=np.int32),
            "translations": np.array([translations[k] for k in ranks], dtype=np.float64),
            "equivalent_atoms": equivalent}


def install_spglib_stand_in(force: bool = False) -> bool:
    """Make ``import spglib`` resolve to this module's ``get_symmetry`` when spglib is not installed
    (``force=True``: even when it is).  Returns True if the stand-in is (now) what ``spglib`` names."""
    if not force:
        try:
            import spglib  # noqa: F401  pylint: disable=import-outside-toplevel,unused-import
            return getattr(sys.modules["spglib"], "__ramannoodle_b200_stand_in__", False)
        except ImportError:
            pass
    module = types.ModuleType("spglib")
    module.get_symmetry = get_symmetry
    module.__ramannoodle_b200_stand_in__ = True
    module.__doc__ = "stand-in for spglib.get_symmetry provided by ramannoodle_b200.symmetry"
    sys.modules["spglib"] = module
    return True
