"""Host-side container for the state an ``InterpolationModel``/``ARTModel`` evaluates.

Model *construction* (symmetry expansion, spline fitting, file parsing) stays in the
reference; this class only carries what ``calc_polarizabilities`` reads
(``ramannoodle/pmodel/_interpolation.py:110-115``): the reference structure's fractional
positions and lattice, the reference polarizability, J Cartesian basis vectors, J vector
valued B-splines (knots ``t``, coefficients ``c`` of shape (n,3,3), degree ``k``) and the
boolean mask.  ``tables()`` flattens it into the ragged arrays the C-ABI
(``include/ramannoodle_b200.h: rn_model_create``) takes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .exceptions import get_type_error, verify_ndarray_shape


@dataclass
class ModelState:
    """Plain arrays; no device state (that lives in ``pmodel._DeviceModel``)."""

    ref_positions: np.ndarray  # (N,3) fractional
    lattice: np.ndarray  # (3,3) rows are lattice vectors (Å)
    ref_polarizability: np.ndarray  # (3,3)
    basis_vectors: list = field(default_factory=list)  # J x (N,3)
    splines: list = field(default_factory=list)  # J x (t, c(n,3,3), k)
    mask: np.ndarray = field(default_factory=lambda: np.array([], dtype=bool))
    is_dummy_model: bool = False
    atomic_numbers: list | None = None  # (N,) — only ARTModel.get_dof_indexes(symbol) needs them

    def __post_init__(self) -> None:
        self.ref_positions = np.ascontiguousarray(self.ref_positions, dtype=np.float64)
        self.lattice = np.ascontiguousarray(self.lattice, dtype=np.float64)
        self.ref_polarizability = np.ascontiguousarray(self.ref_polarizability, dtype=np.float64)
        verify_ndarray_shape("ref_positions", self.ref_positions, (None, 3))
        verify_ndarray_shape("lattice", self.lattice, (3, 3))
        verify_ndarray_shape("ref_polarizability", self.ref_polarizability, (3, 3))
        self.mask = np.asarray(self.mask, dtype=bool)

    @property
    def num_atoms(self) -> int:
        return int(self.ref_positions.shape[0])

    @property
    def num_dofs(self) -> int:
        return len(self.basis_vectors)

    @classmethod
    def from_reference(cls, model) -> "ModelState":
        """Snapshot a reference ``InterpolationModel``/``ARTModel`` (duck-typed).

        Reads exactly the attributes ``calc_polarizabilities`` uses
        (``_interpolation.py:217-252``); nothing is recomputed.
        """
        structure = model._ref_structure  # pylint: disable=protected-access
        return cls(
            ref_positions=np.array(structure.positions, dtype=np.float64),
            lattice=np.array(structure.lattice, dtype=np.float64),
            ref_polarizability=np.array(model._ref_polarizability, dtype=np.float64),
            basis_vectors=[np.array(v, dtype=np.float64) for v in model._cart_basis_vectors],
            splines=[(np.array(s.t, dtype=np.float64), np.array(s.c, dtype=np.float64), int(s.k))
                     for s in model._interpolations],
            mask=np.array(model._mask, dtype=bool),
            is_dummy_model=bool(getattr(model, "_is_dummy_model", False)),
            atomic_numbers=[int(z) for z in getattr(structure, "atomic_numbers", [])] or None,
        )

    def add_dof(self, basis_vector, t, c, k) -> None:
        """Append one DOF exactly as ``_construct_and_add_interpolations`` leaves it
        (``_interpolation.py:403-407``)."""
        basis_vector = np.asarray(basis_vector, dtype=np.float64).reshape(self.num_atoms, 3)
        c = np.asarray(c, dtype=np.float64)
        verify_ndarray_shape("c", c, (None, 3, 3))
        t = np.asarray(t, dtype=np.float64)
        if t.shape != (c.shape[0] + int(k) + 1,):
            raise ValueError("knots/coefficients/degree are inconsistent")
        self.basis_vectors.append(basis_vector)
        self.splines.append((t, c, int(k)))
        self.mask = np.append(self.mask, False)

    def tables(self) -> dict:
        """Ragged C tables: basis (J,3N), degree (J,), knot_off/coef_off (J+1,), knots, coefs (sum n,9),
        weight = 1 - mask (``_interpolation.py:242``)."""
        num_dofs = self.num_dofs
        if len(self.splines) != num_dofs or self.mask.shape != (num_dofs,):
            raise ValueError("basis vectors, interpolations and mask have different lengths")
        dim = 3 * self.num_atoms
        basis = np.empty((num_dofs, dim), dtype=np.float64)
        for j, vector in enumerate(self.basis_vectors):
            basis[j] = np.asarray(vector, dtype=np.float64).reshape(dim)
        degree = np.array([s[2] for s in self.splines], dtype=np.int32)
        knot_off = np.zeros(num_dofs + 1, dtype=np.int64)
        coef_off = np.zeros(num_dofs + 1, dtype=np.int64)
        for j, (t, c, _) in enumerate(self.splines):
            knot_off[j + 1] = knot_off[j] + len(t)
            coef_off[j + 1] = coef_off[j] + c.shape[0]
        if num_dofs:
            knots = np.ascontiguousarray(np.concatenate([s[0] for s in self.splines]), dtype=np.float64)
            coefs = np.ascontiguousarray(
                np.concatenate([np.asarray(s[1]).reshape(-1, 9) for s in self.splines]), dtype=np.float64)
        else:
            knots = np.zeros(0)
            coefs = np.zeros((0, 9))
        weight = np.ascontiguousarray(1.0 - self.mask.astype(np.float64))
        return {"basis": basis, "degree": degree, "knot_off": knot_off, "knots": knots,
                "coef_off": coef_off, "coefs": coefs, "weight": weight}

    def fingerprint(self) -> tuple:
        """Cheap change detector for the mutable parts (mask edits, added DOFs)."""
        return (self.num_dofs, len(self.splines), self.mask.tobytes())

    def get_atom_indexes(self, atom_symbols) -> list:
        """``ReferenceStructure.get_atom_indexes`` (``ramannoodle/structure/_reference.py:344-364``)."""
        if self.atomic_numbers is None:
            raise ValueError("this model state carries no atomic numbers (pass atomic_numbers=...)")
        symbols = [ATOM_SYMBOLS[number] for number in self.atomic_numbers]
        if isinstance(atom_symbols, str):
            atom_symbols = [atom_symbols]
        try:
            return [index for index, symbol in enumerate(symbols) if symbol in atom_symbols]
        except TypeError as err:
            raise get_type_error("atom_symbols", atom_symbols, "list") from err


# element symbols by atomic number (the periodic table; ramannoodle/constants.py:125-246 holds the same map)
ATOM_SYMBOLS = dict(enumerate(
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
    "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
    "Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl "
    "Mc Lv Ts Og".split(), start=1))
