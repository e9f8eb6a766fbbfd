"""Same attributes as ramannoodle/pmodel/_interpolation.py:110-115; evaluation by the oracle."""
import copy

import numpy as np
from scipy.interpolate import BSpline

from oracle import numpy_port as ora
from ramannoodle.exceptions import UserError


class ReferenceStructure:
    def __init__(self, atomic_numbers, lattice, positions):
        self.atomic_numbers = list(atomic_numbers)
        self.lattice = np.array(lattice)
        self.positions = np.array(positions)


class InterpolationModel:
    def __init__(self, ref_structure, ref_polarizability, is_dummy_model=False):
        self._ref_structure = ref_structure
        self._ref_polarizability = np.array(ref_polarizability)
        self._is_dummy_model = is_dummy_model
        self._cart_basis_vectors = []
        self._interpolations = []
        self._mask = np.array([], dtype="bool")

    @property
    def mask(self):
        return self._mask.copy()

    @mask.setter
    def mask(self, value):
        self._mask = value

    def add_dof(self, basis_vector, knots, coefs, degree):
        self._cart_basis_vectors.append(np.array(basis_vector))
        self._interpolations.append(BSpline(knots, coefs, degree, extrapolate=True))
        self._mask = np.append(self._mask, False)

    def get_masked_model(self, dof_indexes_to_mask):
        result = copy.deepcopy(self)
        new_mask = result.mask
        new_mask[:] = False
        new_mask[dof_indexes_to_mask] = True
        result.mask = new_mask
        return result

    def calc_polarizabilities(self, positions_batch):
        if len(self._cart_basis_vectors) != len(self._interpolations) and self._is_dummy_model:
            raise UserError("dummy model cannot calculate polarizabilities")
        model = ora.OracleModel(self._ref_structure.positions, self._ref_structure.lattice, self._ref_polarizability,
                                list(self._cart_basis_vectors),
                                [(s.t, s.c, s.k) for s in self._interpolations], np.asarray(self._mask))
        return ora.calc_polarizabilities(model, positions_batch)
