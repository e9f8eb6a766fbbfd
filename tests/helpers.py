"""Shared test helpers (oracle side)."""
import os

import numpy as np

from oracle.numpy_port import OracleModel
from ramannoodle_b200.state import ModelState

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


def oracle_model(state: ModelState) -> OracleModel:
    return OracleModel(
        ref_positions=state.ref_positions, lattice=state.lattice,
        ref_polarizability=state.ref_polarizability,
        basis_vectors=[np.asarray(v) for v in state.basis_vectors],
        splines=list(state.splines), mask=np.asarray(state.mask, dtype=bool))


def state_from_tables(data, prefix: str) -> ModelState:
    """Rebuild a ModelState from the ragged tables stored in a golden npz."""
    basis = data[f"{prefix}_basis"]
    degree = data[f"{prefix}_degree"]
    knot_off = data[f"{prefix}_knot_off"]
    coef_off = data[f"{prefix}_coef_off"]
    knots = data[f"{prefix}_knots"]
    coefs = data[f"{prefix}_coefs"]
    weight = data[f"{prefix}_weight"]
    state = ModelState(data["ref_positions"], data["lattice"], data["ref_polarizability"])
    for j in range(basis.shape[0]):
        state.add_dof(basis[j], knots[knot_off[j]:knot_off[j + 1]],
                      coefs[coef_off[j]:coef_off[j + 1]].reshape(-1, 3, 3), int(degree[j]))
    state.mask = weight == 0.0
    return state


def rel_err(new, ref) -> float:
    """max|new-ref| / max|ref| — the parity metric of BASELINE.json's north_star."""
    new = np.asarray(new)
    ref = np.asarray(ref)
    scale = np.max(np.abs(ref))
    return float(np.max(np.abs(new - ref)) / (scale if scale > 0 else 1.0))


def pointwise_rel_err(new, ref) -> float:
    new = np.asarray(new)
    ref = np.asarray(ref)
    return float(np.max(np.abs(new - ref) / np.abs(ref)))
