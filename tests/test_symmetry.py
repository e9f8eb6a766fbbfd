"""Next row N4: the space-group search that stands in for ``spglib.get_symmetry`` during model
construction (``ramannoodle/structure/_reference.py:114-122``), on the packaged workload structures and —
when the reference tree is present — through the reference's own ``ReferenceStructure`` / ``ARTModel``
against the pins of ``test/tests/test_structure.py:131-157`` and ``test/tests/test_art.py:18-47,267-321``."""
import os

import numpy as np
import pytest

from oracle.ref_bootstrap import REFERENCE_ROOT, import_reference, reference_available
from ramannoodle_b200 import symmetry, synthetic


def _cell(name):
    s = synthetic.load_structure(name)
    return s["lattice"], s["positions"], s["atomic_numbers"]


@pytest.mark.parametrize("name, operations, nonequivalent", [("TiO2", 288, 2), ("STO", 1, 135), ("LLZO", 32, 9)])
def test_counts_and_group_properties(name, operations, nonequivalent):
    """Operation / orbit counts (``test_structure.py:135-137``: 2, 135 and 9 nonequivalent atoms) and the
    properties any space group has: identity first, closure under composition, every operation maps the
    structure onto itself, orbit labels are the smallest index of the orbit."""
    lattice, positions, numbers = _cell(name)
    sym = symmetry.get_symmetry((lattice, positions, numbers))
    rotations, translations, equivalent = sym["rotations"], sym["translations"], sym["equivalent_atoms"]
    assert rotations.shape == (operations, 3, 3) and translations.shape == (operations, 3)
    assert len(set(equivalent.tolist())) == nonequivalent
    assert np.array_equal(rotations[0], np.eye(3)) and np.allclose(translations[0], 0)
    assert all(equivalent[label] == label for label in set(equivalent.tolist()))
    assert all(equivalent[i] <= i for i in range(len(equivalent)))
    keys = {(tuple(r.reshape(-1)), tuple(np.round(t % 1.0, 6) % 1.0)) for r, t in zip(rotations, translations)}
    assert len(keys) == operations
    rng = np.random.default_rng(3)
    for a, b in rng.integers(0, operations, size=(40, 2)):
        r = rotations[a] @ rotations[b]
        t = (rotations[a] @ translations[b] + translations[a]) % 1.0
        t = np.round(t, 6) % 1.0
        assert (tuple(r.reshape(-1)), tuple(t)) in keys
    for k in rng.integers(0, operations, size=10):
        image = positions @ rotations[k].T + translations[k]
        delta = image[:, None, :] - positions[None, :, :]
        delta -= np.rint(delta)
        dist = np.linalg.norm(delta @ lattice, axis=-1)
        partner = dist.argmin(axis=1)
        assert dist.min(axis=1).max() < 1e-5 and np.array_equal(numbers[partner], numbers)
        assert np.array_equal(equivalent[partner], equivalent)


def test_supercell_and_failure_values():
    """The 1536-atom 2x2x2 LLZO supercell (c5's structure): 32 x 8 operations, still 9 orbits; malformed
    cells give None (what ``ReferenceStructure`` turns into ``SymmetryException``)."""
    sym = symmetry.get_symmetry(_cell("LLZO_2x2x2"))
    assert len(sym["rotations"]) == 256 and len(set(sym["equivalent_atoms"].tolist())) == 9
    lattice, positions, numbers = _cell("TiO2")
    assert symmetry.get_symmetry((lattice[:2], positions, numbers)) is None
    assert symmetry.get_symmetry((lattice, positions[:5], numbers)) is None
    assert symmetry.get_symmetry((np.zeros((3, 3)), positions, numbers)) is None
    # a rattled copy loses everything but the identity
    rattled = positions + np.random.default_rng(0).normal(scale=1e-4, size=positions.shape)
    assert len(symmetry.get_symmetry((lattice, rattled, numbers))["rotations"]) == 1
    # a looser tolerance finds the group again
    assert len(symmetry.get_symmetry((lattice, rattled, numbers), symprec=0.03)["rotations"]) == 288


live = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture()
def reference_with_search(monkeypatch):
    import_reference()
    import ramannoodle.structure._reference as ref_module

    monkeypatch.setattr(ref_module.spglib, "get_symmetry", symmetry.get_symmetry)
    return ref_module


def _outcar_structure(relative):
    from ramannoodle.io.generic import read_ref_structure

    return read_ref_structure(os.path.join(REFERENCE_ROOT, relative), file_format="outcar")


@live
@pytest.mark.parametrize("path, nonequivalent, orthogonal, shape", [
    ("test/data/TiO2/phonons_OUTCAR", 2, 36, [2] * 36),
    ("test/data/STO_RATTLED_OUTCAR", 135, 1, [1]),
    ("test/data/LLZO/LLZO_OUTCAR", 9, 32, [1] * 32),
])
def test_reference_structure_pins(reference_with_search, path, nonequivalent, orthogonal, shape):
    """``test/tests/test_structure.py:131-157`` with the search in place of spglib."""
    structure = _outcar_structure(path)
    assert structure.num_nonequivalent_atoms == nonequivalent
    displacement = structure.positions * 0
    displacement[0, 2] += 0.1
    displacements = structure.get_equivalent_displacements(displacement)
    assert len(displacements) == orthogonal
    assert [len(d["displacements"]) for d in displacements] == shape


@live
def test_reference_art_model_pins(reference_with_search):
    """``test/tests/test_art.py:18-47`` (72 DOFs from one TiO2 displacement, 1 for rattled STO) and
    ``:267-321`` (specification tuples: 2 left for TiO2, atoms 1..35 equivalent to atom 0)."""
    from ramannoodle.pmodel._art import ARTModel

    structure = _outcar_structure("test/data/TiO2/phonons_OUTCAR")
    model = ARTModel(structure, np.zeros((3, 3)))
    model.add_art(0, np.array([1, 0, 0]), np.array([0.01]), np.zeros((1, 3, 3)))
    assert len(model.cart_basis_vectors) == 72
    assert np.isclose(np.linalg.norm(model.cart_basis_vectors[0]), 1)
    tuples = model.get_specification_tuples()
    assert len(tuples) == 2 and tuples[0][0] == 0 and tuples[0][1] == list(range(1, 36))
    structure = _outcar_structure("test/data/STO_RATTLED_OUTCAR")
    model = ARTModel(structure, np.zeros((3, 3)))
    model.add_art(0, np.array([1, 0, 0]), np.array([-0.01, 0.01]), np.zeros((2, 3, 3)))
    assert len(model.cart_basis_vectors) == 1
    assert len(model.get_specification_tuples()) == 135


def test_spglib_stand_in_registration(monkeypatch):
    import sys

    monkeypatch.delitem(sys.modules, "spglib", raising=False)
    assert symmetry.install_spglib_stand_in(force=True)
    import spglib  # pylint: disable=import-outside-toplevel

    assert spglib.get_symmetry is symmetry.get_symmetry
    monkeypatch.delitem(sys.modules, "spglib", raising=False)


@live
def test_full_model_through_reference_construction():
    """tools/build_model_reference.py: a complete 324-DOF TiO2 ARTModel built by the reference's own ``add_art``
    with the space-group search in place of spglib and the vectorised scans of ``construction.py``."""
    import json
    import subprocess
    import sys

    from helpers import REPO

    res = subprocess.run([sys.executable, os.path.join(REPO, "tools", "build_model_reference.py"), "TiO2", "--art"],
                         capture_output=True, text=True, timeout=600, cwd=REPO, check=False)
    assert res.returncode == 0, res.stderr[-2000:]
    report = json.loads(res.stdout.strip().splitlines()[-1])
    assert report["dofs"] == 324 and report["operations"] == 288 and report["nonequivalent_atoms"] == 2
