"""Next row N4 (host-side part): the vectorised basis-vector scans against the reference's loops
(``ramannoodle/structure/_symmetry_utils.py:42-133``).  Runs without a GPU; the comparison with the
live reference is skipped where ``/root/reference`` is absent."""
import time

import numpy as np
import pytest

from oracle.ref_bootstrap import import_reference, reference_available
from ramannoodle_b200 import construction


def _loop_orthogonal(vector_1, vectors):
    """The reference rule, restated (``_symmetry_utils.py:61-75``)."""
    vector_1 = vector_1 / float(np.linalg.norm(vector_1))
    for index, vector_2 in enumerate(vectors):
        vector_2 = vector_2 / np.linalg.norm(vector_2)
        if not np.allclose(np.dot(vector_1.flatten(), vector_2.flatten()) + 1, 1):
            return index
    return -1


def _basis(rng, count, atoms):
    """Orthogonal (N,3) displacement patterns, as a model under construction holds them."""
    q, _ = np.linalg.qr(rng.normal(size=(3 * atoms, count)))
    return [0.1 * (j + 1) * q[:, j].reshape(atoms, 3) for j in range(count)]


def test_orthogonal_scan_on_a_growing_list():
    rng = np.random.default_rng(0)
    vectors = []
    pool = _basis(rng, 60, 40)
    for j, vector in enumerate(pool):
        assert construction.is_orthogonal_to_all(vector, vectors) == -1 == _loop_orthogonal(vector, vectors)
        vectors.append(vector)  # the list add_dof keeps appending to: rows are cached, not rebuilt
        if j % 7 == 3:
            tilted = pool[min(j + 1, 59)] + 1e-3 * pool[j // 2]
            assert construction.is_orthogonal_to_all(tilted, vectors) == _loop_orthogonal(tilted, vectors) == j // 2
    # a different list with the same length, and a shrunken one, must not hit the stale cache
    other = _basis(rng, 60, 40)
    probe = other[5] + other[17]
    assert construction.is_orthogonal_to_all(probe, other) == 5
    del vectors[30:]
    assert construction.is_orthogonal_to_all(pool[45], vectors) == -1
    assert construction.is_orthogonal_to_all(pool[12].flatten(), [v.flatten() for v in vectors]) == 12
    assert construction.is_orthogonal_to_all(pool[0], []) == -1


def test_threshold_matches_allclose():
    unit = np.array([1.0, 0.0, 0.0])
    for eps in (0.0, 5e-6, 1.0009e-5, 1.0011e-5, 2e-5, -1.0009e-5, -1.0011e-5):
        other = np.array([eps, np.sqrt(1 - eps * eps), 0.0])
        assert construction.is_orthogonal_to_all(unit, [other]) == _loop_orthogonal(unit, [other])
    assert construction.is_orthogonal_to_all(unit, [np.array([np.nan, 1.0, 0.0])]) == 0


def test_collinear_scans():
    rng = np.random.default_rng(1)
    base = rng.normal(size=(12, 3))
    same = [2.5 * base, -0.3 * base, base * (1 + 1e-12)]
    assert construction.is_collinear_with_all(base, same) == -1
    assert construction.is_non_collinear_with_all(base, same) == 0
    mixed = [rng.normal(size=(12, 3)), -base, rng.normal(size=(12, 3))]
    assert construction.is_collinear_with_all(base, mixed) == 0
    assert construction.is_non_collinear_with_all(base, mixed) == 1
    assert construction.is_collinear_with_all(base, []) == -1
    assert construction.is_non_collinear_with_all(base, [mixed[0], mixed[2]]) == -1
    with pytest.raises(TypeError, match="vector_1 should have type ndarray"):
        construction.is_collinear_with_all("x", same)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_against_live_reference_and_patching():
    import_reference()
    from ramannoodle.pmodel import _interpolation
    from ramannoodle.structure import _symmetry_utils as ref

    rng = np.random.default_rng(2)
    pool = _basis(rng, 150, 64)
    vectors = pool[:120]
    probes = [pool[130], pool[7] * 3, pool[140] + 1e-6 * pool[33], pool[140] + 1e-4 * pool[33],
              rng.normal(size=(64, 3))]
    original = (ref.is_orthogonal_to_all, ref.is_collinear_with_all, ref.is_non_collinear_with_all)
    for probe in probes:
        assert construction.is_orthogonal_to_all(probe, vectors) == original[0](probe, vectors)
        assert construction.is_collinear_with_all(probe, vectors[:9]) == original[1](probe, vectors[:9])
        assert construction.is_non_collinear_with_all(probe, vectors) == original[2](probe, vectors)
    t0 = time.perf_counter()
    for probe in probes * 4:
        original[0](probe, vectors)
    slow = time.perf_counter() - t0
    t0 = time.perf_counter()
    for probe in probes * 4:
        construction.is_orthogonal_to_all(probe, vectors)
    fast = time.perf_counter() - t0
    assert fast < slow
    try:
        patched = construction.accelerate_construction()
        assert "ramannoodle.pmodel._interpolation.is_orthogonal_to_all" in patched
        assert _interpolation.is_orthogonal_to_all is construction.is_orthogonal_to_all
        assert ref.is_non_collinear_with_all is construction.is_non_collinear_with_all
    finally:
        construction.restore_construction()
    assert ref.is_orthogonal_to_all is original[0] and _interpolation.is_collinear_with_all is original[1]
