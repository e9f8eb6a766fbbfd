"""``install()`` / ``accelerate()`` on the GPU through a stand-in of the reference package
(tests/standin/ramannoodle: the reference's module layout, class names and private attributes, every
method evaluated by the CPU oracle — /root/reference does not exist on the GPU box).  The same live,
mutable objects are evaluated unpatched (oracle) and patched (CUDA): 1e-10 on polarizabilities, 1e-8 on
intensities, wavenumbers bit for bit."""
import os
import sys

import numpy as np
import pytest
from scipy.interpolate import BSpline

import ramannoodle_b200 as rb
from ramannoodle_b200 import synthetic

from gpu_helpers import ALPHA_RTOL, INTENSITY_RTOL
from helpers import REPO, pointwise_rel_err, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def standin():
    """Import the stand-in ``ramannoodle`` for one test and remove it again (another ``ramannoodle`` — the
    real reference in the authoring container — must not be mixed with it)."""
    saved = {name: module for name, module in sys.modules.items() if name == "ramannoodle" or name.startswith("ramannoodle.")}
    for name in saved:
        del sys.modules[name]
    path = os.path.join(REPO, "tests", "standin")
    sys.path.insert(0, path)
    try:
        import ramannoodle  # noqa: F401  pylint: disable=import-outside-toplevel,unused-import

        assert ramannoodle.__file__.startswith(path)
        yield ramannoodle
    finally:
        rb.uninstall()
        sys.path.remove(path)
        for name in [n for n in sys.modules if n == "ramannoodle" or n.startswith("ramannoodle.")]:
            del sys.modules[name]
        sys.modules.update(saved)


def _standin_model(structure, kind, art, num_dofs=None):
    from ramannoodle.pmodel._art import ARTModel
    from ramannoodle.pmodel._interpolation import InterpolationModel, ReferenceStructure

    state = synthetic.make_model(structure, kind, num_dofs=num_dofs, seed=321)
    geom = synthetic.load_structure(structure)
    model = (ARTModel if art else InterpolationModel)(
        ReferenceStructure(geom["atomic_numbers"], state.lattice, state.ref_positions), state.ref_polarizability)
    for vector, (knots, coefs, degree) in zip(state.basis_vectors, state.splines):
        model.add_dof(vector, knots, coefs, degree)
    assert isinstance(model._interpolations[0], BSpline)  # pylint: disable=protected-access
    return model


@pytest.mark.parametrize("structure,kind,art,dofs", [("LLZO", "art", True, None), ("STO", "cubic", False, 90)])
def test_installed_path_matches_unpatched_objects(standin, structure, kind, art, dofs):
    from ramannoodle.dynamics._trajectory import Trajectory
    from ramannoodle.spectrum import utils

    model = _standin_model(structure, kind, art, dofs)
    positions = synthetic.make_trajectory(structure, 300, seed=8, lattice_hops=True)
    trajectory = Trajectory(positions, 2.0)
    want_alpha = model.calc_polarizabilities(trajectory._positions_ts)  # pylint: disable=protected-access
    want_spectrum = trajectory.get_raman_spectrum(model)
    want_wn, want_int = want_spectrum.measure(laser_correction=True, bose_einstein_correction=True)
    want_smooth = utils.convolve_spectrum(want_wn, want_int, "lorentzian", 6.0)

    patched = rb.install(pin_trajectories=True)
    assert len(patched) == 4
    launches = rb._lib.launch_count()  # pylint: disable=protected-access
    alpha = model.calc_polarizabilities(trajectory._positions_ts)  # pylint: disable=protected-access
    assert isinstance(alpha, np.ndarray) and rel_err(alpha, want_alpha) <= ALPHA_RTOL
    spectrum = trajectory.get_raman_spectrum(model)
    assert type(spectrum) is type(want_spectrum)  # still the reference's own class
    assert rel_err(spectrum.polarizability_ts, want_alpha) <= ALPHA_RTOL
    wn, inten = spectrum.measure(laser_correction=True, bose_einstein_correction=True)
    assert np.array_equal(wn, want_wn) and pointwise_rel_err(inten, want_int) <= INTENSITY_RTOL
    smooth = utils.convolve_spectrum(wn, inten, "lorentzian", 6.0)
    assert np.array_equal(smooth[0], want_smooth[0]) and rel_err(smooth[1], want_smooth[1]) <= 1e-10
    assert rb._lib.launch_count() > launches  # pylint: disable=protected-access  (the CUDA path ran)
    second = trajectory.get_raman_spectrum(model)  # the page-locked trajectory is reused
    assert np.array_equal(second.polarizability_ts, spectrum.polarizability_ts)
    with pytest.raises(ValueError, match="polarizability_model and trajectory are incompatible"):
        Trajectory(np.zeros((3, 5, 3)), 1.0).get_raman_spectrum(model)

    # mutation after the first (cached) evaluation: mask setter, in-place edit, unmask, deep copies, new DOF
    rb.uninstall()
    variants = []
    mask = model.mask
    mask[[0, 3, 7]] = True
    model.mask = mask
    variants.append(model.calc_polarizabilities(positions))
    model._mask[1] = True  # pylint: disable=protected-access
    variants.append(model.calc_polarizabilities(positions))
    clone = model.get_masked_model([2, 5])
    variants.append(clone.calc_polarizabilities(positions))
    model.mask = np.zeros_like(mask)
    spline = model._interpolations[4]  # pylint: disable=protected-access
    model.add_dof(model._cart_basis_vectors[4][::-1].copy(), spline.t, 2.0 * spline.c, spline.k)  # pylint: disable=protected-access
    variants.append(model.calc_polarizabilities(positions))

    rb.install()
    model.mask = np.zeros(len(model._cart_basis_vectors) - 1, dtype=bool)  # pylint: disable=protected-access
    del model._cart_basis_vectors[-1], model._interpolations[-1]  # pylint: disable=protected-access
    assert rel_err(model.calc_polarizabilities(positions), want_alpha_unwrapped(model, positions)) <= ALPHA_RTOL
    mask = model.mask
    mask[[0, 3, 7]] = True
    model.mask = mask
    assert rel_err(model.calc_polarizabilities(positions), variants[0]) <= ALPHA_RTOL
    model._mask[1] = True  # pylint: disable=protected-access
    assert rel_err(model.calc_polarizabilities(positions), variants[1]) <= ALPHA_RTOL
    clone = model.get_masked_model([2, 5])
    assert rel_err(clone.calc_polarizabilities(positions), variants[2]) <= ALPHA_RTOL
    model.mask = np.zeros_like(mask)
    spline = model._interpolations[4]  # pylint: disable=protected-access
    model.add_dof(model._cart_basis_vectors[4][::-1].copy(), spline.t, 2.0 * spline.c, spline.k)  # pylint: disable=protected-access
    assert rel_err(model.calc_polarizabilities(positions), variants[3]) <= ALPHA_RTOL


def want_alpha_unwrapped(model, positions):
    """The unpatched evaluation of a stand-in model (install() must be active: goes through the originals)."""
    from ramannoodle_b200 import dropin

    original = dropin._ORIGINALS[("ramannoodle.pmodel._interpolation", "InterpolationModel", "calc_polarizabilities")]  # pylint: disable=protected-access
    return original(model, positions)


def test_accelerate_snapshot_and_dummy_model(standin):
    from ramannoodle.exceptions import UserError
    from ramannoodle.pmodel._art import ARTModel
    from ramannoodle.pmodel._interpolation import ReferenceStructure

    model = _standin_model("TiO2", "art", True)
    positions = synthetic.make_trajectory("TiO2", 64, seed=2)
    want = model.calc_polarizabilities(positions)
    wrapped = rb.accelerate(model)
    assert type(wrapped) is rb.ARTModel
    assert rel_err(wrapped.calc_polarizabilities(positions), want) <= ALPHA_RTOL
    # accelerate() snapshots: a later mask change on the reference object is NOT seen (documented) ...
    mask = model.mask
    mask[:50] = True
    model.mask = mask
    assert rel_err(wrapped.calc_polarizabilities(positions), want) <= ALPHA_RTOL
    # ... wrapping again (or install()) is
    assert rel_err(rb.accelerate(model).calc_polarizabilities(positions), model.calc_polarizabilities(positions)) <= ALPHA_RTOL
    # symbols travel with the snapshot
    titanium = wrapped.get_dof_indexes("Ti")
    assert len(titanium) > 0 and all(wrapped.state.atomic_numbers[j // 3] == 22 for j in titanium)
    # dummy models raise the REFERENCE's UserError through the patched entry
    geom = synthetic.load_structure("TiO2")
    dummy = ARTModel(ReferenceStructure(geom["atomic_numbers"], geom["lattice"], geom["positions"]), np.zeros((3, 3)),
                     is_dummy_model=True)
    dummy._cart_basis_vectors.append(np.zeros((108, 3)))  # pylint: disable=protected-access
    rb.install()
    with pytest.raises(UserError, match="dummy model cannot calculate polarizabilities"):
        dummy.calc_polarizabilities(positions)
