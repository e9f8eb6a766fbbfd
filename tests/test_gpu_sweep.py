"""Next row N3: mask sweeps — several ``get_masked_model`` copies
(``ramannoodle/pmodel/_interpolation.py:697-708``) evaluated in one pass over a trajectory —
against the oracle's per-copy ``calc_polarizabilities`` and against this package's single-model
path.  Tolerance: 1e-10 relative on polarizabilities (north_star), 1e-8 on intensities."""
import copy
import ctypes

import numpy as np
import pytest

import ramannoodle_b200 as rb
from oracle import numpy_port as ora
from ramannoodle_b200 import _lib, synthetic

from gpu_helpers import ALPHA_RTOL, INTENSITY_RTOL, to_cuda
from helpers import oracle_model, pointwise_rel_err, rel_err

pytestmark = pytest.mark.gpu


def _masks(num_dofs, count, seed):
    rng = np.random.default_rng(seed)
    masks = np.zeros((count, num_dofs), dtype=bool)
    for g in range(count):
        if g % 3 == 0:  # a block of DOFs (one species in the masking tutorial)
            lo = rng.integers(0, num_dofs // 2)
            masks[g, lo:lo + num_dofs // 3] = True
        elif g % 3 == 1:  # everything except a block
            masks[g] = True
            lo = rng.integers(0, num_dofs // 2)
            masks[g, lo:lo + num_dofs // 4] = False
        else:
            masks[g] = rng.random(num_dofs) < 0.3
    return masks


def _oracle_sweep(state, masks, positions):
    out = []
    for row in masks:
        state_g = copy.deepcopy(state)
        state_g.mask = row.copy()
        out.append(ora.calc_polarizabilities(oracle_model(state_g), positions))
    return np.stack(out)


@pytest.fixture(autouse=True)
def _fuse_pairs_too():
    """The library fuses runs of three or four models by default (pairs are faster one by one);
    the tests also drive the two-mask instantiation."""
    hook = _lib.lib().rn_debug_set_sweep_min_run
    hook.argtypes = [ctypes.c_int]
    hook.restype = None
    hook(2)
    yield
    hook(3)


@pytest.mark.parametrize("structure,count,frames", [("LLZO", 4, 2051), ("LLZO", 2, 777), ("LLZO", 3, 1024),
                                                     ("LLZO", 5, 1531), ("LLZO", 9, 264), ("TiO2", 4, 1999),
                                                     ("STO", 3, 515)])
def test_art_sweep_matches_oracle(structure, count, frames):
    state = synthetic.make_model(structure, "art")
    model = rb.ARTModel(state)
    masks = _masks(state.num_dofs, count, seed=count * 31 + frames)
    positions = synthetic.make_trajectory(structure, frames, seed=99)
    want = _oracle_sweep(state, masks, positions)
    got_host = model.calc_polarizabilities_masked(positions, masks)
    assert got_host.shape == (count, frames, 3, 3)
    got_dev = model.calc_polarizabilities_masked(to_cuda(positions), masks).cpu().numpy()
    for g in range(count):
        assert rel_err(got_host[g], want[g]) <= ALPHA_RTOL
        assert rel_err(got_dev[g], want[g]) <= ALPHA_RTOL
        single = model.get_masked_model(np.flatnonzero(masks[g])).calc_polarizabilities(positions)
        assert rel_err(got_host[g], single) <= 1e-13


def test_art_sweep_unwrapped_positions():
    """Positions outside [0,1) (lattice hops, large shifts) take the general wrap formula; exact
    minimum-image ties (ill-conditioned against the reference, whose answer there depends on the
    rounding of ``p // 1``) must at least resolve like this package's single-model kernel."""
    state = synthetic.make_model("LLZO", "art")
    model = rb.ARTModel(state)
    masks = _masks(state.num_dofs, 4, seed=2)
    positions = synthetic.make_trajectory("LLZO", 300, seed=3, lattice_hops=True)
    ref = np.asarray(state.ref_positions)
    positions[7, :, 2] += 2.0
    positions[8, 3, :] -= 7.0
    positions[9] += 1.0
    want = _oracle_sweep(state, masks, positions)
    got = model.calc_polarizabilities_masked(to_cuda(positions), masks).cpu().numpy()
    for g in range(4):
        assert rel_err(got[g], want[g]) <= ALPHA_RTOL
    ties = np.repeat(ref[None], 16, axis=0)
    ties[1, :, 0] += 0.5
    ties[2, :, 1] -= 0.5
    ties[3, ::2, 2] += 0.5
    ties[4, 1::2, :] -= 0.5
    ties[5] += 1.5
    got = model.calc_polarizabilities_masked(to_cuda(ties), masks).cpu().numpy()
    for g in range(4):
        single = model.get_masked_model(np.flatnonzero(masks[g])).calc_polarizabilities(to_cuda(ties)).cpu().numpy()
        assert rel_err(got[g], single) <= 1e-13


def test_sweep_fallback_is_the_single_model_path():
    """With the fused kernel switched off the sweep is bit-identical to one call per model."""
    state = synthetic.make_model("LLZO", "art")
    model = rb.ARTModel(state)
    masks = _masks(state.num_dofs, 4, seed=3)
    positions = to_cuda(synthetic.make_trajectory("LLZO", 600, seed=5))
    hook = _lib.lib().rn_debug_set_sweep_fused
    hook.argtypes = [ctypes.c_int]
    hook.restype = None
    hook(0)
    try:
        got = model.calc_polarizabilities_masked(positions, masks).cpu().numpy()
    finally:
        hook(1)
    for g in range(4):
        single = model.get_masked_model(np.flatnonzero(masks[g])).calc_polarizabilities(positions).cpu().numpy()
        assert np.array_equal(got[g], single)


@pytest.mark.parametrize("structure,kind,num_dofs,count,frames", [
    ("TiO2", "cubic", 60, 3, 700), ("TiO2", "mixed", 60, 4, 700), ("STO", "cubic", None, 4, 1300),
    ("STO", "cubic", None, 2, 129), ("LLZO", "quadratic", 200, 5, 2500), ("TiO2", "mixed", None, 9, 300),
    ("LLZO", "linear5", 130, 4, 20_000)])
def test_spline_models_sweep(structure, kind, num_dofs, count, frames):
    """Masked copies of a spline model share the projection onto the basis (dense_kernel_tp with 2 or 4
    masks per launch, the epilogue once per mask); "mixed" models also carry linear DOFs, whose
    affine part is written first and accumulated onto.  Odd atom counts (STO), the unit-balanced
    schedule (short trajectories) and whole-tile schedule (20k frames) are covered."""
    state = synthetic.make_model(structure, kind, num_dofs=num_dofs, seed=5)
    model = rb.InterpolationModel(state)
    masks = _masks(state.num_dofs, count, seed=17 + count)
    positions = synthetic.make_trajectory(structure, frames, seed=11)
    sel = np.r_[0:min(frames, 200), max(0, frames - 200):frames]
    want = _oracle_sweep(state, masks, positions[sel])
    got_dev = model.calc_polarizabilities_masked(to_cuda(positions), masks).cpu().numpy()
    got_host = model.calc_polarizabilities_masked(positions, masks)
    for g in range(count):
        assert rel_err(got_dev[g][sel], want[g]) <= ALPHA_RTOL
        assert rel_err(got_host[g][sel], want[g]) <= ALPHA_RTOL
        single = model.get_masked_model(np.flatnonzero(masks[g])).calc_polarizabilities(positions)
        assert rel_err(got_host[g], single) <= 1e-12


def test_sweep_of_different_models_and_chunked_host_stream():
    """The free function takes arbitrary models of one structure; small chunks exercise the
    double-buffered host pipeline."""
    art = synthetic.make_model("TiO2", "art")
    other = synthetic.make_model("TiO2", "art", seed=77)
    cubic = synthetic.make_model("TiO2", "cubic", num_dofs=40)
    models = [rb.ARTModel(art), rb.ARTModel(other), rb.InterpolationModel(cubic),
              rb.ARTModel(art).get_masked_model([0, 1, 2, 50])]
    positions = synthetic.make_trajectory("TiO2", 1234, seed=21)
    natives = [m._native_model() for m in models]  # pylint: disable=protected-access
    import torch

    alpha = torch.empty((4, 1234, 3, 3), dtype=torch.float64, device="cuda:0")
    handles = (ctypes.c_void_p * 4)(*[n.handle for n in natives])
    outs = (ctypes.c_void_p * 4)(*[ctypes.c_void_p(alpha[g].data_ptr()) for g in range(4)])
    status = _lib.lib().rn_calc_polarizabilities_host_sweep(handles, 4, ctypes.c_void_p(positions.ctypes.data), 1234,
                                                            outs, 100)
    _lib.check(status, "rn_calc_polarizabilities_host_sweep")
    got = alpha.cpu().numpy()
    via_api = rb.calc_polarizabilities_sweep(models, positions)
    for g, model in enumerate(models):
        want = ora.calc_polarizabilities(oracle_model(model.state), positions)
        assert rel_err(got[g], want) <= ALPHA_RTOL
        assert rel_err(via_api[g], want) <= ALPHA_RTOL


def test_trajectory_get_raman_spectra():
    state = synthetic.make_model("LLZO", "art")
    model = rb.ARTModel(state)
    masks = _masks(state.num_dofs, 3, seed=8)
    copies = [model.get_masked_model(np.flatnonzero(row)) for row in masks]
    positions = synthetic.make_trajectory("LLZO", 3000, seed=14)
    for resident in (False, True):
        trajectory = rb.Trajectory(to_cuda(positions) if resident else positions, 2.0)
        spectra = trajectory.get_raman_spectra(copies)
        assert len(spectra) == 3
        for g, spectrum in enumerate(spectra):
            alpha = ora.calc_polarizabilities(oracle_model(copies[g].state), ora.trajectory_positions(positions))
            wn_ref, inten_ref = ora.md_measure(alpha, 2.0)
            wn, inten = spectrum.measure()
            assert np.array_equal(wn, wn_ref)
            assert pointwise_rel_err(inten, inten_ref) <= INTENSITY_RTOL


def test_sweep_error_behaviour():
    model = rb.ARTModel(synthetic.make_model("TiO2", "art", num_dofs=6))
    other = rb.ARTModel(synthetic.make_model("STO", "art", num_dofs=6))
    good = synthetic.make_trajectory("TiO2", 4)
    with pytest.raises(ValueError, match="same structure"):
        rb.calc_polarizabilities_sweep([model, other], good)
    with pytest.raises(ValueError, match="positions has wrong shape"):
        rb.calc_polarizabilities_sweep([model, model], synthetic.make_trajectory("STO", 4))
    with pytest.raises(TypeError, match="positions should have type ndarray"):
        rb.calc_polarizabilities_sweep([model], [[1.0]])
    with pytest.raises(ValueError, match="at least one model"):
        rb.calc_polarizabilities_sweep([], good)
    with pytest.raises(ValueError, match="masks has wrong shape"):
        model.calc_polarizabilities_masked(good, np.zeros((2, 5), dtype=bool))
    with pytest.raises(ValueError, match="incompatible"):
        rb.Trajectory(synthetic.make_trajectory("STO", 4), 1.0).get_raman_spectra([model, model])
    empty = model.calc_polarizabilities_masked(good[:0], np.zeros((2, 6), dtype=bool))
    assert empty.shape == (2, 0, 3, 3)
