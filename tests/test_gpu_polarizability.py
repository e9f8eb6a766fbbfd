"""Parity of the CUDA polarizability kernels (through the C-ABI) against the CPU oracle and
the golden vectors generated from the unmodified reference.  Tolerance: max|new-ref| /
max|ref| <= 1e-10 (BASELINE.json north_star)."""
import copy

import ctypes

import numpy as np
import pytest
import torch

import ramannoodle_b200 as rb
from oracle import numpy_port as ora
from oracle.make_golden import SYNTHETIC_CASES
from ramannoodle_b200 import _lib, synthetic

from gpu_helpers import ALPHA_RTOL, to_cuda
from helpers import GOLDEN, oracle_model, rel_err, state_from_tables

pytestmark = pytest.mark.gpu


def _check(model, positions, want, tol=ALPHA_RTOL):
    got_host = model.calc_polarizabilities(positions)
    assert got_host.shape == want.shape and got_host.dtype == np.float64
    assert rel_err(got_host, want) <= tol
    got_dev = model.calc_polarizabilities(to_cuda(positions))
    assert got_dev.is_cuda
    assert rel_err(got_dev.cpu().numpy(), want) <= tol
    return got_host


@pytest.mark.parametrize("prefix", ["k1", "k2", "k3", "art"])
def test_real_tio2_golden(prefix):
    """25 real DFT geometries (``test/tests/test_phonon_spectrum.py:33-45``) through the GPU."""
    with np.load(f"{GOLDEN}/real_tio2.npz") as data:
        state = state_from_tables(data, prefix)
        model = rb.InterpolationModel(state)
        got = _check(model, data["positions"], data[f"{prefix}_alpha"])
        if prefix != "art":
            assert np.allclose(got, data["known_polarizabilities"], atol=1e-4)
        # S=1 batches, as Phonons.get_raman_spectrum issues them (dynamics/_phonon.py:96-101)
        one = model.calc_polarizabilities(data["positions"][3:4])
        assert rel_err(one, data[f"{prefix}_alpha"][3:4]) <= ALPHA_RTOL
        if prefix == "art":
            model.mask = data["art_masked_mask"]
            _check(model, data["positions"], data["art_masked_alpha"])


@pytest.mark.parametrize("case", SYNTHETIC_CASES, ids=[c[0] for c in SYNTHETIC_CASES])
def test_synthetic_golden(case):
    name, structure, kind, num_dofs, noisy, masked, frames, hops, art = case
    state = synthetic.make_model(structure, kind, num_dofs=num_dofs, noisy_basis=noisy, masked_fraction=masked)
    positions = synthetic.make_trajectory(structure, frames, timestep=1.0, seed=4242, lattice_hops=hops)
    model = (rb.ARTModel if art else rb.InterpolationModel)(state)
    with np.load(f"{GOLDEN}/synthetic_cases.npz") as data:
        _check(model, positions, data[f"{name}_alpha"])
    info = model.path_info()
    if kind == "art":
        assert info["affine_dofs"] == state.num_dofs and info["dense_dofs"] == 0
    elif kind in ("cubic", "quadratic"):
        assert info["dense_dofs"] == state.num_dofs


@pytest.mark.parametrize("frames", [1, 7, 8, 9, 63, 129, 1000])
@pytest.mark.parametrize("structure,kind", [("LLZO", "art"), ("TiO2", "art"), ("STO", "art"), ("STO", "cubic"),
                                            ("LLZO", "mixed")])
def test_against_oracle_ragged_sizes(structure, kind, frames):
    """Tile tails (S not a multiple of 8 / 128), odd atom counts (STO: 135), mixed degrees."""
    num_dofs = None if kind == "art" else 140
    state = synthetic.make_model(structure, kind, num_dofs=num_dofs, masked_fraction=0.1, seed=7)
    positions = synthetic.make_trajectory(structure, frames, seed=frames, lattice_hops=(frames % 2 == 1))
    want = ora.calc_polarizabilities(oracle_model(state), positions)
    _check(rb.InterpolationModel(state), positions, want)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("structure,kind,frames", [("STO", "cubic", 1000), ("TiO2", "mixed", 2500),
                                                    ("LLZO", "quadratic", 19_100), ("STO", "cubic", 127)])
def test_dense_schedules_agree(structure, kind, frames, mode):
    """The dense kernel's three schedules — whole frame tiles per CTA (0), automatic (1), always
    balanced (frame tile, DOF tile) units with atomically finished shared tiles (2) — against the
    oracle, for a pure spline model and for one whose linear DOFs ran through the affine kernel."""
    hook = _lib.lib().rn_debug_set_dense_split
    hook.argtypes = [ctypes.c_int]
    hook.restype = None
    state = synthetic.make_model(structure, kind, masked_fraction=0.05, seed=3)
    positions = synthetic.make_trajectory(structure, frames, seed=frames)
    sel = np.r_[0:min(frames, 300), max(0, frames - 300):frames]
    want = ora.calc_polarizabilities(oracle_model(state), positions[sel])
    hook(mode)
    try:
        got = rb.InterpolationModel(state).calc_polarizabilities(to_cuda(positions)).cpu().numpy()
    finally:
        hook(1)
    assert rel_err(got[sel], want) <= ALPHA_RTOL


def test_affine_kernels_agree_and_force_dense():
    """TMA affine kernel == generic affine kernel == dense DMMA path on an ARTModel."""
    state = synthetic.make_model("LLZO", "art", masked_fraction=0.1)
    positions = synthetic.make_trajectory("LLZO", 300, seed=11)
    want = ora.calc_polarizabilities(oracle_model(state), positions)
    d_pos = to_cuda(positions)
    model = rb.ARTModel(state)
    assert model.path_info()["tma_affine"]
    tma = model.calc_polarizabilities(d_pos).cpu().numpy()
    _lib.lib().rn_debug_force_generic_affine(1)
    try:
        generic = model.calc_polarizabilities(d_pos).cpu().numpy()
    finally:
        _lib.lib().rn_debug_force_generic_affine(0)
    dense = rb.ARTModel(state, force_dense=True)
    assert dense.path_info()["dense_dofs"] == state.num_dofs
    forced = dense.calc_polarizabilities(d_pos).cpu().numpy()
    for got in (tma, generic, forced):
        assert rel_err(got, want) <= ALPHA_RTOL
    # unaligned device pointer (view starting 8 bytes into an allocation) -> generic fallback
    flat = torch.empty(d_pos.numel() + 1, dtype=torch.float64, device=d_pos.device)
    flat[1:] = d_pos.reshape(-1)
    shifted = flat[1:].view(d_pos.shape)
    assert shifted.data_ptr() % 16 == 8
    assert rel_err(model.calc_polarizabilities(shifted).cpu().numpy(), want) <= ALPHA_RTOL


@pytest.mark.parametrize("structure,kind", [("LLZO", "art"), ("STO", "cubic"), ("TiO2", "mixed")])
def test_get_polarizability_from_cart_displacements(structure, kind):
    """The north_star entry ``get_polarizability(cart_displacements)`` (= _interpolation.py:233-252)."""
    state = synthetic.make_model(structure, kind, num_dofs=None if kind == "art" else 90, seed=3)
    positions = synthetic.make_trajectory(structure, 77, seed=5)
    omodel = oracle_model(state)
    cart = ora.calc_cart_displacements(omodel, positions)
    want = ora.get_polarizability(omodel, cart)
    model = rb.InterpolationModel(state)
    got = model.get_polarizability(cart)
    assert rel_err(got, want) <= ALPHA_RTOL
    got3 = model.get_polarizability(to_cuda(cart.reshape(77, -1, 3))).cpu().numpy()
    assert rel_err(got3, want) <= ALPHA_RTOL


def test_minimum_image_edge_cases():
    """Ties of the wrap (d % 1 == 0.5 stays +0.5), exact zeros, tiny negatives, far images."""
    state = synthetic.make_model("TiO2", "mixed", num_dofs=60, seed=9)
    state.ref_positions = np.round(state.ref_positions * 8) / 8  # exactly representable
    base = np.broadcast_to(state.ref_positions, (12,) + state.ref_positions.shape).copy()
    offsets = np.array([0.5, -0.5, 1.5, -1.5, 0.25, -0.25, 1.0, -1.0, 1e-20, -1e-20, 7.5, 0.0])
    positions = base + offsets[:, None, None]
    want = ora.calc_polarizabilities(oracle_model(state), positions)
    got = rb.InterpolationModel(state).calc_polarizabilities(positions)
    assert rel_err(got, want) <= ALPHA_RTOL
    art = synthetic.make_model("TiO2", "art", seed=9)
    art.ref_positions = state.ref_positions
    want = ora.calc_polarizabilities(oracle_model(art), positions)
    assert rel_err(rb.ARTModel(art).calc_polarizabilities(positions), want) <= ALPHA_RTOL


def test_extrapolation_and_nan():
    """Amplitudes far outside the knot range use the end polynomials; NaN in -> NaN out."""
    state = synthetic.make_model("STO", "cubic", num_dofs=100, seed=21)
    positions = synthetic.make_trajectory("STO", 16, seed=2, noise=0.3)  # amplitudes >> 0.2 Å
    want = ora.calc_polarizabilities(oracle_model(state), positions)
    model = rb.InterpolationModel(state)
    assert rel_err(model.calc_polarizabilities(positions), want) <= ALPHA_RTOL
    positions[3, 5, 1] = np.nan
    got = model.calc_polarizabilities(positions)
    assert np.isnan(got[3]).all() and np.isfinite(np.delete(got, 3, axis=0)).all()


def test_empty_batch_and_no_dofs():
    state = synthetic.make_model("TiO2", "art", num_dofs=10)
    model = rb.ARTModel(state)
    assert model.calc_polarizabilities(np.zeros((0, 108, 3))).shape == (0, 3, 3)
    empty = rb.ModelState(state.ref_positions, state.lattice, state.ref_polarizability)
    got = rb.InterpolationModel(empty).calc_polarizabilities(synthetic.make_trajectory("TiO2", 5))
    assert np.array_equal(got, np.broadcast_to(state.ref_polarizability, (5, 3, 3)))


def test_mask_mutation_and_deepcopy():
    """Models are mutable (mask setter, unmask) and deep-copied by get_masked_model
    (_interpolation.py:174-189,697-712): the device tables must follow."""
    state = synthetic.make_model("LLZO", "mixed", num_dofs=50, seed=4)
    positions = synthetic.make_trajectory("LLZO", 20, seed=8)
    model = rb.InterpolationModel(state)
    omodel = oracle_model(state)
    assert rel_err(model.calc_polarizabilities(positions), ora.calc_polarizabilities(omodel, positions)) <= ALPHA_RTOL
    masked = model.get_masked_model([0, 3, 17])
    omasked = copy.deepcopy(omodel)
    omasked.mask = masked.mask
    assert rel_err(masked.calc_polarizabilities(positions), ora.calc_polarizabilities(omasked, positions)) <= ALPHA_RTOL
    assert not model.mask.any()
    mask = model.mask
    mask[5] = True
    model.mask = mask
    omodel.mask = mask
    assert rel_err(model.calc_polarizabilities(positions), ora.calc_polarizabilities(omodel, positions)) <= ALPHA_RTOL
    model.unmask()
    omodel.mask = np.zeros(50, dtype=bool)
    assert rel_err(model.calc_polarizabilities(positions), ora.calc_polarizabilities(omodel, positions)) <= ALPHA_RTOL


def test_error_behaviour():
    """Messages pinned by the reference's tests."""
    state = synthetic.make_model("TiO2", "art", num_dofs=6)
    model = rb.ARTModel(state)
    with pytest.raises(TypeError, match="positions should have type ndarray, not list"):
        model.calc_polarizabilities([[1.0, 2.0, 3.0]])
    with pytest.raises(ValueError, match=r"positions has wrong shape: \(4,5,3\) != \(_,108,3\)"):
        model.calc_polarizabilities(np.zeros((4, 5, 3)))
    with pytest.raises(ValueError, match=r"positions has wrong shape: \(108,3\) != \(_,108,3\)"):
        model.calc_polarizabilities(np.zeros((108, 3)))
    dummy = rb.ModelState(state.ref_positions, state.lattice, np.zeros((3, 3)), is_dummy_model=True)
    dummy.basis_vectors.append(np.zeros((108, 3)))
    dummy.mask = np.array([False])
    with pytest.raises(rb.UserError, match="dummy model cannot calculate polarizabilities"):
        rb.ARTModel(dummy).calc_polarizabilities(np.zeros((1, 108, 3)))


def test_large_batch_linearity_property():
    """Full-size property check (no oracle at this size): for an ARTModel alpha - alpha0 is
    linear in the wrapped displacement, so alpha(p_ref + 2d) - alpha0 == 2 (alpha(p_ref + d) - alpha0)."""
    state = synthetic.make_model("LLZO", "art")
    model = rb.ARTModel(state)
    frames = 200_000
    pos = synthetic.make_trajectory_cuda("LLZO", frames, "cuda:0", seed=123, noise=0.01)
    ref = to_cuda(state.ref_positions)
    disp = pos - ref
    disp = disp - torch.round(disp)
    a1 = model.calc_polarizabilities(ref + disp)
    a2 = model.calc_polarizabilities(ref + 2 * disp)
    a0 = model.calc_polarizabilities(ref[None])
    lhs = (a2 - a0).cpu().numpy()
    rhs = 2 * (a1 - a0).cpu().numpy()
    assert rel_err(lhs, rhs) <= 1e-10
    # and the first / last frames against the oracle
    sel = np.r_[0:64, frames - 64:frames]
    want = ora.calc_polarizabilities(oracle_model(state), pos[sel].cpu().numpy())
    got = model.calc_polarizabilities(pos)[sel].cpu().numpy()
    assert rel_err(got, want) <= ALPHA_RTOL


def test_c5_llzo_supercell_1536_atoms():
    """BASELINE.json configs[4] shape: 1536-atom LLZO 2x2x2 supercell, InterpolationModel with
    ~4600 DOFs of mixed degree 1-3 (170 MB basis, streamed from HBM/L2), on a few hundred frames."""
    state = synthetic.make_model("LLZO_2x2x2", "mixed", num_dofs=4600, masked_fraction=0.05, seed=5)
    positions = synthetic.make_trajectory("LLZO_2x2x2", 300, seed=12)
    model = rb.InterpolationModel(state)
    got = model.calc_polarizabilities(to_cuda(positions)).cpu().numpy()
    sel = np.r_[0:24, 150:158, 292:300]
    want = ora.calc_polarizabilities(oracle_model(state), positions[sel])
    assert rel_err(got[sel], want) <= ALPHA_RTOL
    info = model.path_info()
    assert info["num_atoms"] == 1536 and info["affine_dofs"] + info["dense_dofs"] == 4600


@pytest.mark.parametrize("variant", [0, 64, 80, 3])
@pytest.mark.parametrize("structure,offset", [("STO", 0), ("STO", 1), ("TiO2", 0), ("TiO2", 1)])
def test_dense_row_alignment_and_variants(structure, offset, variant):
    """Row alignments of the dense producers (odd row lengths — STO, 3N = 405 — and views that start on an
    odd element take the 8-byte copy path) against the oracle, for the A/B variants of the kernel
    (rn_debug_set_dense_config: automatic ring depth, 4 and 5 slots, branchy wrap + computed padding)."""
    hook = _lib.lib().rn_debug_set_dense_config
    frames = 777
    state = synthetic.make_model(structure, "cubic", masked_fraction=0.05, seed=11)
    positions = synthetic.make_trajectory(structure, frames, seed=5, lattice_hops=True)
    want = ora.calc_polarizabilities(oracle_model(state), positions)
    flat = torch.zeros(positions.size + 2, dtype=torch.float64, device="cuda:0")
    view = flat[offset:offset + positions.size].view(positions.shape)
    view.copy_(torch.from_numpy(positions))
    assert view.data_ptr() % 16 == 8 * offset
    hook(4, variant)
    try:
        got = rb.InterpolationModel(state).calc_polarizabilities(view).cpu().numpy()
    finally:
        hook(4, 0)
    assert rel_err(got, want) <= ALPHA_RTOL


@pytest.mark.parametrize("first_tile, tiles, stripe", [(37, 300, 1024), (0, 1000, 2048), (129, 100, 1024)])
def test_routed_evaluation_in_two_phases(first_tile, tiles, stripe):
    """rn_calc_polarizabilities_routed_phase: phases 0 and 1 together store exactly what the one-launch
    routed evaluation stores — locally and in the (here: emulated, same GPU) series buffers of both ranks."""
    lib = _lib.lib()
    frames, first = 16 * tiles, 16 * first_tile
    state = synthetic.make_model("LLZO", "art")
    model = rb.ARTModel(state)
    positions = to_cuda(synthetic.make_trajectory("LLZO", frames, seed=3, lattice_hops=True))
    total = first + frames + 64
    assert model.routed_phases_supported(positions, first, stripe)
    assert not model.routed_phases_supported(positions, first + 8, stripe)
    results = []
    for phased in (False, True):
        series = [torch.full((total * 9,), float("nan"), dtype=torch.float64, device="cuda:0") for _ in range(2)]
        local_ptr = series[0].data_ptr() + first * 72
        peers = [0, series[1].data_ptr()]
        if phased:
            for phase in (1, 0):
                model.calc_polarizabilities_routed(positions, local_ptr, peers, first, 2 * stripe, stripe,
                                                   stripe=stripe, phase=phase)
        else:
            model.calc_polarizabilities_routed(positions, local_ptr, peers, first, 2 * stripe, stripe)
        torch.cuda.synchronize()
        results.append([t.cpu().numpy() for t in series])
    for plain, phased in zip(*results):
        assert np.array_equal(plain, phased, equal_nan=True)
    local = results[1][0].reshape(total, 9)[first:first + frames]
    assert not np.isnan(local).any()
    want = ora.calc_polarizabilities(oracle_model(state), positions[:256].cpu().numpy())
    assert rel_err(local[:256].reshape(-1, 3, 3), want) <= ALPHA_RTOL
    hits = lib.rn_debug_phase_tiles(frames, first, stripe, 0, (ctypes.c_int64 * 1)(), 1)
    assert 0 < hits < tiles
