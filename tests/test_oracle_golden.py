"""Pins the numpy oracle against golden vectors produced by the unmodified reference
(``oracle/make_golden.py``).  CPU only; these run everywhere, including the GPU box."""
import numpy as np
import pytest

from oracle import numpy_port as ora
from oracle.make_golden import SYNTHETIC_CASES
from ramannoodle_b200 import synthetic

from helpers import GOLDEN, oracle_model, state_from_tables


@pytest.mark.parametrize("prefix", ["k1", "k2", "k3", "art"])
def test_real_tio2_models(prefix):
    """Replays ``test/tests/test_phonon_spectrum.py:33-45``: the model reproduces every DFT
    tensor (atol 1e-4) and the oracle reproduces the reference bit for bit."""
    with np.load(f"{GOLDEN}/real_tio2.npz") as data:
        model = oracle_model(state_from_tables(data, prefix))
        alpha = ora.calc_polarizabilities(model, data["positions"])
        assert np.array_equal(alpha, data[f"{prefix}_alpha"])
        if prefix != "art":  # ART keeps only the antisymmetrised linear response: not a DFT interpolant
            assert np.allclose(alpha, data["known_polarizabilities"], atol=1e-4)


def test_real_tio2_masked_art():
    with np.load(f"{GOLDEN}/real_tio2.npz") as data:
        state = state_from_tables(data, "art")
        state.mask = data["art_masked_mask"]
        alpha = ora.calc_polarizabilities(oracle_model(state), data["positions"])
        assert np.array_equal(alpha, data["art_masked_alpha"])


@pytest.mark.parametrize("case", SYNTHETIC_CASES, ids=[c[0] for c in SYNTHETIC_CASES])
def test_synthetic_cases(case):
    name, structure, kind, num_dofs, noisy, masked, frames, hops, _ = case
    state = synthetic.make_model(structure, kind, num_dofs=num_dofs, noisy_basis=noisy, masked_fraction=masked)
    positions = synthetic.make_trajectory(structure, frames, timestep=1.0, seed=4242, lattice_hops=hops)
    with np.load(f"{GOLDEN}/synthetic_cases.npz") as data:
        checksum = np.array([positions.sum(), (positions**2).sum()])
        assert np.array_equal(checksum, data[f"{name}_positions_checksum"]), "synthetic generator drifted"
        alpha = ora.calc_polarizabilities(oracle_model(state), positions)
        assert np.array_equal(alpha, data[f"{name}_alpha"])
        if not hops:
            wn, inten = ora.md_measure(alpha, 1.0, laser_correction=True, laser_wavelength=532,
                                       bose_einstein_correction=True, temperature=300)
            assert np.array_equal(wn, data[f"{name}_wavenumbers"])
            assert np.array_equal(inten, data[f"{name}_intensities"])


@pytest.mark.parametrize("frames", [41, 52, 258, 1000, 4097])
def test_spectrum_cases(frames):
    key = f"s{frames}"
    with np.load(f"{GOLDEN}/spectrum_cases.npz") as data:
        alpha, dt = data[f"{key}_alpha"], float(data[f"{key}_timestep"])
        wn, inten = ora.md_measure(alpha, dt)
        assert np.array_equal(wn, data[f"{key}_raw_wavenumbers"])
        assert np.array_equal(inten, data[f"{key}_raw_intensities"])
        assert wn.shape == (int(np.ceil((frames - 1) / 2)) - 1,)
        wn, inten = ora.md_measure(alpha, dt, laser_correction=True, laser_wavelength=532,
                                   bose_einstein_correction=True, temperature=300)
        assert np.array_equal(inten, data[f"{key}_corr_intensities"])
        swn, sint = ora.calc_signal_spectrum(np.diff(alpha, axis=0)[:, 0, 1], dt)
        assert np.array_equal(swn, data[f"{key}_signal_wavenumbers"])
        assert np.array_equal(sint, data[f"{key}_signal_intensities"])
        if frames == 1000:
            for function in ("gaussian", "lorentzian"):
                cw, ci = ora.convolve_spectrum(wn, inten, function, 7.5)
                assert np.array_equal(cw, data[f"{key}_{function}_wavenumbers"])
                assert np.array_equal(ci, data[f"{key}_{function}_intensities"])


def test_reference_smearing_goldens():
    """``test/tests/test_phonon_spectrum.py:403-449`` replayed on the oracle."""
    with np.load(f"{GOLDEN}/smearing.npz") as data:
        wn, inten = data["known_spectrum_wavenumbers"], data["known_spectrum_intensities"]
        for function in ("gaussian", "lorentzian"):
            cw, ci = ora.convolve_spectrum(wn, inten, function)
            assert np.allclose(cw, data[f"known_{function}_spectrum_wavenumbers"])
            assert np.allclose(ci, data[f"known_{function}_spectrum_intensities"])


@pytest.mark.parametrize("signal_len", [40, 51])
def test_calc_signal_spectrum_shape(signal_len):
    """``test/tests/test_trajectory_spectrum.py:18-32``."""
    signal = np.random.default_rng(0).random(signal_len)
    wn, inten = ora.calc_signal_spectrum(signal, 1.0)
    assert wn.shape == (int(np.ceil(signal_len / 2)),)
    assert inten.shape == wn.shape


@pytest.mark.parametrize("positions, known", [
    (np.array([0.2, 0.3, 0]), np.array([0.2, 0.3, 0])),
    (np.array([1.2, 1.3, 1.8]), np.array([0.2, 0.3, 0.8])),
    (np.array([-6.2, -0.3, -0.4]), np.array([0.8, 0.7, 0.6])),
])
def test_apply_pbc(positions, known):
    """``test/tests/test_structure.py:186-197``."""
    assert np.allclose(ora.apply_pbc(positions), known)


@pytest.mark.parametrize("displacement, known", [
    (np.array([0.2, 0.3, 0.4]), np.array([0.2, 0.3, 0.4])),
    (np.array([1.8, -0.6, 0]), np.array([-0.2, 0.4, 0])),
    (np.array([-4.51, -0.3, 9.6]), np.array([0.49, -0.3, -0.4])),
])
def test_apply_pbc_displacement(displacement, known):
    """``test/tests/test_structure.py:216-230``."""
    assert np.allclose(ora.apply_pbc_displacement(displacement), known)


def test_phonon_path_golden():
    """Next row N1: ``Phonons.get_raman_spectrum`` + ``PhononRamanSpectrum.measure``
    of the reference on TiO2 phonons x the P1 cubic model, replayed on the oracle."""
    with np.load(f"{GOLDEN}/phonons_tio2.npz") as ph, np.load(f"{GOLDEN}/real_tio2.npz") as data:
        model = oracle_model(state_from_tables(data, "k3"))
        tensors = ora.phonon_raman_tensors(model, ph["ref_positions"], ph["displacements"])
        assert np.array_equal(tensors, ph["raman_tensors"])
        wn, inten = ora.phonon_measure(ph["wavenumbers"], tensors, laser_correction=True, laser_wavelength=532,
                                       bose_einstein_correction=True, temperature=300)
        assert np.array_equal(wn, ph["measure_wavenumbers"])
        assert np.array_equal(inten, ph["measure_intensities"])
