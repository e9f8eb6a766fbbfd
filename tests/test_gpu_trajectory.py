"""Drop-in ``Trajectory.get_raman_spectrum`` -> ``measure`` on the GPU vs the reference
goldens and the oracle (host-resident and HBM-resident trajectories)."""
import numpy as np
import pytest

import ramannoodle_b200 as rb
from oracle import numpy_port as ora
from oracle.make_golden import SYNTHETIC_CASES
from ramannoodle_b200 import synthetic

from gpu_helpers import ALPHA_RTOL, INTENSITY_RTOL, to_cuda
from helpers import GOLDEN, oracle_model, pointwise_rel_err, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [c for c in SYNTHETIC_CASES if not c[7]], ids=lambda c: c[0])
def test_end_to_end_golden(case):
    name, structure, kind, num_dofs, noisy, masked, frames, _, art = case
    state = synthetic.make_model(structure, kind, num_dofs=num_dofs, noisy_basis=noisy, masked_fraction=masked)
    positions = synthetic.make_trajectory(structure, frames, timestep=1.0, seed=4242)
    model = (rb.ARTModel if art else rb.InterpolationModel)(state)
    with np.load(f"{GOLDEN}/synthetic_cases.npz") as data:
        for resident in (False, True):
            trajectory = rb.Trajectory(to_cuda(positions) if resident else positions, 1.0)
            assert trajectory.is_device_resident == resident
            assert len(trajectory) == frames
            spectrum = trajectory.get_raman_spectrum(model)
            assert rel_err(spectrum.polarizability_ts, data[f"{name}_alpha"]) <= ALPHA_RTOL
            wn, inten = spectrum.measure(laser_correction=True, laser_wavelength=532,
                                         bose_einstein_correction=True, temperature=300)
            assert np.array_equal(wn, data[f"{name}_wavenumbers"])
            assert pointwise_rel_err(inten, data[f"{name}_intensities"]) <= INTENSITY_RTOL


def test_trajectory_wraps_like_reference():
    """Trajectory stores apply_pbc(positions) (dynamics/_trajectory.py:45), on host or device."""
    raw = synthetic.make_trajectory("STO", 33, seed=1, lattice_hops=True)
    want = ora.trajectory_positions(raw)
    host = rb.Trajectory(raw, 2.0)
    assert np.array_equal(host.positions_ts, want)
    dev = rb.Trajectory(to_cuda(raw), 2.0)
    assert np.array_equal(dev.positions_ts, want)
    assert np.array_equal(host[3], want[3]) and np.array_equal(dev[3:5], want[3:5])
    with pytest.raises(IndexError, match="trajectory index out of bounds"):
        host[99]  # pylint: disable=pointless-statement


def test_c1_tio2_art_10k_frames_spectrum_on_cpu():
    """BASELINE.json configs[0]: TiO2 ARTModel, 10k-frame trajectory, polarizabilities on the
    GPU, MDRamanSpectrum.measure of the ORACLE on the CPU fed with the GPU series — and the
    GPU spectrum against it."""
    state = synthetic.make_model("TiO2", "art")
    positions = synthetic.make_trajectory("TiO2", 10_000, timestep=1.0, seed=2024)
    spectrum = rb.Trajectory(positions, 1.0).get_raman_spectrum(rb.ARTModel(state))
    alpha = spectrum.polarizability_ts
    sel = np.r_[0:128, 5000:5128, 9872:10_000]
    want = ora.calc_polarizabilities(oracle_model(state), positions[sel])
    assert rel_err(alpha[sel], want) <= ALPHA_RTOL
    ref_wn, ref_inten = ora.md_measure(alpha, 1.0)
    wn, inten = spectrum.measure()
    assert np.array_equal(wn, ref_wn)
    assert pointwise_rel_err(inten, ref_inten) <= INTENSITY_RTOL


def test_c2_sto_cubic_slice():
    """BASELINE.json configs[1] (InterpolationModel cubic, SrTiO3) on a 4096-frame slice."""
    state = synthetic.make_model("STO", "cubic")
    positions = synthetic.make_trajectory("STO", 4096, seed=77)
    model = rb.InterpolationModel(state)
    alpha = model.calc_polarizabilities(positions)
    sel = np.r_[0:96, 2000:2096, 4000:4096]
    want = ora.calc_polarizabilities(oracle_model(state), positions[sel])
    assert rel_err(alpha[sel], want) <= ALPHA_RTOL


def test_trajectory_error_behaviour():
    """``test/tests/test_trajectory_spectrum.py:96-139``."""
    state = synthetic.make_model("TiO2", "art", num_dofs=6)
    model = rb.ARTModel(state)
    sto = synthetic.make_trajectory("STO", 4)
    with pytest.raises(ValueError, match="polarizability_model and trajectory are incompatible"):
        rb.Trajectory(sto, 5).get_raman_spectrum(model)
    with pytest.raises(ValueError, match="timestep must be positive"):
        rb.Trajectory(sto, -1)
    with pytest.raises(TypeError, match="timestep should have type float, not list"):
        rb.Trajectory(sto, [1, 2])
    with pytest.raises(ValueError, match=r"positions_ts has wrong shape: \(4,135\) != \(_,_,3\)"):
        rb.Trajectory(sto[:, :, 0], 1.0)


def test_phonon_path_batched_on_gpu():
    """Next row N1: the 2*M central-difference geometries of ``Phonons.get_raman_spectrum``
    (dynamics/_phonon.py:93-106) as ONE batched GPU call, against the reference golden."""
    from helpers import state_from_tables

    with np.load(f"{GOLDEN}/phonons_tio2.npz") as ph, np.load(f"{GOLDEN}/real_tio2.npz") as data:
        model = rb.InterpolationModel(state_from_tables(data, "k3"))
        phonons = rb.Phonons(ph["ref_positions"], ph["wavenumbers"], ph["displacements"])
        spectrum = phonons.get_raman_spectrum(model)
        # Raman tensors are central differences of nearly equal polarizabilities (step 1e-3):
        # the 1e-15 evaluation error is amplified by ~1e3 relative to the tensor scale
        assert rel_err(spectrum.raman_tensors, ph["raman_tensors"]) <= 1e-9
        wn, inten = spectrum.measure(laser_correction=True, laser_wavelength=532,
                                     bose_einstein_correction=True, temperature=300)
        assert np.array_equal(wn, ph["measure_wavenumbers"])
        assert rel_err(inten, ph["measure_intensities"]) <= 1e-8
        sto = rb.Phonons(np.zeros((5, 3)), np.ones(2), np.zeros((2, 5, 3)))
        with pytest.raises(ValueError, match="polarizability_model and phonons are incompatible"):
            sto.get_raman_spectrum(model)
        with pytest.raises(ValueError, match=r"displacements has wrong shape: \(2,4,3\) != \(2,5,3\)"):
            rb.Phonons(np.zeros((5, 3)), np.ones(2), np.zeros((2, 4, 3)))
