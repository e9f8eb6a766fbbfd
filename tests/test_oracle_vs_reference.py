"""Live pin of the oracle against the UNMODIFIED reference, imported from /root/reference.

Runs only where the reference tree exists (the authoring container); on the GPU box these
tests skip and ``test_oracle_golden.py`` (same reference outputs, committed) stands in."""
import os

import numpy as np
import pytest

from oracle import numpy_port as ora
from oracle.ref_bootstrap import import_reference, reference_available
from ramannoodle_b200 import synthetic

from helpers import oracle_model

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return import_reference()


def _reference_model(state, structure, art):
    from oracle.make_golden import reference_model_from_state

    state.atomic_numbers = [int(z) for z in synthetic.load_structure(structure)["atomic_numbers"]]
    return reference_model_from_state(state, art=art)


@pytest.mark.parametrize("structure,kind,art", [("LLZO", "art", True), ("STO", "cubic", False),
                                                ("TiO2", "mixed", False)])
def test_full_path_bit_identical(ref, structure, kind, art):
    from ramannoodle.dynamics._trajectory import Trajectory
    from ramannoodle.spectrum.utils import convolve_spectrum

    state = synthetic.make_model(structure, kind, num_dofs=150, masked_fraction=0.1, seed=99)
    model = _reference_model(state, structure, art)
    raw = synthetic.make_trajectory(structure, 120, seed=31, lattice_hops=True)
    trajectory = Trajectory(raw, 2.0)
    assert np.array_equal(trajectory.positions_ts, ora.trajectory_positions(raw))
    spectrum = trajectory.get_raman_spectrum(model)
    alpha = ora.calc_polarizabilities(oracle_model(state), ora.trajectory_positions(raw))
    assert np.array_equal(spectrum.polarizability_ts, alpha)
    wn, inten = spectrum.measure(laser_correction=True, bose_einstein_correction=True)
    own, ointen = ora.md_measure(alpha, 2.0, laser_correction=True, bose_einstein_correction=True)
    assert np.array_equal(wn, own) and np.array_equal(inten, ointen)
    for function in ("gaussian", "lorentzian"):
        a = convolve_spectrum(wn, inten, function, 4.0)
        b = ora.convolve_spectrum(own, ointen, function, 4.0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_pbc_helpers_bit_identical(ref):
    from ramannoodle.structure.utils import apply_pbc, apply_pbc_displacement, calc_displacement

    rng = np.random.default_rng(1)
    x = rng.uniform(-9, 9, size=(50, 7, 3))
    y = rng.uniform(-9, 9, size=(7, 3))
    assert np.array_equal(apply_pbc(x), ora.apply_pbc(x))
    assert np.array_equal(apply_pbc_displacement(x), ora.apply_pbc_displacement(x))
    assert np.array_equal(calc_displacement(y, x), ora.calc_displacement(y, x))


def test_corrections_bit_identical(ref):
    from ramannoodle.spectrum._raman import get_bose_einstein_correction, get_laser_correction

    wn = np.linspace(1.0, 4000.0, 777)
    assert np.array_equal(get_bose_einstein_correction(wn, 300), ora.get_bose_einstein_correction(wn, 300))
    assert np.array_equal(get_laser_correction(wn, 1e7 / 532), ora.get_laser_correction(wn, 1e7 / 532))


def test_vasprun_reader_matches_reference_live(ref, tmp_path):
    """Next row N2: the native vasprun.xml reader against the unmodified reference reader
    (``io/vasp/vasprun.py:298-330``, defusedxml stood in by the stdlib ElementTree) on a synthesised
    file and on the reference's own malformed fixture."""
    from ramannoodle.exceptions import InvalidFileException
    from ramannoodle.io.vasp.vasprun import read_trajectory

    from ramannoodle_b200 import io as rio
    from test_ingest import _vasprun_text

    rng = np.random.default_rng(11)
    frames = rng.uniform(-2.0, 3.0, size=(23, 9, 3))
    path = tmp_path / "vasprun.xml"
    path.write_text(_vasprun_text(frames, potim="2.25"))
    expected = read_trajectory(path)
    positions, timestep = rio.read_vasprun_positions_ts(path, wrap=True)
    assert timestep == expected.timestep == 2.25
    assert np.array_equal(positions, expected.positions_ts)
    malformed = os.path.join(os.path.dirname(ref.__file__), "..", "test", "data", "malformed", "vasprun.xml")
    with pytest.raises(InvalidFileException, match="no trajectory found"):
        read_trajectory(malformed)
    with pytest.raises(rio.InvalidFileException, match="no trajectory found"):
        rio.read_vasprun_positions_ts(malformed)


def test_cartesian_xdatcar_matches_reference_live(ref, tmp_path):
    """XDATCAR frames in Cartesian coordinates: the reference converts them with
    ``positions @ inv(lattice)`` (``io/vasp/poscar.py:118-119``); so does this package, bit for bit."""
    from ramannoodle.io.vasp.xdatcar import read_positions_ts

    from ramannoodle_b200 import io as rio

    rng = np.random.default_rng(2)
    lines = ["synthetic", "  1.5", "  4.0 0.1 0.0", "  0.0 4.2 0.2", "  0.3 0.0 3.9", " Ti O", " 1 2"]
    for frame in range(5):
        lines.append("Cartesian configuration= %d" % (frame + 1) if frame % 2 == 0 else "Direct configuration= %d" % (frame + 1))
        for row in rng.uniform(-1.0, 7.0, size=(3, 3)):
            lines.append("  " + " ".join(repr(float(x)) for x in row))
    path = tmp_path / "XDATCAR_cart"
    path.write_text("\n".join(lines) + "\n")
    expected = read_positions_ts(path)
    assert np.array_equal(rio.read_positions_ts(path), expected)
    trajectory = rio.read_trajectory(path, 1.0)
    assert np.array_equal(np.asarray(trajectory.positions_ts), expected - expected // 1)
