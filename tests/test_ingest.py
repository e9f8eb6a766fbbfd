"""Next row N2: the native XDATCAR reader against the reference reader's output
(``ramannoodle/io/vasp/xdatcar.py:21-56``).  Host-only: runs without a GPU."""
import os
import time

import numpy as np
import pytest

from ramannoodle_b200 import io as rio

from helpers import GOLDEN


def test_reference_fixture_bit_identical():
    got = rio.read_positions_ts(os.path.join(GOLDEN, "sto_xdatcar.txt"))
    with np.load(os.path.join(GOLDEN, "sto_xdatcar_positions.npz")) as data:
        assert got.shape == (4, 135, 3)
        assert np.array_equal(got, data["positions_ts"])
    lattice = rio.read_lattice(os.path.join(GOLDEN, "sto_xdatcar.txt"))
    assert np.array_equal(lattice, np.eye(3) * 11.823067)


def _write_xdatcar(path, positions, scale=1.0, fmt="%12.8f"):
    frames, atoms, _ = positions.shape
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("synthetic\n")
        fh.write(f"   {scale}\n")
        for row in np.eye(3) * 10.0:
            fh.write("  " + "  ".join(f"{x:.6f}" for x in row) + "\n")
        fh.write("   A   B\n")
        fh.write(f"   {atoms - atoms // 3}   {atoms // 3}\n")
        for s in range(frames):
            fh.write(f"Direct configuration= {s + 1:5d}\n")
            for a in range(atoms):
                fh.write(" ".join(fmt % x for x in positions[s, a]) + "\n")


def _python_parse(path, atoms):
    """The reference's parsing rule for coordinate lines: ``float(item) for item in line.split()[0:3]``."""
    rows = []
    with open(path, encoding="utf-8") as fh:
        lines = fh.readlines()[7:]
    for line in lines:
        if line[:1] in "Dd":
            continue
        rows.append([float(item) for item in line.split()[0:3]])
    return np.array(rows).reshape(-1, atoms, 3)


@pytest.mark.parametrize("fmt", ["%12.8f", "%.17g", "%+.6e"])
def test_synthetic_file_matches_python_float(tmp_path, fmt):
    rng = np.random.default_rng(5)
    positions = rng.uniform(-1.5, 2.5, size=(257, 31, 3))
    path = tmp_path / "XDATCAR"
    _write_xdatcar(path, positions, scale=1.5, fmt=fmt)
    want = _python_parse(path, 31)
    for threads in (1, 3, 0):
        got = rio.read_positions_ts(path, num_threads=threads)
        assert np.array_equal(got, want)
    assert np.array_equal(rio.read_lattice(path), np.eye(3) * 15.0)
    traj = rio.read_trajectory(path, 2.0)
    assert len(traj) == 257 and traj.timestep == 2.0
    assert np.array_equal(traj.positions_ts, want - want // 1)
    assert np.array_equal(rio.read_positions_ts(path, wrap=True), want - want // 1)


def test_large_file_throughput(tmp_path):
    rng = np.random.default_rng(6)
    positions = rng.uniform(0, 1, size=(2000, 192, 3))
    path = tmp_path / "XDATCAR"
    _write_xdatcar(path, positions)
    t0 = time.perf_counter()
    got = rio.read_positions_ts(path)
    native = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = _python_parse(path, 192)
    python = time.perf_counter() - t0
    assert np.array_equal(got, want)
    assert native < python  # the Python rule is the reference's per-line float() parsing


def test_malformed_files(tmp_path):
    positions = np.random.default_rng(0).uniform(0, 1, size=(3, 4, 3))
    good = tmp_path / "good"
    _write_xdatcar(good, positions)
    text = open(good, encoding="utf-8").read().splitlines()
    # file ends inside the last frame
    (tmp_path / "short").write_text("\n".join(text[:-2]) + "\n")
    with pytest.raises(rio.InvalidFileException, match="positions could not be parsed"):
        rio.read_positions_ts(tmp_path / "short")
    # a coordinate that is not a number
    bad = list(text)
    bad[9] = "   0.1   oops   0.3"
    (tmp_path / "nan").write_text("\n".join(bad) + "\n")
    with pytest.raises(rio.InvalidFileException, match="positions could not be parsed in frame 1"):
        rio.read_positions_ts(tmp_path / "nan")
    # bad scale factor / counts
    bad = list(text)
    bad[1] = "  scale"
    (tmp_path / "scale").write_text("\n".join(bad) + "\n")
    with pytest.raises(rio.InvalidFileException, match="scale factor could not be parsed"):
        rio.read_positions_ts(tmp_path / "scale")
    bad = list(text)
    bad[6] = "   4"
    (tmp_path / "counts").write_text("\n".join(bad) + "\n")
    with pytest.raises(rio.InvalidFileException, match="wrong number of ion counts: 1 != 2"):
        rio.read_positions_ts(tmp_path / "counts")
    # Cartesian frames are refused, a trailing blank line ends the series like in the reference
    bad = [line.replace("Direct", "Cartesian") for line in text]
    (tmp_path / "cart").write_text("\n".join(bad) + "\n")
    with pytest.raises(rio.InvalidFileException, match="Cartesian"):
        rio.read_positions_ts(tmp_path / "cart")
    (tmp_path / "blank").write_text("\n".join(text) + "\n\n\n")
    assert rio.read_positions_ts(tmp_path / "blank").shape == (3, 4, 3)
    with pytest.raises(FileNotFoundError):
        rio.read_positions_ts(tmp_path / "missing")
    with pytest.raises(ValueError, match="unsupported format"):
        rio.read_trajectory(good, 1.0, file_format="outcar")


@pytest.mark.gpu
def test_file_to_spectrum_matches_oracle(tmp_path):
    """File -> pinned Trajectory -> get_raman_spectrum, against the oracle fed by Python parsing."""
    import ramannoodle_b200 as rb
    from ramannoodle_b200 import synthetic
    from helpers import oracle_model, pointwise_rel_err
    from oracle import numpy_port as ora

    state = synthetic.make_model("STO", "cubic")
    positions = synthetic.make_trajectory("STO", 600, seed=8)
    path = tmp_path / "XDATCAR"
    _write_xdatcar(path, positions, fmt="%.10f")
    parsed = _python_parse(path, positions.shape[1])
    traj = rio.read_trajectory(path, 1.5)
    assert np.array_equal(traj.positions_ts, parsed - parsed // 1)
    wn, inten = traj.get_raman_spectrum(rb.InterpolationModel(state)).measure()
    alpha = ora.calc_polarizabilities(oracle_model(state), parsed - parsed // 1)
    wn_ref, inten_ref = ora.md_measure(alpha, 1.5)
    assert np.array_equal(wn, wn_ref)
    assert pointwise_rel_err(inten, inten_ref) <= 1e-8
