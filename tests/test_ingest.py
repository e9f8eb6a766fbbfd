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
    # Cartesian frames are converted like in the reference (poscar.py:118-119: positions @ inv(lattice)); a
    # trailing blank line ends the series like in the reference
    cart = [line.replace("Direct", "Cartesian") for line in text]
    (tmp_path / "cart").write_text("\n".join(cart) + "\n")
    direct = rio.read_positions_ts(good)
    lattice = rio.read_lattice(good)
    converted = rio.read_positions_ts(tmp_path / "cart")
    assert converted.shape == direct.shape
    assert np.array_equal(converted, np.array([frame @ np.linalg.inv(lattice) for frame in direct]))
    assert np.array_equal(rio.read_positions_ts(tmp_path / "cart", wrap=True), converted - converted // 1)
    (tmp_path / "blank").write_text("\n".join(text) + "\n\n\n")
    assert rio.read_positions_ts(tmp_path / "blank").shape == (3, 4, 3)
    with pytest.raises(FileNotFoundError):
        rio.read_positions_ts(tmp_path / "missing")
    with pytest.raises(ValueError, match="unsupported format"):
        rio.read_trajectory(good, 1.0, file_format="poscar")
    with pytest.raises(ValueError, match="timestep is required"):
        rio.read_trajectory(good)


def _outcar_fixture(tmp_path):
    import gzip
    import shutil

    path = tmp_path / "OUTCAR_trajectory"
    with gzip.open(os.path.join(GOLDEN, "llzo_outcar_trajectory.txt.gz"), "rb") as fin, open(path, "wb") as fout:
        shutil.copyfileobj(fin, fout)
    return path


def test_outcar_reference_fixture_bit_identical(tmp_path):
    """``test/tests/test_outcar.py:76-94`` (15 kept frames, last position) and, beyond that pin, the
    whole array the reference reader returns for its fixture (machine-learned + ab-initio steps)."""
    path = _outcar_fixture(tmp_path)
    with np.load(os.path.join(GOLDEN, "llzo_outcar_trajectory.npz")) as data:
        want, timestep = data["positions_ts"], float(data["timestep"])
    for threads in (1, 4, 0):
        trajectory = rio.read_trajectory(path, file_format="outcar", num_threads=threads)
        assert len(trajectory) == 15 and trajectory.timestep == timestep == 1.0
        assert np.array_equal(trajectory.positions_ts, want)
    assert np.allclose(trajectory[-1][-1], np.array([0.83330583, 0.83331287, 0.29209206]))
    cart, lattice, _ = rio.read_outcar_positions_ts(path, cartesian=True)
    frac, _, _ = rio.read_outcar_positions_ts(path)
    assert np.allclose(cart @ np.linalg.inv(lattice), frac, rtol=0, atol=1e-15)
    assert rio.read_trajectory(path, 2.5, file_format="outcar").timestep == 2.5


def _write_outcar(path, lattice, cart_ts, ml_pattern=None, timestep=2.0):
    """A minimal OUTCAR with the markers the reference reader walks through (outcar.py:46-86,
    212-241, 481-538)."""
    frames, atoms, _ = cart_ts.shape
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(" vasp.6.4.2 synthetic\n")
        fh.write(" POTCAR:    PAW_PBE Ti_pv 07Sep2000\n POTCAR:    PAW_PBE O 08Apr2002\n")
        fh.write(" POTCAR:    PAW_PBE Ti_pv 07Sep2000\n   VRHFIN =Ti: 3p4s3d\n   LEXCH  = PE\n")
        fh.write(" POTCAR:    PAW_PBE O 08Apr2002\n   VRHFIN =O: s2p4\n")
        fh.write(f"   ions per type =              {atoms - atoms // 3}  {atoms // 3}\n")
        fh.write(f"   POTIM  = {timestep:.4f}    time-step for ionic-motion\n")
        fh.write("      direct lattice vectors                 reciprocal lattice vectors\n")
        for row in np.eye(3):  # symmetry-reduced cell written before the flags: must be skipped
            fh.write("  " + "  ".join(f"{x:.9f}" for x in row) + "  0 0 0\n")
        fh.write(" Write flags\n")
        fh.write("      direct lattice vectors                 reciprocal lattice vectors\n")
        for row in lattice:
            fh.write("  " + "  ".join(f"{x:.9f}" for x in row) + "   0.1 0.2 0.3\n")
        for s in range(frames):
            tags = ml_pattern[s] if ml_pattern else [""]
            for tag in tags:
                fh.write(f" POSITION                                       TOTAL-FORCE (eV/Angst){tag}\n")
                fh.write(" -----------------------------------------------------------------------------------\n")
                for a in range(atoms):
                    x, y, z = cart_ts[s, a]
                    fh.write(f"  {x:12.5f} {y:12.5f} {z:12.5f}     0.1 -0.2 0.3\n")
                fh.write(" -----------------------------------------------------------------------------------\n")
                fh.write("    total drift:   0.0 0.0 0.0\n")


def test_outcar_synthetic_triclinic_and_ml_steps(tmp_path):
    rng = np.random.default_rng(12)
    lattice = np.array([[9.0, 0.3, -0.2], [1.1, 8.0, 0.4], [-0.7, 0.9, 10.0]])
    cart = rng.uniform(-3, 12, size=(40, 12, 3))
    # per stored step: an ML block alone, an ML block followed by its ab-initio repeat (dropped),
    # or a plain ab-initio block
    pattern = [[" (ML)"], [" (ML)", ""], [""]]
    ml_pattern = [pattern[s % 3] for s in range(40)]
    path = tmp_path / "OUTCAR"
    _write_outcar(path, lattice, cart, ml_pattern)
    written = np.array([[[float(f"{x:12.5f}") for x in atom] for atom in frame] for frame in cart])
    got_cart, got_lattice, timestep = rio.read_outcar_positions_ts(path, cartesian=True)
    assert timestep == 2.0 and np.array_equal(got_lattice, np.round(lattice, 9))
    assert np.array_equal(got_cart, written)
    want = np.array([frame @ np.linalg.inv(got_lattice) for frame in written])  # outcar.py:529
    got, _, _ = rio.read_outcar_positions_ts(path, num_threads=3)
    assert np.allclose(got, want, rtol=0, atol=4e-16)
    trajectory = rio.read_trajectory(path, file_format="outcar")
    assert np.allclose(trajectory.positions_ts, want - want // 1, rtol=0, atol=4e-16)


def test_outcar_malformed(tmp_path):
    lattice = np.eye(3) * 5
    cart = np.random.default_rng(1).uniform(0, 5, size=(3, 6, 3))
    good = tmp_path / "good"
    _write_outcar(good, lattice, cart)
    text = open(good, encoding="utf-8").read()
    assert rio.read_outcar_positions_ts(good)[0].shape == (3, 6, 3)
    (tmp_path / "nopotcar").write_text(text.replace("POTCAR:    ", "POTCAR: "))
    with pytest.raises(rio.InvalidFileException, match="POTCAR block not found"):
        rio.read_trajectory(tmp_path / "nopotcar", file_format="outcar")
    (tmp_path / "badsym").write_text(text.replace("PAW_PBE O 08Apr2002", "PAW_PBE Qq 08Apr2002", 1))
    with pytest.raises(rio.InvalidFileException, match="POTCAR block could not be parsed"):
        rio.read_trajectory(tmp_path / "badsym", file_format="outcar")
    (tmp_path / "notime").write_text(text.replace("time-step for ionic-motion", "time step"))
    with pytest.raises(rio.InvalidFileException, match="timestep not found"):
        rio.read_trajectory(tmp_path / "notime", file_format="outcar")
    (tmp_path / "noflags").write_text(text.replace("Write flags", "flags"))
    with pytest.raises(rio.InvalidFileException, match="outcar does not have expected format"):
        rio.read_trajectory(tmp_path / "noflags", file_format="outcar")
    (tmp_path / "nomd").write_text(text.replace("TOTAL-FORCE", "TOTAL FORCE"))
    with pytest.raises(rio.InvalidFileException, match="no trajectory found"):
        rio.read_trajectory(tmp_path / "nomd", file_format="outcar")
    lines = text.splitlines()
    (tmp_path / "short").write_text("\n".join(lines[:-5]) + "\n")
    with pytest.raises(rio.InvalidFileException, match="Cartesian positions could not be parsed"):
        rio.read_trajectory(tmp_path / "short", file_format="outcar")


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


@pytest.mark.gpu
def test_outcar_file_to_spectrum_matches_oracle(tmp_path):
    """The reference's OUTCAR MD fixture (108 atoms: the rutile TiO2 supercell) -> pinned Trajectory ->
    ARTModel spectrum, against the oracle fed with the reference reader's positions."""
    import ramannoodle_b200 as rb
    from ramannoodle_b200 import synthetic
    from helpers import oracle_model, pointwise_rel_err, rel_err
    from oracle import numpy_port as ora

    path = _outcar_fixture(tmp_path)
    with np.load(os.path.join(GOLDEN, "llzo_outcar_trajectory.npz")) as data:
        want_positions = data["positions_ts"]
    state = synthetic.make_model("TiO2", "art")
    trajectory = rio.read_trajectory(path, file_format="outcar")
    spectrum = trajectory.get_raman_spectrum(rb.ARTModel(state))
    alpha = ora.calc_polarizabilities(oracle_model(state), want_positions)
    assert rel_err(spectrum.polarizability_ts, alpha) <= 1e-10
    wn, inten = spectrum.measure()
    wn_ref, inten_ref = ora.md_measure(alpha, trajectory.timestep)
    assert np.array_equal(wn, wn_ref)
    assert pointwise_rel_err(inten, inten_ref) <= 1e-8


def test_host_apply_pbc_matches_numpy():
    """``Trajectory.__init__`` wraps host arrays with the threaded ``rn_host_apply_pbc``: bit-identical
    to ``positions - positions // 1`` (``structure/utils.py:27``) incl. -0.0, NaN, Inf, huge values."""
    import ramannoodle_b200 as rb

    rng = np.random.default_rng(3)
    x = rng.uniform(-3, 3, (700, 37, 3))
    x[0, 0, :] = [-0.0, 0.0, 1.0]
    x[0, 1, :] = [np.nan, np.inf, -np.inf]
    x[0, 2, :] = [-1e-20, 1 - 1e-17, 2.0 ** 52 + 0.5]
    x[0, 3, :] = [-5.0, 7.0, -1e-300]
    with np.errstate(invalid="ignore"):
        want = x - x // 1
    got = rb.Trajectory(x, 1.0).positions_ts
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(np.signbit(got), np.signbit(want))
    # non-contiguous and non-float64 inputs
    view = x[::2, :, ::-1]
    with np.errstate(invalid="ignore"):
        assert np.array_equal(rb.Trajectory(view, 1.0).positions_ts, view - view // 1, equal_nan=True)
    ints = rng.integers(-2, 3, (5, 4, 3)).astype(np.float32)
    assert np.array_equal(rb.Trajectory(ints, 1.0).positions_ts, ints - ints // 1)


def test_parallel_frame_scan_keeps_reference_semantics(tmp_path):
    """Files above 8 MB take the threaded frame scan when they are perfectly regular; trailing blank
    lines are fine, anything irregular falls back to the line-by-line walk of the reference
    (``xdatcar.py:33-56``): a blank label line ends the series, a truncated last frame is an error."""
    rng = np.random.default_rng(9)
    frames, atoms = 800, 400
    positions = np.round(rng.random((frames, atoms, 3)), 8)

    def write(path, blank_at=None, tail=""):
        with open(path, "w", encoding="utf-8") as fh:
            fh.write("t\n 1.0\n 10 0 0\n 0 10 0\n 0 0 10\n A B\n 200 200\n")
            for s in range(frames):
                if s == blank_at:
                    fh.write("\n")
                fh.write(f"Direct configuration= {s + 1:5d}\n")
                fh.write("".join(f"  {a:.8f}  {b:.8f}  {c:.8f}\n" for a, b, c in positions[s]))
            fh.write(tail)

    write(tmp_path / "regular")
    assert os.path.getsize(tmp_path / "regular") > (8 << 20)
    want = _python_parse(tmp_path / "regular", atoms)  # float() of every token, as the reference parses
    assert want.shape == positions.shape and np.allclose(want, positions, rtol=0, atol=1e-15)
    assert np.array_equal(rio.read_positions_ts(tmp_path / "regular"), want)
    write(tmp_path / "tail", tail="\n  \n\n")
    assert np.array_equal(rio.read_positions_ts(tmp_path / "tail"), want)
    write(tmp_path / "blank", blank_at=500)
    assert np.array_equal(rio.read_positions_ts(tmp_path / "blank"), want[:500])
    write(tmp_path / "truncated", tail="Direct configuration=   801\n  0.1 0.2 0.3\n")
    with pytest.raises(rio.InvalidFileException, match="file ends inside frame 801"):
        rio.read_positions_ts(tmp_path / "truncated")


def _vasprun_fixture(tmp_path):
    import gzip
    import shutil

    path = tmp_path / "md_run_vasprun.xml"
    with gzip.open(os.path.join(GOLDEN, "tio2_md_run_vasprun.xml.gz"), "rb") as fin, open(path, "wb") as fout:
        shutil.copyfileobj(fin, fout)
    return path


def test_vasprun_reference_fixture_bit_identical(tmp_path):
    """``test/tests/test_vasprun.py:102-126`` (19 frames, last position, POTIM 1 fs) and, beyond that pin,
    the whole array the reference reader (``io/vasp/vasprun.py:298-330``, run on the standard library's
    ElementTree by ``oracle/make_golden.py``) returns for its molecular-dynamics fixture."""
    path = _vasprun_fixture(tmp_path)
    known = np.load(os.path.join(GOLDEN, "tio2_md_run_vasprun.npz"))
    for threads in (1, 4):
        positions, timestep = rio.read_vasprun_positions_ts(path, num_threads=threads)
        assert positions.shape == (19, 108, 3)
        assert timestep == 1.0
        assert np.allclose(positions[-1][-1], [0.83414850, 0.82850374, 0.30051845])
        raw = positions - positions // 1  # the golden holds Trajectory.positions_ts (wrapped)
        assert np.array_equal(raw, known["positions_ts"])
    etree_positions, etree_timestep = rio._vasprun_positions_etree(str(path))  # pylint: disable=protected-access
    assert np.array_equal(etree_positions, positions) and etree_timestep == timestep
    trajectory = rio.read_trajectory(path, file_format="vasprun.xml")
    assert len(trajectory) == 19 and trajectory.timestep == float(known["timestep"])
    assert np.array_equal(np.asarray(trajectory.positions_ts), known["positions_ts"])
    assert rio.read_trajectory(path, timestep=2.5, file_format="vasprun.xml").timestep == 2.5


_VASPRUN_HEAD = """<?xml version="1.0" encoding="ISO-8859-1"?>
<!-- synthetic -->
<modeling>
 <generator><i name="program" type="string">vasp </i></generator>
 <parameters>
  <separator name="electronic"><i name="POTIM">  9.0</i></separator>
  <separator name="ionic">
   <i type="int" name="NSW">     3</i>
   <i name="POTIM">      {potim}</i>
  </separator>
 </parameters>
 <structure name="initialpos">
  <crystal><varray name="basis"><v> 4.0 0.0 0.0 </v><v> 0.0 4.0 0.0 </v><v> 0.0 0.0 4.0 </v></varray></crystal>
  <varray name="positions"><v> 0.0 0.0 0.0 </v><v> 0.5 0.5 0.5 </v></varray>
 </structure>
"""


def _vasprun_text(frames, potim="1.50000000", tail=""):
    body = [_VASPRUN_HEAD.format(potim=potim)]
    for frame in frames:
        rows = "\n".join("   <v> " + " ".join(repr(float(x)) for x in row) + " </v>" for row in frame)
        body.append(" <calculation><scstep><energy><i name='e_fr_energy'> -1.0 </i></energy></scstep>\n"
                    "  <structure><crystal><varray name=\"basis\"><v> 4.0 0.0 0.0 </v><v> 0.0 4.0 0.0 </v>"
                    "<v> 0.0 0.0 4.0 </v></varray></crystal><varray name=\"positions\" >\n" + rows +
                    "\n  </varray></structure></calculation>\n")
    for frame in frames:  # the frames ramannoodle reads: unnamed structures directly under the root
        rows = "\n".join("   <v> " + " ".join(repr(float(x)) for x in row) + " </v>" for row in frame)
        body.append(" <structure>\n  <crystal>\n   <varray name=\"basis\" >\n    <v> 4.0 0.0 0.0 </v>\n    <v> 0.0 4.0 0.0 </v>\n"
                    "    <v> 0.0 0.0 4.0 </v>\n   </varray>\n  </crystal>\n  <varray name=\"positions\" >\n" + rows +
                    "\n  </varray>\n  <varray name=\"velocities\"><v> 9.0 9.0 9.0 </v></varray>\n </structure>\n")
    body.append(" <structure name=\"finalpos\"><varray name=\"positions\"><v> 0.1 0.1 0.1 </v></varray></structure>\n")
    body.append(tail + "</modeling>\n")
    return "".join(body)


def test_vasprun_synthetic_against_etree_rules(tmp_path):
    """A synthesised vasprun.xml (nested calculation structures, named structures, a second varray per
    frame, a decoy POTIM, comments) through the native tokenizer against the ElementTree walk that
    restates the reference's rules; values compare bit for bit with Python's float()."""
    rng = np.random.default_rng(5)
    frames = rng.uniform(-1.5, 2.5, size=(37, 11, 3))
    frames[3, 2] = [1e-20, -0.0, 123456789.123456789]
    path = tmp_path / "vasprun.xml"
    path.write_text(_vasprun_text(frames))
    etree_positions, etree_timestep = rio._vasprun_positions_etree(str(path))  # pylint: disable=protected-access
    assert etree_positions.shape == (37, 11, 3) and etree_timestep == 1.5
    assert np.array_equal(etree_positions, frames)
    for threads in (1, 3):
        positions, timestep = rio.read_vasprun_positions_ts(path, num_threads=threads)
        assert np.array_equal(positions, frames) and timestep == 1.5
    wrapped, _ = rio.read_vasprun_positions_ts(path, wrap=True)
    assert np.array_equal(wrapped, frames - frames // 1)
    # a larger file exercises the threaded row conversion
    big = rng.uniform(0.0, 1.0, size=(400, 64, 3))
    path.write_text(_vasprun_text(big, potim="0.5"))
    positions, timestep = rio.read_vasprun_positions_ts(path)
    assert np.array_equal(positions, big) and timestep == 0.5


def test_vasprun_error_behaviour(tmp_path):
    """``test/tests/test_vasprun.py:155-205``: a file that is not XML -> "root xml element could not be
    found"; no unnamed root-level structure -> "no trajectory found"; plus the reader's other checks."""
    frames = np.zeros((2, 2, 3))
    not_xml = tmp_path / "POSCAR"
    not_xml.write_text("TiO2\n 1.0\n 4.0 0.0 0.0\n")
    with pytest.raises(rio.InvalidFileException, match="root xml element could not be found"):
        rio.read_vasprun_positions_ts(not_xml)
    empty = tmp_path / "empty.xml"
    empty.write_text(_vasprun_text(frames[:0]))
    with pytest.raises(rio.InvalidFileException, match="no trajectory found"):
        rio.read_vasprun_positions_ts(empty)
    with pytest.raises(rio.InvalidFileException, match="no trajectory found"):
        rio.read_trajectory(empty, file_format="vasprun.xml")
    no_potim = tmp_path / "no_potim.xml"
    no_potim.write_text(_vasprun_text(frames).replace('name="POTIM">      1.5', 'name="TEBEG">      1.5'))
    with pytest.raises(rio.InvalidFileException, match="timestep not found"):
        rio.read_vasprun_positions_ts(no_potim)
    no_varray = tmp_path / "no_varray.xml"
    no_varray.write_text(_vasprun_text(frames, tail=" <structure><crystal></crystal></structure>\n"))
    with pytest.raises(rio.InvalidFileException, match="structure varray not found"):
        rio.read_vasprun_positions_ts(no_varray)
    truncated = tmp_path / "truncated.xml"
    truncated.write_text(_vasprun_text(frames)[:-30])
    with pytest.raises(rio.InvalidFileException, match="root xml element could not be found"):
        rio.read_vasprun_positions_ts(truncated)
    with pytest.raises(FileNotFoundError):
        rio.read_vasprun_positions_ts(tmp_path / "missing.xml")
    # markup the tokenizer leaves to ElementTree: an entity inside a row
    entity = tmp_path / "entity.xml"
    entity.write_text(_vasprun_text(frames).replace("<v> 0.0 0.0 0.0 </v>\n   <v> 0.0 0.0 0.0 </v>\n  </varray>\n  <varray name=\"velocities\">",
                                                    "<v> 0.0 0.0 0.0 </v>\n   <v> 0.25&#32;0.5 0.75 </v>\n  </varray>\n  <varray name=\"velocities\">", 1))
    positions, _ = rio.read_vasprun_positions_ts(entity)
    assert positions.shape == (2, 2, 3) and list(positions[0, 1]) == [0.25, 0.5, 0.75]
