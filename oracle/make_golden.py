"""Generate golden vectors by running the UNMODIFIED reference (authoring container only).

Usage:  python -m oracle.make_golden          (from the repo root; needs /root/reference)

Writes
  ramannoodle_b200/data/structures.npz   reference structures read from the reference's fixtures
  tests/golden/real_tio2.npz             25 real DFT geometries/tensors + P1 models (k=1,2,3) + ART
  tests/golden/synthetic_cases.npz       reference outputs for seeded synthetic models/trajectories
  tests/golden/spectrum_cases.npz        reference measure()/calc_signal_spectrum outputs
  tests/golden/smearing.npz              the reference's own known_{gaussian,lorentzian}_spectrum goldens
  tests/golden/phonons_tio2.npz          Phonons.get_raman_spectrum / PhononRamanSpectrum.measure (next row N1)
  tests/golden/sto_xdatcar*.{txt,npz}    the reference's XDATCAR fixture + its read_positions_ts output (next row N2)
  tests/golden/tio2_md_run_vasprun.*     the reference's vasprun.xml MD fixture + its read_trajectory output (N2)
Everything here is produced by importing the reference through ``oracle/ref_bootstrap.py``
(spglib/defusedxml stubbed; hot-path arithmetic untouched).  The GPU box has no reference
tree: tests there read only these files.
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle.ref_bootstrap import REFERENCE_ROOT, import_reference  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
DATA = os.path.join(REPO, "ramannoodle_b200", "data")


def reference_model_from_state(state, art: bool = False):
    """Build a reference InterpolationModel/ARTModel holding ``state`` (SURVEY.md App. A.4:
    populate exactly what ``_construct_and_add_interpolations`` leaves behind,
    ``ramannoodle/pmodel/_interpolation.py:403-407``)."""
    from scipy.interpolate import BSpline
    from ramannoodle.pmodel._art import ARTModel
    from ramannoodle.pmodel._interpolation import InterpolationModel
    from ramannoodle.structure._reference import ReferenceStructure

    structure = ReferenceStructure(state.atomic_numbers, state.lattice, state.ref_positions)
    cls = ARTModel if art else InterpolationModel
    model = cls(structure, state.ref_polarizability)
    model._cart_basis_vectors = [np.array(v) for v in state.basis_vectors]
    model._interpolations = [BSpline(t, c, k, extrapolate=True) for (t, c, k) in state.splines]
    model._mask = np.array(state.mask, dtype=bool)
    return model


SYNTHETIC_CASES = [
    # name, structure, kind, num_dofs, noisy, masked, frames, hops, art
    ("art_tio2", "TiO2", "art", None, True, 0.0, 40, False, True),
    ("art_llzo_masked", "LLZO", "art", None, True, 0.1, 24, False, True),
    ("art_llzo_onehot", "LLZO", "art", None, False, 0.0, 16, False, True),
    ("cubic_sto", "STO", "cubic", None, True, 0.0, 24, False, False),
    ("cubic_sto_hops", "STO", "cubic", 120, True, 0.1, 21, True, False),
    ("quadratic_tio2", "TiO2", "quadratic", 50, True, 0.0, 16, False, False),
    ("mixed_llzo", "LLZO", "mixed", 97, True, 0.1, 33, False, False),
    ("linear5_sto", "STO", "linear5", 64, True, 0.0, 17, False, False),
]


def make_structures() -> None:
    import ramannoodle.io.generic as generic_io

    os.makedirs(DATA, exist_ok=True)
    out = {}
    for name, path in (("TiO2", "test/data/TiO2/phonons_OUTCAR"),
                       ("STO", "test/data/STO_RATTLED_OUTCAR"),
                       ("LLZO", "test/data/LLZO/LLZO_OUTCAR")):
        structure = generic_io.read_ref_structure(os.path.join(REFERENCE_ROOT, path), file_format="outcar")
        out[f"{name}_positions"] = np.array(structure.positions)
        out[f"{name}_lattice"] = np.array(structure.lattice)
        out[f"{name}_atomic_numbers"] = np.array(structure.atomic_numbers, dtype=np.int64)
    np.savez_compressed(os.path.join(DATA, "structures.npz"), **out)


def make_real_tio2() -> None:
    """Replayable form of ``test/tests/test_phonon_spectrum.py:33-45``."""
    import ramannoodle.io.generic as generic_io
    from ramannoodle.pmodel._art import ARTModel
    from ramannoodle.pmodel._interpolation import InterpolationModel

    data_dir = os.path.join(REFERENCE_ROOT, "test/data/TiO2")
    structure = generic_io.read_ref_structure(f"{data_dir}/phonons_OUTCAR", file_format="outcar")
    _, ref_pol = generic_io.read_positions_and_polarizability(f"{data_dir}/ref_eps_OUTCAR", file_format="outcar")
    names = sorted(os.path.basename(p) for p in glob.glob(f"{data_dir}/*eps_OUTCAR"))
    positions, known = [], []
    for name in names:
        pos, pol = generic_io.read_positions_and_polarizability(f"{data_dir}/{name}", file_format="outcar")
        positions.append(pos)
        known.append(pol)
    positions = np.array(positions)
    known = np.array(known)
    out = {"file_names": np.array(names), "positions": positions, "known_polarizabilities": known,
           "ref_positions": np.array(structure.positions), "lattice": np.array(structure.lattice),
           "ref_polarizability": np.array(ref_pol)}
    import warnings
    for order in (1, 2, 3):
        model = InterpolationModel(structure, ref_pol)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for atom in ("Ti5", "O43"):
                for direction in "xyz":
                    files = [f"{data_dir}/{atom}_{s}{direction}_eps_OUTCAR" for s in ("0.1", "0.2", "m0.1", "m0.2")]
                    model.add_dof_from_files(files, file_format="outcar", interpolation_order=order)
        _dump_model(out, f"k{order}", model)
        out[f"k{order}_alpha"] = model.calc_polarizabilities(positions)
    art = ARTModel(structure, ref_pol)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for atom in ("Ti5", "O43"):
            for direction in "xyz":
                files = [f"{data_dir}/{atom}_{s}{direction}_eps_OUTCAR" for s in ("0.1", "m0.1")]
                art.add_art_from_files(files, file_format="outcar")
    _dump_model(out, "art", art)
    out["art_alpha"] = art.calc_polarizabilities(positions)
    mask = art.mask
    mask[[1, 4]] = True
    art.mask = mask
    out["art_masked_mask"] = mask
    out["art_masked_alpha"] = art.calc_polarizabilities(positions)
    np.savez_compressed(os.path.join(GOLDEN, "real_tio2.npz"), **out)


def _dump_model(out: dict, prefix: str, model) -> None:
    from ramannoodle_b200.state import ModelState

    tables = ModelState.from_reference(model).tables()
    for key, value in tables.items():
        out[f"{prefix}_{key}"] = value


def make_synthetic() -> None:
    from ramannoodle.dynamics._trajectory import Trajectory
    from ramannoodle_b200 import synthetic

    out = {}
    for name, structure, kind, num_dofs, noisy, masked, frames, hops, art in SYNTHETIC_CASES:
        state = synthetic.make_model(structure, kind, num_dofs=num_dofs, noisy_basis=noisy, masked_fraction=masked)
        state.atomic_numbers = [int(z) for z in synthetic.load_structure(structure)["atomic_numbers"]]
        model = reference_model_from_state(state, art=art)
        positions = synthetic.make_trajectory(structure, frames, timestep=1.0, seed=4242, lattice_hops=hops)
        alpha = model.calc_polarizabilities(positions)
        out[f"{name}_alpha"] = alpha
        out[f"{name}_positions_checksum"] = np.array([positions.sum(), (positions**2).sum()])
        if not hops:
            spectrum = Trajectory(positions, 1.0).get_raman_spectrum(model)
            assert np.array_equal(spectrum.polarizability_ts, alpha)
            wn, inten = spectrum.measure(laser_correction=True, laser_wavelength=532,
                                         bose_einstein_correction=True, temperature=300)
            out[f"{name}_wavenumbers"] = wn
            out[f"{name}_intensities"] = inten
    np.savez_compressed(os.path.join(GOLDEN, "synthetic_cases.npz"), **out)


def make_spectrum() -> None:
    from ramannoodle.spectrum._raman import MDRamanSpectrum
    from ramannoodle.spectrum.utils import calc_signal_spectrum, convolve_spectrum

    rng = np.random.default_rng(77)
    out = {}
    for frames, dt in ((41, 1.0), (52, 2.5), (258, 1.0), (1000, 5.0), (4097, 0.5)):
        steps = np.arange(frames)[:, None, None]
        alpha = (6.0 * np.eye(3)[None] + 0.05 * np.sin(0.07 * steps + rng.uniform(0, 6, (1, 3, 3)))
                 + 0.02 * np.sin(0.31 * steps + rng.uniform(0, 6, (1, 3, 3))) + 0.01 * rng.normal(size=(frames, 3, 3)))
        key = f"s{frames}"
        out[f"{key}_alpha"] = alpha
        out[f"{key}_timestep"] = np.array(dt)
        spectrum = MDRamanSpectrum(alpha, dt)
        wn, inten = spectrum.measure()
        out[f"{key}_raw_wavenumbers"], out[f"{key}_raw_intensities"] = wn, inten
        wn, inten = spectrum.measure(laser_correction=True, laser_wavelength=532,
                                     bose_einstein_correction=True, temperature=300)
        out[f"{key}_corr_wavenumbers"], out[f"{key}_corr_intensities"] = wn, inten
        signal = np.diff(alpha, axis=0)[:, 0, 1]
        swn, sint = calc_signal_spectrum(signal, dt)
        out[f"{key}_signal_wavenumbers"], out[f"{key}_signal_intensities"] = swn, sint
        if frames == 1000:
            for function in ("gaussian", "lorentzian"):
                cw, ci = convolve_spectrum(wn, inten, function, 7.5)
                out[f"{key}_{function}_wavenumbers"], out[f"{key}_{function}_intensities"] = cw, ci
            grid = np.linspace(-50.0, 900.0, 333)
            cw, ci = convolve_spectrum(wn, inten, "gaussian", 3.0, grid)
            out[f"{key}_grid_wavenumbers"], out[f"{key}_grid_intensities"] = cw, ci
    np.savez_compressed(os.path.join(GOLDEN, "spectrum_cases.npz"), **out)


def make_phonons() -> None:
    """Phonon path through the same evaluator: TiO2 phonons (every 4th mode of
    ``test/data/TiO2/phonons_OUTCAR``) x the P1 cubic model of ``make_real_tio2``."""
    import warnings

    import ramannoodle.io.generic as generic_io
    from ramannoodle.dynamics._phonon import Phonons
    from ramannoodle.pmodel._interpolation import InterpolationModel

    data_dir = os.path.join(REFERENCE_ROOT, "test/data/TiO2")
    structure = generic_io.read_ref_structure(f"{data_dir}/phonons_OUTCAR", file_format="outcar")
    _, ref_pol = generic_io.read_positions_and_polarizability(f"{data_dir}/ref_eps_OUTCAR", file_format="outcar")
    phonons = generic_io.read_phonons(f"{data_dir}/phonons_OUTCAR", file_format="outcar")
    keep = slice(0, None, 4)
    sub = Phonons(phonons.ref_positions, phonons.wavenumbers[keep], phonons.displacements[keep])
    model = InterpolationModel(structure, ref_pol)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for atom in ("Ti5", "O43"):
            for direction in "xyz":
                files = [f"{data_dir}/{atom}_{s}{direction}_eps_OUTCAR" for s in ("0.1", "0.2", "m0.1", "m0.2")]
                model.add_dof_from_files(files, file_format="outcar", interpolation_order=3)
    spectrum = sub.get_raman_spectrum(model)
    wn, inten = spectrum.measure(laser_correction=True, laser_wavelength=532, bose_einstein_correction=True,
                                 temperature=300)
    out = {"ref_positions": sub.ref_positions, "wavenumbers": sub.wavenumbers, "displacements": sub.displacements,
           "raman_tensors": spectrum.raman_tensors, "measure_wavenumbers": wn, "measure_intensities": inten}
    np.savez_compressed(os.path.join(GOLDEN, "phonons_tio2.npz"), **out)


def make_ingest() -> None:
    """Next row N2: the reference's own XDATCAR fixture (``test/data/STO/XDATCAR``, 4 frames x 135
    atoms; copied as DATA) and what the reference reader makes of it
    (``io/vasp/xdatcar.py:21-56``; pinned by the reference at ``test/tests/test_xdatcar.py``)."""
    import shutil

    from ramannoodle.io.vasp.xdatcar import read_positions_ts

    src = os.path.join(REFERENCE_ROOT, "test/data/STO/XDATCAR")
    shutil.copyfile(src, os.path.join(GOLDEN, "sto_xdatcar.txt"))
    np.savez_compressed(os.path.join(GOLDEN, "sto_xdatcar_positions.npz"), positions_ts=read_positions_ts(src))

    # OUTCAR molecular-dynamics fixture (``test/data/LLZO/OUTCAR_trajectory``: 15 kept frames x 108
    # atoms, machine-learned + ab-initio steps; pinned by ``test/tests/test_outcar.py:76-94``),
    # gzip-compressed DATA copy, and the reference reader's output (``io/vasp/outcar.py:497-538``)
    import gzip

    from ramannoodle.io.vasp.outcar import read_trajectory

    src = os.path.join(REFERENCE_ROOT, "test/data/LLZO/OUTCAR_trajectory")
    with open(src, "rb") as fin, gzip.GzipFile(os.path.join(GOLDEN, "llzo_outcar_trajectory.txt.gz"), "wb",
                                               mtime=0) as fout:
        shutil.copyfileobj(fin, fout)
    trajectory = read_trajectory(src)
    np.savez_compressed(os.path.join(GOLDEN, "llzo_outcar_trajectory.npz"), positions_ts=trajectory.positions_ts,
                        timestep=trajectory.timestep)


def make_vasprun() -> None:
    """Next row N2, vasprun.xml: the reference's molecular-dynamics fixture
    (``test/data/TiO2/md_run_vasprun.xml``, 19 frames x 108 atoms, POTIM 1 fs; pinned by
    ``test/tests/test_vasprun.py:102-126``), gzip-compressed DATA copy, and what the reference reader
    (``io/vasp/vasprun.py:298-330``, running on the standard library's ElementTree) makes of it."""
    import gzip
    import shutil

    from ramannoodle.io.vasp.vasprun import read_trajectory

    src = os.path.join(REFERENCE_ROOT, "test/data/TiO2/md_run_vasprun.xml")
    with open(src, "rb") as fin, gzip.GzipFile(os.path.join(GOLDEN, "tio2_md_run_vasprun.xml.gz"), "wb",
                                               mtime=0) as fout:
        shutil.copyfileobj(fin, fout)
    trajectory = read_trajectory(src)
    np.savez_compressed(os.path.join(GOLDEN, "tio2_md_run_vasprun.npz"), positions_ts=trajectory.positions_ts,
                        timestep=trajectory.timestep)


def make_smearing() -> None:
    """The reference's own goldens for ``convolve_spectrum``
    (``test/tests/test_phonon_spectrum.py:403-449``)."""
    data_dir = os.path.join(REFERENCE_ROOT, "test/data/TiO2")
    out = {}
    for name in ("known_spectrum", "known_gaussian_spectrum", "known_lorentzian_spectrum"):
        with np.load(f"{data_dir}/{name}.npz") as data:
            out[f"{name}_wavenumbers"] = data["wavenumbers"]
            out[f"{name}_intensities"] = data["intensities"]
    np.savez_compressed(os.path.join(GOLDEN, "smearing.npz"), **out)


def main() -> None:
    import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    make_structures()
    make_smearing()
    make_real_tio2()
    make_synthetic()
    make_spectrum()
    make_phonons()
    make_ingest()
    make_vasprun()
    for path in sorted(glob.glob(os.path.join(GOLDEN, "*.npz")) + glob.glob(os.path.join(DATA, "*.npz"))):
        print(f"{os.path.getsize(path):>9d}  {os.path.relpath(path, REPO)}")


if __name__ == "__main__":
    main()
