"""Parity of the CUDA spectrum path (Bluestein FFT, orientational average, corrections,
smearing) against the oracle and the reference-generated goldens.
Tolerance: pointwise relative error <= 1e-8 on intensities (BASELINE.json north_star);
wavenumbers are reproduced bit for bit."""
import numpy as np
import pytest
import torch

import ramannoodle_b200 as rb
from oracle import numpy_port as ora

from gpu_helpers import INTENSITY_RTOL, to_cuda
from helpers import GOLDEN, pointwise_rel_err, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("frames", [41, 52, 258, 1000, 4097])
def test_measure_golden(frames):
    key = f"s{frames}"
    with np.load(f"{GOLDEN}/spectrum_cases.npz") as data:
        alpha, dt = data[f"{key}_alpha"], float(data[f"{key}_timestep"])
        spectrum = rb.MDRamanSpectrum(alpha, dt)
        wn, inten = spectrum.measure()
        assert np.array_equal(wn, data[f"{key}_raw_wavenumbers"])
        assert pointwise_rel_err(inten, data[f"{key}_raw_intensities"]) <= INTENSITY_RTOL
        wn, inten = spectrum.measure(laser_correction=True, laser_wavelength=532,
                                     bose_einstein_correction=True, temperature=300)
        assert np.array_equal(wn, data[f"{key}_corr_wavenumbers"])
        assert pointwise_rel_err(inten, data[f"{key}_corr_intensities"]) <= INTENSITY_RTOL
        # device-resident series (what Trajectory.get_raman_spectrum hands over)
        wn2, inten2 = rb.MDRamanSpectrum(to_cuda(alpha), dt).measure()
        assert np.array_equal(wn2, data[f"{key}_raw_wavenumbers"])
        assert pointwise_rel_err(inten2, data[f"{key}_raw_intensities"]) <= INTENSITY_RTOL


@pytest.mark.parametrize("length", [1, 2, 3, 4, 5, 7, 8, 16, 17, 40, 51, 64, 127, 128, 129, 257, 1021, 4096, 9999])
def test_signal_spectrum_vs_oracle(length):
    """calc_signal_spectrum for powers of two, primes and composites (test_trajectory_spectrum.py:18-32
    pins the ceil(S/2) output length)."""
    signal = np.random.default_rng(length).normal(size=length)
    wn, inten = rb.calc_signal_spectrum(signal, 1.5)
    ref_wn, ref_inten = ora.calc_signal_spectrum(signal, 1.5)
    assert wn.shape == (int(np.ceil(length / 2)),) and inten.shape == wn.shape
    assert np.array_equal(wn, ref_wn)
    # I[k] >= sum(x^2)/2 > 0, so pointwise relative error is well conditioned
    assert pointwise_rel_err(inten, ref_inten) <= INTENSITY_RTOL


def test_measure_large_series_vs_oracle():
    """S = 200001 (M = 2^6 * 5^5, Bluestein length 2^19) against the scipy-based oracle."""
    rng = np.random.default_rng(5)
    frames = 200_001
    steps = np.arange(frames)[:, None, None]
    alpha = (6.0 * np.eye(3)[None] + 0.05 * np.sin(0.013 * steps + rng.uniform(0, 6, (1, 3, 3)))
             + 0.01 * rng.normal(size=(frames, 3, 3)))
    wn, inten = rb.MDRamanSpectrum(alpha, 1.0).measure(laser_correction=True, bose_einstein_correction=True)
    ref_wn, ref_inten = ora.md_measure(alpha, 1.0, laser_correction=True, bose_einstein_correction=True)
    assert np.array_equal(wn, ref_wn)
    assert pointwise_rel_err(inten, ref_inten) <= INTENSITY_RTOL


def test_parseval_property_full_size():
    """Size-independent check at S = 1e6+1 (no CPU oracle needed): for a real signal
    sum_k |X_k|^2 = M sum_n x_n^2, hence sum over ALL M bins of I = (|X|^2+E)/2 is M*E.
    With the first ceil(M/2) bins returned and Hermitian symmetry this pins the FFT scale."""
    M = 1_000_000
    x = torch.randn(M, dtype=torch.float64, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(3))
    wn, inten = rb.calc_signal_spectrum(x.cpu().numpy(), 1.0)
    energy = float((x * x).sum())
    # bins 1..M/2-1 appear twice in the full spectrum, bins 0 and M/2 once
    xf = torch.fft.fft(x)
    p_half = float((xf[M // 2].abs() ** 2 + energy) / 2)
    total = inten[0] + 2 * inten[1:].sum() + p_half
    assert abs(total - M * energy) / (M * energy) < 1e-10
    ref = ((xf[: M // 2].abs() ** 2 + energy) / 2).cpu().numpy()  # independent FFT (cuFFT via torch) as a cross-check
    assert pointwise_rel_err(inten, ref) <= INTENSITY_RTOL


@pytest.mark.parametrize("function", ["gaussian", "lorentzian"])
def test_reference_smearing_goldens(function):
    """The reference's own goldens (test/tests/test_phonon_spectrum.py:403-449)."""
    with np.load(f"{GOLDEN}/smearing.npz") as data:
        wn, inten = data["known_spectrum_wavenumbers"], data["known_spectrum_intensities"]
        cw, ci = rb.convolve_spectrum(wn, inten, function)
        assert np.allclose(cw, data[f"known_{function}_spectrum_wavenumbers"])
        assert np.allclose(ci, data[f"known_{function}_spectrum_intensities"])
        ow, oi = ora.convolve_spectrum(wn, inten, function)
        assert np.array_equal(cw, ow)
        assert rel_err(ci, oi) <= 1e-12


def test_smearing_md_golden_and_custom_grid():
    with np.load(f"{GOLDEN}/spectrum_cases.npz") as data:
        wn, inten = data["s1000_corr_wavenumbers"], data["s1000_corr_intensities"]
        for function in ("gaussian", "lorentzian"):
            cw, ci = rb.convolve_spectrum(wn, inten, function, 7.5)
            assert np.array_equal(cw, data[f"s1000_{function}_wavenumbers"])
            assert rel_err(ci, data[f"s1000_{function}_intensities"]) <= 1e-12
        grid = np.linspace(-50.0, 900.0, 333)
        cw, ci = rb.convolve_spectrum(wn, inten, "gaussian", 3.0, grid)
        assert cw is grid
        assert rel_err(ci, data["s1000_grid_intensities"]) <= 1e-12


def test_smearing_many_points_with_tile_skipping():
    """Enough input points that far-apart (input chunk, output tile) pairs are skipped: the
    skipped Gaussian factors underflow to exactly 0, so the result still matches."""
    rng = np.random.default_rng(8)
    wn = np.sort(rng.uniform(1.0, 4000.0, 30_000))
    inten = rng.uniform(0.0, 2.0, 30_000)
    grid = np.linspace(-100.0, 4100.0, 1500)
    for function, width in (("gaussian", 2.0), ("lorentzian", 6.0)):
        _, got = rb.convolve_spectrum(wn, inten, function, width, grid)
        _, want = ora.convolve_spectrum(wn[:3000], inten[:3000], function, width, grid)
        _, got_small = rb.convolve_spectrum(wn[:3000], inten[:3000], function, width, grid)
        assert rel_err(got_small, want) <= 1e-12
        # full-size: compare against a float64 torch evaluation of the same sum
        d_wn, d_in, d_grid = to_cuda(wn), to_cuda(inten), to_cuda(grid)
        dx = d_wn[None, :] - d_grid[:, None]
        if function == "gaussian":
            factor = (1 / width) * (1 / np.sqrt(2 * np.pi)) * torch.exp(-(dx**2) / (2 * width**2))
        else:
            factor = (1 / np.pi) * (0.5 * width / (dx**2 + (0.5 * width) ** 2))
        ref = (factor * d_in[None, :]).sum(dim=1).cpu().numpy()
        assert rel_err(got, ref) <= 1e-12


def test_spectrum_error_behaviour():
    alpha = np.random.default_rng(0).normal(size=(30, 3, 3))
    spectrum = rb.MDRamanSpectrum(alpha, 1.0)
    with pytest.raises(NotImplementedError, match="only polycrystalline spectra are supported for now"):
        spectrum.measure(orientation="single crystal")
    with pytest.raises(ValueError, match="invalid temperature: -1 <= 0"):
        spectrum.measure(bose_einstein_correction=True, temperature=-1)
    with pytest.raises(ValueError, match="invalid laser_wavenumber"):
        spectrum.measure(laser_correction=True, laser_wavelength=-5)
    with pytest.raises(ValueError, match=r"polarizability_ts has wrong shape: \(30,3\) != \(_,3,3\)"):
        rb.MDRamanSpectrum(alpha[:, 0], 1.0)
    wn, inten = spectrum.measure()
    with pytest.raises(ValueError, match="invalid width: -1 <= 0"):
        rb.convolve_spectrum(wn, inten, "gaussian", -1)
    with pytest.raises(ValueError, match="unsupported convolution type: blah"):
        rb.convolve_spectrum(wn, inten, "blah", 3)
    with pytest.raises(TypeError, match="intensities should have type ndarray, not list"):
        rb.convolve_spectrum(np.array([1.0, 2.0, 3.0]), [0, 3, 0], "gaussian", 5)
    with pytest.raises(ValueError, match=r"intensities has wrong shape: \(2,\) != \(3,\)"):
        rb.convolve_spectrum(np.array([1.0, 2.0, 3.0]), np.array([0.0, 3.0]), "gaussian", 5)


@pytest.mark.parametrize("log2l", [3, 4, 5, 6, 7, 9, 10, 12, 13, 14, 16, 17, 20, 21, 22, 23, 24])
def test_tiled_fft_against_torch(log2l):
    """The hand-written tiled Stockham FFT (1, 2 and 3 global passes; first sub-pass radix 2, 4
    and 8) against torch.fft (cuFFT) in both directions."""
    import ctypes

    from ramannoodle_b200 import _lib
    from ramannoodle_b200.spectrum import _get_plan

    length = 1 << log2l
    frames = (length >> 1) + 1  # M = L/2 -> Bluestein length L
    plan = _get_plan(frames, 0)
    gen = torch.Generator("cuda:0").manual_seed(log2l)
    x = torch.randn(length, 2, dtype=torch.float64, device="cuda:0", generator=gen)
    out = torch.empty_like(x)
    lib = _lib.lib()
    lib.rn_debug_fft.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.rn_debug_fft.restype = ctypes.c_int
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    xc = torch.view_as_complex(x)
    for sign, ref in ((-1, torch.fft.fft(xc)), (1, torch.fft.ifft(xc) * length)):
        assert lib.rn_debug_fft(plan.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), sign,
                                stream) == 0
        got = torch.view_as_complex(out)
        err = float((got - ref).abs().max() / ref.abs().max())
        assert err < 1e-13, f"L=2^{log2l} sign={sign}: {err}"


@pytest.mark.parametrize("frames", [41, 1000, 4097, 50_001])
def test_sharded_parts_sum_to_measure(frames):
    """The three packed transforms of the sharded measure (rn_md_spectrum_part) add up to
    MDRamanSpectrum.measure (oracle parity for the multi-GPU spectrum path on one GPU)."""
    import ctypes

    from ramannoodle_b200 import _lib
    from ramannoodle_b200.distributed import spectrum_parts
    from ramannoodle_b200.spectrum import _get_plan

    rng = np.random.default_rng(frames)
    steps = np.arange(frames)[:, None, None]
    alpha = (6.0 * np.eye(3)[None] + 0.05 * np.sin(0.011 * steps + rng.uniform(0, 6, (1, 3, 3)))
             + 0.01 * rng.normal(size=(frames, 3, 3)))
    ref_wn, ref_inten = ora.md_measure(alpha, 2.0, laser_correction=True, laser_wavelength=532,
                                       bose_einstein_correction=True, temperature=250)
    d_alpha = to_cuda(alpha)
    lib = _lib.lib()
    plan = _get_plan(frames, 0)
    points = int(lib.rn_spectrum_num_points(frames))
    total = torch.zeros(points, dtype=torch.float64, device="cuda:0")
    part = torch.empty_like(total)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert sorted(spectrum_parts(2, 0) + spectrum_parts(2, 1)) == [0, 1, 2]
    assert [spectrum_parts(8, r) for r in range(4)] == [[0], [1], [2], []]
    for index in range(3):
        assert lib.rn_md_spectrum_part(plan.handle, ctypes.c_void_p(d_alpha.data_ptr()), index,
                                       ctypes.c_void_p(part.data_ptr()), stream) == 0
        total += part
    wn = torch.empty_like(total)
    inten = torch.empty_like(total)
    assert lib.rn_md_spectrum_finish(frames, ctypes.c_void_p(total.data_ptr()), 2.0, 1, 532.0, 1, 250.0,
                                     ctypes.c_void_p(wn.data_ptr()), ctypes.c_void_p(inten.data_ptr()), stream) == 0
    assert np.array_equal(wn.cpu().numpy(), ref_wn)
    assert pointwise_rel_err(inten.cpu().numpy(), ref_inten) <= INTENSITY_RTOL


@pytest.mark.parametrize("frames", [41, 1000, 4097, 50_001, 300_000])
def test_split_transforms_sum_to_measure(frames):
    """The two-rank split of every packed transform (rn_md_spectrum_half + rn_md_spectrum_half_combine:
    residues 0/1, each finishing half of the bins) emulated on one GPU: the six partial spectra add up
    to MDRamanSpectrum.measure (oracle, 1e-8) and to rn_md_spectrum_part's (1e-12)."""
    import ctypes

    from ramannoodle_b200 import _lib
    from ramannoodle_b200.distributed import spectrum_half_units
    from ramannoodle_b200.spectrum import _get_plan

    rng = np.random.default_rng(frames + 1)
    steps = np.arange(frames)[:, None, None]
    alpha = (6.0 * np.eye(3)[None] + 0.05 * np.sin(0.013 * steps + rng.uniform(0, 6, (1, 3, 3)))
             + 0.01 * rng.normal(size=(frames, 3, 3)))
    ref_wn, ref_inten = ora.md_measure(alpha, 1.5, laser_correction=True, laser_wavelength=532)
    d_alpha = to_cuda(alpha)
    lib = _lib.lib()
    plan = _get_plan(frames, 0)
    points = int(lib.rn_spectrum_num_points(frames))
    half = int(lib.rn_spectrum_half_length(plan.handle))
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    # schedules: every (part, residue) unit is owned exactly once and partners are mutual
    for world in (2, 6, 8):
        owned = []
        for rank in range(world):
            units, partner = spectrum_half_units(world, rank)
            owned += units
            if units:
                assert spectrum_half_units(world, partner)[1] == rank
                assert [(p, 1 - r) for p, r in units] == spectrum_half_units(world, partner)[0]
        assert sorted(owned) == [(p, r) for p in range(3) for r in range(2)]
    assert spectrum_half_units(4, 1) is None and spectrum_half_units(1, 0) is None

    total = torch.zeros(points, dtype=torch.float64, device="cuda:0")
    by_part = torch.zeros(points, dtype=torch.float64, device="cuda:0")
    for part in range(3):
        z = torch.full((2, half, 2), float("nan"), dtype=torch.float64, device="cuda:0")
        for residue in range(2):
            assert lib.rn_md_spectrum_half(plan.handle, ptr(d_alpha), part, residue, ptr(z[residue]), residue, stream) == 0
        piece = torch.full((points,), float("nan"), dtype=torch.float64, device="cuda:0")
        assert lib.rn_md_spectrum_half_combine(plan.handle, part, 0, ptr(z[0]), ptr(z[1]), ptr(piece), 0, stream) == 0
        assert lib.rn_md_spectrum_half_combine(plan.handle, part, 1, ptr(z[0]), ptr(z[1]), ptr(piece), 1, stream) == 0
        assert lib.rn_md_spectrum_part(plan.handle, ptr(d_alpha), part, ptr(by_part), stream) == 0
        assert rel_err(piece.cpu().numpy(), by_part.cpu().numpy()) <= 1e-12
        total += piece
    wn = torch.empty_like(total)
    inten = torch.empty_like(total)
    assert lib.rn_md_spectrum_finish(frames, ptr(total), 1.5, 1, 532.0, 0, 0.0, ptr(wn), ptr(inten), stream) == 0
    assert np.array_equal(wn.cpu().numpy(), ref_wn)
    assert pointwise_rel_err(inten.cpu().numpy(), ref_inten) <= INTENSITY_RTOL


@pytest.mark.parametrize("frames", [41, 4097, 120_001])
def test_sharded_energy_constant(frames):
    """Energy mode 1 (multi-GPU measure): parts and half transforms without the series energies plus
    the constants of three frame shards (rn_series_energy_constant) equal the default computation."""
    import ctypes

    from ramannoodle_b200 import _lib
    from ramannoodle_b200.distributed import shard_bounds
    from ramannoodle_b200.spectrum import _get_plan

    rng = np.random.default_rng(frames + 7)
    alpha = 6.0 * np.eye(3)[None] + 0.02 * rng.normal(size=(frames, 3, 3)).cumsum(axis=0) / np.sqrt(frames)
    d_alpha = to_cuda(alpha)
    lib = _lib.lib()
    plan = _get_plan(frames, 0)
    points = int(lib.rn_spectrum_num_points(frames))
    half = int(lib.rn_spectrum_half_length(plan.handle))
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    def parts_total():
        total = torch.zeros(points, dtype=torch.float64, device="cuda:0")
        piece = torch.empty_like(total)
        for part in range(3):
            assert lib.rn_md_spectrum_part(plan.handle, ptr(d_alpha), part, ptr(piece), stream) == 0
            total += piece
        return total

    def halves_total():
        total = torch.zeros(points, dtype=torch.float64, device="cuda:0")
        z = torch.empty((2, half, 2), dtype=torch.float64, device="cuda:0")
        for part in range(3):
            for residue in range(2):
                assert lib.rn_md_spectrum_half(plan.handle, ptr(d_alpha), part, residue, ptr(z[residue]),
                                               0 if (part == 0 and residue == 0) else 1, stream) == 0
            for residue in range(2):
                assert lib.rn_md_spectrum_half_combine(plan.handle, part, residue, ptr(z[0]), ptr(z[1]), ptr(total), 1,
                                                       stream) == 0
        return total

    want = parts_total()
    assert lib.rn_spectrum_set_energy_mode(plan.handle, 1) == 0
    try:
        bare_parts = parts_total()
        bare_halves = halves_total()
        constant = torch.zeros(3, dtype=torch.float64, device="cuda:0")
        for rank in range(3):
            begin, end = shard_bounds(frames - 1, 3, rank)
            assert lib.rn_series_energy_constant(plan.handle, ptr(d_alpha), begin, end, ptr(constant[rank:]), stream) == 0
    finally:
        assert lib.rn_spectrum_set_energy_mode(plan.handle, 0) == 0
    shift = constant.sum()
    assert float(shift) > 0
    assert rel_err((bare_parts + shift).cpu().numpy(), want.cpu().numpy()) <= 1e-12
    assert rel_err((bare_halves + shift).cpu().numpy(), want.cpu().numpy()) <= 1e-12
    assert rel_err(parts_total().cpu().numpy(), want.cpu().numpy()) == 0.0  # mode restored
