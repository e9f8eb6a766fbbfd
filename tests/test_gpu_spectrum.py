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


def _random_complex(length, seed):
    gen = torch.Generator("cuda:0").manual_seed(seed)
    return torch.randn(length, 2, dtype=torch.float64, device="cuda:0", generator=gen)


def _plan_for_length(log2l):
    from ramannoodle_b200.spectrum import _get_plan

    frames = ((1 << log2l) >> 1) + 1  # M = L/2 -> chirp-z length L (L >= 4096 always)
    return _get_plan(frames, 0), frames - 1


@pytest.mark.parametrize("log2l", [12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24])
def test_fft_forward_against_torch(log2l):
    """The in-place DIF forward transform (no level, one level with last radix 2 / 4 / 8, two levels)
    against torch.fft (cuFFT), through the digit-reversal map of tests/fft_model.py."""
    import ctypes

    import fft_model
    from ramannoodle_b200 import _lib

    plan, _ = _plan_for_length(log2l)
    info = (ctypes.c_int64 * 8)()
    assert _lib.lib().rn_spectrum_plan_info(plan.handle, info) == 0
    assert info[0] == log2l and info[1] == 1 and info[2] == log2l
    levels = fft_model.plan_levels(log2l)
    assert info[3] == len(levels) and [info[4], info[5]][: len(levels)] == levels
    x = _random_complex(1 << log2l, log2l)
    out = torch.empty_like(x)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert _lib.lib().rn_debug_fft_forward(plan.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                           stream) == 0
    ref = torch.fft.fft(torch.view_as_complex(x))
    perm = torch.from_numpy(fft_model.frequency_of_position(log2l)).to("cuda:0")
    stored = torch.from_numpy(fft_model.filter_layout(log2l)).to("cuda:0")
    got = torch.view_as_complex(out)[stored]  # position p of the transform is stored at stored[p]
    err = float((got - ref[perm]).abs().max() / ref.abs().max())
    assert err < 1e-13, f"L=2^{log2l}: {err}"


@pytest.mark.parametrize("log2l", [12, 13, 15, 16, 18, 21, 22, 23])
def test_fft_convolve_against_torch(log2l):
    """Forward transform, filter multiply and mirrored inverse (the whole convolution core) against
    ifft(fft(x) fft(h)) with the chirp filter h built in torch."""
    import ctypes

    from ramannoodle_b200 import _lib

    plan, m_len = _plan_for_length(log2l)
    length = 1 << log2l
    x = _random_complex(length, 100 + log2l)
    out = torch.empty_like(x)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert _lib.lib().rn_debug_fft_convolve(plan.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                            stream) == 0
    idx = torch.arange(length, device="cuda:0", dtype=torch.int64)
    m = torch.where(idx < m_len, idx, torch.where(length - idx < m_len, length - idx, torch.full_like(idx, -1)))
    phase = ((m * m) % (2 * m_len)).to(torch.float64) / m_len
    h = torch.where(m >= 0, torch.exp(1j * torch.pi * phase), torch.zeros((), dtype=torch.complex128, device="cuda:0"))
    ref = torch.fft.ifft(torch.fft.fft(torch.view_as_complex(x)) * torch.fft.fft(h)) * length
    err = float((torch.view_as_complex(out) - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, f"L=2^{log2l}: {err}"


def _emulated_shared_measure(alpha, timestep, world, by_sequence=False, **corrections):
    """The multi-GPU measure (rn_spectrum_dist_*: one transform shared by `world` ranks) with every
    rank emulated on cuda:0 — the same kernels and buffers, peers being local pointers."""
    import ctypes

    from ramannoodle_b200 import _lib

    lib = _lib.lib()
    frames = alpha.shape[0]
    d_alpha = to_cuda(alpha)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    plans = []
    for rank in range(world):
        handle = ctypes.c_void_p()
        assert lib.rn_spectrum_plan_create_dist(frames, 0, world, rank, ctypes.byref(handle)) == 0
        plans.append(handle)
    try:
        sizes = [ctypes.c_int64() for _ in range(3)]
        assert lib.rn_spectrum_dist_sizes(plans[0], *[ctypes.byref(v) for v in sizes]) == 0
        work = [torch.full((sizes[0].value // 8,), float("nan"), dtype=torch.float64, device="cuda:0") for _ in range(world)]
        recv = [torch.full((sizes[1].value // 8,), float("nan"), dtype=torch.float64, device="cuda:0") for _ in range(world)]
        spec = [torch.full((sizes[2].value // 8,), float("nan"), dtype=torch.float64, device="cuda:0") for _ in range(world)]
        info = (ctypes.c_int64 * 8)()
        assert lib.rn_spectrum_plan_info(plans[0], info) == 0
        group = int(info[1])

        def table(tensors):
            return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

        # all three sequences in one launch, or one sequence per launch (what the pipelined schedule does)
        stripe = ctypes.c_int64()
        assert lib.rn_spectrum_dist_stripe(plans[0], ctypes.byref(stripe)) == 0
        # ... and, where the blocks are large enough, the pack in the two halves of the overlapped schedule
        phases = (1, 0) if (stripe.value > 0 and not by_sequence) else (-1,)
        for seq in ((-1,) if not by_sequence else (2, 0, 1)):
            for rank in range(world):
                for phase in phases:
                    assert lib.rn_spectrum_dist_pack(plans[rank], ctypes.c_void_p(d_alpha.data_ptr()),
                                                     table(work[:group]), table(spec), world, seq, phase, stream) == 0
            for rank in range(world):
                assert lib.rn_spectrum_dist_transform(plans[rank], ctypes.c_void_p(work[rank].data_ptr()),
                                                      table(recv[:group]), seq, stream) == 0
        points = int(lib.rn_spectrum_num_points(frames))
        laser = (1 if "laser_wavelength" in corrections else 0, float(corrections.get("laser_wavelength", 0.0)))
        bose = (1 if "temperature" in corrections else 0, float(corrections.get("temperature", 0.0)))
        outputs = [(torch.full((points,), float("nan"), dtype=torch.float64, device="cuda:0"),
                    torch.full((points,), float("nan"), dtype=torch.float64, device="cuda:0")) for _ in range(world)]
        for rank in range(world):
            assert lib.rn_spectrum_dist_final(plans[rank], ctypes.c_void_p(recv[rank].data_ptr()),
                                              ctypes.c_void_p(spec[rank].data_ptr()), table(spec), world, float(timestep),
                                              *laser, *bose,
                                              ctypes.c_void_p(outputs[rank][0].data_ptr()) if rank < group else None,
                                              stream) == 0
        results = []
        for rank in range(world):
            wn, inten = outputs[rank]
            assert lib.rn_spectrum_dist_finish(plans[rank], ctypes.c_void_p(spec[rank].data_ptr()), float(timestep),
                                               None if rank < group else ctypes.c_void_p(wn.data_ptr()),
                                               ctypes.c_void_p(inten.data_ptr()), stream) == 0
            results.append((wn.cpu().numpy(), inten.cpu().numpy()))
        return results
    finally:
        for handle in plans:
            lib.rn_spectrum_plan_destroy(handle)


@pytest.mark.parametrize("frames,world", [(41, 2), (41, 8), (1000, 4), (4097, 2), (9000, 8), (50_001, 2), (50_001, 3),
                                          (50_001, 4), (50_001, 8), (300_000, 8), (300_000, 6)])
def test_shared_transform_matches_measure(frames, world):
    """One chirp-z transform decimated over 2 / 4 / 8 ranks (3 and 6 ranks: the largest power of two
    transforms, the rest only receive) reproduces MDRamanSpectrum.measure on every rank."""
    rng = np.random.default_rng(frames + world)
    steps = np.arange(frames)[:, None, None]
    alpha = (6.0 * np.eye(3)[None] + 0.05 * np.sin(0.013 * steps + rng.uniform(0, 6, (1, 3, 3)))
             + 0.01 * rng.normal(size=(frames, 3, 3)))
    ref_wn, ref_inten = ora.md_measure(alpha, 1.5, laser_correction=True, laser_wavelength=532,
                                       bose_einstein_correction=True, temperature=250)
    results = _emulated_shared_measure(alpha, 1.5, world, by_sequence=(frames % 2 == 0), laser_wavelength=532,
                                       temperature=250)
    single = rb.MDRamanSpectrum(alpha, 1.5).measure(laser_correction=True, laser_wavelength=532,
                                                    bose_einstein_correction=True, temperature=250)[1]
    for wn, inten in results:
        assert np.array_equal(wn, ref_wn)
        assert pointwise_rel_err(inten, ref_inten) <= INTENSITY_RTOL
        assert pointwise_rel_err(inten, single) <= 1e-11
    for wn, inten in results[1:]:
        assert np.array_equal(inten, results[0][1])  # every rank holds the same bits
