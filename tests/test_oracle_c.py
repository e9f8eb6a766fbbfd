"""Pins the plain-C oracle (``oracle/oracle.c``) against scipy and the numpy oracle."""
import numpy as np
import pytest
from scipy.interpolate import BSpline, make_interp_spline

from oracle import c_port
from oracle import numpy_port as ora
from ramannoodle_b200 import synthetic

from helpers import GOLDEN, oracle_model, rel_err, state_from_tables


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 9])
def test_deboor_bit_exact_vs_scipy(k, n):
    """The de Boor restatement equals ``BSpline.__call__`` bit for bit: inside the base
    interval, at knots, at data points, extrapolating, and NaN -> NaN."""
    if n <= k:
        pytest.skip("needs n >= k+1")
    rng = np.random.default_rng(100 * k + n)
    x = np.sort(rng.uniform(-0.3, 0.3, size=n))
    y = rng.normal(size=(n, 3, 3))
    spline = make_interp_spline(x, y, k=k, bc_type=None)
    probe = np.concatenate([rng.uniform(-0.3, 0.3, 200), x, spline.t, [-1.0, 1.0, -0.0, 0.0],
                            rng.uniform(-5, 5, 20)])
    got = c_port.eval_bspline(spline.t, spline.c, spline.k, probe)
    want = BSpline(spline.t, spline.c, spline.k, extrapolate=True)(probe)
    assert np.array_equal(got, want)
    nan = c_port.eval_bspline(spline.t, spline.c, spline.k, np.array([np.nan]))
    assert np.isnan(nan).all()


def test_pbc_helpers_match_numpy():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-7, 7, 5000), [0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 1.0, -1.0, 1e-20, -1e-20,
                                                    0.49999999999999994, 0.5000000000000001]])
    assert np.array_equal(c_port.apply_pbc(x), ora.apply_pbc(x))
    assert np.array_equal(c_port.apply_pbc_displacement(x), ora.apply_pbc_displacement(x))


@pytest.mark.parametrize("kind,structure", [("art", "TiO2"), ("cubic", "STO"), ("mixed", "LLZO")])
def test_calc_polarizabilities_matches_numpy_port(kind, structure):
    state = synthetic.make_model(structure, kind, num_dofs=60, masked_fraction=0.1)
    positions = synthetic.make_trajectory(structure, 9, seed=5, lattice_hops=True)
    model = oracle_model(state)
    want = ora.calc_polarizabilities(model, positions)
    got = c_port.calc_polarizabilities(model, positions)
    assert rel_err(got, want) < 1e-13
    cart = c_port.cart_displacements(model, positions)
    assert np.max(np.abs(cart - ora.calc_cart_displacements(model, positions))) < 1e-14


def test_real_tio2_c_port():
    with np.load(f"{GOLDEN}/real_tio2.npz") as data:
        model = oracle_model(state_from_tables(data, "k3"))
        got = c_port.calc_polarizabilities(model, data["positions"])
        assert rel_err(got, data["k3_alpha"]) < 1e-13
        assert np.allclose(got, data["known_polarizabilities"], atol=1e-4)


@pytest.mark.parametrize("length", [40, 51, 128])
def test_signal_spectrum_by_definition(length):
    """The O(M^2) definition (autocorrelation, then DFT) agrees with the reference's
    scipy correlate + fftpack route."""
    signal = np.random.default_rng(length).normal(size=length)
    wn, inten = c_port.signal_spectrum_direct(signal, 2.0)
    ref_wn, ref_inten = ora.calc_signal_spectrum(signal, 2.0)
    assert np.allclose(wn, ref_wn, rtol=1e-14)
    assert rel_err(inten, ref_inten) < 1e-12


@pytest.mark.parametrize("function", ["gaussian", "lorentzian"])
def test_convolve_matches_numpy_port(function):
    with np.load(f"{GOLDEN}/smearing.npz") as data:
        wn, inten = data["known_spectrum_wavenumbers"], data["known_spectrum_intensities"]
        grid, want = ora.convolve_spectrum(wn, inten, function)
        _, got = c_port.convolve_spectrum(wn, inten, function, 5.0, grid)
        assert rel_err(got, want) < 1e-13
