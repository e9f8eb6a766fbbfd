"""ctypes wrapper around ``oracle.c`` (plain-C oracle) — TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile ``oracle.c`` with gcc (see ``oracle/Makefile``)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "_build/liboracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_eval_bspline.restype = ctypes.c_int
        _lib.orc_get_polarizability.restype = ctypes.c_int
        _lib.orc_calc_polarizabilities.restype = ctypes.c_int
        _lib.orc_convolve_spectrum.restype = ctypes.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def apply_pbc(x):
    x = _c(x)
    out = np.empty_like(x)
    lib().orc_apply_pbc(_d(x), _d(out), ctypes.c_int64(x.size))
    return out


def apply_pbc_displacement(x):
    x = _c(x)
    out = np.empty_like(x)
    lib().orc_apply_pbc_displacement(_d(x), _d(out), ctypes.c_int64(x.size))
    return out


def eval_bspline(t, c, k, x):
    """``BSpline(t, c, k, extrapolate=True)(x)`` for c of shape (n, ...)."""
    t = _c(t)
    c = _c(c)
    x = _c(x)
    n = c.shape[0]
    m = int(np.prod(c.shape[1:])) if c.ndim > 1 else 1
    out = np.empty((x.size, m))
    rc = lib().orc_eval_bspline(_d(t), ctypes.c_int(t.size), _d(c.reshape(n, m)), ctypes.c_int(m),
                                ctypes.c_int(int(k)), _d(x), ctypes.c_int64(x.size), _d(out))
    if rc != 0:
        raise ValueError("degree too large for the C oracle")
    return out.reshape(x.shape + c.shape[1:])


def _tables(model):
    """Flatten an ``oracle.numpy_port.OracleModel`` into the ragged C tables."""
    J = len(model.basis_vectors)
    N = model.num_atoms
    basis = _c(np.array([v.reshape(-1) for v in model.basis_vectors]).reshape(J, 3 * N))
    k = np.array([s[2] for s in model.splines], dtype=np.int32)
    t_off = np.zeros(J + 1, dtype=np.int64)
    c_off = np.zeros(J + 1, dtype=np.int64)
    for j, (t, c, _) in enumerate(model.splines):
        t_off[j + 1] = t_off[j] + len(t)
        c_off[j + 1] = c_off[j] + c.shape[0]
    t = _c(np.concatenate([s[0] for s in model.splines])) if J else np.zeros(0)
    c = _c(np.concatenate([s[1].reshape(-1, 9) for s in model.splines])) if J else np.zeros((0, 9))
    weight = _c(1 - np.asarray(model.mask, dtype=np.float64))
    return basis, k, t_off, t, c_off, c, weight


def calc_polarizabilities(model, positions_batch):
    pos = _c(positions_batch)
    S, N = pos.shape[0], pos.shape[1]
    basis, k, t_off, t, c_off, c, weight = _tables(model)
    alpha = np.empty((S, 3, 3))
    rc = lib().orc_calc_polarizabilities(
        _d(_c(model.ref_positions)), _d(_c(model.lattice)), ctypes.c_int64(N), _d(pos),
        ctypes.c_int64(S), _d(basis), ctypes.c_int64(len(k)), k.ctypes.data_as(_i32p),
        t_off.ctypes.data_as(_i64p), _d(t), c_off.ctypes.data_as(_i64p), _d(c), _d(weight),
        _d(_c(model.ref_polarizability)), _d(alpha))
    if rc != 0:
        raise RuntimeError(f"orc_calc_polarizabilities failed: {rc}")
    return alpha


def cart_displacements(model, positions_batch):
    pos = _c(positions_batch)
    S, N = pos.shape[0], pos.shape[1]
    out = np.empty((S, 3 * N))
    lib().orc_cart_displacements(_d(_c(model.ref_positions)), _d(_c(model.lattice)), _d(pos),
                                 ctypes.c_int64(S), ctypes.c_int64(N), _d(out))
    return out


def signal_spectrum_direct(x, dt):
    x = _c(x)
    M = x.size
    nk = (M + 1) // 2
    wn = np.empty(nk)
    inten = np.empty(nk)
    lib().orc_signal_spectrum_direct(_d(x), ctypes.c_int64(M), ctypes.c_double(dt), _d(wn), _d(inten))
    return wn, inten


def convolve_spectrum(wn, inten, function, width, out_wn):
    wn, inten, out_wn = _c(wn), _c(inten), _c(out_wn)
    kind = {"gaussian": 0, "lorentzian": 1}[function]
    out = np.empty_like(out_wn)
    rc = lib().orc_convolve_spectrum(_d(wn), _d(inten), ctypes.c_int64(wn.size), ctypes.c_int(kind),
                                     ctypes.c_double(width), _d(out_wn), ctypes.c_int64(out_wn.size),
                                     _d(out))
    if rc != 0:
        raise ValueError("unsupported convolution type")
    return out_wn, out
