"""Deterministic synthetic workloads for the BASELINE.json configs.

Reference structures (TiO2 108 atoms, rattled SrTiO3 135 atoms, LLZO 192 atoms) were read
once from the reference's test fixtures by ``oracle/make_golden.py`` and are stored in
``ramannoodle_b200/data/structures.npz`` (positions, lattice, atomic numbers only).  The
reference ships no displaced-polarizability data for STO or LLZO, so the models for those
configs are synthetic: spline tables are built with ``scipy.interpolate.make_interp_spline``
exactly as the reference builds them (``ramannoodle/pmodel/_interpolation.py:394-401``) from
random polarizability data on the reference structure.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.interpolate import make_interp_spline

from .state import ModelState

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "structures.npz")
MODEL_SEED = 20240


def load_structure(name: str) -> dict:
    """``name`` in {"TiO2", "STO", "LLZO", "LLZO_2x2x2"} -> positions (N,3), lattice (3,3), atomic_numbers."""
    base = "LLZO" if name == "LLZO_2x2x2" else name
    with np.load(_DATA) as data:
        positions = data[f"{base}_positions"].copy()
        lattice = data[f"{base}_lattice"].copy()
        numbers = data[f"{base}_atomic_numbers"].copy()
    if name == "LLZO_2x2x2":  # 1536-atom supercell: lattice x2, positions/2 + shifts
        shifts = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], dtype=np.float64)
        positions = np.concatenate([(positions + s) / 2.0 for s in shifts])
        numbers = np.tile(numbers, 8)
        lattice = lattice * 2.0
    return {"positions": positions, "lattice": lattice, "atomic_numbers": numbers}


def _basis(num_atoms: int, num_dofs: int, rng, noisy: bool) -> np.ndarray:
    """Rows of a (J,3N) basis: one-hot (atom, direction) DOFs; ``noisy`` adds N(0,1e-6)
    off-entries and renormalises, mimicking file-built models (SURVEY.md §7 "hard parts")."""
    dim = 3 * num_atoms
    basis = np.zeros((num_dofs, dim))
    basis[np.arange(num_dofs), np.arange(num_dofs) % dim] = 1.0
    if noisy:
        basis += rng.normal(0.0, 1e-6, size=basis.shape)
        basis /= np.linalg.norm(basis, axis=1, keepdims=True)
    return basis


def make_model(structure: str = "LLZO", kind: str = "art", num_dofs: int | None = None,
               noisy_basis: bool = True, masked_fraction: float = 0.0,
               seed: int = MODEL_SEED) -> ModelState:
    """Synthetic model on a real reference structure.

    kind="art":    degree-1 splines through amplitudes (-a, +a), Δα only (ARTModel,
                   ``ramannoodle/pmodel/_art.py:187-197``) -> knots [-a,-a,a,a], c (2,3,3).
    kind="cubic":  degree-3 ``make_interp_spline`` through x=[-.2,-.1,0,.1,.2] with the
                   (0, 0) reference point included (InterpolationModel) -> c (5,3,3).
    kind="quadratic"/"linear5": degree 2 / degree 1 through the same five amplitudes.
    kind="mixed":  DOF j uses degree 1 + (j % 3) on the five-point grid.
    """
    geom = load_structure(structure)
    num_atoms = geom["positions"].shape[0]
    if num_dofs is None:
        num_dofs = 3 * num_atoms
    rng = np.random.default_rng(seed)
    ref_pol = np.diag(rng.uniform(5.0, 7.0, size=3)) + 0.01 * _sym(rng.normal(size=(3, 3)))
    state = ModelState(geom["positions"], geom["lattice"], ref_pol)
    basis = _basis(num_atoms, num_dofs, rng, noisy_basis)
    for j in range(num_dofs):
        if kind == "art":
            amp = 0.1
            half = 0.05 * _symmetric_ish(rng)
            x = np.array([-amp, amp])
            y = np.array([-half, half])
            spline = make_interp_spline(x=x, y=y, k=1, bc_type=None)
        else:
            degree = {"cubic": 3, "quadratic": 2, "linear5": 1, "mixed": 1 + (j % 3)}[kind]
            x = np.array([-0.2, -0.1, 0.0, 0.1, 0.2])
            lin = 0.5 * _symmetric_ish(rng)
            quad = 0.8 * _symmetric_ish(rng)
            cub = 2.0 * _symmetric_ish(rng)
            y = (x[:, None, None] * lin + x[:, None, None] ** 2 * quad + x[:, None, None] ** 3 * cub
                 + 1e-3 * rng.normal(size=(5, 3, 3)))
            y[2] = 0.0  # the (0, 0) point added when include_ref_polarizability=True (:312-314)
            spline = make_interp_spline(x=x, y=y, k=degree, bc_type=None)
        state.add_dof(basis[j].reshape(num_atoms, 3), spline.t, spline.c, spline.k)
    if masked_fraction > 0:
        mask = np.zeros(num_dofs, dtype=bool)
        mask[rng.choice(num_dofs, size=int(round(masked_fraction * num_dofs)), replace=False)] = True
        state.mask = mask
    return state


def _sym(a: np.ndarray) -> np.ndarray:
    return 0.5 * (a + a.T)


def _symmetric_ish(rng) -> np.ndarray:
    """Symmetric tensor plus a small antisymmetric part (rotated DFT tensors are only
    symmetric to rounding, so the kernels may not assume symmetry)."""
    a = rng.normal(size=(3, 3))
    return _sym(a) + 1e-9 * (a - a.T)


def _mode_parameters(num_atoms: int, seed: int, num_modes: int = 8):
    rng = np.random.default_rng(seed)
    amplitude = rng.uniform(0.02, 0.08, size=num_modes)  # Å
    frequency = rng.uniform(2.0, 25.0, size=num_modes)  # THz
    phase = rng.uniform(0.0, 2 * np.pi, size=num_modes)
    vectors = rng.normal(size=(num_modes, 3 * num_atoms))
    vectors /= np.linalg.norm(vectors, axis=1, keepdims=True)
    # unit vectors over 3N coordinates would make per-atom motion tiny; scale so a typical
    # atom moves by ~amplitude
    vectors *= np.sqrt(num_atoms)
    return amplitude, frequency, phase, vectors


def make_trajectory(structure: str, num_frames: int, timestep: float = 1.0, seed: int = 1000,
                    noise: float = 0.02, first_frame: int = 0, lattice_hops: bool = False) -> np.ndarray:
    """Host (numpy) trajectory, already ``apply_pbc``-wrapped; frames
    ``first_frame .. first_frame+num_frames``.

    p_s = p_ref + (sum_m A_m e_m sin(2π f_m s dt + φ_m) + σ ξ_s) · L⁻¹  (SURVEY.md §8d).
    ``lattice_hops`` shifts every 7th frame by random integer lattice translations so the
    unwrapped-input branch is exercised.
    """
    geom = load_structure(structure)
    num_atoms = geom["positions"].shape[0]
    amplitude, frequency, phase, vectors = _mode_parameters(num_atoms, seed)
    steps = np.arange(first_frame, first_frame + num_frames, dtype=np.float64)
    angles = 2 * np.pi * frequency[None, :] * 1e-3 * steps[:, None] * timestep + phase[None, :]
    cart = (np.sin(angles) * amplitude[None, :]) @ vectors  # (S,3N) Å
    rng = np.random.default_rng([seed, first_frame, 7])
    cart += noise * rng.normal(size=cart.shape)
    frac = cart.reshape(num_frames, num_atoms, 3) @ np.linalg.inv(geom["lattice"])
    positions = geom["positions"][None, :, :] + frac
    if lattice_hops:
        hops = rng.integers(-3, 4, size=(num_frames, num_atoms, 3)).astype(np.float64)
        hops[np.arange(num_frames) % 7 != 0] = 0.0
        return positions + hops  # deliberately NOT wrapped
    return positions - np.floor(positions)


def make_trajectory_cuda(structure: str, num_frames: int, device, timestep: float = 1.0,
                         seed: int = 1000, noise: float = 0.02, first_frame: int = 0,
                         chunk: int = 65536):
    """Same recipe generated on the GPU with torch (data generation only — plumbing, not
    the product path).  Returns a (S,N,3) float64 CUDA tensor, wrapped into [0,1)."""
    import torch  # pylint: disable=import-outside-toplevel

    geom = load_structure(structure)
    num_atoms = geom["positions"].shape[0]
    amplitude, frequency, phase, vectors = _mode_parameters(num_atoms, seed)
    dev = torch.device(device)
    t_amp = torch.tensor(amplitude, device=dev)
    t_freq = torch.tensor(frequency, device=dev)
    t_phase = torch.tensor(phase, device=dev)
    t_vec = torch.tensor(vectors, device=dev)
    t_inv = torch.tensor(np.linalg.inv(geom["lattice"]), device=dev)
    t_ref = torch.tensor(geom["positions"], device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed + 7919 * first_frame)
    out = torch.empty((num_frames, num_atoms, 3), dtype=torch.float64, device=dev)
    for start in range(0, num_frames, chunk):
        stop = min(num_frames, start + chunk)
        steps = torch.arange(first_frame + start, first_frame + stop, device=dev, dtype=torch.float64)
        angles = 2 * np.pi * t_freq[None, :] * 1e-3 * steps[:, None] * timestep + t_phase[None, :]
        cart = (torch.sin(angles) * t_amp[None, :]) @ t_vec
        cart += noise * torch.randn(cart.shape, generator=gen, device=dev, dtype=torch.float64)
        frac = cart.view(stop - start, num_atoms, 3) @ t_inv
        pos = t_ref[None, :, :] + frac
        out[start:stop] = pos - torch.floor(pos)
    return out
