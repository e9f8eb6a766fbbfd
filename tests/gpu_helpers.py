"""Helpers for the ``-m gpu`` parity tests (CUDA path through the C-ABI vs the oracle)."""
import numpy as np
import torch

ALPHA_RTOL = 1e-10      # BASELINE.json north_star: <= 1e-10 relative on polarizabilities
INTENSITY_RTOL = 1e-8   # <= 1e-8 relative on spectrum intensities


def cuda_device():
    return torch.device("cuda:0")


def to_cuda(array):
    return torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64)).to(cuda_device())
