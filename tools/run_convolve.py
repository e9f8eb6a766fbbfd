"""Timing of convolve_spectrum on synthetic spectra: python tools/run_convolve.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402

for points, smooth in ((4999, True), (49_999, True), (49_999, False), (499_999, True), (499_999, False)):
    wn = np.arange(1, points + 1) * (16678.0 / points)
    rng = np.random.default_rng(points)
    inten = np.exp(-((wn - 600.0) / 300.0) ** 2) * 1e3 + 1.0 if smooth else rng.uniform(0.0, 1e6, points)
    for function in ("gaussian", "lorentzian"):
        rb.convolve_spectrum(wn, inten, function, 5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            rb.convolve_spectrum(wn, inten, function, 5)
        torch.cuda.synchronize()
        print(f"K={points} smooth={smooth} {function}: {(time.perf_counter() - t0) / 3 * 1e3:.3f} ms", flush=True)
