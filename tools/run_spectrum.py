"""Run MDRamanSpectrum.measure a few times on a synthetic (S,3,3) series (profiling helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
alpha = torch.randn(frames, 3, 3, dtype=torch.float64, device="cuda:0")
spec = rb.MDRamanSpectrum(alpha, 1.0)
for _ in range(2):
    spec.measure_device()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    spec.measure_device()
e1.record()
torch.cuda.synchronize()
print(f"measure: {e0.elapsed_time(e1) / reps:.4f} ms for {frames} frames")
