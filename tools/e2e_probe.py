"""Probe the host->device path: raw pinned H2D bandwidth vs the chunked host pipeline."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import synthetic  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
state = synthetic.make_model("LLZO", "art")
model = rb.ARTModel(state)
pos = synthetic.make_trajectory_cuda("LLZO", frames, "cuda:0")
host = torch.empty(pos.shape, dtype=torch.float64, pin_memory=True)
host.copy_(pos)
torch.cuda.synchronize()
dev = torch.empty_like(pos)
for _ in range(2):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"raw pinned H2D: {host.numel() * 8 / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)")
arr = host.numpy()
traj = rb.Trajectory(arr, 1.0)
print("trajectory pinned:", traj._pinned_owner is not None)
for name, fn in (("calc_polarizabilities_to_device", lambda: model.calc_polarizabilities_to_device(traj._positions_ts)),
                 ("get_raman_spectrum+measure", lambda: traj.get_raman_spectrum(model).measure())):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt * 1e3:.1f} ms  {frames / dt / 1e6:.2f} Mframes/s  {host.numel() * 8 / dt / 1e9:.1f} GB/s")
