"""Mask-sweep timing: G masked copies of the LLZO ARTModel on a device-resident
trajectory, fused kernel vs one call per model.  python tools/run_sweep.py [FRAMES] [STRUCTURE] [COUNTS e.g. 2,4]
Prints one JSON line per G (CUDA-event timing, 3 warm-ups, 10 timed repeats)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import _lib, synthetic  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
structure = sys.argv[2] if len(sys.argv) > 2 else "LLZO"
state = synthetic.make_model(structure, "art")
model = rb.ARTModel(state)
pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
K = 3 * state.num_atoms
hook = _lib.lib().rn_debug_set_sweep_fused
hook.argtypes = [ctypes.c_int]
hook.restype = None
peaks = {"hbm_gbs": 6549.4, "dmma_tf": 37.16}


def timed(fn, repeats=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(repeats):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / repeats


rng = np.random.default_rng(0)
counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3, 4, 8]
for count in counts:
    masks = rng.random((count, state.num_dofs)) < 0.5
    copies = [model.get_masked_model(np.flatnonzero(row)) for row in masks]
    for m in copies:
        m.calc_polarizabilities(pos[:8])
    hook(1)
    fused = timed(lambda: rb.calc_polarizabilities_sweep(copies, pos))
    hook(0)
    separate = timed(lambda: rb.calc_polarizabilities_sweep(copies, pos))
    hook(1)
    runs = [min(4, count - i) for i in range(0, count, 4)]
    flops = sum(2.0 * K * 8 * ((9 * r + 7) // 8) for r in runs if r > 1) * frames
    print(json.dumps({
        "structure": structure, "frames": frames, "masks": count, "fused_ms": round(fused, 4),
        "separate_ms": round(separate, 4), "speedup": round(separate / fused, 3),
        "mask_frames_per_s": round(count * frames / fused * 1e3, 1),
        "fused_read_gbs": round(len(runs) * frames * K * 8 / fused / 1e6, 1),
        "fused_dmma_tf": round(flops / fused / 1e9, 2),
        "fused_dmma_frac": round(flops / fused / 1e9 / peaks["dmma_tf"], 3)}), flush=True)

# host-resident trajectory: one pass over PCIe for all masks vs one pass per mask
import time  # noqa: E402

host_frames = min(frames, 200_000)
traj = rb.Trajectory(pos[:host_frames].cpu().numpy(), 1.0)
masks = rng.random((4, state.num_dofs)) < 0.5
copies = [model.get_masked_model(np.flatnonzero(row)) for row in masks]
for label, fn in (("sweep", lambda: traj.get_raman_spectra(copies)),
                  ("one_by_one", lambda: [traj.get_raman_spectrum(m) for m in copies])):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(json.dumps({"host_trajectory_frames": host_frames, "masks": 4, "mode": label,
                      "ms": round((time.perf_counter() - t0) / 3 * 1e3, 3)}), flush=True)
