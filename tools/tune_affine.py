"""Time the TMA affine kernel configurations (frames-per-tile multiplier MT, pipeline depth) on
the c3 workload; verifies every configuration against the generic kernel first."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import _lib, synthetic  # noqa: E402

structure = sys.argv[1] if len(sys.argv) > 1 else "LLZO"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
state = synthetic.make_model(structure, "art")
model = rb.ARTModel(state)
pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
lib = _lib.lib()
lib.rn_debug_set_affine_config.argtypes = [ctypes.c_int, ctypes.c_int]
lib.rn_debug_force_generic_affine(1)
ref = model.calc_polarizabilities(pos[:100_000]).clone()
lib.rn_debug_force_generic_affine(0)
nbytes = (24 * state.num_atoms + 72) * frames
for mt, stages in [(0, 2), (1, 2), (2, 2)]:
    lib.rn_debug_set_affine_config(mt, stages)
    out = model.calc_polarizabilities(pos)
    err = float((out[:100_000] - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(20):
        e0.record()
        model.calc_polarizabilities(pos)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = float(np.median(times))
    print(f"MT={mt} STAGES={stages}: {t:.4f} ms  {nbytes / t / 1e6:.0f} GB/s  {frames / t / 1e6:.1f} Mframes/s  err={err:.1e}")
