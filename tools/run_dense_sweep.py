"""Mask-sweep timing for spline models (shared projection, epilogue per mask) vs one call per model:
python tools/run_dense_sweep.py  -> one JSON line per case (CUDA events, 3 warm-ups + 10 repeats)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import _lib, synthetic  # noqa: E402

hook = _lib.lib().rn_debug_set_sweep_fused
hook.argtypes = [ctypes.c_int]
hook.restype = None


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
for structure, kind, frames in (("STO", "cubic", 100_000), ("LLZO", "cubic", 200_000), ("TiO2", "cubic", 10_000)):
    state = synthetic.make_model(structure, kind)
    model = rb.InterpolationModel(state)
    pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
    for count in (2, 4, 8):
        masks = rng.random((count, state.num_dofs)) < 0.5
        copies = [model.get_masked_model(np.flatnonzero(row)) for row in masks]
        for m in copies:
            m.calc_polarizabilities(pos[:8])
        hook(1)
        fused = timed(lambda: rb.calc_polarizabilities_sweep(copies, pos))
        hook(0)
        separate = timed(lambda: rb.calc_polarizabilities_sweep(copies, pos))
        hook(1)
        print(json.dumps({"structure": structure, "kind": kind, "frames": frames, "masks": count,
                          "fused_ms": round(fused, 4), "separate_ms": round(separate, 4),
                          "speedup": round(separate / fused, 3)}), flush=True)
