"""A/B timing of the dense kernels (v1 / v2, tile shapes) on the c2 and forced-dense c3 workloads."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import _lib, synthetic  # noqa: E402

lib = _lib.lib()
lib.rn_debug_set_dense_config.argtypes = [ctypes.c_int, ctypes.c_int]
for structure, kind, frames, force in (("STO", "cubic", 100_000, False), ("LLZO", "art", 200_000, True),
                                       ("LLZO", "cubic", 200_000, False)):
    state = synthetic.make_model(structure, kind)
    model = rb.InterpolationModel(state, force_dense=force)
    pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
    flops = 2.0 * 3 * state.num_atoms * state.num_dofs * frames
    ref = None
    for version, wn in ((1, 0), (3, 0), (4, 0)):
        lib.rn_debug_set_dense_config(version, wn)
        out = model.calc_polarizabilities(pos)
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(5):
            e0.record()
            model.calc_polarizabilities(pos)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        t = float(np.median(times))
        print(f"{structure}/{kind} S={frames} v{version} wn={wn}: {t:.3f} ms  {flops / t / 1e9:.2f} TFLOP/s  "
              f"{frames / t / 1e3:.2f} Mframes/s  diff_vs_v1={err:.1e}", flush=True)
    lib.rn_debug_set_dense_config(4, 0)
