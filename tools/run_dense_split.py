"""Dense-kernel schedule timing: whole frame tiles per CTA (mode 0) vs balanced (frame tile, DOF tile)
units (mode 1 = automatic).  python tools/run_dense_split.py  -> one JSON line per case."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import _lib, synthetic  # noqa: E402

hook = _lib.lib().rn_debug_set_dense_split
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


for structure, kind, frames in (("STO", "cubic", 100_000), ("STO", "cubic", 10_000), ("TiO2", "cubic", 10_000),
                                ("LLZO", "cubic", 100_000), ("LLZO", "cubic", 1_000_000)):
    state = synthetic.make_model(structure, kind)
    model = rb.InterpolationModel(state)
    pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
    K = 3 * state.num_atoms
    out = {"structure": structure, "kind": kind, "frames": frames}
    for mode in (0, 1):
        hook(mode)
        ms = timed(lambda: model.calc_polarizabilities(pos))
        out[f"mode{mode}_ms"] = round(ms, 4)
        out[f"mode{mode}_tf"] = round(2.0 * K * state.num_dofs * frames / ms / 1e9, 2)
    hook(1)
    print(json.dumps(out), flush=True)
