"""Time the dense (spline) evaluation on the dense bench shapes: python tools/run_dense_cases.py [label]
-> one JSON line per case (c2 = STO cubic 100k frames, c3dense = LLZO ARTModel forced dense 200k frames,
LLZO cubic 200k, c5 slice = 1536-atom supercell 4600 DOFs 16k frames)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import _lib, synthetic  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else ""
VARIANTS = [int(v) for v in os.environ.get("RN_VARIANTS", "0,64,96,112").split(",")]


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


CASES = [("c2", "STO", "cubic", 100_000, False, 10), ("c3dense", "LLZO", "art", 200_000, True, 10),
         ("llzo_cubic", "LLZO", "cubic", 200_000, False, 10), ("tio2_cubic", "TiO2", "cubic", 100_000, False, 10)]
if os.environ.get("RN_C5"):
    CASES.append(("c5_slice", "LLZO_2x2x2", "cubic4600", 16_384, False, 3))
for name, structure, kind, frames, force, repeats in CASES:
    state = synthetic.make_model(structure, kind)
    model = (rb.ARTModel if kind == "art" else rb.InterpolationModel)(state, force_dense=force)
    pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
    flops = 2.0 * 3 * state.num_atoms * state.num_dofs * frames
    # A/B in one process (boxes differ by 2-3 %): variant bit 0 = branchy one-DADD wrap, bit 1 = DOF padding
    # computed, bits 4-7 = ring slots (include/ramannoodle_b200_debug.h)
    out = {"label": label, "case": name, "frames": frames, "K": 3 * state.num_atoms, "J": state.num_dofs}
    for variant in VARIANTS + VARIANTS:
        _lib.lib().rn_debug_set_dense_config(4, variant)
        ms = timed(lambda: model.calc_polarizabilities(pos), repeats)
        key = f"v{variant}"
        out[key + "_ms"] = round(min(ms, out.get(key + "_ms", 1e9)), 4)
        out[key + "_tf"] = round(flops / out[key + "_ms"] / 1e9, 2)
    _lib.lib().rn_debug_set_dense_config(4, 0)
    print(json.dumps(out), flush=True)
    del pos, model
