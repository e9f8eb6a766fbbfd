"""Single-GPU timing of the split-transform pieces (rn_md_spectrum_half / _combine) against whole parts
(rn_md_spectrum_part) on an S-frame series:  python tools/run_split_spectrum.py [S]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ramannoodle_b200 import _lib  # noqa: E402
from ramannoodle_b200.spectrum import _get_plan  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
alpha = torch.randn(frames, 3, 3, dtype=torch.float64, device="cuda:0")
lib = _lib.lib()
plan = _get_plan(frames, 0)
points = int(lib.rn_spectrum_num_points(frames))
half = int(lib.rn_spectrum_half_length(plan.handle))
z = torch.zeros(2, half, 2, dtype=torch.float64, device="cuda:0")
out = torch.zeros(points, dtype=torch.float64, device="cuda:0")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"S={frames} L/2={half}")
print(f"  whole part          : {timed(lambda: lib.rn_md_spectrum_part(plan.handle, ptr(alpha), 1, ptr(out), stream)):.4f} ms")
print(f"  half (residue 0)    : {timed(lambda: lib.rn_md_spectrum_half(plan.handle, ptr(alpha), 1, 0, ptr(z[0]), 0, stream)):.4f} ms")
print(f"  half (residue 1)    : {timed(lambda: lib.rn_md_spectrum_half(plan.handle, ptr(alpha), 1, 1, ptr(z[1]), 1, stream)):.4f} ms")
print(f"  combine (local)     : {timed(lambda: lib.rn_md_spectrum_half_combine(plan.handle, 1, 0, ptr(z[0]), ptr(z[1]), ptr(out), 0, stream)):.4f} ms")
