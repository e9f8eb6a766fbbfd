import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb
from ramannoodle_b200 import _lib, synthetic
frames = 1_000_000
state = synthetic.make_model("LLZO", "art"); model = rb.ARTModel(state)
pos = synthetic.make_trajectory_cuda("LLZO", frames, "cuda:0")
host = torch.empty(pos.shape, dtype=torch.float64, pin_memory=True); host.copy_(pos); torch.cuda.synchronize()
arr = host.numpy()
native = model._native_model(0)
alpha = torch.empty((frames, 3, 3), dtype=torch.float64, device="cuda:0")
lib = _lib.lib()
for chunk in (0, 14563, 58254, 233016, 1000000):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = lib.rn_calc_polarizabilities_host(native.handle, ctypes.c_void_p(arr.ctypes.data), frames, None, ctypes.c_void_p(alpha.data_ptr()), chunk)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"chunk={chunk} rep={rep}: {dt*1e3:.1f} ms rc={rc}", flush=True)
