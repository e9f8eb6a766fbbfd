"""Run the dense kernel a few times (profiling helper): python tools/run_dense.py STRUCT KIND FRAMES FORCE"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from ramannoodle_b200 import synthetic  # noqa: E402

structure, kind, frames, force = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4] == "1"
state = synthetic.make_model(structure, kind)
model = rb.InterpolationModel(state, force_dense=force)
pos = synthetic.make_trajectory_cuda(structure, frames, "cuda:0")
for _ in range(3):
    model.calc_polarizabilities(pos)
torch.cuda.synchronize()
print("ok")
