"""Multi-GPU parity check (launch with torchrun, one rank per GPU): ShardedTrajectory ->
all-gather (NCCL) -> measure, against the CPU oracle on the full trajectory."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ramannoodle_b200 as rb  # noqa: E402
from oracle import numpy_port as ora  # noqa: E402
from ramannoodle_b200 import synthetic  # noqa: E402
from ramannoodle_b200.distributed import ShardedTrajectory, shard_bounds  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ok = True
for structure, kind, frames in (("LLZO", "art", 5001), ("STO", "cubic", 1237)):
    state = synthetic.make_model(structure, kind, num_dofs=None if kind == "art" else 64)
    positions = synthetic.make_trajectory(structure, frames, seed=99)
    start, stop = shard_bounds(frames, world, rank)
    model = (rb.ARTModel if kind == "art" else rb.InterpolationModel)(state, device=local)
    for resident, fused in ((False, True), (True, True), (True, False), (False, False)):
        block = positions[start:stop]
        if resident:
            block = torch.from_numpy(block).to(f"cuda:{local}")
        spectrum = ShardedTrajectory(block, 1.0, frames).get_raman_spectrum(model, fused=fused)
        wn, inten = spectrum.measure(laser_correction=True, bose_einstein_correction=True)
        omodel = ora.OracleModel(state.ref_positions, state.lattice, state.ref_polarizability,
                                 list(state.basis_vectors), list(state.splines), state.mask)
        want_alpha = ora.calc_polarizabilities(omodel, positions)
        want_wn, want_int = ora.md_measure(want_alpha, 1.0, laser_correction=True, bose_einstein_correction=True)
        alpha = spectrum.polarizability_ts
        e_a = np.max(np.abs(alpha - want_alpha)) / np.max(np.abs(want_alpha))
        e_i = np.max(np.abs(inten - want_int) / np.abs(want_int))
        good = alpha.shape == (frames, 3, 3) and e_a <= 1e-10 and e_i <= 1e-8 and np.array_equal(wn, want_wn)
        ok = ok and good
        from ramannoodle_b200 import distributed as rdist
        used_symm = bool(rdist._SYMMETRIC_SERIES)
        split = bool(rdist._SYMMETRIC_HALVES)
        print(f"rank {rank}/{world} {structure}/{kind} resident={resident} fused={fused} symm={used_symm} split={split}: alpha {e_a:.1e} intensity {e_i:.1e} ok={good}",
              flush=True)
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
