"""Multi-GPU parity check (launch with torchrun, one rank per GPU; tests/test_gpu_multi.py does):
ShardedTrajectory -> routed evaluation -> shared chirp-z transform, against the CPU oracle on the
full trajectory (small cases) and against the single-GPU path on the gathered trajectory (a case
large enough for strided FFT levels)."""
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
device = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=device)
ok = True
big_frames = int(os.environ.get("RN_CHECK_BIG_FRAMES", "600000"))

for structure, kind, frames in (("LLZO", "art", 5001), ("STO", "cubic", 1237), ("TiO2", "art", 40)):
    state = synthetic.make_model(structure, kind, num_dofs=None if kind == "art" else 64)
    positions = synthetic.make_trajectory(structure, frames, seed=99)
    start, stop = shard_bounds(frames, world, rank)
    model = (rb.ARTModel if kind == "art" else rb.InterpolationModel)(state, device=local)
    omodel = ora.OracleModel(state.ref_positions, state.lattice, state.ref_polarizability,
                             list(state.basis_vectors), list(state.splines), state.mask)
    want_alpha = ora.calc_polarizabilities(omodel, ora.trajectory_positions(positions))
    want_wn, want_int = ora.md_measure(want_alpha, 1.0, laser_correction=True, bose_einstein_correction=True)
    for resident, shared in ((False, True), (True, True), (True, False), (False, False)):
        block = positions[start:stop]
        if resident:
            block = torch.from_numpy(block).to(device)
        spectrum = ShardedTrajectory(block, 1.0, frames).get_raman_spectrum(model, shared=shared)
        wn, inten = spectrum.measure(laser_correction=True, bose_einstein_correction=True)
        wn2, inten2 = spectrum.measure(laser_correction=True, bose_einstein_correction=True)  # buffers are reusable
        alpha = spectrum.polarizability_ts
        e_a = np.max(np.abs(alpha - want_alpha)) / np.max(np.abs(want_alpha))
        e_i = np.max(np.abs(inten - want_int) / np.abs(want_int))
        good = (alpha.shape == (frames, 3, 3) and e_a <= 1e-10 and e_i <= 1e-8 and np.array_equal(wn, want_wn)
                and np.array_equal(inten, inten2) and np.array_equal(wn, wn2))
        ok = ok and good
        used = getattr(spectrum, "_context", None) is not None
        print(f"rank {rank}/{world} {structure}/{kind} S={frames} resident={resident} shared={shared} (used={used}): "
              f"alpha {e_a:.1e} intensity {e_i:.1e} ok={good}", flush=True)

# large cases: strided levels in the local transforms, rows routed across many tiles.  The first has an uneven
# tail; the blocks of the second start and end on multiples of 16 frames, which lets the evaluation run in two
# phases with the first half of the pack overlapped (the overlapped schedule must have been used there).
state = synthetic.make_model("LLZO", "art")
model = rb.ARTModel(state, device=local)
for frames, overlapped in ((big_frames + 7, False), (world * 16 * (big_frames // (16 * world)), True)):
    os.environ["RN_DIST_OVERLAP"] = "1" if overlapped else "0"  # read when the shared context of this size is created
    start, stop = shard_bounds(frames, world, rank)
    block = synthetic.make_trajectory_cuda("LLZO", stop - start, device, seed=1000 + rank, first_frame=start)
    sharded = ShardedTrajectory(block, 0.5, frames)
    spectrum = sharded.get_raman_spectrum(model)
    ctx = spectrum._context  # pylint: disable=protected-access
    was_overlapped = ctx is not None and ctx.packed_generation == ctx.generation
    wn, inten = spectrum.measure_device(laser_correction=True, laser_wavelength=532)
    wn_again, inten_again = spectrum.measure_device(laser_correction=True, laser_wavelength=532)  # packs everything anew
    series = spectrum._gathered()  # pylint: disable=protected-access
    # reference: the gathered trajectory through the single-GPU path on this rank
    per = -(-frames // world)
    padded = torch.zeros((per,) + tuple(block.shape[1:]), dtype=torch.float64, device=device)
    padded[: stop - start] = sharded.local._positions_ts  # pylint: disable=protected-access
    full = torch.empty((world * per,) + tuple(block.shape[1:]), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, padded)
    single = rb.Trajectory(full[:frames], 0.5).get_raman_spectrum(model)
    swn, sint = single.measure_device(laser_correction=True, laser_wavelength=532)
    e_a = float((series - single._polarizability_ts).abs().max() / single._polarizability_ts.abs().max())  # pylint: disable=protected-access
    e_i = float(((inten - sint).abs() / sint.abs()).max())
    good = (e_a <= 1e-13 and e_i <= 1e-10 and bool(torch.equal(wn, swn)) and bool(torch.equal(inten, inten_again))
            and was_overlapped == overlapped)
    ok = ok and good
    print(f"rank {rank}/{world} LLZO/art S={frames} resident shared (used={ctx is not None}) overlapped={was_overlapped}: "
          f"alpha {e_a:.1e} intensity {e_i:.1e} vs single GPU ok={good}", flush=True)
    del block, sharded, spectrum, series, padded, full, single

flag = torch.tensor([1 if ok else 0], device=device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
