"""world_size-2 gloo tests of the frame-sharding host logic (no GPU): shard bounds and the
single all-gather that assembles the (S,3,3) series, with per-shard series produced by the
oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ramannoodle_b200.distributed import shard_bounds


def test_spectrum_schedules():
    """Host-side schedules of the sharded measure: whole parts (``spectrum_parts``) and the two-rank
    split of every packed transform (``spectrum_half_units``)."""
    from ramannoodle_b200.distributed import spectrum_half_units, spectrum_parts

    for world in range(1, 9):
        assert sorted(sum((spectrum_parts(world, r) for r in range(world)), [])) == [0, 1, 2]
    for world in (2, 6, 7, 8):
        owned = []
        for rank in range(world):
            units, partner = spectrum_half_units(world, rank)
            owned += units
            if units:
                other_units, other_partner = spectrum_half_units(world, partner)
                assert other_partner == rank and other_units == [(p, 1 - r) for p, r in units]
            else:
                assert partner is None and world > 6 and rank >= 6
        assert sorted(owned) == [(p, r) for p in range(3) for r in range(2)]
    assert spectrum_half_units(2, 1) == ([(0, 1), (1, 1), (2, 1)], 0)
    assert spectrum_half_units(8, 5) == ([(2, 1)], 4)
    for world in (1, 3, 4, 5):
        assert spectrum_half_units(world, 0) is None
    with pytest.raises(ValueError):
        spectrum_half_units(2, 2)


def test_shard_bounds_partition_frames():
    for frames in (0, 1, 7, 8, 1000, 1001, 10_000_019):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(frames, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == frames
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
            sizes = [b - a for a, b in blocks]
            assert max(sizes) == -(-frames // world) or frames == 0
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _worker(rank, world, port, frames, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import numpy_port as ora
        from ramannoodle_b200 import synthetic
        from ramannoodle_b200.distributed import allgather_series
        from ramannoodle_b200.spectrum import MDRamanSpectrum

        state = synthetic.make_model("TiO2", "art", num_dofs=12)
        omodel = ora.OracleModel(state.ref_positions, state.lattice, state.ref_polarizability,
                                 list(state.basis_vectors), list(state.splines), state.mask)
        positions = synthetic.make_trajectory("TiO2", frames, seed=3)
        start, stop = shard_bounds(frames, world, rank)
        local = torch.from_numpy(ora.calc_polarizabilities(omodel, positions[start:stop]))
        full = allgather_series(local, frames)
        want = ora.calc_polarizabilities(omodel, positions)
        ok = tuple(full.shape) == (frames, 3, 3) and np.array_equal(full.numpy(), want)
        # the gathered series is what MDRamanSpectrum owns on every rank
        spectrum = MDRamanSpectrum(full, 1.0)
        ok = ok and np.array_equal(spectrum.polarizability_ts, want)
        try:
            allgather_series(local[:-1], frames)
            ok = False
        except ValueError:
            pass
        queue.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("frames", [64, 37])
def test_allgather_series_world2(frames):
    """Even and uneven (padded tail) shards."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, queue)) for r in range(2)]
    for proc in procs:
        proc.start()
    results = [queue.get(timeout=120) for _ in procs]
    for proc in procs:
        proc.join(timeout=60)
        assert proc.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
