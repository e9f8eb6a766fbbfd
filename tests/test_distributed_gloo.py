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


def test_transform_group_and_routing():
    """Host-side schedule of the shared spectrum: which ranks share the transform and which rank
    consumes which difference signal (``rn_spectrum_dist_route``; mirrored by ``route_owner``)."""
    from ramannoodle_b200.distributed import route_owner, transform_group_size

    assert [transform_group_size(w) for w in range(1, 10)] == [1, 2, 2, 4, 4, 4, 4, 8, 8]
    with pytest.raises(ValueError):
        transform_group_size(0)
    # 8 ranks, S = 8e6: L = 2^24, period = L/8 = 2^21, width = 2^18; every difference signal n < M has
    # exactly one owner and the owners' blocks tile every period
    period, width = 1 << 21, 1 << 18
    for frame in (0, 1, width - 1, width, period - 1, period, 3 * period + 5 * width + 17, 7_999_998):
        owner = route_owner(frame, period, width)
        assert 0 <= owner < 8 and owner == ((frame % period) // width)
    owners = [route_owner(n, 4096, 512) for n in range(0, 3 * 4096, 512)]
    assert owners == [0, 1, 2, 3, 4, 5, 6, 7] * 3


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


class _OracleBackedModel:
    """Duck-typed polarizability model evaluated by the oracle (no GPU): drives the sharded trajectory
    through the all-gather path."""

    def __init__(self, omodel):
        self._omodel = omodel

    def calc_polarizabilities(self, positions_batch):
        from oracle import numpy_port as ora

        return ora.calc_polarizabilities(self._omodel, np.asarray(positions_batch))


def _trajectory_worker(rank, world, port, frames, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import numpy_port as ora
        from ramannoodle_b200 import synthetic
        from ramannoodle_b200.distributed import ShardedMDRamanSpectrum, ShardedTrajectory

        state = synthetic.make_model("TiO2", "art", num_dofs=12)
        omodel = ora.OracleModel(state.ref_positions, state.lattice, state.ref_polarizability,
                                 list(state.basis_vectors), list(state.splines), state.mask)
        positions = synthetic.make_trajectory("TiO2", frames, seed=5)
        start, stop = shard_bounds(frames, world, rank)
        sharded = ShardedTrajectory(positions[start:stop], 2.0, frames)
        spectrum = sharded.get_raman_spectrum(_OracleBackedModel(omodel))  # gloo -> all-gather path
        want = ora.calc_polarizabilities(omodel, ora.trajectory_positions(positions))
        ok = isinstance(spectrum, ShardedMDRamanSpectrum) and spectrum.timestep == 2.0
        ok = ok and np.array_equal(spectrum.polarizability_ts, want)
        ok = ok and np.array_equal(np.asarray(spectrum.local_polarizability_ts), want[start:stop])
        try:
            ShardedTrajectory(positions[start:stop][:-1], 2.0, frames)
            ok = False
        except ValueError:
            pass
        queue.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("frames", [50, 33])
def test_sharded_trajectory_allgather_path_world2(frames):
    """ShardedTrajectory.get_raman_spectrum on a gloo group: local evaluation, one all-gather, every
    rank holds the series the reference's MDRamanSpectrum would own."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trajectory_worker, args=(r, 2, port, frames, queue)) for r in range(2)]
    for proc in procs:
        proc.start()
    results = [queue.get(timeout=180) for _ in procs]
    for proc in procs:
        proc.join(timeout=60)
        assert proc.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


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


@pytest.mark.parametrize("num_frames, first_frame, stripe", [(1_000_000, 3_000_000, 131072), (1_000_000, 0, 262144),
                                                             (147_456, 294_912, 65536), (16, 2048, 1024),
                                                             (4096, 1024 * 7 + 16, 1024)])
def test_phase_tile_selection_partitions_the_block(num_frames, first_frame, stripe):
    """The two launches of the overlapped multi-GPU schedule (rn_calc_polarizabilities_routed_phase) between them
    evaluate every 16-frame tile of a block exactly once, phase 0 exactly the frames n with
    (n mod 2*stripe) < stripe + 16 — every row the first half of any rank's pack reads (difference signal n
    needs rows n and n + 1, n mod 2*stripe < stripe) — in increasing order (host arithmetic of the native library)."""
    import ctypes

    from ramannoodle_b200 import _lib

    lib = _lib.lib()
    tiles = num_frames // 16
    selected = []
    for phase in (0, 1):
        buffer = (ctypes.c_int64 * tiles)()
        count = lib.rn_debug_phase_tiles(num_frames, first_frame, stripe, phase, buffer, tiles)
        assert 0 <= count <= tiles
        chosen = list(buffer[:count])
        assert chosen == sorted(chosen)
        selected.append(chosen)
        for tile in chosen:
            early = ((first_frame + 16 * tile) % (2 * stripe)) < stripe + 16
            assert early == (phase == 0)
    assert sorted(selected[0] + selected[1]) == list(range(tiles))
    needed = {(n - first_frame) // 16 for n in range(first_frame, first_frame + num_frames)
              if n % (2 * stripe) < stripe or (n - 1) % (2 * stripe) < stripe} if num_frames <= 200_000 else set()
    assert needed <= set(selected[0])
