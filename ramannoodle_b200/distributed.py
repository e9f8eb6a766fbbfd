"""Frame sharding across GPUs (one process per GPU, ``torch.distributed``).

``calc_polarizabilities`` has no cross-frame term, so frames shard naturally: rank r
evaluates the contiguous block ``shard_bounds(S, G, r)`` with the model tables replicated.
The only exchange step is one all-gather of the per-rank (S_r,3,3) blocks (72 B/frame), after
which every rank holds the full series that ``MDRamanSpectrum`` owns
(``ramannoodle/spectrum/_raman.py:212-216``); ``np.diff`` across shard boundaries needs no
halo because it runs after the gather.
"""
from __future__ import annotations

from .abstract import Dynamics
from .dynamics import Trajectory
from .spectrum import MDRamanSpectrum


def shard_bounds(num_frames: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of rank ``rank``: blocks of ceil(S/G) frames."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("invalid rank/world_size")
    block = -(-num_frames // world_size)
    start = min(rank * block, num_frames)
    return start, min(start + block, num_frames)


def allgather_series(local_series, num_frames: int, group=None):
    """All-gather per-rank (S_r,3,3) blocks into the full (S,3,3) series on every rank.

    One collective: blocks are padded to ceil(S/G) frames so a single
    ``all_gather_into_tensor`` (NCCL over NVLink on GPUs, gloo in CPU tests) suffices; the
    padding of the tail rank is dropped afterwards.
    """
    import torch  # pylint: disable=import-outside-toplevel
    import torch.distributed as dist  # pylint: disable=import-outside-toplevel

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    block = -(-num_frames // world)
    start, stop = shard_bounds(num_frames, world, rank)
    if tuple(local_series.shape) != (stop - start, 3, 3):
        raise ValueError(f"rank {rank}: local series has shape {tuple(local_series.shape)}, "
                         f"expected {(stop - start, 3, 3)}")
    if world == 1:
        return local_series
    padded = local_series
    if stop - start != block:
        padded = torch.zeros((block, 3, 3), dtype=local_series.dtype, device=local_series.device)
        padded[: stop - start] = local_series
    full = torch.empty((world * block, 3, 3), dtype=local_series.dtype, device=local_series.device)
    dist.all_gather_into_tensor(full, padded.contiguous(), group=group)
    return full[:num_frames]


class ShardedTrajectory(Dynamics):
    """The local frame block of a trajectory that is sharded over the ranks of a process group.

    Parameters
    ----------
    local_positions_ts
        (fractional) (S_r,N,3) block of this rank, frames ``shard_bounds(num_frames, G, r)``;
        numpy (host) or CUDA tensor (HBM-resident).
    timestep
        (fs)
    num_frames
        Total number of frames S over all ranks.
    """

    def __init__(self, local_positions_ts, timestep: float, num_frames: int, group=None) -> None:
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel

        self._local = Trajectory(local_positions_ts, timestep)
        self._num_frames = int(num_frames)
        self._group = group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        start, stop = shard_bounds(self._num_frames, world, rank)
        if len(self._local) != stop - start:
            raise ValueError(f"rank {rank} holds {len(self._local)} frames, expected {stop - start}")

    @property
    def local(self) -> Trajectory:
        return self._local

    @property
    def num_frames(self) -> int:
        return self._num_frames

    def get_raman_spectrum(self, polarizability_model) -> MDRamanSpectrum:
        """Evaluate the local block, all-gather the series, return the full-series spectrum."""
        local = self._local.get_raman_spectrum(polarizability_model)
        series = local._polarizability_ts  # pylint: disable=protected-access
        if not hasattr(series, "data_ptr"):
            import torch  # pylint: disable=import-outside-toplevel

            series = torch.from_numpy(series)
        full = allgather_series(series, self._num_frames, self._group)
        return MDRamanSpectrum(full, self._local.timestep)
