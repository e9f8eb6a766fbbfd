"""Frame sharding across GPUs (one process per GPU, ``torch.distributed``).

``calc_polarizabilities`` has no cross-frame term, so frames shard naturally: rank r
evaluates the contiguous block ``shard_bounds(S, G, r)`` with the model tables replicated.
The only exchange step is one all-gather of the per-rank (S_r,3,3) blocks (72 B/frame), after
which every rank holds the full series that ``MDRamanSpectrum`` owns
(``ramannoodle/spectrum/_raman.py:212-216``); ``np.diff`` across shard boundaries needs no
halo because it runs after the gather.
"""
from __future__ import annotations

import ctypes
import os

from . import _lib
from .abstract import Dynamics
from .dynamics import Trajectory
from .exceptions import get_type_error
from .spectrum import MDRamanSpectrum, _get_plan, _stream


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


_SYMMETRIC_SERIES: dict = {}


def symmetric_series(num_frames: int, device, group=None):
    """A (S,3,3) fp64 series buffer in symmetric memory (the same allocation on every rank of
    ``group``, each rank's copy mapped into every other rank's address space over NVLink), plus its
    rendezvous handle.  Cached per (S, device, group): the rendezvous is a collective."""
    import torch  # pylint: disable=import-outside-toplevel
    import torch.distributed as dist  # pylint: disable=import-outside-toplevel
    import torch.distributed._symmetric_memory as symm_mem  # pylint: disable=import-outside-toplevel

    pg = group if group is not None else dist.group.WORLD
    key = (int(num_frames), str(device), pg.group_name)
    entry = _SYMMETRIC_SERIES.get(key)
    if entry is None:
        tensor = symm_mem.empty((int(num_frames), 3, 3), dtype=torch.float64, device=device)
        handle = symm_mem.rendezvous(tensor, group=pg.group_name)
        entry = (tensor, handle)
        _SYMMETRIC_SERIES.clear()  # one live buffer: they are as large as the series
        _SYMMETRIC_SERIES[key] = entry
    return entry


def spectrum_parts(world_size: int, rank: int) -> list[int]:
    """Which of the three packed transforms of ``measure`` rank ``rank`` computes
    (``include/ramannoodle_b200.h: rn_md_spectrum_part``): round-robin over the ranks."""
    return [part for part in range(3) if part % world_size == rank]


def spectrum_half_units(world_size: int, rank: int):
    """Split-transform schedule of ``measure`` (``rn_md_spectrum_half``): each of the three packed
    transforms is two half-length transforms (output residues 0/1) run by a pair of ranks that read
    each other's result.  Returns ``(units, partner)`` with ``units`` the ``(part, residue)`` pairs
    of this rank (slot order) and ``partner`` the rank holding the other residue of every unit, or
    ``None`` when the world size does not profit (then whole parts are dealt out, ``spectrum_parts``).

    2 ranks: rank r runs residue r of all three parts (3 half transforms each instead of 2 + 1 full
    ones).  6 or more ranks: ranks 2p and 2p+1 run the residues of part p (one half transform each
    instead of one full transform on three ranks); further ranks only take part in the barriers."""
    if not 0 <= rank < world_size:
        raise ValueError("invalid rank/world_size")
    if world_size == 2:
        return [(part, rank) for part in range(3)], 1 - rank
    if world_size >= 6:
        if rank < 6:
            return [(rank // 2, rank % 2)], rank ^ 1
        return [], None
    return None


_SYMMETRIC_HALVES: dict = {}


def symmetric_halves(slots: int, half_length: int, device, group=None):
    """``(slots, L/2, 2)`` fp64 buffer in symmetric memory for the half-transform results, plus its
    rendezvous handle (cached: the rendezvous is a collective)."""
    import torch  # pylint: disable=import-outside-toplevel
    import torch.distributed as dist  # pylint: disable=import-outside-toplevel
    import torch.distributed._symmetric_memory as symm_mem  # pylint: disable=import-outside-toplevel

    pg = group if group is not None else dist.group.WORLD
    key = (int(slots), int(half_length), str(device), pg.group_name)
    entry = _SYMMETRIC_HALVES.get(key)
    if entry is None:
        tensor = symm_mem.empty((int(slots), int(half_length), 2), dtype=torch.float64, device=device)
        handle = symm_mem.rendezvous(tensor, group=pg.group_name)
        entry = (tensor, handle)
        _SYMMETRIC_HALVES.clear()
        _SYMMETRIC_HALVES[key] = entry
    return entry


class ShardedMDRamanSpectrum(MDRamanSpectrum):
    """``MDRamanSpectrum`` whose ``measure`` is spread over the ranks of a process group.

    Every rank holds the full (S,3,3) series (after the all-gather).  The orientational average
    45 a^2 + 7 g^2 (``ramannoodle/spectrum/_raman.py:286-297``) is a sum of three independent
    packed chirp-z transforms; rank r computes parts ``spectrum_parts(G, r)``, the (P,) partial
    intensities are summed with one all-reduce, and every rank applies the wavenumber grid and
    the optional corrections.  With G >= 3 the spectrum stage costs one transform instead of three.
    """

    def __init__(self, polarizability_ts, timestep: float, group=None, split_transforms: bool = True):
        super().__init__(polarizability_ts, timestep)
        self._group = group
        # RN_SPLIT_TRANSFORMS=0: A/B switch back to whole transforms per rank
        self._split_transforms = bool(split_transforms) and os.environ.get("RN_SPLIT_TRANSFORMS", "1") != "0"

    def _measure_split(self, series, plan, total, world, rank, device) -> bool:
        """Half-transform schedule (``spectrum_half_units``); False if it does not apply here."""
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel

        schedule = spectrum_half_units(world, rank) if self._split_transforms else None
        if schedule is None or dist.get_backend(self._group) != "nccl":
            return False
        units, partner = schedule
        half_length = int(_lib.lib().rn_spectrum_half_length(plan.handle))
        slots = 3 if world == 2 else 1
        try:
            zbuf, handle = symmetric_halves(slots, half_length, series.device, self._group)
        except Exception:  # pylint: disable=broad-except  (no symmetric-memory support on this system)
            return False
        stream = _stream(device)
        handle.barrier()  # the partner is done reading the previous contents
        for slot, (part, residue) in enumerate(units):
            status = _lib.lib().rn_md_spectrum_half(plan.handle, ctypes.c_void_p(series.data_ptr()), part, residue,
                                                    ctypes.c_void_p(zbuf[slot].data_ptr()), 1 if slot > 0 else 0, stream)
            _lib.check(status, "rn_md_spectrum_half")
        handle.barrier()  # both residues of every unit are complete
        for slot, (part, residue) in enumerate(units):
            own = int(zbuf[slot].data_ptr())
            other = int(handle.buffer_ptrs[partner]) + slot * half_length * 16
            res0, res1 = (own, other) if residue == 0 else (other, own)
            status = _lib.lib().rn_md_spectrum_half_combine(plan.handle, part, residue, ctypes.c_void_p(res0),
                                                            ctypes.c_void_p(res1), ctypes.c_void_p(total.data_ptr()),
                                                            1 if slot > 0 else 0, stream)
            _lib.check(status, "rn_md_spectrum_half_combine")
        return True

    # pylint: disable=too-many-arguments,too-many-positional-arguments,too-many-locals
    def measure_device(self, orientation="polycrystalline", laser_correction=False, laser_wavelength=522,
                       bose_einstein_correction=False, temperature=300):
        import torch  # pylint: disable=import-outside-toplevel
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel

        if orientation != "polycrystalline":
            raise NotImplementedError("only polycrystalline spectra are supported for now")
        if laser_correction:
            laser_wavenumber = 10000000 / laser_wavelength
            try:
                if laser_wavenumber <= 0:
                    raise ValueError(f"invalid laser_wavenumber: {laser_wavenumber} <= 0")
            except TypeError as exc:
                raise get_type_error("laser_wavenumber", laser_wavenumber, "float") from exc
        if bose_einstein_correction:
            try:
                if temperature <= 0:
                    raise ValueError(f"invalid temperature: {temperature} <= 0")
            except TypeError as exc:
                raise get_type_error("temperature", temperature, "float") from exc
        series = self._device_series()
        num_frames = int(series.shape[0])
        if num_frames < 2:
            raise ValueError("polarizability_ts must contain at least 2 configurations")
        device = int(series.device.index or 0)
        world, rank = dist.get_world_size(self._group), dist.get_rank(self._group)
        points = int(_lib.lib().rn_spectrum_num_points(num_frames))
        with torch.cuda.device(device):
            # element `points` of the reduced vector carries the sharded series-energy constant
            total = torch.zeros(points + 1, dtype=torch.float64, device=series.device)
            wavenumbers = torch.empty(points, dtype=torch.float64, device=series.device)
            intensities = torch.empty(points, dtype=torch.float64, device=series.device)
            if points > 0:
                plan = _get_plan(num_frames, device)
                lib = _lib.lib()
                # the energies (one pass over the whole series) shard over ranks: every rank sums the
                # difference signals of its block and the constant rides along with the all-reduce
                shard_energy = world > 1
                if shard_energy:
                    _lib.check(lib.rn_spectrum_set_energy_mode(plan.handle, 1), "rn_spectrum_set_energy_mode")
                try:
                    split = world > 1 and self._measure_split(series, plan, total, world, rank, device)
                    parts = [] if split else spectrum_parts(world, rank)
                    if parts:
                        partial = torch.empty(points, dtype=torch.float64, device=series.device)
                        for part in parts:
                            status = lib.rn_md_spectrum_part(plan.handle, ctypes.c_void_p(series.data_ptr()), part,
                                                             ctypes.c_void_p(partial.data_ptr()), _stream(device))
                            _lib.check(status, "rn_md_spectrum_part")
                            total[:points] += partial
                    if shard_energy:
                        begin, end = shard_bounds(num_frames - 1, world, rank)
                        status = lib.rn_series_energy_constant(
                            plan.handle, ctypes.c_void_p(series.data_ptr()), begin, end,
                            ctypes.c_void_p(total.data_ptr() + 8 * points), _stream(device))
                        _lib.check(status, "rn_series_energy_constant")
                finally:
                    if shard_energy:
                        lib.rn_spectrum_set_energy_mode(plan.handle, 0)
                if world > 1:
                    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self._group)
                    total[:points] += total[points]
                status = lib.rn_md_spectrum_finish(
                    num_frames, ctypes.c_void_p(total.data_ptr()), float(self._timestep),
                    1 if laser_correction else 0, float(laser_wavelength) if laser_correction else 0.0,
                    1 if bose_einstein_correction else 0, float(temperature) if bose_einstein_correction else 0.0,
                    ctypes.c_void_p(wavenumbers.data_ptr()), ctypes.c_void_p(intensities.data_ptr()), _stream(device))
                _lib.check(status, "rn_md_spectrum_finish")
        return wavenumbers, intensities


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

    def get_raman_spectrum(self, polarizability_model, fused: bool = True,
                           reuse_series_buffer: bool = False) -> MDRamanSpectrum:
        """Evaluate the local block and assemble the full (S,3,3) series on every rank; returns a
        ``ShardedMDRamanSpectrum`` (its ``measure`` is spread over the ranks too).

        ``fused=True`` (default, CUDA + NCCL groups with this package's models): the series lives
        in symmetric memory and the evaluation kernels store every row to all ranks' copies over
        NVLink themselves, so the all-gather overlaps the evaluation; two device-side barriers
        bracket the stores.  Otherwise (or if symmetric memory is unavailable) the local block is
        evaluated first and one ``all_gather_into_tensor`` assembles the series.
        ``reuse_series_buffer=True`` hands out the symmetric buffer itself (overwritten by the next
        call) instead of a copy."""
        if fused:
            spectrum = self._get_raman_spectrum_fused(polarizability_model, reuse_series_buffer)
            if spectrum is not None:
                return spectrum
        local = self._local.get_raman_spectrum(polarizability_model)
        series = local._polarizability_ts  # pylint: disable=protected-access
        if not hasattr(series, "data_ptr"):
            import torch  # pylint: disable=import-outside-toplevel

            series = torch.from_numpy(series)
        full = allgather_series(series, self._num_frames, self._group)
        return ShardedMDRamanSpectrum(full, self._local.timestep, self._group)

    def _get_raman_spectrum_fused(self, polarizability_model, reuse_series_buffer: bool):
        import torch  # pylint: disable=import-outside-toplevel
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel

        multi = getattr(polarizability_model, "calc_polarizabilities_multi", None)
        if multi is None or not torch.cuda.is_available() or dist.get_backend(self._group) != "nccl":
            return None
        world, rank = dist.get_world_size(self._group), dist.get_rank(self._group)
        if world == 1 or world > 8:
            return None
        positions = self._local._positions_ts  # pylint: disable=protected-access
        device = positions.device if hasattr(positions, "data_ptr") else torch.device("cuda", torch.cuda.current_device())
        try:
            series, handle = symmetric_series(self._num_frames, device, self._group)
        except Exception:  # pylint: disable=broad-except  (no symmetric-memory support on this system)
            return None
        start, _ = shard_bounds(self._num_frames, world, rank)
        order = [rank] + [r for r in range(world) if r != rank]  # local copy first
        ptrs = [int(handle.buffer_ptrs[r]) + start * 72 for r in order]
        dup = int(os.environ.get("RN_DEBUG_DUP_PEERS", "0"))  # timing aid: emulate more peers on a 2-GPU box
        while dup and len(ptrs) < min(dup + 1, 8):
            ptrs.append(ptrs[1])
        handle.barrier()  # every rank is done with the previous contents of the buffers
        try:
            multi(positions, ptrs)
        except ValueError as exc:
            raise ValueError("polarizability_model and trajectory are incompatible") from exc
        handle.barrier()  # every rank's rows have landed everywhere
        full = series if reuse_series_buffer else series.clone()
        return ShardedMDRamanSpectrum(full, self._local.timestep, self._group)
