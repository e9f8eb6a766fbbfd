#!/usr/bin/env python
"""Benchmark of the MD-Raman hot path (BASELINE.json metric: MD frames/s, polarizability +
spectrum, at 1/2/4/8 B200 vs the reference CPU path; % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|c1|c3dense]

A step = one pass of the hot path over one synthetic trajectory block that is already
resident in HBM: ``calc_polarizabilities`` over every frame of the rank's block, (N>1) one
NCCL all-gather of the (S,3,3) series, ``MDRamanSpectrum.measure`` on the full series.
N>1 is launched with torchrun (one process per GPU); frames are sharded with no data-path
collective other than that all-gather ("weak": frames per GPU fixed).  Prints ONE JSON line.

Default workload "c3" = BASELINE.json configs[2], the config north_star's target is quoted
on: ARTModel of 192-atom LLZO, 1M-frame synthetic trajectory per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (structure, kind, frames per GPU, description)
    "c1": ("TiO2", "art", 10_000, "ARTModel rutile TiO2 (108 atoms, 324 DOFs), 10k-frame synthetic trajectory"),
    "c2": ("STO", "cubic", 100_000, "InterpolationModel cubic BSpline, SrTiO3 (135 atoms, 405 DOFs), 100k frames"),
    "c3": ("LLZO", "art", 1_000_000, "ARTModel LLZO (192 atoms, 576 DOFs), 1M-frame synthetic trajectory per GPU"),
    "c5": ("LLZO_2x2x2", "cubic4600", 37_888,
           "InterpolationModel (cubic BSpline) of the 1536-atom LLZO 2x2x2 supercell, 4600 DOFs, 37,888 frames per GPU"),
    "c3dense": ("LLZO", "art", 200_000, "LLZO ARTModel forced through the dense DMMA projection, 200k frames per GPU"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (override)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=8192, help="frames in the CPU-baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md "clocks line")
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x0000000000000004: "sw_power_cap", 0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown", 0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # pylint: disable=broad-except
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # pylint: disable=broad-except
                pass
            self._stop.wait(0.002)

    def start(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks() -> tuple[float, str]:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path, encoding="utf-8") as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak() -> tuple[float, str]:
    path = os.path.join(REPO, "profiles", "fp64_peaks.json")
    if os.path.exists(path):
        with open(path, encoding="utf-8") as fh:
            return float(json.load(fh)["dmma_m8n8k4_tflops"]), "measured (profiles/fp64_peaks.json, DMMA m8n8k4)"
    return 37.0, "fallback (tools/fp64_peaks.cu, earlier run)"


def ncu_traffic(kernel: str, workload: str):
    path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path, encoding="utf-8") as fh:
            return json.load(fh).get(f"{kernel}:{workload}")
    return None


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle numpy port on the host cores
# --------------------------------------------------------------------------------------
def _cpu_chunk(args):
    structure, kind, first, count = args
    from oracle import numpy_port as ora
    from ramannoodle_b200 import synthetic

    state = (synthetic.make_model(structure, "cubic", num_dofs=4600) if kind == "cubic4600"
             else synthetic.make_model(structure, kind))
    omodel = ora.OracleModel(ref_positions=state.ref_positions, lattice=state.lattice,
                             ref_polarizability=state.ref_polarizability, basis_vectors=list(state.basis_vectors),
                             splines=list(state.splines), mask=state.mask)
    positions = synthetic.make_trajectory(structure, count, seed=1000, first_frame=first)
    t0 = time.perf_counter()
    alpha = ora.calc_polarizabilities(omodel, positions)
    return alpha, time.perf_counter() - t0


def cpu_path_rate(structure, kind, sample, processes):
    """frames/s of the oracle port (= the reference's numpy/scipy statements) on `sample` frames:
    calc_polarizabilities split over `processes` worker processes, then md_measure on the series."""
    from oracle import numpy_port as ora

    per = max(1, sample // processes)
    jobs = [(structure, kind, i * per, per) for i in range(processes)]
    if processes == 1:
        results = [_cpu_chunk(jobs[0])]
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(processes) as pool:
            results = pool.map(_cpu_chunk, jobs)
    alpha = np.concatenate([r[0] for r in results])
    # workers run concurrently: the stage takes as long as the slowest worker's evaluation
    # (model construction and trajectory synthesis are outside the timed calls)
    t_poly = max(r[1] for r in results)
    t1 = time.perf_counter()
    ora.md_measure(alpha, 1.0)
    t_meas = time.perf_counter() - t1
    frames = per * processes
    return frames / (t_poly + t_meas), {"frames": frames, "t_polarizability_s": round(t_poly, 3),
                                       "t_measure_s": round(t_meas, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    structure, kind, frames, desc = WORKLOADS[args.workload]
    kind = "art" if args.workload == "c3dense" else kind
    cores = os.cpu_count() or 1
    sample = max(cores * 256, min(args.cpu_sample * 2, cores * 1024))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_path_rate(structure, kind, cores * 64, cores)
    rates, detail = [], None
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 5))):
        rate, detail = cpu_path_rate(structure, kind, sample, cores)
        rates.append(rate)
        if time.perf_counter() - t0 > 120:
            break
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "md_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(rates), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * detail["frames"] / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sample_frames_per_step": detail["frames"]},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{detail['frames']} frames/step of the same synthetic workload: oracle numpy/scipy port "
                                   f"(statement-for-statement restatement of the pure-Python reference) — "
                                   f"calc_polarizabilities frame-chunked over {cores} processes + md_measure; {detail}"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import ramannoodle_b200 as rb
    from ramannoodle_b200 import _lib, synthetic
    from ramannoodle_b200.distributed import ShardedMDRamanSpectrum, ShardedTrajectory, allgather_series

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL prints its version banner to stdout when NCCL_DEBUG is VERSION/INFO; keep stdout = one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("RN_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.require_device(local_rank)

    structure, kind, frames, desc = WORKLOADS[args.workload]
    if args.frames > 0:
        frames = args.frames
    force_dense = args.workload == "c3dense"
    if kind == "cubic4600":
        state = synthetic.make_model(structure, "cubic", num_dofs=4600)
    else:
        state = synthetic.make_model(structure, kind)
    model = (rb.ARTModel if kind == "art" else rb.InterpolationModel)(state, device=local_rank, force_dense=force_dense)
    num_atoms = state.num_atoms
    total_frames = frames * world
    positions = synthetic.make_trajectory_cuda(structure, frames, device, seed=1000 + rank, first_frame=rank * frames)
    trajectory = rb.Trajectory(positions, 1.0)  # HBM-resident (wrap runs on the device)
    del positions
    info = model.path_info()
    # N > 1: the public multi-GPU API.  Frames are sharded; the evaluation kernels store every row
    # of the (S,3,3) series to all ranks' symmetric-memory copies over NVLink (fused all-gather;
    # falls back to one NCCL all-gather), and measure() is spread over the ranks.
    sharded = ShardedTrajectory(trajectory._positions_ts, 1.0, total_frames) if world > 1 else None  # pylint: disable=protected-access
    fused_gather = None
    split_transforms = None

    def step():
        if world > 1:
            return sharded.get_raman_spectrum(model).measure_device()
        return trajectory.get_raman_spectrum(model).measure_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        wn, inten = step()
    barrier()

    # ---- timed region: K steps, CUDA events, max over ranks ----
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    sampler.start()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        wn, inten = step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms, float(launches)], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        elapsed_ms, launches = float(tmax[0]), int(t[1])
    ms_per_step = elapsed_ms / args.steps
    value = total_frames / (ms_per_step * 1e-3)

    # ---- per-stage device times (same stream, CUDA events) for the roofline ----
    reps = max(5, min(args.steps, 20))
    stage = {}
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = np.zeros(3)
    for _ in range(reps):
        # local evaluation alone (the roofline numerator), then the sharded path, then measure.
        # A primer launch of the same evaluation runs first: the timed launches are enqueued while it
        # executes, so the events bracket kernel time, not the host's launch latency.
        trajectory.get_raman_spectrum(model)
        evs[0].record()
        spectrum = trajectory.get_raman_spectrum(model)
        evs[1].record()
        if world > 1:
            spectrum = sharded.get_raman_spectrum(model)
        evs[2].record()
        spectrum.measure_device()
        evs[3].record()
        torch.cuda.synchronize()
        acc += [evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])]
    stage = dict(zip(["polarizability_ms", "polarizability_plus_gather_ms", "spectrum_ms"], (acc / reps).round(4).tolist()))
    if world > 1:
        from ramannoodle_b200 import distributed as rdist
        fused_gather = bool(rdist._SYMMETRIC_SERIES)  # pylint: disable=protected-access
        split_transforms = bool(rdist._SYMMETRIC_HALVES)  # pylint: disable=protected-access

    hbm_peak, hbm_src = measured_peaks()
    if info["dense_dofs"] == 0:
        alg_bytes = (24 * num_atoms + 72) * frames  # read positions once, write alpha once
        achieved = alg_bytes / (stage["polarizability_ms"] * 1e-3) / 1e9
        kernel_name = "affine_tma_kernel" if info["tma_affine"] else "affine_generic_kernel"
        roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "peak_source": hbm_src,
                    "algorithmic_bytes_per_frame": 24 * num_atoms + 72,
                    "traffic": ncu_traffic(kernel_name, args.workload)}
    else:
        peak_tf, peak_src = fp64_peak()
        flops = 2.0 * 3 * num_atoms * info["dense_dofs"] * frames
        achieved = flops / (stage["polarizability_ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "dense_kernel", "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf, "peak_source": peak_src + " (FP64 tensor pipe)",
                    "algorithmic_flops_per_frame": 2 * 3 * num_atoms * info["dense_dofs"],
                    "traffic": ncu_traffic("dense_kernel", args.workload)}

    # ---- end to end through the public API with host buffers (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host_traj = rb.Trajectory(trajectory.positions_ts, 1.0)  # pinned host copy (wrap is idempotent)
        assert not host_traj.is_device_resident

        host_sharded = ShardedTrajectory(host_traj._positions_ts, 1.0, total_frames) if world > 1 else None  # pylint: disable=protected-access

        def e2e_step():
            # chunked H2D overlapped with evaluation (and, N > 1, with the fused all-gather)
            spectrum = (host_sharded if world > 1 else host_traj).get_raman_spectrum(model)
            return spectrum.measure()  # numpy results: D2H of the spectrum

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            wn_h, inten_h = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_frames / float(tt[0]), "unit": "frames/s",
               "h2d_bytes_per_step": int(frames * num_atoms * 24 * world),
               "d2h_bytes_per_step": int((wn_h.nbytes + inten_h.nbytes) * world),
               "ms_per_step": float(tt[0]) * 1e3, "steps": args.e2e_steps,
               "api": "Trajectory(host pinned).get_raman_spectrum(model).measure() -> numpy"}
        del host_traj

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, detail = cpu_path_rate(structure, kind, args.cpu_sample, 1)
        cpu_baseline = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
                        "sample": f"{detail['frames']} frames of the same synthetic workload through the oracle "
                                  f"numpy/scipy port (single process, as the reference ships): {detail}"}

    if rank == 0:
        line = {
            "metric": "md_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "frames_per_gpu": frames, "total_frames": total_frames,
                       "atoms": num_atoms, "dofs": state.num_dofs, "path": info,
                       "l2": "inputs larger than L2 (no flush needed)" if frames * num_atoms * 24 > 2 * 126e6
                             else "inputs smaller than L2: cache-resident between steps",
                       "stages": stage, "fused_allgather": fused_gather,
                       "split_transforms": split_transforms},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
