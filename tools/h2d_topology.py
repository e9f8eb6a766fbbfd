#!/usr/bin/env python
"""Host-to-device bandwidth against NUMA placement, N ranks at once (torchrun, one rank per GPU).

    python -m torch.distributed.run --nproc-per-node N tools/h2d_topology.py [--mb 1024]

Why: the end-to-end bench line is PCIe/host bound and its per-GPU H2D rate halves from 1 to 4-8 GPUs.
Every rank reports the NUMA node of its GPU (sysfs), the node its pinned pages landed on
(move_pages query) and its H2D rate, alone and with all ranks copying at once, for three placements of
the pinned buffer: wherever the allocating thread happens to run (default), after pinning the thread
to the GPU's node, and interleaved over all nodes.  Rank 0 prints one JSON object.
"""
from __future__ import annotations

import argparse
import ctypes
import glob
import json
import os
import subprocess

import torch
import torch.distributed as dist

LIBC = ctypes.CDLL(None, use_errno=True)
SYS_MOVE_PAGES, SYS_SET_MEMPOLICY = 279, 238
MPOL_DEFAULT, MPOL_BIND, MPOL_INTERLEAVE = 0, 2, 3


def read(path: str) -> str | None:
    try:
        with open(path, encoding="ascii") as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in (text or "").split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part.strip():
            cpus.add(int(part))
    return cpus


def nodes() -> list[int]:
    return sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))


def gpu_node(index: int) -> tuple[str, int | None]:
    props = torch.cuda.get_device_properties(index)
    bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    text = read(f"/sys/bus/pci/devices/{bus}/numa_node")
    return bus, (int(text) if text is not None else None)


def set_mempolicy(mode: int, node_list: list[int]) -> int:
    if not node_list:
        return LIBC.syscall(SYS_SET_MEMPOLICY, MPOL_DEFAULT, None, 0)
    mask = 0
    for n in node_list:
        mask |= 1 << n
    arr = (ctypes.c_ulong * 16)(*[(mask >> (64 * i)) & (2 ** 64 - 1) for i in range(16)])
    rc = LIBC.syscall(SYS_SET_MEMPOLICY, mode, arr, 16 * 64 + 1)
    return rc if rc == 0 else -ctypes.get_errno()


def pages_on_nodes(tensor: torch.Tensor, samples: int = 64) -> dict:
    size = tensor.numel() * tensor.element_size()
    step = max(4096, (size // samples) // 4096 * 4096)
    addrs = [tensor.data_ptr() // 4096 * 4096 + i * step for i in range(samples) if i * step < size]
    pages = (ctypes.c_void_p * len(addrs))(*addrs)
    status = (ctypes.c_int * len(addrs))()
    rc = LIBC.syscall(SYS_MOVE_PAGES, 0, len(addrs), pages, None, status, 0)
    if rc != 0:
        return {"error": -ctypes.get_errno()}
    out: dict = {}
    for s in status:
        out[str(s)] = out.get(str(s), 0) + 1
    return out


def h2d_rate(host: torch.Tensor, dev: torch.Tensor, reps: int) -> float:
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    return host.numel() * host.element_size() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    bus, node = gpu_node(local)
    all_cpus = sorted(os.sched_getaffinity(0))
    node_cpus = sorted(parse_cpulist(read(f"/sys/devices/system/node/node{node}/cpulist")) & set(all_cpus)) \
        if node is not None and node >= 0 else []
    dev = torch.empty(args.mb << 20, dtype=torch.uint8, device=device)
    report = {"rank": rank, "bus": bus, "gpu_node": node, "allowed_cpus": f"{all_cpus[0]}-{all_cpus[-1]} ({len(all_cpus)})",
              "node_cpus_allowed": len(node_cpus), "placements": {}}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for name in ("default", "local", "interleave"):
        note = ""
        if name == "local":
            if node_cpus:
                os.sched_setaffinity(0, node_cpus)
            rc = set_mempolicy(MPOL_BIND, [node]) if node is not None and node >= 0 else 0
            note = f"bind rc={rc}"
        elif name == "interleave":
            os.sched_setaffinity(0, all_cpus)
            rc = set_mempolicy(MPOL_INTERLEAVE, nodes())
            note = f"interleave rc={rc}"
        host = torch.empty(args.mb << 20, dtype=torch.uint8).pin_memory()
        host.fill_(1)
        set_mempolicy(MPOL_DEFAULT, [])
        entry = {"note": note, "pages": pages_on_nodes(host)}
        # alone: ranks take turns
        for turn in range(world):
            barrier()
            if turn == rank:
                entry["alone_gbs"] = round(h2d_rate(host, dev, args.reps), 1)
        barrier()
        entry["together_gbs"] = round(h2d_rate(host, dev, args.reps * 2), 1)
        barrier()
        report["placements"][name] = entry
        del host
    os.sched_setaffinity(0, all_cpus)
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, report)
    else:
        gathered = [report]
    if rank == 0:
        def run(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30, check=False).stdout
            except Exception as exc:  # pylint: disable=broad-except
                return f"{cmd}: {exc}"
        system = {
            "nodes": {n: {"cpulist": read(f"/sys/devices/system/node/node{n}/cpulist"),
                          "meminfo": (read(f"/sys/devices/system/node/node{n}/meminfo") or "").splitlines()[:2]}
                      for n in nodes()},
            "cpuset_cpus": read("/sys/fs/cgroup/cpuset.cpus.effective"),
            "cpuset_mems": read("/sys/fs/cgroup/cpuset.mems.effective"),
            "cpu_count": os.cpu_count(),
            "topo": run("nvidia-smi topo -m"),
            "lscpu": [l for l in run("lscpu").splitlines() if "NUMA" in l or "Model name" in l or "Socket" in l],
        }
        print(json.dumps({"world": world, "mb": args.mb, "system": system, "ranks": gathered}, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
