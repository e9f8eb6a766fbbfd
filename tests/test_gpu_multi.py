"""Multi-GPU parity (skipped below 2 GPUs): one process per GPU under torchrun runs
tools/check_sharded.py — sharded trajectories (host and HBM-resident), routed evaluation, the chirp-z
transform shared by the ranks, and the all-gather fallback, against the oracle and the single-GPU path.
Tolerances: 1e-10 on polarizabilities, 1e-8 on intensities (BASELINE.json north_star)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _run_check(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tools", "check_sharded.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=REPO)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"check_sharded_n{world}.log"), "w", encoding="utf-8") as log:
        log.write(res.stdout + "\n--- stderr ---\n" + res.stderr)
    assert res.returncode == 0, res.stdout[-6000:] + res.stderr[-6000:]
    # 3 small cases x (host | HBM-resident) x (shared transform | all-gather) + 2 large cases, per rank
    assert res.stdout.count("ok=True") == world * 14 and "ok=False" not in res.stdout, res.stdout[-6000:]
    # the shared transform really ran (not the all-gather fallback)
    assert res.stdout.count("shared=True (used=True)") == world * 6
    assert res.stdout.count("resident shared (used=True)") == 2 * world
    # ... and the aligned large case took the schedule that overlaps the first pack half with the evaluation
    assert res.stdout.count("overlapped=True") == world


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_spectrum_two_ranks():
    _run_check(2)


@pytest.mark.skipif(torch.cuda.device_count() < 3, reason="needs more than 2 GPUs")
def test_sharded_spectrum_all_ranks():
    _run_check(min(torch.cuda.device_count(), 8))
