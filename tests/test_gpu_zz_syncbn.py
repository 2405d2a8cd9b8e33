"""MinkowskiSyncBatchNorm (native kernels + all-reduced statistics) against torch BatchNorm1d
over the full batch — the DDP path of BASELINE configs[3] (reference:
MinkowskiNormalization.py:101-192, examples/multigpu_ddp.py:91-95).  Runs last (file name): the
workers are separate processes with their own NCCL process group; one rank always (exercises the
device-side count and the all-reduce plumbing), two ranks when the box has two GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    if world == 1:
        cmd = [sys.executable, os.path.join(HERE, "syncbn_worker.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", env["MASTER_PORT"], os.path.join(HERE, "syncbn_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert r.stdout.count("SYNCBN_OK") == world, r.stdout[-2000:]


def test_syncbn_one_rank(cuda):
    _run(1)


def test_syncbn_two_ranks(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2)
