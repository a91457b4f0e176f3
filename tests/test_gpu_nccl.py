"""Needs >= 2 B200s (skipped otherwise): the data-parallel training step over real NCCL.  One process per GPU via
torch.distributed.run; see tests/helpers/nccl_worker.py for what is asserted."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_gradient_equals_mean_of_single_gpu_gradients():
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "helpers", "nccl_worker.py")]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = os.path.join(out_dir, "nccl_worker_pytest.log")
    with open(log, "w") as f:        # to a file, not a pipe: a hang still leaves the ranks' output (and their stack dumps) behind
        r = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=420, cwd=ROOT)
    text = open(log).read()
    assert r.returncode == 0 and "nccl_worker: OK" in text, text[-4000:]
