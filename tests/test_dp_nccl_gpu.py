"""GPU (>= 2 devices): data parallel over NCCL -- tools/dp_check.py under torchrun: DP gradients (overlapped bucketed
all-reduce) against the one-GPU full-batch gradients, and the CUDA-graph step with the all-reduce captured inside
against the eager DP path (reference trainer.py:1185 wraps the model in DDP).  Skipped on a one-GPU box; the host-side
bucket logic is covered on CPU by tests/test_dp_gloo.py."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_nccl_two_gpus():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "tools/dp_check.py"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DP OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
