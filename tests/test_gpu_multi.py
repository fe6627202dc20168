"""On-hardware multi-GPU equivalence (skipped below 2 GPUs): the frame-sharded, exchanged step equals the single-GPU step on
the same global batch, for the NCCL exchange and for the peer-memory exchange kernel.  The host-side sharding / exchange
arithmetic is covered on CPU by tests/test_dist_cpu.py (gloo, world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_step_equals_single_gpu_step():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(HERE, "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
