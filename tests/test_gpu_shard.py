"""Multi-GPU parity (needs >= 2 B200s on the box; skipped on a single-GPU box): launches
tools/shard_check.py with one process per GPU over NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_search_matches_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "shard_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "shard_check OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
