"""The exchange step on a ONE-GPU box: two and three ranks, one process each, all on cuda:0.
CUDA IPC maps the mailboxes across the processes, so the fused peer-memory kernel (`exchange_merge_kernel`) and the
all-gather + `merge_hits_kernel` fallback both run and are compared with the oracle over the whole table
(tools/exchange_check.py).  The multi-GPU form of the same check is tests/test_gpu_shard.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_between_processes_sharing_one_gpu(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29640 + world), os.path.join(ROOT, "tools", "exchange_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "exchange_check OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
