"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): clip-range shards + exchange + merge
must reproduce a single-GPU scan of the whole search set bit for bit, for both exchange implementations."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_scan_equals_single_gpu_scan():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "exchange p2p:" in r.stdout and "exchange p2p-lagged:" in r.stdout and "exchange nccl:" in r.stdout
    assert "rank store scan" in r.stdout and "rank store batched scan" in r.stdout and "rank store labelled sims" in r.stdout
    assert "rank store selection scan" in r.stdout and "host mailbox and NCCL small exchanges agree: True" in r.stdout
