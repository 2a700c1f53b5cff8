"""The N > 1 path on real GPUs: one process per GPU (torchrun), one workload sharded by contiguous window ranges, records
gathered on rank 0 over NCCL through the C ABI (lb2_comm_gather) -- identical to the reference's records for the whole
workload.  Needs at least two GPUs (skipped otherwise; the host logic alone is covered on CPU by test_shard_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600, method="thread")
def test_sharded_workload_gathered_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "multirank_check.py"), "40000"], capture_output=True, text=True, timeout=580)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["same"] and out["records"] > 30 and out["failed"] == 0 and out["world"] == world
