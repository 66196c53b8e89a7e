"""GPU, >= 2 devices (torchrun is spawned by the test): the sharded update must give, on every
rank, exactly the single-GPU result -- permutation, labels, dirty sets -- and the device plan must
equal the host mirror of the dealing rule."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_update_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "scripts", "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    last = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(last)
    assert out["ok"] and out["world"] == 2 and out["exchanged_bytes"] > 0
