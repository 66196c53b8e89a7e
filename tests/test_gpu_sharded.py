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


def test_row_block_update_over_nccl_equals_single_gpu():
    """one mesh over 2 processes / GPUs, the library's own NCCL communicator (scripts/rowblock_check.py)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); tests/test_gpu_rowblock.py covers the same logic on one")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "rowblock_check.py"), "24", "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["world"] == 2
