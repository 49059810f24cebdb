"""World > 1 on real GPUs: spawns min(2, device_count) ranks (one process per GPU, NCCL + CUDA IPC peer memory) and
runs tests/multirank_worker.py in them.  Skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multirank.py -m gpu` is the call that exercises it (log: profiles/r02_multirank_2gpu.txt)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_distributed_transform_on_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTIRANK OK" in r.stdout
