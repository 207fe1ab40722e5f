"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): tests/multi_gpu_check.py under torchrun, world size 2."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_ranks_match_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900)
    print(run.stdout[-3000:], run.stderr[-3000:])
    assert run.returncode == 0
    fused = "inside the cost kernel: True" in run.stdout
    assert run.stdout.count("-> ok") == (12 if fused else 6)   # 3 cases x 2 ranks, fused and NCCL reductions
