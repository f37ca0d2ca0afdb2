"""-m gpu: the multi-rank pass equals the single-GPU pass and - with injected draws - the oracle (needs >= 2 GPUs on the box;
skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_reproduce_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_GPU_CHECK_OK" in out.stdout and "MULTI_GPU_ORACLE_OK" in out.stdout, \
        out.stdout[-3000:] + out.stderr[-3000:]
