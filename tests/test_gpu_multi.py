"""Partitioned vmult / CG on >= 2 GPUs of one box (NCCL and NVLink peer-memory halo) against the CPU oracle: spawns
tests/multi_gpu_check.py under torchrun with 2 ranks.  Skipped on boxes with a single GPU (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{}, {"EXADG_B200_CART_KERNEL": "pipe"}], ids=["default_kernels", "pipelined_kernel"])
def test_two_rank_vmult_and_cg_match_oracle(env):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29650 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env), cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
