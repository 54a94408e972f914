"""exadg_b200_vmult_host_pipelined (upload, operator and download overlapped chunk by chunk) must give the device vmult bit for bit.
The stream/event choreography was written after the GPU budget of round 1 was spent, so it runs in a child process (a fault there
cannot poison the CUDA context of the other tests) and is marked xfail(strict=False) until it has been seen to pass on hardware;
its host-side plan is covered on the CPU by tests/test_host_pipeline.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
import exadg_b200
torch.cuda.set_device(0)
# (degree, n_sub, refine, deformation): affine fast path with 24-, 32- and 16-cell batches (9 / 3 chunks), general path (cell lists)
for (degree, n_sub, refine, deformation) in [(4, 3, 3, 0.0), (4, 5, 2, 0.0), (3, 3, 3, 0.0), (5, 3, 2, 0.0), (2, 3, 2, 0.1), (3, 5, 1, 0.1)]:
    op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, 1, deformation, 2, (0,) * 6, 1.0)
    src = torch.rand(op.local_size(), dtype=torch.float64, device="cuda") * 2 - 1
    dst = op.initialize_dof_vector()
    op.vmult(dst, src)
    h_src = torch.empty(op.local_size(), dtype=torch.float64).pin_memory(); h_src.copy_(src.cpu())
    h_dst = torch.empty(op.local_size(), dtype=torch.float64).pin_memory()
    for rep in range(3):
        h_dst.fill_(float("nan"))
        op.vmult_host_pipelined(h_dst, h_src)
        assert (h_dst.cuda() - dst).abs().max().item() == 0.0, (degree, n_sub, refine, deformation, rep)
    del op
print("PIPELINED_OK")
""" % ROOT


@pytest.mark.xfail(strict=False, reason="stream choreography not yet seen on hardware (written after the round-1 GPU budget was spent)")
def test_pipelined_host_vmult_is_bitwise_the_device_vmult():
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=150)
    assert r.returncode == 0 and "PIPELINED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
