"""exadg_b200_vmult_host_pipelined (upload, operator and download overlapped chunk by chunk) must give the device vmult bit for bit -
eagerly on the first call with a pair of host buffers, as one CUDA graph launch from the second call on, and again after the buffers
change.  It runs in a child process (a fault in the stream / graph choreography cannot poison the CUDA context of the other tests); its
host-side plan is covered on the CPU by tests/test_host_pipeline.py."""
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
    for rep in range(4):   # eager, capture + graph launch, graph launch, graph launch
        h_dst.fill_(float("nan"))
        op.vmult_host_pipelined(h_dst, h_src)
        assert (h_dst.cuda() - dst).abs().max().item() == 0.0, (degree, n_sub, refine, deformation, rep)
    # another pair of buffers and another vector: the graph of the first pair must not be reused
    src2 = torch.rand(op.local_size(), dtype=torch.float64, device="cuda") * 2 - 1
    op.vmult(dst, src2)
    h_src2 = torch.empty(op.local_size(), dtype=torch.float64).pin_memory(); h_src2.copy_(src2.cpu())
    h_dst2 = torch.empty(op.local_size(), dtype=torch.float64).pin_memory()
    for rep in range(3):
        h_dst2.fill_(float("nan"))
        op.vmult_host_pipelined(h_dst2, h_src2)
        assert (h_dst2.cuda() - dst).abs().max().item() == 0.0, (degree, n_sub, refine, deformation, "second pair", rep)
    # same buffers, new contents: the graph copies from the host memory at launch time
    h_src2.copy_(src.cpu())
    op.vmult(dst, src)
    op.vmult_host_pipelined(h_dst2, h_src2)
    assert (h_dst2.cuda() - dst).abs().max().item() == 0.0
    del op
print("PIPELINED_OK")
""" % ROOT


def test_pipelined_host_vmult_is_bitwise_the_device_vmult():
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PIPELINED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
