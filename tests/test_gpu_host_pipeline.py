"""exadg_b200_vmult_host_pipelined (upload, operator and download overlapped inside the call) must give the device vmult bit for bit in
both of its variants - chunk by chunk with a copy-engine download ("staged"), and piece-wise upload with the kernels storing dst
straight into the pinned host buffer ("direct") - on repeated calls and again after the buffers change.  It runs in a child process (a fault in the stream / graph choreography cannot poison the CUDA context of the other tests); its
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
    # both variants: "staged" (chunk plan, copy-engine download per chunk) and "direct" (piece-wise upload, the kernels store dst
    # straight into the pinned host tensor); then the automatic choice, which the remaining checks run with
    for mode in ("staged", "direct", "auto"):
        op.set_host_pipeline_mode(mode)
        for rep in range(3):   # plan and events are built by the first call and reused
            h_dst.fill_(float("nan"))
            op.vmult_host_pipelined(h_dst, h_src)
            assert (h_dst.cuda() - dst).abs().max().item() == 0.0, (degree, n_sub, refine, deformation, mode, rep)
    # direct mode needs a host buffer the GPU can address: pageable memory is refused there and takes the staged path in automatic mode
    p_dst = torch.full((op.local_size(),), float("nan"), dtype=torch.float64)
    op.vmult_host_pipelined(p_dst, h_src)
    assert (p_dst.cuda() - dst).abs().max().item() == 0.0
    op.set_host_pipeline_mode("direct")
    try:
        op.vmult_host_pipelined(p_dst, h_src)
        raise SystemExit("direct mode accepted pageable memory")
    except exadg_b200.ExaDGError:
        pass
    op.set_host_pipeline_mode("auto")
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
