"""Warp-specialised kernel of the affine fast path (k = 4, exadg_b200/csrc/cart_ws.hpp) against the oracle and against the
pipelined kernel: relative l2 <= 1e-12.  The CPU emulation of the same kernel body is tests/test_ws_emulation.py."""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleOperator, synthetic_vector

pytestmark = pytest.mark.gpu
TOL = 1e-12


# every kernel variant of the k=4 affine fast path that bench.py may time (include/exadg_b200.h: exadg_b200_set_kernel_variant)
VARIANTS = [1, 2, 3, 4, 6]


@pytest.fixture(params=VARIANTS, ids=["ws_depth8", "ws_depth12", "ws_4producers", "warp_private", "ws_staged"])
def ws_kernel(request):
    return request.param


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


# 4^3 = 64 cells (3 batches, the last with 16 cells), 6^3 = 216 (9 batches), 5^3 = 125 (odd: no bulk copy for the last batch),
# 8^3 = 512, 12^3 = 1728 (72 batches), 16^3 = 4096 cells (171 batches: more than one batch per CTA only on small grids ...)
@pytest.mark.parametrize("grid", [(1, 2), (3, 1), (5, 0), (1, 3), (3, 2), (1, 4)])
def test_ws_kernel_matches_oracle(ws_kernel, grid):
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(4, grid[0], grid[1])
    op.set_kernel_variant(ws_kernel)  # per operator, not process-wide
    assert op.is_cartesian_path == 1 and op.get_kernel_variant() == ws_kernel
    ref = OracleOperator(4, grid[0], grid[1])
    x = synthetic_vector(ref.n_dofs, seed=11)
    src = torch.from_numpy(x).cuda()
    dst = op.initialize_dof_vector()
    op.vmult(dst, src)
    y_ref = ref.vmult_cellwise(x)
    assert rel(dst.cpu().numpy(), y_ref) < TOL
    op.vmult_add(dst, src)  # bulk add-reduction
    assert rel(dst.cpu().numpy(), 2 * y_ref) < TOL


def test_ws_kernel_persistent_ctas_equal_pipelined_kernel(ws_kernel):
    """48^3 cells = 4608 batches on 296 CTAs: every CTA runs ~16 batches through the trace area; two operators of one process with
    different kernels (the variant is a property of the operator)"""
    import exadg_b200
    ops = [exadg_b200.LaplaceOperator.hypercube(4, 3, 4), exadg_b200.LaplaceOperator.hypercube(4, 3, 4)]
    ops[0].set_kernel_variant(0)
    ops[1].set_kernel_variant(ws_kernel)
    src = torch.rand(ops[0].local_size(), dtype=torch.float64, device="cuda") * 2 - 1
    y = [op.initialize_dof_vector() for op in ops]
    for v in (0, 1):
        for _ in range(3):  # repeated launches: no state may leak between them
            ops[v].vmult(y[v], src)
    assert ((y[1] - y[0]).norm() / y[0].norm()).item() < TOL
    # size-independent properties at this size: A 1 = 0 on the periodic box, symmetry
    one = torch.ones_like(src)
    ops[1].vmult(y[1], one)
    assert y[1].abs().max().item() < 1e-9 * y[0].abs().max().item()
    v2 = torch.rand_like(src)
    av = ops[1].initialize_dof_vector()
    ops[1].vmult(av, v2)
    assert abs((torch.dot(y[0], v2) - torch.dot(src, av)).item()) < 1e-11 * abs(torch.dot(y[0], v2).item()) + 1e-6
