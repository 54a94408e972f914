"""Warp-specialised kernel of the affine fast path (k = 4, exadg_b200/csrc/cart_ws.hpp) against the oracle and against the
pipelined kernel: relative l2 <= 1e-12.  The CPU emulation of the same kernel body is tests/test_ws_emulation.py."""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleOperator, synthetic_vector

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture
def ws_kernel():
    import exadg_b200
    previous = exadg_b200.cartesian_kernel(1)
    yield
    exadg_b200.cartesian_kernel(previous)


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


# 4^3 = 64 cells (3 batches, the last with 16 cells), 6^3 = 216 (9 batches), 5^3 = 125 (odd: no bulk copy for the last batch),
# 8^3 = 512, 12^3 = 1728 (72 batches), 16^3 = 4096 cells (171 batches: more than one batch per CTA only on small grids ...)
@pytest.mark.parametrize("grid", [(1, 2), (3, 1), (5, 0), (1, 3), (3, 2), (1, 4)])
def test_ws_kernel_matches_oracle(ws_kernel, grid):
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(4, grid[0], grid[1])
    assert op.is_cartesian_path == 1 and exadg_b200.cartesian_kernel() == 1
    ref = OracleOperator(4, grid[0], grid[1])
    x = synthetic_vector(ref.n_dofs, seed=11)
    src = torch.from_numpy(x).cuda()
    dst = op.initialize_dof_vector()
    op.vmult(dst, src)
    y_ref = ref.vmult_cellwise(x)
    assert rel(dst.cpu().numpy(), y_ref) < TOL
    op.vmult_add(dst, src)  # bulk add-reduction
    assert rel(dst.cpu().numpy(), 2 * y_ref) < TOL


def test_ws_kernel_persistent_ctas_equal_pipelined_kernel():
    """48^3 cells = 4608 batches on 296 CTAs: every CTA runs ~16 batches through the double-buffered trace area"""
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(4, 3, 4)
    src = torch.rand(op.local_size(), dtype=torch.float64, device="cuda") * 2 - 1
    y = [op.initialize_dof_vector(), op.initialize_dof_vector()]
    previous = exadg_b200.cartesian_kernel(-1)
    try:
        for v in (0, 1):
            exadg_b200.cartesian_kernel(v)
            for _ in range(3):  # repeated launches: no state may leak between them
                op.vmult(y[v], src)
    finally:
        exadg_b200.cartesian_kernel(previous)
    assert ((y[1] - y[0]).norm() / y[0].norm()).item() < TOL
