"""Matrix-free oracle vs an independent dense assembly of the SIPG bilinear form, plus the
self-checks of SURVEY Appendix A.6 (symmetry, null space, positivity, diagonal)."""
import numpy as np
import pytest

from oracle import dense_sipg
from oracle.oracle import OracleOperator, synthetic_vector

P6 = (0,) * 6
CASES = [
    # degree, n_sub, refine, mapping degree, deformation, bc
    (1, 3, 0, 1, 0.1, P6),
    (2, 2, 0, 1, 0.0, P6),
    (2, 3, 0, 1, 0.1, P6),
    (2, 1, 2, 1, 0.1, P6),
    (3, 2, 0, 3, 0.15, (1, 2, 1, 1, 1, 1)),
    (1, 3, 0, 2, 0.1, (1, 1, 0, 0, 2, 1)),
    (4, 2, 0, 1, 0.0, (1, 2, 1, 1, 1, 1)),
    (3, 3, 0, 1, 0.1, P6),
]


@pytest.mark.parametrize("case", CASES)
def test_matrix_free_equals_dense(case):
    k, nsub, ref, m, deform, bc = case
    op = OracleOperator(k, nsub, ref, m, deform, 2, bc)
    xmap, nb, nbface, bt = op.mesh()
    A, tau = dense_sipg.assemble(k, xmap, nb, nbface, bt, m)
    assert np.abs(tau - op.tau()).max() < 1e-12 * np.abs(tau).max()
    x = synthetic_vector(op.n_dofs)
    yd = A @ x
    y = op.vmult(x)
    assert np.linalg.norm(y - yd) / np.linalg.norm(yd) < 1e-13
    # cell-wise evaluation (threaded baseline) is the same operator
    yc = op.vmult_cellwise(x)
    assert np.linalg.norm(y - yc) / np.linalg.norm(y) < 1e-14
    # symmetric, positive semi-definite
    assert np.abs(A - A.T).max() < 1e-13 * np.abs(A).max()
    ev = np.linalg.eigvalsh(0.5 * (A + A.T))
    assert ev.min() > -1e-12 * ev.max()
    # diagonal = unit-vector columns (verify_calculation_of_diagonal.h:57-92 idea)
    d = op.diagonal()
    assert np.abs(d - np.diag(A)).max() < 1e-13 * np.abs(d).max()
    if all(b == 0 for b in bc):
        # constants are in the null space on the periodic box
        assert np.abs(op.vmult(np.ones(op.n_dofs))).max() < 1e-11 * np.abs(A).max()
    else:
        assert ev.min() > 0


def test_vmult_add_accumulates():
    op = OracleOperator(2, 2, 0, 1, 0.0, 2, P6)
    x = synthetic_vector(op.n_dofs)
    y = op.vmult(x)
    z = np.full(op.n_dofs, 3.0)
    op.vmult_add(z, x)
    assert np.allclose(z, y + 3.0, rtol=0, atol=1e-13 * np.abs(y).max())


def test_penalty_uniform_cartesian():
    # tau_K = 3/h on the uniform periodic box
    op = OracleOperator(3, 1, 2, 1, 0.0, 2, P6)
    assert np.allclose(op.tau(), 3.0 / 0.5, rtol=1e-14)


def test_invert_diagonal_guard():
    import ctypes as C
    from oracle.oracle import lib
    d = np.array([2.0, 1e-11, -4.0, 0.0])
    lib().orc_invert_diagonal.argtypes = [C.POINTER(C.c_double), C.c_long]
    lib().orc_invert_diagonal(d.ctypes.data_as(C.POINTER(C.c_double)), 4)
    assert np.allclose(d, [0.5, 1.0, -0.25, 1.0])
