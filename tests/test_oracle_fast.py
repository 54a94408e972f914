"""The vectorised CPU baseline (oracle/sipg_fast.inc: 8 cells per SIMD batch, compile-time degree, OpenMP) must be the same operator as
the face-centric oracle that is pinned to the reference's golden fixtures; outside its scope (curved cells, boundary faces) it must
fall back to the scalar cell-wise path.  It is what bench.py's cpu_baseline / --impl reference legs time."""
import numpy as np
import pytest

from oracle.oracle import OracleOperator, synthetic_vector


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("grid", [(1, 2), (3, 1), (5, 0)])
def test_fast_path_equals_face_centric_oracle_on_periodic_boxes(degree, grid):
    op = OracleOperator(degree, grid[0], grid[1])
    x = synthetic_vector(op.n_dofs, seed=5)
    for threads in (1, 3):  # tail batches (cells not a multiple of 8) and thread-range boundaries
        y = op.vmult_fast(x, threads)
        assert op.fast_path_used
        ref = op.vmult(x)
        assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("kwargs", [dict(deformation=0.1), dict(bc=(1, 2, 1, 1, 1, 1)), dict(mapping_degree=3, deformation=0.15, bc=(1, 2, 1, 1, 1, 1))])
def test_fast_path_falls_back_outside_its_scope(kwargs):
    op = OracleOperator(3, 2, 1, kwargs.get("mapping_degree", 1), kwargs.get("deformation", 0.0), 2, kwargs.get("bc", (0,) * 6))
    x = synthetic_vector(op.n_dofs, seed=6)
    y = op.vmult_fast(x)
    assert not op.fast_path_used
    ref = op.vmult(x)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13
