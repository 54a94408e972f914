"""Pins the CPU oracle against the reference's own golden fixtures.

The reference has no fixture for a bare vmult; what it pins are the relative L2 errors of
the full DG Poisson solve of applications/poisson/sine (tests/cartesian.output,
tests/curvilinear.output).  The error depends on the discrete operator (cell, interior
face, Dirichlet and Neumann terms, penalty, MappingQ(3) geometry), the rhs and the error
quadrature - not on the preconditioner - so solving the oracle's system with its own CG
and reproducing all printed digits pins every piece of the oracle's operator.
"""
import json
import os

import pytest

from oracle.oracle import OracleOperator

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sine_l2_errors.json")))


@pytest.mark.parametrize("mesh", ["cartesian", "curvilinear"])
@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
def test_sine_l2_error_matches_reference_output(mesh, degree):
    cfg = GOLD["config"]
    deform = cfg["deformation_curvilinear"] if mesh == "curvilinear" else 0.0
    op = OracleOperator(degree, cfg["n_cells_1d_coarse"], cfg["refine"], cfg["mapping_degree"], deform,
                        cfg["frequency"], tuple(cfg["bc"]))
    assert op.n_cells == 512
    b = op.rhs_sine()
    x, its, hist, converged = op.cg(b, jacobi=True, abs_tol=1e-20, rel_tol=cfg["cg_rel_tol"], max_it=10000)
    assert converged
    err = op.l2_error_sine(x)
    gold = GOLD[mesh][degree - 1]
    # the fixture prints 6 significant digits
    assert abs(err / gold - 1.0) < 6e-6, (err, gold)
