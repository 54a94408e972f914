"""Multigrid preconditioner on the GPU: V-cycle of multigrid_algorithm.h:173-243 with Chebyshev(point Jacobi)
smoothers, p- and h-transfers and the CG coarse solver, against the CPU restatement (oracle/multigrid.py), and the
multigrid-preconditioned solve of applications/poisson/sine against the reference's golden L2 errors."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle.multigrid import OracleMultigrid, initialize_levels
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_rhs_error import gpu_rhs, solution  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sine_l2_errors.json")))
SINE_BC = (1, 2, 1, 1, 1, 1)


def make_pair(mg_type, degree, n_sub, refine, mapping_degree, deformation, bc, **kw):
    import exadg_b200
    args = dict(degree=degree, n_subdivisions=n_sub, n_refinements=refine, mapping_degree=mapping_degree, deformation=deformation, boundary=bc)
    mg = exadg_b200.MultigridPreconditioner.hypercube(args, mg_type, "Bisect", **kw)
    ref = OracleMultigrid(initialize_levels(mg_type, "Bisect", degree, refine + 1), n_sub=n_sub, mapping_degree=mapping_degree, deformation=deformation, bc=bc, **kw)
    assert mg.levels == ref.levels
    return mg, ref


@pytest.mark.parametrize("case", [("pMG", 4, 2, 1, 1, 0.0, SINE_BC), ("hMG", 2, 1, 2, 1, 0.1, SINE_BC), ("phMG", 3, 1, 2, 2, 0.15, SINE_BC),
                                  ("hpMG", 4, 1, 1, 1, 0.0, SINE_BC), ("phMG", 4, 1, 2, 1, 0.0, (0,) * 6)])
def test_v_cycle_matches_the_restatement(case):
    mg_type, degree, n_sub, refine, m, deformation, bc = case
    # coarse solve converged on both sides (an iteration more or less at rel 1e-3 would change the result at that level)
    singular = all(b != 1 for b in bc)   # CG on the singular coarse system stagnates in round-off below ~1e-12: stop earlier there
    mg, ref = make_pair(mg_type, degree, n_sub, refine, m, deformation, bc, coarse_rel_tol=1e-9 if singular else 1e-13, coarse_abs_tol=1e-14)
    for level in range(1, len(mg.levels)):
        lmin, lmax, theta, delta = mg.smoother_interval(level)
        assert abs(lmax / ref.smoothers[level].lambda_max_est - 1.0) < 1e-7
        mg.set_smoother_interval(level, ref.smoothers[level].theta, ref.smoothers[level].delta)   # strict comparison of the cycle itself
    rng = np.random.default_rng(11)
    src = rng.uniform(-1, 1, ref.ops[-1].n_dofs)
    if singular:
        src -= src.mean()
    dst = mg.op.initialize_dof_vector()
    for cycle in range(2):   # the second cycle starts the coarse CG from the previous coarse solution, as the reference does
        if cycle == 1:
            src = np.roll(src, 17) * 0.7 + 0.1 * src
        mg.vmult(dst, torch.from_numpy(src).cuda())
        y_ref = ref.vmult(src)
        y = dst.cpu().numpy()
        assert np.linalg.norm(y - y_ref) < (1e-7 if singular else 1e-9) * np.linalg.norm(y_ref), np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref)
    assert mg.info()["cycles"] == 2 and mg.info()["coarse_iterations"] > 0


@pytest.mark.parametrize("mesh,degree", [("cartesian", 2), ("cartesian", 4), ("curvilinear", 3), ("curvilinear", 4), ("curvilinear", 7)])
def test_multigrid_preconditioned_gpu_solve_reproduces_the_reference_golden_l2_errors(mesh, degree):
    """applications/poisson/sine as configured by the reference (CG, rel 1e-10, multigrid with Chebyshev(5) smoothers, CG + point
    Jacobi to 1e-3 on the coarse level, p-sequence bisect; application.h:160-177), phMG instead of cphMG; everything on the GPU."""
    import exadg_b200
    cfg = GOLD["config"]
    deform = cfg["deformation_curvilinear"] if mesh == "curvilinear" else 0.0
    args = dict(degree=degree, n_subdivisions=cfg["n_cells_1d_coarse"], n_refinements=cfg["refine"], mapping_degree=cfg["mapping_degree"],
                deformation=deform, boundary=tuple(cfg["bc"]))
    op = exadg_b200.LaplaceOperator.hypercube(**args)
    mg = exadg_b200.MultigridPreconditioner.hypercube(args, "phMG", "Bisect", fine_operator=op)
    b = gpu_rhs(op, degree)
    solver = exadg_b200.KrylovSolverCG(op, mg, exadg_b200.SolverData(10000, 1e-20, cfg["cg_rel_tol"]))
    x = op.initialize_dof_vector()
    its = solver.solve(x, b)
    assert 0 < its < 25, its
    err = op.l2_error(x, solution(op.cell_quadrature_points(degree + 3)))
    assert abs(err / GOLD[mesh][degree - 1] - 1.0) < 6e-6, (err, GOLD[mesh][degree - 1])
    # the restatement needs the same number of iterations
    ref = OracleMultigrid(initialize_levels("phMG", "Bisect", degree, cfg["refine"] + 1), n_sub=cfg["n_cells_1d_coarse"], mapping_degree=cfg["mapping_degree"],
                          deformation=deform, bc=tuple(cfg["bc"]))
    _, its_ref, hist = ref.pcg(ref.ops[-1], b.cpu().numpy(), rel_tol=cfg["cg_rel_tol"])
    assert its == its_ref, (its, its_ref)
    assert np.allclose(solver.residuals[:5], hist[:5], rtol=1e-5)


def test_multigrid_on_the_singular_periodic_box():
    """pressure-Poisson-like: all-periodic box, singular operator; the coarse solver removes the mean (coarse_grid_solvers.h:169-171)"""
    import exadg_b200
    args = dict(degree=3, n_subdivisions=1, n_refinements=3)
    op = exadg_b200.LaplaceOperator.hypercube(**args)
    assert op.operator_is_singular() and op.is_cartesian_path
    mg = exadg_b200.MultigridPreconditioner.hypercube(args, "hpMG", "Bisect", fine_operator=op)
    rng = np.random.default_rng(2)
    b = torch.from_numpy(rng.uniform(-1, 1, op.local_size())).cuda()
    op.subtract_mean_value(b)
    x = op.initialize_dof_vector()
    its = exadg_b200.KrylovSolverCG(op, mg, exadg_b200.SolverData(1000, 1e-20, 1e-8)).solve(x, b)
    assert 0 < its < 25, its
    r = op.initialize_dof_vector()
    op.vmult(r, x)
    assert ((r - b).norm() / b.norm()).item() < 1e-7


def test_level_operators_must_nest():
    import exadg_b200
    a = exadg_b200.LaplaceOperator.hypercube(2, 1, 1)
    b = exadg_b200.LaplaceOperator.hypercube(3, 1, 2)   # degree and mesh change at once
    with pytest.raises(exadg_b200.ExaDGError, match="only one type of transfer"):
        exadg_b200.MultigridPreconditioner([a, b])
