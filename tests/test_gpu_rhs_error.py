"""rhs / evaluate (inhomogeneous boundary integrals), the volume source term and the L2 error on the GPU:
OperatorBase::rhs / evaluate (operator_base.cpp:509-606), weak_boundary_conditions.h:72-134, RHSOperator,
calculate_error (error_calculation.cpp:36-115) - against the oracle and against the reference's own golden L2 errors of
applications/poisson/sine, with NOTHING of the solve taken from the oracle: boundary data, source term and exact solution are
evaluated here (numpy, at the quadrature points the library hands out), everything else runs on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.oracle import OracleOperator

pytestmark = pytest.mark.gpu
SINE_BC = (1, 2, 1, 1, 1, 1)
W = 3.0 * np.pi  # applications/poisson/sine/application.h:32
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sine_l2_errors.json")))


def solution(x):
    return np.sin(W * x[..., 0]) * np.sin(W * x[..., 1]) * np.sin(W * x[..., 2])


def neumann(x):  # d u / d n on the face x = +1 (outward normal +x): application.h:221-233
    return W * np.cos(W * x[..., 0]) * np.sin(W * x[..., 1]) * np.sin(W * x[..., 2])


def source(x):  # -laplace u
    return 3.0 * W * W * solution(x)


def gpu_rhs(op, degree):
    """Poisson::Operator::rhs (operator.cpp:414-423): -(inhomogeneous boundary integrals) + (f, v), all on the GPU"""
    import exadg_b200
    xyz, bt = op.boundary_quadrature_points()
    values = np.where((bt == exadg_b200.DIRICHLET)[:, None], solution(xyz), neumann(xyz))
    op.set_boundary_values(values)
    b = op.initialize_dof_vector()
    op.rhs(b)
    op.integrate_source_add(b, source(op.cell_quadrature_points(degree + 1)))
    return b


@pytest.mark.parametrize("case", [(2, 2, 1, 1, 0.0), (3, 2, 1, 3, 0.15), (4, 2, 2, 3, 0.0), (5, 2, 1, 2, 0.1)])
def test_rhs_and_l2_error_match_oracle(case):
    import exadg_b200
    degree, n_sub, refine, m, deformation = case
    op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, m, deformation, 2, SINE_BC)
    ref = OracleOperator(degree, n_sub, refine, m, deformation, 2, SINE_BC)
    b = gpu_rhs(op, degree).cpu().numpy()
    b_ref = ref.rhs_sine()
    assert np.linalg.norm(b - b_ref) < 1e-12 * np.linalg.norm(b_ref)
    # evaluate = homogeneous operator + inhomogeneous boundary integrals (operator_base.cpp:548-606); rhs = -(evaluate(0))
    u = np.random.default_rng(3).uniform(-1, 1, ref.n_dofs)
    src = torch.from_numpy(u).cuda()
    ev, au, rb = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
    op.evaluate(ev, src)
    op.vmult(au, src)
    op.rhs(rb)
    assert ((ev - (au - rb)).norm() / ev.norm()).item() < 1e-13
    assert np.linalg.norm(au.cpu().numpy() - ref.vmult(u)) < 1e-12 * np.linalg.norm(au.cpu().numpy())
    op.evaluate_add(ev, src)
    assert ((ev - 2 * (au - rb)).norm() / ev.norm()).item() < 1e-13
    # L2 error of an arbitrary vector, Gauss(k+3)
    err = op.l2_error(src, solution(op.cell_quadrature_points(degree + 3)))
    assert abs(err / ref.l2_error_sine(u) - 1.0) < 1e-12


@pytest.mark.parametrize("mesh", ["cartesian", "curvilinear"])
@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
def test_gpu_pipeline_reproduces_reference_golden_l2_errors(mesh, degree):
    """applications/poisson/sine end to end on the GPU: rhs (boundary data + source), CG + point-Jacobi, relative L2 error.  All 14
    numbers of the reference's cartesian.output / curvilinear.output are reproduced to the printed digits; the oracle is not used."""
    import exadg_b200
    cfg = GOLD["config"]
    deform = cfg["deformation_curvilinear"] if mesh == "curvilinear" else 0.0
    op = exadg_b200.LaplaceOperator.hypercube(degree, cfg["n_cells_1d_coarse"], cfg["refine"], cfg["mapping_degree"], deform, 2, tuple(cfg["bc"]))
    assert not op.operator_is_singular()
    b = gpu_rhs(op, degree)
    solver = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), exadg_b200.SolverData(10000, 1e-20, cfg["cg_rel_tol"]))
    x = op.initialize_dof_vector()
    solver.solve(x, b)
    err = op.l2_error(x, solution(op.cell_quadrature_points(degree + 3)))
    assert abs(err / GOLD[mesh][degree - 1] - 1.0) < 6e-6, (err, GOLD[mesh][degree - 1])


def test_operator_is_singular_follows_the_boundary_conditions():
    import exadg_b200
    assert exadg_b200.LaplaceOperator.hypercube(2, 1, 1).operator_is_singular()                      # all-periodic box
    assert exadg_b200.LaplaceOperator.hypercube(2, 1, 1, boundary=(2,) * 6).operator_is_singular()   # pure Neumann
    assert not exadg_b200.LaplaceOperator.hypercube(2, 1, 1, boundary=SINE_BC).operator_is_singular()


def test_singular_pressure_poisson_system_is_solved_after_subtracting_the_mean():
    """the pressure Poisson operator without Dirichlet boundary is singular (operator_projection_methods.cpp:88-153);
    its callers subtract the mean of the right-hand side (time_int_bdf_dual_splitting.cpp:655-656) and CG then converges to the
    mean-free solution.  Periodic box, k = 3, Jacobi-preconditioned CG as configured for the pressure solve."""
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(3, 1, 2)
    assert op.operator_is_singular()
    ref = OracleOperator(3, 1, 2)
    rng = np.random.default_rng(5)
    b = torch.from_numpy(rng.uniform(-1, 1, ref.n_dofs) + 0.3).cuda()   # inconsistent: mean != 0 is not in the range of A
    op.subtract_mean_value(b)
    assert abs(b.sum().item()) < 1e-9
    x = op.initialize_dof_vector()
    its = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), exadg_b200.SolverData(2000, 1e-20, 1e-8)).solve(x, b)
    assert 0 < its < 2000
    r = op.initialize_dof_vector()
    op.vmult(r, x)
    assert ((r - b).norm() / b.norm()).item() < 1e-7
    op.subtract_mean_value(x)   # adjust_pressure_level_if_undefined: mean-free representative
    assert abs(x.sum().item()) < 1e-8 * x.abs().sum().item()
