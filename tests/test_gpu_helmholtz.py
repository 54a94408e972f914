"""Helmholtz / viscous operator: scaling_factor_mass * M + viscosity * A_SIPG on every component of a vector-valued
DG field (momentum_operator.cpp:376-480, viscous_operator.h:365-560, mass_kernel.h:32-93), its diagonal, the cell-wise inverse
mass operator and a Jacobi-preconditioned CG solve of the viscous step, against the CPU restatement."""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleHelmholtz, OracleOperator

pytestmark = pytest.mark.gpu
WALLS = (1, 1, 1, 1, 2, 2)   # velocity Dirichlet on four sides, Neumann (outflow-like) on two


@pytest.mark.parametrize("case", [(2, 3, 1, 2, 1, 0.0, (0,) * 6, 30.0, 0.01), (3, 3, 2, 1, 2, 0.1, WALLS, 2.5, 0.3), (5, 3, 1, 1, 1, 0.0, (0,) * 6, 100.0, 1e-3),
                                  (4, 1, 1, 2, 3, 0.15, WALLS, 1.0, 1.0), (1, 2, 2, 1, 1, 0.05, WALLS, 0.0, 2.0),
                                  # uniform periodic boxes: the affine fast kernels on the (cell, component) blocks, every kernel family
                                  # (plane kernel n = 2, 4, 5; line kernel n = 7, 8), ragged last batches included
                                  (1, 3, 3, 1, 1, 0.0, (0,) * 6, 7.0, 0.5), (3, 3, 3, 1, 1, 0.0, (0,) * 6, 12.0, 0.02), (4, 2, 1, 2, 1, 0.0, (0,) * 6, 3.0, 1.0),
                                  (6, 3, 1, 1, 1, 0.0, (0,) * 6, 50.0, 0.1), (7, 1, 1, 1, 1, 0.0, (0,) * 6, 0.5, 2.0)])
def test_helmholtz_vmult_diagonal_and_inverse_mass_match_the_restatement(case):
    import exadg_b200
    degree, ncomp, n_sub, refine, m, deformation, bc, alpha, nu = case
    op = exadg_b200.LaplaceOperator.hypercube_helmholtz(degree, ncomp, alpha, nu, n_sub, refine, m, deformation, 2, bc)
    ref = OracleHelmholtz(OracleOperator(degree, n_sub, refine, m, deformation, 2, bc), ncomp, alpha, nu)
    assert op.local_size() == ref.n_dofs and not op.is_cartesian_path
    rng = np.random.default_rng(7)
    u = rng.uniform(-1, 1, ref.n_dofs)
    src = torch.from_numpy(u).cuda()
    dst = op.initialize_dof_vector()
    op.vmult(dst, src)
    y_ref = ref.vmult(u)
    assert np.linalg.norm(dst.cpu().numpy() - y_ref) < 1e-12 * np.linalg.norm(y_ref)
    op.vmult_add(dst, src)
    assert np.linalg.norm(dst.cpu().numpy() - 2 * y_ref) < 1e-12 * np.linalg.norm(y_ref)
    d = op.initialize_dof_vector()
    op.calculate_diagonal(d)
    d_ref = ref.diagonal()
    assert np.linalg.norm(d.cpu().numpy() - d_ref) < 1e-12 * np.linalg.norm(d_ref)
    # M (M^-1 v) = v with the restatement's mass operator
    v = rng.uniform(-1, 1, ref.n_dofs)
    w = op.initialize_dof_vector()
    op.inverse_mass_vmult(w, torch.from_numpy(v).cuda())
    assert np.linalg.norm(ref.mass_vmult(w.cpu().numpy()) - v) < 1e-11 * np.linalg.norm(v)
    # a new time step size changes the mass factor only
    op.set_scaling_factor_mass(2 * alpha + 1)
    ref.alpha = 2 * alpha + 1
    op.vmult(dst, src)
    y_ref = ref.vmult(u)
    assert np.linalg.norm(dst.cpu().numpy() - y_ref) < 1e-12 * np.linalg.norm(y_ref)


def test_viscous_step_solve():
    """the viscous step of the dual splitting scheme: (gamma0/dt M + nu A) u = rhs, CG with the inverse-mass-like Jacobi
    preconditioner; k = 3, three velocity components, Dirichlet walls"""
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube_helmholtz(3, 3, 150.0, 0.01, 1, 2, 1, 0.1, 2, WALLS)
    ref = OracleHelmholtz(OracleOperator(3, 1, 2, 1, 0.1, 2, WALLS), 3, 150.0, 0.01)
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, ref.n_dofs)
    x = op.initialize_dof_vector()
    its = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), exadg_b200.SolverData(1000, 1e-20, 1e-10)).solve(x, torch.from_numpy(b).cuda())
    assert 0 < its < 100
    r = ref.vmult(x.cpu().numpy()) - b
    assert np.linalg.norm(r) < 1e-9 * np.linalg.norm(b)
    # p-multigrid on the vector-valued operator (k = 3 -> 1)
    lv = [exadg_b200.LaplaceOperator.hypercube_helmholtz(k, 3, 150.0, 0.01, 1, 2, 1, 0.1, 2, WALLS) for k in (1, 3)]
    mg = exadg_b200.MultigridPreconditioner(lv)
    x2 = lv[-1].initialize_dof_vector()
    its2 = exadg_b200.KrylovSolverCG(lv[-1], mg, exadg_b200.SolverData(1000, 1e-20, 1e-10)).solve(x2, torch.from_numpy(b).cuda())
    assert 0 < its2 <= its
    assert ((x2 - x).norm() / x.norm()).item() < 1e-8


def test_scalar_entry_points_refuse_the_vector_operator():
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube_helmholtz(2, 3, 1.0, 1.0, 1, 1, boundary=WALLS)
    with pytest.raises(exadg_b200.ExaDGError, match="scalar Laplace operator only"):
        op.boundary_quadrature_points()
    lap = exadg_b200.LaplaceOperator.hypercube(2, 1, 1)
    with pytest.raises(exadg_b200.ExaDGError, match="not a Helmholtz operator"):
        lap.set_scaling_factor_mass(1.0)
