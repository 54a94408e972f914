"""Parity of the CUDA path (through the C ABI) with the CPU oracle: relative l2 <= 1e-12 (FP64),
the tolerance BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleOperator, synthetic_vector

pytestmark = pytest.mark.gpu
TOL = 1e-12
P6 = (0,) * 6
SINE_BC = (1, 2, 1, 1, 1, 1)  # applications/poisson/sine: Dirichlet, Neumann on x=+1


def make_pair(degree, n_sub, refine, mapping_degree=1, deformation=0.0, bc=P6, force_general=False):
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, mapping_degree, deformation, 2, bc, 1.0, force_general=force_general)
    ref = OracleOperator(degree, n_sub, refine, mapping_degree, deformation, 2, bc)
    assert op.n() == ref.n_dofs
    return op, ref


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def gpu_vmult(op, x):
    src = torch.from_numpy(x).cuda()
    dst = op.initialize_dof_vector()
    op.vmult(dst, src)
    return dst.cpu().numpy()


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("grid", [(1, 2), (3, 1)])
def test_vmult_cartesian_periodic(degree, grid):
    op, ref = make_pair(degree, *grid)
    assert op.is_cartesian_path == 1
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


@pytest.mark.parametrize("degree", [2, 3, 4])
def test_vmult_cartesian_fast_path_tail_batches_and_larger_grid(degree):
    # 5^3 = 125 cells (ragged last batch) and 8^3 = 512 cells
    for grid in [(5, 0), (1, 3)]:
        op, ref = make_pair(degree, *grid)
        assert op.is_cartesian_path == 1
        x = synthetic_vector(ref.n_dofs, seed=7)
        assert rel(gpu_vmult(op, x), ref.vmult_cellwise(x)) < TOL


@pytest.mark.parametrize("degree", [1, 2, 3, 4])
def test_general_kernel_equals_fast_path_on_cartesian(degree):
    op, ref = make_pair(degree, 1, 2, force_general=True)
    assert op.is_cartesian_path == 0
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
def test_vmult_curved_periodic_trilinear(degree):
    # applications/poisson/throughput with MeshType=Curvilinear: deformation 0.1, mapping degree 1
    op, ref = make_pair(degree, 1, 2, 1, 0.1)
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


@pytest.mark.parametrize("degree", [2, 4, 5, 7])
def test_vmult_sine_case_dirichlet_neumann_mapping_q3(degree):
    # applications/poisson/sine, curvilinear: deformation 0.15, mapping degree 3, Dirichlet + Neumann
    op, ref = make_pair(degree, 2, 1, 3, 0.15, SINE_BC)
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


@pytest.mark.parametrize("bc", [(1, 1, 1, 1, 1, 1), (2, 1, 0, 0, 1, 2), (1, 2, 1, 1, 1, 1)])
def test_vmult_cartesian_with_boundaries(bc):
    op, ref = make_pair(3, 3, 0, 1, 0.0, bc)
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
def test_uniform_box_with_boundaries_takes_the_hybrid_fast_path(degree):
    """Uniform box with Dirichlet / Neumann / periodic faces (the Cartesian sine case, laplace_operator.cpp:221-265): batches whose cells
    see the interior penalty on all faces run the affine fast kernels, the two cell layers next to the boundary the general kernel."""
    import exadg_b200
    bc = (1, 2, 1, 1, 0, 0) if degree % 2 else (1, 1, 2, 1, 1, 1)
    op, ref = make_pair(degree, 1, 4, 1, 0.0, bc)
    assert op.is_cartesian_path == 2
    x = synthetic_vector(ref.n_dofs)
    y_ref = ref.vmult_cellwise(x)
    assert rel(gpu_vmult(op, x), y_ref) < TOL
    # vmult_add, diagonal and the general-only run of the same mesh
    src = torch.from_numpy(x).cuda()
    dst = torch.from_numpy(y_ref).cuda()
    op.vmult_add(dst, src)
    assert rel(dst.cpu().numpy(), 2 * y_ref) < TOL
    gen, _ = make_pair(degree, 1, 4, 1, 0.0, bc, force_general=True)
    assert gen.is_cartesian_path == 0
    assert rel(gpu_vmult(gen, x), y_ref) < TOL
    if degree <= 3:
        d = op.initialize_dof_vector()
        op.calculate_diagonal(d)
        assert rel(d.cpu().numpy(), ref.diagonal()) < TOL


def test_rank_without_cells_is_empty_locally():
    """more ranks than cells: the p4est-style partition leaves ranks empty (OperatorBase::is_empty_locally); their calls are no-ops"""
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(2, 1, 0, rank=0, world=2)   # one cell, owned by rank 1
    assert op.is_empty_locally() and op.local_size() == 0 and op.n() == 27
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    op.vmult(dst, src)
    op.vmult_add(dst, src)
    op.calculate_diagonal(dst)
    assert dst.numel() == 0


def test_single_cell_periodic_is_its_own_neighbour():
    for degree in (2, 4, 5):
        op, ref = make_pair(degree, 1, 0)
        x = synthetic_vector(ref.n_dofs)
        assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


def test_vmult_add_and_aliases():
    op, ref = make_pair(3, 1, 2, 1, 0.1)
    x = synthetic_vector(ref.n_dofs)
    y_ref = ref.vmult(x)
    src = torch.from_numpy(x).cuda()
    dst = torch.full((ref.n_dofs,), 3.0, dtype=torch.float64, device="cuda")
    op.vmult_add(dst, src)
    assert np.abs(dst.cpu().numpy() - (y_ref + 3.0)).max() < 1e-12 * np.abs(y_ref).max()
    dst2 = op.initialize_dof_vector()
    op.apply(dst2, src)
    assert rel(dst2.cpu().numpy(), y_ref) < TOL
    op.vmult_interface_down(dst2, src)
    assert rel(dst2.cpu().numpy(), y_ref) < TOL
    # fast path
    op, ref = make_pair(4, 1, 2)
    y_ref = ref.vmult(x := synthetic_vector(ref.n_dofs))
    dst = torch.full((ref.n_dofs,), -2.0, dtype=torch.float64, device="cuda")
    op.vmult_add(dst, torch.from_numpy(x).cuda())
    assert np.abs(dst.cpu().numpy() - (y_ref - 2.0)).max() < 1e-12 * np.abs(y_ref).max()


def test_from_mesh_matches_hypercube():
    import exadg_b200
    ref = OracleOperator(3, 3, 0, 2, 0.1, 2, (1, 1, 0, 0, 2, 1))
    xmap, nb, nbface, bt = ref.mesh()
    op = exadg_b200.LaplaceOperator.from_mesh(3, 2, xmap, nb, nbface, bt)
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL
    # Cartesian periodic mesh given as arrays is detected and takes the fast path
    ref = OracleOperator(2, 1, 2, 1, 0.0)
    xmap, nb, nbface, bt = ref.mesh()
    op = exadg_b200.LaplaceOperator.from_mesh(2, 1, xmap, nb, nbface, bt)
    assert op.is_cartesian_path == 1
    x = synthetic_vector(ref.n_dofs)
    assert rel(gpu_vmult(op, x), ref.vmult(x)) < TOL


def test_error_behaviour():
    import exadg_b200
    op, ref = make_pair(2, 1, 1)
    with pytest.raises(exadg_b200.ExaDGError):
        op.el(0, 0)
    v = op.initialize_dof_vector()
    with pytest.raises(exadg_b200.ExaDGError):
        op.vmult(v, v)  # dst and src must not alias
    with pytest.raises(exadg_b200.ExaDGError):
        op.vmult(v, torch.zeros(3, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("case", [(2, 1, 2, 1, 0.0, P6), (4, 1, 1, 1, 0.0, P6), (3, 1, 2, 1, 0.1, P6), (2, 2, 1, 3, 0.15, SINE_BC), (5, 2, 0, 1, 0.1, SINE_BC)])
def test_diagonal_and_inverse_diagonal(case):
    op, ref = make_pair(*case)
    d = op.initialize_dof_vector()
    op.calculate_diagonal(d)
    d_ref = ref.diagonal()
    assert np.abs(d.cpu().numpy() - d_ref).max() < 1e-12 * np.abs(d_ref).max()
    op.add_diagonal(d)
    assert np.abs(d.cpu().numpy() - 2 * d_ref).max() < 1e-12 * np.abs(d_ref).max()
    op.calculate_inverse_diagonal(d)
    assert np.abs(d.cpu().numpy() * d_ref - 1.0).max() < 1e-12


# ---- size-independent properties at a size the oracle would not finish quickly -------------------
@pytest.mark.parametrize("degree,n_sub,refine,deformation", [(4, 3, 3, 0.0), (3, 1, 5, 0.0), (4, 1, 4, 0.1), (7, 1, 3, 0.1)])
def test_properties_at_scale(degree, n_sub, refine, deformation):
    import exadg_b200
    op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, 1, deformation)
    n = op.local_size()
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    v = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    Au, Av, Aw = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
    op.vmult(Au, u)
    op.vmult(Av, v)
    scale = (Au.norm() * v.norm()).item()
    # symmetry <Au,v> = <u,Av>
    assert abs(torch.dot(Au, v).item() - torch.dot(u, Av).item()) < 1e-12 * scale
    # linearity
    op.vmult(Aw, 2.0 * u - 3.0 * v)
    assert (Aw - (2.0 * Au - 3.0 * Av)).norm().item() < 1e-12 * (2 * Au.norm() + 3 * Av.norm()).item()
    # constants are in the null space of the periodic box; positive semi-definite
    op.vmult(Aw, torch.ones_like(u))
    assert Aw.abs().max().item() < 1e-10 * Au.abs().max().item()
    assert torch.dot(Au, u).item() > 0
