"""exadg_b200: B200-native SIPG Laplace operator (FP64, k = 1..7) behind ExaDG's operator surface.

The product is libexadg_b200.so (CUDA sm_100a + C ABI, include/exadg_b200.h).  This package is the
thin Python mirror of the reference's operator interface used by the tests and bench.py:
    LaplaceOperator  ~ ExaDG::Poisson::LaplaceOperator / OperatorBase (vmult, vmult_add, apply, apply_add,
                       calculate_diagonal, calculate_inverse_diagonal, initialize_dof_vector, m, n, el)
    KrylovSolverCG   ~ ExaDG::Krylov::KrylovSolver with solver "cg"
    ChebyshevSmoother, JacobiPreconditioner
PyTorch is used only for device memory and streams.  There is no CPU fallback: if the CUDA library
is missing or no GPU is present, constructing an operator raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libexadg_b200.so")

PERIODIC, DIRICHLET, NEUMANN = 0, 1, 2
PRECOND_NONE, PRECOND_POINT_JACOBI, PRECOND_CHEBYSHEV = 0, 1, 2

# every symbol include/exadg_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "exadg_b200_last_error", "exadg_b200_version", "exadg_b200_create_hypercube", "exadg_b200_create", "exadg_b200_destroy",
    "exadg_b200_set_stream", "exadg_b200_synchronize", "exadg_b200_n", "exadg_b200_local_size", "exadg_b200_n_cells_owned",
    "exadg_b200_n_cells_ghost", "exadg_b200_is_cartesian_path", "exadg_b200_kernel_launches", "exadg_b200_initialize_dof_vector",
    "exadg_b200_free_dof_vector", "exadg_b200_vmult", "exadg_b200_vmult_add", "exadg_b200_vmult_host",
    "exadg_b200_vmult_host_pipelined", "exadg_b200_host_pipeline_plan", "exadg_b200_set_host_pipeline_mode", "exadg_b200_host_stream_plan",
    "exadg_b200_calculate_diagonal", "exadg_b200_add_diagonal", "exadg_b200_calculate_inverse_diagonal", "exadg_b200_jacobi_vmult",
    "exadg_b200_cg_solve", "exadg_b200_chebyshev_create", "exadg_b200_chebyshev_destroy", "exadg_b200_chebyshev_get",
    "exadg_b200_chebyshev_set_interval", "exadg_b200_chebyshev_vmult", "exadg_b200_chebyshev_step", "exadg_b200_set_nccl_comm",
    "exadg_b200_nccl_unique_id", "exadg_b200_nccl_init", "exadg_b200_halo_n_peers", "exadg_b200_halo_peer",
    "exadg_b200_halo_send_list", "exadg_b200_ghost_global_ids", "exadg_b200_ghost_buffer", "exadg_b200_halo_pack",
    "exadg_b200_fp64_peak", "exadg_b200_cartesian_kernel", "exadg_b200_plan_create", "exadg_b200_plan_destroy", "exadg_b200_plan_sizes", "exadg_b200_plan_peer",
    "exadg_b200_plan_tables", "exadg_b200_p2p_export", "exadg_b200_p2p_connect",
    "exadg_b200_wait_stream", "exadg_b200_stream_wait_operator", "exadg_b200_operator_is_singular", "exadg_b200_degree",
    "exadg_b200_set_kernel_variant", "exadg_b200_get_kernel_variant",
    "exadg_b200_n_boundary_faces", "exadg_b200_boundary_quadrature_points", "exadg_b200_set_boundary_values", "exadg_b200_rhs", "exadg_b200_rhs_add",
    "exadg_b200_evaluate", "exadg_b200_evaluate_add", "exadg_b200_cell_quadrature_points", "exadg_b200_integrate_source_add", "exadg_b200_l2_error",
    "exadg_b200_subtract_mean_value",
    "exadg_b200_create_hypercube_helmholtz", "exadg_b200_create_helmholtz", "exadg_b200_n_components", "exadg_b200_set_scaling_factor_mass",
    "exadg_b200_inverse_mass_vmult",
    "exadg_b200_multigrid_levels", "exadg_b200_multigrid_create", "exadg_b200_multigrid_destroy", "exadg_b200_multigrid_vmult", "exadg_b200_multigrid_info",
    "exadg_b200_multigrid_smoother", "exadg_b200_cg_solve_multigrid",
]


class HypercubeDesc(C.Structure):
    _fields_ = [("degree", C.c_int), ("n_subdivisions", C.c_int), ("n_refinements", C.c_int), ("mapping_degree", C.c_int),
                ("deformation", C.c_double), ("frequency", C.c_int), ("boundary", C.c_int * 6), ("ip_factor", C.c_double),
                ("rank", C.c_int), ("world", C.c_int), ("force_general", C.c_int)]


class MeshDesc(C.Structure):
    _fields_ = [("degree", C.c_int), ("mapping_degree", C.c_int), ("n_cells_owned", C.c_int64), ("n_cells_ghost", C.c_int64),
                ("mapping_points", C.POINTER(C.c_double)), ("neighbors", C.POINTER(C.c_int32)), ("neighbor_face", C.POINTER(C.c_uint8)),
                ("boundary_type", C.POINTER(C.c_uint8)), ("ip_factor", C.c_double), ("n_global_cells", C.c_int64),
                ("global_cell_offset", C.c_int64), ("force_general", C.c_int), ("operator_is_singular", C.c_int)]


class HelmholtzData(C.Structure):
    _fields_ = [("n_components", C.c_int), ("scaling_factor_mass", C.c_double), ("viscosity", C.c_double)]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("exadg_b200: %s is missing - run `python __graft_entry__.py` (nvcc, sm_100a) first; "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, i64 = C.c_void_p, C.c_void_p, C.c_int64  # device pointers travel as integers
    L.exadg_b200_last_error.restype = C.c_char_p
    L.exadg_b200_create_hypercube.argtypes = [C.POINTER(HypercubeDesc), C.POINTER(vp)]
    L.exadg_b200_create.argtypes = [C.POINTER(MeshDesc), C.POINTER(vp)]
    L.exadg_b200_create_hypercube_helmholtz.argtypes = [C.POINTER(HypercubeDesc), C.POINTER(HelmholtzData), C.POINTER(vp)]
    L.exadg_b200_create_helmholtz.argtypes = [C.POINTER(MeshDesc), C.POINTER(HelmholtzData), C.POINTER(vp)]
    L.exadg_b200_n_components.argtypes = [vp]
    L.exadg_b200_set_scaling_factor_mass.argtypes = [vp, C.c_double]
    L.exadg_b200_inverse_mass_vmult.argtypes = [vp, dp, dp]
    L.exadg_b200_destroy.argtypes = [vp]
    L.exadg_b200_set_stream.argtypes = [vp, vp]
    L.exadg_b200_synchronize.argtypes = [vp]
    L.exadg_b200_wait_stream.argtypes = [vp, vp]
    L.exadg_b200_stream_wait_operator.argtypes = [vp, vp]
    L.exadg_b200_operator_is_singular.argtypes = [vp]
    L.exadg_b200_degree.argtypes = [vp]
    L.exadg_b200_set_kernel_variant.argtypes = [vp, C.c_int]
    L.exadg_b200_get_kernel_variant.argtypes = [vp]
    for name in ("exadg_b200_n", "exadg_b200_local_size", "exadg_b200_n_cells_owned", "exadg_b200_n_cells_ghost"):
        getattr(L, name).restype = i64
        getattr(L, name).argtypes = [vp]
    L.exadg_b200_is_cartesian_path.argtypes = [vp]
    L.exadg_b200_subtract_mean_value.argtypes = [vp, dp]
    L.exadg_b200_n_boundary_faces.argtypes = [vp, C.POINTER(i64)]
    L.exadg_b200_boundary_quadrature_points.argtypes = [vp, vp, vp]
    L.exadg_b200_set_boundary_values.argtypes = [vp, vp]
    L.exadg_b200_rhs.argtypes = [vp, dp]
    L.exadg_b200_rhs_add.argtypes = [vp, dp]
    L.exadg_b200_evaluate.argtypes = [vp, dp, dp]
    L.exadg_b200_evaluate_add.argtypes = [vp, dp, dp]
    L.exadg_b200_cell_quadrature_points.argtypes = [vp, C.c_int, vp]
    L.exadg_b200_integrate_source_add.argtypes = [vp, dp, vp]
    L.exadg_b200_l2_error.argtypes = [vp, dp, vp, C.c_int, C.POINTER(C.c_double)]
    L.exadg_b200_kernel_launches.argtypes = [vp, C.POINTER(i64)]
    L.exadg_b200_initialize_dof_vector.argtypes = [vp, C.POINTER(vp)]
    L.exadg_b200_free_dof_vector.argtypes = [vp]
    L.exadg_b200_vmult.argtypes = [vp, dp, dp]
    L.exadg_b200_vmult_add.argtypes = [vp, dp, dp]
    L.exadg_b200_vmult_host.argtypes = [vp, vp, vp]
    L.exadg_b200_vmult_host_pipelined.argtypes = [vp, vp, vp]
    L.exadg_b200_host_pipeline_plan.argtypes = [C.POINTER(HypercubeDesc), i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.exadg_b200_set_host_pipeline_mode.argtypes = [vp, C.c_int]
    L.exadg_b200_host_stream_plan.argtypes = [C.POINTER(HypercubeDesc), C.c_int, i64, C.POINTER(C.c_int32), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.exadg_b200_calculate_diagonal.argtypes = [vp, dp]
    L.exadg_b200_add_diagonal.argtypes = [vp, dp]
    L.exadg_b200_calculate_inverse_diagonal.argtypes = [vp, dp]
    L.exadg_b200_jacobi_vmult.argtypes = [vp, dp, dp, dp]
    L.exadg_b200_cg_solve.argtypes = [vp, dp, dp, C.c_int, vp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.exadg_b200_chebyshev_create.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.POINTER(vp)]
    L.exadg_b200_chebyshev_destroy.argtypes = [vp]
    L.exadg_b200_chebyshev_get.argtypes = [vp] + [C.POINTER(C.c_double)] * 4
    L.exadg_b200_chebyshev_set_interval.argtypes = [vp, C.c_double, C.c_double]
    L.exadg_b200_chebyshev_vmult.argtypes = [vp, dp, dp]
    L.exadg_b200_chebyshev_step.argtypes = [vp, dp, dp]
    L.exadg_b200_multigrid_levels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.exadg_b200_multigrid_create.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(vp)]
    L.exadg_b200_multigrid_destroy.argtypes = [vp]
    L.exadg_b200_multigrid_vmult.argtypes = [vp, dp, dp]
    L.exadg_b200_multigrid_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(i64)]
    L.exadg_b200_multigrid_smoother.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.exadg_b200_cg_solve_multigrid.argtypes = [vp, dp, dp, vp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.exadg_b200_set_nccl_comm.argtypes = [vp, vp]
    L.exadg_b200_nccl_unique_id.argtypes = [C.c_char_p]
    L.exadg_b200_nccl_init.argtypes = [vp, C.c_char_p]
    L.exadg_b200_halo_n_peers.argtypes = [vp]
    L.exadg_b200_halo_peer.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.exadg_b200_halo_send_list.argtypes = [vp, C.c_int, C.POINTER(C.c_int32)]
    L.exadg_b200_ghost_global_ids.argtypes = [vp, C.POINTER(i64)]
    L.exadg_b200_ghost_buffer.restype = vp
    L.exadg_b200_ghost_buffer.argtypes = [vp]
    L.exadg_b200_halo_pack.argtypes = [vp, C.c_int, dp, dp]
    L.exadg_b200_cartesian_kernel.argtypes = [C.c_int]
    L.exadg_b200_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.exadg_b200_plan_create.argtypes = [C.POINTER(HypercubeDesc), C.POINTER(vp)]
    L.exadg_b200_plan_destroy.argtypes = [vp]
    L.exadg_b200_plan_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int)]
    L.exadg_b200_plan_peer.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32)]
    L.exadg_b200_plan_tables.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(i64)]
    L.exadg_b200_p2p_export.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
    L.exadg_b200_p2p_connect.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
    _lib = L
    return L


from .laplace_operator import (ChebyshevSmoother, ExaDGError, JacobiPreconditioner, KrylovSolverCG, MultigridPreconditioner, PartitionPlan,  # noqa: E402
                               LaplaceOperator, SolverData, cartesian_kernel, fp64_peak, host_pipeline_plan, host_stream_plan, multigrid_levels)

__all__ = ["LaplaceOperator", "KrylovSolverCG", "ChebyshevSmoother", "JacobiPreconditioner", "SolverData", "ExaDGError",
           "fp64_peak", "cartesian_kernel", "host_pipeline_plan", "host_stream_plan", "load_library", "PERIODIC", "DIRICHLET", "NEUMANN"]
