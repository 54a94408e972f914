/*
 * exadg_b200.h -- C ABI of the B200-native SIPG Laplace operator (libexadg_b200.so).
 *
 * The reference (ExaDG) has no FFI for this path: the interface is the C++ class surface of
 * ExaDG::OperatorBase / ExaDG::Poisson::LaplaceOperator consumed by dealii::SolverCG,
 * dealii::PreconditionChebyshev, JacobiPreconditioner and the multigrid V-cycle.  Every entry
 * point below names the reference member it replaces (paths relative to the reference root,
 * I/ = include/exadg/).  A header-only C++ shim with the reference's names lives in
 * exadg_b200/laplace_operator.h; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - all functions return 0 on success, non-zero on error (the reference throws through
 *    AssertThrow; the shim converts the status into an exception); exadg_b200_last_error() gives the
 *    message of the last failure on the calling thread;
 *  - vectors are FP64 DEVICE pointers holding the locally owned DoFs, cell by cell in active-cell
 *    order, (k+1)^3 values per cell, lexicographic inside the cell (x fastest) - the layout of
 *    dealii::LinearAlgebra::distributed::Vector for FE_DGQ(k); pointers must be
 *    16-byte aligned; dst and src must not alias in vmult (as in the reference);
 *  - *_host variants take HOST pointers and copy;
 *  - one operator object per GPU/process; it is bound to one CUDA stream and, like the
 *    reference's operator (mutable state in const methods), not re-entrant.
 */
#ifndef EXADG_B200_H
#define EXADG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct exadg_b200_operator exadg_b200_operator;
typedef struct exadg_b200_chebyshev exadg_b200_chebyshev;
typedef struct exadg_b200_multigrid exadg_b200_multigrid;

enum { EXADG_B200_OK = 0, EXADG_B200_ERR_ARG = 1, EXADG_B200_ERR_CUDA = 2, EXADG_B200_ERR_UNSUPPORTED = 3, EXADG_B200_ERR_NOT_CONVERGED = 4 };
enum { EXADG_B200_PERIODIC = 0, EXADG_B200_DIRICHLET = 1, EXADG_B200_NEUMANN = 2 };
enum { EXADG_B200_PRECOND_NONE = 0, EXADG_B200_PRECOND_POINT_JACOBI = 1, EXADG_B200_PRECOND_CHEBYSHEV = 2, EXADG_B200_PRECOND_MULTIGRID = 3 };
/* MultigridType / PSequenceType of I/solvers_and_preconditioners/multigrid/multigrid_parameters.h:38-63 (the types that stay in the DG space) */
enum { EXADG_B200_MG_H = 0, EXADG_B200_MG_P = 1, EXADG_B200_MG_HP = 2, EXADG_B200_MG_PH = 3 };
enum { EXADG_B200_PSEQ_GO_TO_ONE = 0, EXADG_B200_PSEQ_DECREASE_BY_ONE = 1, EXADG_B200_PSEQ_BISECT = 2 };

/* Hypercube grids of the reference's benchmark/test applications
 * (I/grid/periodic_box.h:35-88, applications/poisson/throughput/application.h:93-165,
 *  applications/poisson/sine/application.h:180-330): subdivided_hyper_cube(n_sub,-1,1), refine_global,
 *  optional sine deformation (I/grid/deformed_cube_manifold.h:47-60), p4est-style partition. */
typedef struct {
  int degree;          /* k = 1..7, FE_DGQ(k), QGauss(k+1) */
  int n_subdivisions;  /* cells per direction of the coarse grid */
  int n_refinements;   /* global refinements */
  int mapping_degree;  /* MappingQ degree (>= 1) */
  double deformation;  /* 0: Cartesian; else amplitude of the sine deformation */
  int frequency;       /* deformation frequency (reference uses 2) */
  int boundary[6];     /* per domain face x-,x+,y-,y+,z-,z+: EXADG_B200_PERIODIC / DIRICHLET / NEUMANN */
  double ip_factor;    /* LaplaceKernelData::IP_factor (laplace_operator.h:40-46) */
  int rank, world;     /* partition (one rank per GPU) */
  int force_general;   /* non-zero: do not use the Cartesian fast path (testing) */
} exadg_b200_hypercube_desc;

/* General mesh: what a reference-side binding extracts from dealii::MatrixFree / Triangulation
 * (I/poisson/spatial_discretization/operator.cpp:261-284).  All arrays are HOST pointers, copied. */
typedef struct {
  int degree;
  int mapping_degree;
  int64_t n_cells_owned, n_cells_ghost;
  const double *mapping_points;   /* [(owned+ghost)][(m+1)^3][3] MappingQ support points, lexicographic */
  const int32_t *neighbors;       /* [owned][6] local cell index (ghosts: owned + i) or -1 on the boundary */
  const uint8_t *neighbor_face;   /* [owned][6] face number seen from the neighbour (standard orientation only) */
  const uint8_t *boundary_type;   /* [(owned+ghost)][6] 0 interior/periodic, 1 Dirichlet, 2 Neumann */
  double ip_factor;
  int64_t n_global_cells, global_cell_offset;
  int force_general;
  int operator_is_singular;       /* OperatorBaseData::operator_is_singular (I/operators/operator_base.h:59-103): 1 / 0, or -1 to derive it
                                     from the local boundary types (no Dirichlet face => singular) */
} exadg_b200_mesh_desc;

const char *exadg_b200_last_error(void);
int exadg_b200_version(void);

/* LaplaceOperator::initialize (I/poisson/spatial_discretization/laplace_operator.cpp:33-51) together with
 * MatrixFree::reinit and IP::calculate_penalty_parameter (I/operators/interior_penalty_parameter.h:43-99) */
int exadg_b200_create_hypercube(const exadg_b200_hypercube_desc *desc, exadg_b200_operator **op);
int exadg_b200_create(const exadg_b200_mesh_desc *desc, exadg_b200_operator **op);
int exadg_b200_destroy(exadg_b200_operator *op);

/* Helmholtz / viscous operator of the incompressible Navier-Stokes module: the momentum operator with the viscous
 * term in Laplace formulation and constant viscosity,  scaling_factor_mass * (v, u) + viscosity * a_SIPG(u, v)  on each of
 * n_components components (I/incompressible_navier_stokes/spatial_discretization/operators/momentum_operator.cpp:376-480,
 * viscous_operator.h:365-386 volume flux nu grad u, :489-560 gradient flux -1/2 nu [u] n and value flux nu ({dn u} - tau [u]),
 * I/operators/mass_kernel.h:32-93; penalty as for the Laplace operator, interior_penalty_parameter.h:43-128).  Vectors hold
 * FESystem(FE_DGQ(k)^n_components) DoFs: cell by cell, component blocks of (k+1)^3 values inside a cell.  Boundary types apply to
 * every component (velocity Dirichlet walls / Neumann outflow / periodic).  vmult, vmult_add, calculate_(inverse_)diagonal, the
 * Jacobi / Chebyshev / CG entry points and the multigrid apply.  Kernels: the general kernel; on a uniform periodic box (one GPU) the
 * affine fast kernels on the (cell, component) blocks.  Partitions (world > 1): the ghost import moves whole cell blocks. */
typedef struct { int n_components; double scaling_factor_mass; double viscosity; } exadg_b200_helmholtz_data;
int exadg_b200_create_hypercube_helmholtz(const exadg_b200_hypercube_desc *desc, const exadg_b200_helmholtz_data *data, exadg_b200_operator **op);
int exadg_b200_create_helmholtz(const exadg_b200_mesh_desc *desc, const exadg_b200_helmholtz_data *data, exadg_b200_operator **op);
int exadg_b200_n_components(const exadg_b200_operator *op);
/* MomentumOperator::set_scaling_factor_mass_operator (momentum_operator.cpp): gamma_0 / dt of the current time step */
int exadg_b200_set_scaling_factor_mass(exadg_b200_operator *op, double scaling_factor_mass);
/* InverseMassOperator::apply (I/operators/inverse_mass_operator.h): dst = M^-1 src, cell-wise exact inverse of the DG mass matrix
 * (Gauss(k+1) quadrature: M_K = S^T diag(JxW) S); the InverseMassPreconditioner of the momentum equation */
int exadg_b200_inverse_mass_vmult(exadg_b200_operator *op, double *dst, const double *src);
/* bind to a CUDA stream (cudaStream_t passed as void*); default: a stream owned by the operator */
int exadg_b200_set_stream(exadg_b200_operator *op, void *cuda_stream);
int exadg_b200_synchronize(exadg_b200_operator *op);
/* stream ordering against the caller's stream without a host synchronisation: the operator's stream waits for the work queued
 * on cuda_stream so far (call before vmult & co. when src/dst were produced on another stream) / cuda_stream waits for the
 * operator's work queued so far (call before consuming results there) */
int exadg_b200_wait_stream(exadg_b200_operator *op, void *cuda_stream);
int exadg_b200_stream_wait_operator(exadg_b200_operator *op, void *cuda_stream);

/* OperatorBase::m()/n() (I/operators/operator_base.cpp:199-214): global number of DoFs */
int64_t exadg_b200_n(const exadg_b200_operator *op);
int64_t exadg_b200_local_size(const exadg_b200_operator *op);   /* locally owned DoFs */
int64_t exadg_b200_n_cells_owned(const exadg_b200_operator *op);
int64_t exadg_b200_n_cells_ghost(const exadg_b200_operator *op);
/* 1: the affine fast kernels run on all cells (uniform box, all faces interior/periodic); 2: uniform box with Dirichlet / Neumann
 * faces - the fast kernels run on the batches whose cells see the interior penalty on all their faces, the general kernel on the two
 * cell layers next to the boundary; 0: general kernel */
int exadg_b200_is_cartesian_path(const exadg_b200_operator *op);
/* OperatorBase::operator_is_singular (I/operators/operator_base.h:196; OperatorBaseData::operator_is_singular): 1 if constants
 * lie in the kernel (no Dirichlet face: all-periodic or pure Neumann box) */
int exadg_b200_operator_is_singular(const exadg_b200_operator *op);
int exadg_b200_degree(const exadg_b200_operator *op);              /* k of FE_DGQ(k) */
int exadg_b200_kernel_launches(const exadg_b200_operator *op, int64_t *count); /* kernels launched so far by this operator */

/* OperatorBase::initialize_dof_vector (operator_base.cpp:232-237): allocate a zeroed device vector */
int exadg_b200_initialize_dof_vector(const exadg_b200_operator *op, double **vec);
int exadg_b200_free_dof_vector(double *vec);

/* OperatorBase::vmult / apply (operator_base.cpp:156-168, 264-310): dst = A src */
int exadg_b200_vmult(exadg_b200_operator *op, double *dst, const double *src);
/* OperatorBase::vmult_add / apply_add (operator_base.cpp:170-182, 312-354): dst += A src */
int exadg_b200_vmult_add(exadg_b200_operator *op, double *dst, const double *src);
/* same through host buffers (H2D copy, vmult, D2H copy) */
int exadg_b200_vmult_host(exadg_b200_operator *op, double *dst_host, const double *src_host);

/* same result, but upload, operator and download overlap chunk by chunk inside the call (PCIe is full duplex): a chunk of cells
 * is applied once the chunks holding its face neighbours have arrived.  The host buffers should be pinned.  Two variants, see
 * exadg_b200_set_host_pipeline_mode; returns EXADG_B200_ERR_UNSUPPORTED where neither applies. */
int exadg_b200_vmult_host_pipelined(exadg_b200_operator *op, double *dst_host, const double *src_host);
/* host-only view of the chunk plan of exadg_b200_vmult_host_pipelined (no CUDA call; CPU tests): n_chunks with null arrays, then
 * upload order, compute order and, per chunk, the chunk whose upload makes it computable; model = duration of one call in units
 * of a one-direction transfer (2 = no overlap). cells_per_chunk <= 0 selects the library's default for 24-cell batches. */
int exadg_b200_host_pipeline_plan(const exadg_b200_hypercube_desc *desc, int64_t cells_per_chunk, int32_t *n_chunks, int32_t *upload_order,
                                  int32_t *compute_order, int32_t *ready_chunk, double *model);

/* Variants of exadg_b200_vmult_host_pipelined (no reference counterpart; returns the previous mode, a value outside 0..2 only queries):
 * 0 automatic, 1 "staged" - the chunk plan above with a copy-engine download per chunk, 2 "direct" - src is uploaded piece by piece in
 * address order, behind every piece one launch applies the kernel units (cell batches) whose cells and face neighbours are complete,
 * and the kernels store dst straight into dst_host through its device mapping (needs cudaHostAlloc / cudaHostRegister memory; every
 * DoF of dst is written exactly once).  Automatic = direct on the affine fast path when dst_host is device-accessible, else staged.
 * The direct variant also serves partitioned operators whose ghost import the library owns (hypercube partitions; all ranks call
 * it together like vmult): the units that touch ghost cells follow the ghost import behind the last upload.
 * Environment override of the automatic choice: EXADG_B200_HOST_PIPELINE=staged|direct. */
int exadg_b200_set_host_pipeline_mode(exadg_b200_operator *op, int mode);
/* host-only view of the plan of the direct variant (no CUDA call; CPU tests): *n_steps pieces / launches and *n_units kernel units of
 * `unit` cells with null arrays, then piece_begin[n_steps + 1] (cell ranges in upload order), step_begin[n_steps + 1] and units[n_units]
 * (the units applied behind upload i are units[step_begin[i] .. step_begin[i + 1])); model as above (1 = perfect overlap).
 * On a partition (desc->world > 1) the units with a ghost neighbour are not listed: they follow the ghost import.
 * cells_per_piece <= 0 selects the library's default. */
int exadg_b200_host_stream_plan(const exadg_b200_hypercube_desc *desc, int unit, int64_t cells_per_piece, int32_t *n_steps, int64_t *n_units,
                                int64_t *piece_begin, int64_t *step_begin, int32_t *units, double *model);

/* OperatorBase::calculate_diagonal / add_diagonal / calculate_inverse_diagonal
 * (operator_base.cpp:608-646, 249-262; invert_diagonal.h:35-46) */
int exadg_b200_calculate_diagonal(exadg_b200_operator *op, double *diagonal);
int exadg_b200_add_diagonal(exadg_b200_operator *op, double *diagonal);
int exadg_b200_calculate_inverse_diagonal(exadg_b200_operator *op, double *diagonal);

/* dealii::VectorTools::subtract_mean_value on a device vector (global mean over all ranks): what the callers of a singular
 * operator (operator_is_singular, e.g. the pressure Poisson operator of the dual splitting scheme without pressure Dirichlet
 * boundary) apply to the right-hand side to make the system consistent
 * (I/incompressible_navier_stokes/time_integration/time_int_bdf_dual_splitting.cpp:655-656) and to the start vector of the
 * eigenvalue estimate (I/solvers_and_preconditioners/utilities/compute_eigenvalues.h:52-53) */
int exadg_b200_subtract_mean_value(exadg_b200_operator *op, double *vec);

/* Inhomogeneous boundary data, right-hand side and error norms on the GPU.  The reference evaluates
 * dealii::Function objects at quadrature points; the C ABI hands out the physical coordinates of its quadrature points and takes
 * the function values back as HOST arrays (the binding evaluates BoundaryDescriptor / FieldFunctions there).
 *   boundary faces of the owned cells, cell-major then face number; (k+1)^2 Gauss points per face, x fastest within the face:
 *   xyz_host [n_faces][(k+1)^2][3], type_host [n_faces] (EXADG_B200_DIRICHLET / NEUMANN); values: g on Dirichlet faces, h on
 *   Neumann faces (I/poisson/user_interface/boundary_descriptor.h) */
int exadg_b200_n_boundary_faces(exadg_b200_operator *op, int64_t *n_faces);
int exadg_b200_boundary_quadrature_points(exadg_b200_operator *op, double *xyz_host, uint8_t *type_host);
int exadg_b200_set_boundary_values(exadg_b200_operator *op, const double *values_host);
/* OperatorBase::rhs / rhs_add (operator_base.cpp:509-546): dst (+)= -(inhomogeneous boundary face integrals), exterior values per
 * weak_boundary_conditions.h:72-134, 188-234 with OperatorType::inhomogeneous */
int exadg_b200_rhs(exadg_b200_operator *op, double *dst);
int exadg_b200_rhs_add(exadg_b200_operator *op, double *dst);
/* OperatorBase::evaluate / evaluate_add (operator_base.cpp:548-606): homogeneous operator + inhomogeneous boundary integrals */
int exadg_b200_evaluate(exadg_b200_operator *op, double *dst, const double *src);
int exadg_b200_evaluate_add(exadg_b200_operator *op, double *dst, const double *src);
/* Gauss(n_q_points_1d) points of the owned cells, xyz_host [owned][n_q^3][3], x fastest */
int exadg_b200_cell_quadrature_points(exadg_b200_operator *op, int n_q_points_1d, double *xyz_host);
/* RHSOperator (I/poisson/spatial_discretization/operator.cpp:414-423): dst += (f, phi_i), f_host = f at the Gauss(k+1) points */
int exadg_b200_integrate_source_add(exadg_b200_operator *op, double *dst, const double *f_host);
/* calculate_error (I/postprocessor/error_calculation.cpp:36-115): L2 norm of u - u_exact with Gauss(k+3), relative to the L2 norm
 * of u_exact if `relative`; exact_host = u_exact at exadg_b200_cell_quadrature_points(op, k + 3, .); reduced over all ranks */
int exadg_b200_l2_error(exadg_b200_operator *op, const double *u, const double *exact_host, int relative, double *error);

/* JacobiPreconditioner::vmult (I/solvers_and_preconditioners/preconditioners/jacobi_preconditioner.h:50-62) */
int exadg_b200_jacobi_vmult(exadg_b200_operator *op, double *dst, const double *src, const double *inverse_diagonal);

/* Krylov::KrylovSolver::solve with solver "cg" = dealii::SolverCG + ReductionControl(max_iter, abs_tol, rel_tol)
 * (I/solvers_and_preconditioners/solvers/iterative_solvers_dealii_wrapper.h:137-221).  x holds the initial
 * guess.  n_iter = ReductionControl::last_step(); residuals (optional, host, length >= max_iter+1) gets
 * the l2 residual history.  Returns EXADG_B200_ERR_NOT_CONVERGED when max_iter is reached (the reference
 * throws SolverControl::NoConvergence). */
int exadg_b200_cg_solve(exadg_b200_operator *op, double *x, const double *b, int preconditioner, exadg_b200_chebyshev *cheb,
                        double abs_tol, double rel_tol, int max_iter, int *n_iter, double *residuals);

/* ChebyshevSmoother (I/solvers_and_preconditioners/multigrid/smoothers/chebyshev_smoother.h:41-42,149-172):
 * dealii::PreconditionChebyshev with point-Jacobi; lambda_max from eig_cg_n_iterations CG steps. */
int exadg_b200_chebyshev_create(exadg_b200_operator *op, int degree, double smoothing_range, int eig_cg_n_iterations, exadg_b200_chebyshev **cheb);
int exadg_b200_chebyshev_destroy(exadg_b200_chebyshev *cheb);
int exadg_b200_chebyshev_get(const exadg_b200_chebyshev *cheb, double *lambda_min_est, double *lambda_max_est, double *theta, double *delta);
int exadg_b200_chebyshev_set_interval(exadg_b200_chebyshev *cheb, double theta, double delta);
/* SmootherBase::vmult (zero initial guess) and ::step (chebyshev_smoother.h:79-119) */
int exadg_b200_chebyshev_vmult(exadg_b200_chebyshev *cheb, double *dst, const double *src);
int exadg_b200_chebyshev_step(exadg_b200_chebyshev *cheb, double *dst, const double *src);

/* Multigrid preconditioner on a hierarchy of DG level operators.
 *  - exadg_b200_multigrid_levels: MultigridPreconditionerBase::initialize_levels
 *    (I/solvers_and_preconditioners/multigrid/multigrid_preconditioner_base.cpp:97-323) for hMG / pMG / hpMG / phMG with is_dg = true:
 *    level l (coarse -> fine) lives on h-level h_level[l] (0 = coarsest of n_h_levels global refinement levels) with degree
 *    level_degree[l]; call with null arrays to query n_levels.  The c-transfer types (cphMG, ...) need a continuous FE_Q Laplace
 *    operator, which this library does not have: they return EXADG_B200_ERR_ARG.
 *  - exadg_b200_multigrid_create: takes the level operators the caller created for that list (as the reference creates one
 *    operator per level, multigrid_preconditioner_base.cpp:593-640), coarse -> fine.  Between consecutive levels either the degree
 *    changes on the same cells (p-transfer) or every cell c of the coarser level has the children 8 c .. 8 c + 7 (h-transfer of a
 *    global refinement; partitions must be aligned) - dealii::MGTwoLevelTransfer as set up by multigrid/transfer.cpp:28-69
 *    (prolongation = embedding, restriction = its transpose).  Smoother on every level > 0: ChebyshevSmoother with point Jacobi
 *    (multigrid_preconditioner_base.cpp:706-733; defaults degree 5, smoothing range 20, 20 CG iterations for the eigenvalue
 *    estimate, multigrid_parameters.h:169-178); coarse solver: MGCoarseKrylov = CG + point Jacobi to coarse_rel_tol
 *    (coarse_grid_solvers.h:62-232; defaults abs 1e-12, rel 1e-3, 1e4 iterations), mean value removed for singular operators.
 *    The level operators are re-bound to the stream of the finest one and must outlive the multigrid object.
 *    Level arithmetic is FP64 (the reference instantiates the level operators in float, multigrid_preconditioner_base.h:60).
 *  - exadg_b200_multigrid_vmult: MultigridPreconditionerBase::vmult = MultigridAlgorithm::vmult, one V-cycle with zero initial guess
 *    (multigrid_algorithm.h:88-109, 173-243).
 *  - exadg_b200_cg_solve_multigrid: exadg_b200_cg_solve with Preconditioner::Multigrid. */
int exadg_b200_multigrid_levels(int mg_type, int p_sequence, int degree, int n_h_levels, int max_levels, int *n_levels, int *h_level, int *level_degree);
int exadg_b200_multigrid_create(int n_levels, exadg_b200_operator *const *level_operators, int smoother_degree, double smoothing_range, int eig_cg_n_iterations,
                                double coarse_abs_tol, double coarse_rel_tol, int coarse_max_iter, exadg_b200_multigrid **mg);
int exadg_b200_multigrid_destroy(exadg_b200_multigrid *mg);
int exadg_b200_multigrid_vmult(exadg_b200_multigrid *mg, double *dst, const double *src);
int exadg_b200_multigrid_info(const exadg_b200_multigrid *mg, int *n_levels, int64_t *coarse_iterations, int64_t *cycles);
int exadg_b200_multigrid_smoother(const exadg_b200_multigrid *mg, int level, exadg_b200_chebyshev **smoother); /* owned by mg */
int exadg_b200_cg_solve_multigrid(exadg_b200_operator *op, double *x, const double *b, exadg_b200_multigrid *mg, double abs_tol, double rel_tol, int max_iter,
                                  int *n_iter, double *residuals);

/* Multi-GPU halo exchange of src (the update_ghost_values of MatrixFree::loop).
 * The library packs the owned cells each peer needs; transport is pluggable:
 *  - exadg_b200_set_nccl_comm: pass an initialised ncclComm_t (as void*); ghost import uses grouped
 *    ncclSend/ncclRecv and dot products use ncclAllReduce on the operator's stream;
 *  - without a communicator and world > 1 the halo can be driven from outside with
 *    exadg_b200_halo_* (used by the gloo CPU tests of the plan). */
int exadg_b200_set_nccl_comm(exadg_b200_operator *op, void *nccl_comm);
/* convenience: create the communicator inside the library (id from rank 0, broadcast by the caller) */
int exadg_b200_nccl_unique_id(char *id128);
int exadg_b200_nccl_init(exadg_b200_operator *op, const char *id128);
/* NVLink peer-memory halo (no NCCL in the data path): the pack kernel stores the cells each peer needs straight
 * into that peer's ghost buffer (CUDA IPC mapping), followed by a release flag; see csrc/c_api.cu.
 *   1. every rank: exadg_b200_p2p_export -> 64-byte IPC handle + recv_begin_by_rank[world + 1] (last entry: ghost buffer bytes)
 *   2. all-gather both over the ranks (the caller's transport, e.g. torch.distributed / MPI)
 *   3. every rank: exadg_b200_p2p_connect(handles[world][64], recv_begin_table[world][world + 1]) */
int exadg_b200_p2p_export(exadg_b200_operator *op, char *handle64, int64_t *recv_begin_by_rank);
int exadg_b200_p2p_connect(exadg_b200_operator *op, const char *handles, const int64_t *recv_begin_table);
int exadg_b200_halo_n_peers(const exadg_b200_operator *op);
int exadg_b200_halo_peer(const exadg_b200_operator *op, int i, int *peer_rank, int64_t *send_cells, int64_t *recv_cell_begin, int64_t *recv_cells);
int exadg_b200_halo_send_list(const exadg_b200_operator *op, int i, int32_t *cells_host);
int exadg_b200_ghost_global_ids(const exadg_b200_operator *op, int64_t *ids_host);
double *exadg_b200_ghost_buffer(exadg_b200_operator *op);        /* device, [n_ghost][(k+1)^3] */
int exadg_b200_halo_pack(exadg_b200_operator *op, int i, const double *src, double *send_buffer); /* device buffers */

/* Host-only view of the partition and halo plan of a hypercube grid (no CUDA call is made): the p4est-style
 * partition (I/grid/grid_utilities.h:188-207), the owned cells each peer needs and the ghost ordering. */
typedef struct exadg_b200_plan exadg_b200_plan;
int exadg_b200_plan_create(const exadg_b200_hypercube_desc *desc, exadg_b200_plan **plan);
int exadg_b200_plan_destroy(exadg_b200_plan *plan);
int exadg_b200_plan_sizes(const exadg_b200_plan *plan, int64_t *n_owned, int64_t *n_ghost, int64_t *global_offset, int *n_peers);
int exadg_b200_plan_peer(const exadg_b200_plan *plan, int i, int *peer_rank, int64_t *n_send, int64_t *recv_begin, int64_t *recv_count, int32_t *send_cells);
int exadg_b200_plan_tables(const exadg_b200_plan *plan, int32_t *neighbors, int64_t *ghost_global_ids);

/* Tuning switch without a reference counterpart: kernel of the affine fast path for degree 4 (0: pipelined 4-warp kernel,
 * 1 (default) / 2: warp-specialised kernel with producer warps fetching 8 / 12 neighbour cells per round, 3: the same with four
 * producer warps and register re-allocation between the roles (experimental); -1 only queries).
 * Process-wide; returns the previous value. All kernels compute the same operator (OperatorBase::apply,
 * operator_base.cpp:264-310); the environment variable EXADG_B200_CART_KERNEL=pipe / ws / ws12 / ws4p selects 0 / 1 / 2 / 3 at start-up. */
int exadg_b200_cartesian_kernel(int variant);
/* the same switch per operator (-1: follow the process-wide default); two operators of one process can differ */
int exadg_b200_set_kernel_variant(exadg_b200_operator *op, int variant);
int exadg_b200_get_kernel_variant(const exadg_b200_operator *op);

/* FP64 pipe microbenchmarks used for the roofline denominators (DFMA and DMMA rates) */
int exadg_b200_fp64_peak(double *dfma_tflops, double *dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* EXADG_B200_H */
