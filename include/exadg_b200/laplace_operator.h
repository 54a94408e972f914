// Header-only C++ shim over the C ABI (include/exadg_b200.h) that mirrors the public surface of
// ExaDG::OperatorBase / ExaDG::Poisson::LaplaceOperator for the matrix-free DG Laplace path, so that
// dealii::SolverCG / dealii::PreconditionChebyshev-style templates (which only need vmult and a
// vector type) and ExaDG's JacobiPreconditioner / MultigridOperator wrappers can call it unchanged.
//
//   reference member (I/ = include/exadg/)                                   -> shim
//   OperatorBase::vmult / vmult_add            I/operators/operator_base.cpp:156-182   -> vmult / vmult_add
//   OperatorBase::apply / apply_add            operator_base.cpp:264-354               -> apply / apply_add
//   OperatorBase::vmult_interface_down / _up   operator_base.cpp:184-198               -> same names
//   OperatorBase::m / n / el                   operator_base.cpp:199-222               -> m / n / el (el throws)
//   OperatorBase::initialize_dof_vector        operator_base.cpp:232-237               -> initialize_dof_vector
//   OperatorBase::calculate_diagonal / add_diagonal / calculate_inverse_diagonal
//                                              operator_base.cpp:249-262, 608-646      -> same names
//   LaplaceOperator::initialize                I/poisson/spatial_discretization/laplace_operator.cpp:33-51 -> constructors
//   typedef Number value_type                  laplace_operator.h:248                  -> value_type
// Errors: the reference throws dealii::ExcMessage via AssertThrow; the shim throws std::runtime_error
// carrying exadg_b200_last_error().
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../exadg_b200.h"

namespace ExaDG
{
namespace B200
{
inline void check(int status)
{
  if (status != EXADG_B200_OK) throw std::runtime_error(std::string("exadg_b200: ") + exadg_b200_last_error());
}

// Device vector with the locally owned DoFs (the role of dealii::LinearAlgebra::distributed::Vector<double>;
// ghost entries live inside the operator).  Movable, not copyable.
class DeviceVector
{
public:
  DeviceVector() = default;
  DeviceVector(DeviceVector const &) = delete;
  DeviceVector & operator=(DeviceVector const &) = delete;
  DeviceVector(DeviceVector && o) noexcept : ptr(o.ptr), n(o.n) { o.ptr = nullptr; o.n = 0; }
  DeviceVector & operator=(DeviceVector && o) noexcept { std::swap(ptr, o.ptr); std::swap(n, o.n); return *this; }
  ~DeviceVector() { if (ptr) exadg_b200_free_dof_vector(ptr); }
  double * data() { return ptr; }
  double const * data() const { return ptr; }
  std::int64_t locally_owned_size() const { return n; }

private:
  friend class LaplaceOperator;
  double * ptr = nullptr;
  std::int64_t n = 0;
};

class LaplaceOperator
{
public:
  typedef double value_type;
  typedef DeviceVector VectorType;

  // benchmark / test grids of the reference (periodic box, sine case)
  explicit LaplaceOperator(exadg_b200_hypercube_desc const & desc) { check(exadg_b200_create_hypercube(&desc, &op)); }
  // general mesh extracted from dealii::MatrixFree by the reference-side binding
  explicit LaplaceOperator(exadg_b200_mesh_desc const & desc) { check(exadg_b200_create(&desc, &op)); }
  // IncNS::MomentumOperator with the viscous term in Laplace formulation (momentum_operator.cpp:376-480, viscous_operator.h:365-560):
  // scaling_factor_mass * M + viscosity * A_SIPG on every component
  LaplaceOperator(exadg_b200_hypercube_desc const & desc, exadg_b200_helmholtz_data const & data) { check(exadg_b200_create_hypercube_helmholtz(&desc, &data, &op)); }
  LaplaceOperator(exadg_b200_mesh_desc const & desc, exadg_b200_helmholtz_data const & data) { check(exadg_b200_create_helmholtz(&desc, &data, &op)); }
  LaplaceOperator(LaplaceOperator const &) = delete;
  LaplaceOperator & operator=(LaplaceOperator const &) = delete;
  ~LaplaceOperator() { exadg_b200_destroy(op); }

  void vmult(VectorType & dst, VectorType const & src) const { check(exadg_b200_vmult(op, dst.data(), src.data())); }
  void vmult_add(VectorType & dst, VectorType const & src) const { check(exadg_b200_vmult_add(op, dst.data(), src.data())); }
  void apply(VectorType & dst, VectorType const & src) const { vmult(dst, src); }
  void apply_add(VectorType & dst, VectorType const & src) const { vmult_add(dst, src); }
  void vmult_interface_down(VectorType & dst, VectorType const & src) const { vmult(dst, src); }
  void vmult_add_interface_up(VectorType & dst, VectorType const & src) const { vmult_add(dst, src); }
  // host vectors (a binding that keeps LinearAlgebra::distributed::Vector on the host): upload, vmult, download inside the call;
  // overlap = true: the three overlap chunk by chunk (unpartitioned operators, pinned host memory)
  void vmult_host(double * dst, double const * src, bool overlap = false) const
  {
    check(overlap ? exadg_b200_vmult_host_pipelined(op, dst, src) : exadg_b200_vmult_host(op, dst, src));
  }

  // inhomogeneous boundary data (OperatorBase::rhs / rhs_add / evaluate / evaluate_add, operator_base.h:314-344): the binding
  // evaluates its BoundaryDescriptor functions at boundary_quadrature_points() and hands the values to set_boundary_values()
  std::int64_t n_boundary_faces() const { std::int64_t n = 0; check(exadg_b200_n_boundary_faces(op, &n)); return n; }
  void boundary_quadrature_points(double * xyz, std::uint8_t * type) const { check(exadg_b200_boundary_quadrature_points(op, xyz, type)); }
  void set_boundary_values(double const * values) const { check(exadg_b200_set_boundary_values(op, values)); }
  void rhs(VectorType & dst) const { check(exadg_b200_rhs(op, dst.data())); }
  void rhs_add(VectorType & dst) const { check(exadg_b200_rhs_add(op, dst.data())); }
  void evaluate(VectorType & dst, VectorType const & src) const { check(exadg_b200_evaluate(op, dst.data(), src.data())); }
  void evaluate_add(VectorType & dst, VectorType const & src) const { check(exadg_b200_evaluate_add(op, dst.data(), src.data())); }

  // MomentumOperator::set_scaling_factor_mass_operator, InverseMassOperator::apply (Helmholtz operators)
  void set_scaling_factor_mass_operator(double const factor) const { check(exadg_b200_set_scaling_factor_mass(op, factor)); }
  void apply_inverse_mass(VectorType & dst, VectorType const & src) const { check(exadg_b200_inverse_mass_vmult(op, dst.data(), src.data())); }
  unsigned int n_components() const { return (unsigned int)exadg_b200_n_components(op); }

  // dealii::VectorTools::subtract_mean_value for the singular (pressure Poisson) system
  void subtract_mean_value(VectorType & v) const { check(exadg_b200_subtract_mean_value(op, v.data())); }

  std::int64_t m() const { return n(); }
  std::int64_t n() const { return exadg_b200_n(op); }
  double el(unsigned int, unsigned int) const { throw std::runtime_error("Matrix-free does not allow for entry access"); }
  bool is_empty_locally() const { return exadg_b200_n_cells_owned(op) == 0; }
  bool operator_is_singular() const { return exadg_b200_operator_is_singular(op) != 0; } // operator_base.h:196
  // getters of operator_base.h:144-197 that have a meaning without deal.II objects; get_matrix_free and
  // get_affine_constraints have no counterpart: the mesh lives inside the operator and FE_DGQ has no constraints
  // (I/solvers_and_preconditioners/multigrid/constraints.h:120-125)
  unsigned int get_level() const { return static_cast<unsigned int>(-1); } // dealii::numbers::invalid_unsigned_int = active level
  unsigned int get_dof_index() const { return 0; }
  unsigned int get_quad_index() const { return 0; }
  unsigned int get_degree() const { return (unsigned int)exadg_b200_degree(op); }
  void set_time(double const t) const { time = t; } // no time-dependent coefficient in the Laplace operator; kept for the interface
  double get_time() const { return time; }
  // order the operator's stream behind / ahead of a caller stream (cudaStream_t as void*)
  void wait_stream(void * cuda_stream) const { check(exadg_b200_wait_stream(op, cuda_stream)); }
  void stream_wait_operator(void * cuda_stream) const { check(exadg_b200_stream_wait_operator(op, cuda_stream)); }

  void initialize_dof_vector(VectorType & v) const
  {
    VectorType fresh;
    check(exadg_b200_initialize_dof_vector(op, &fresh.ptr));
    fresh.n = exadg_b200_local_size(op);
    v = std::move(fresh);
  }

  void calculate_diagonal(VectorType & diagonal) const
  {
    if (diagonal.locally_owned_size() == 0) initialize_dof_vector(diagonal);
    check(exadg_b200_calculate_diagonal(op, diagonal.data()));
  }
  void add_diagonal(VectorType & diagonal) const { check(exadg_b200_add_diagonal(op, diagonal.data())); }
  void calculate_inverse_diagonal(VectorType & diagonal) const
  {
    if (diagonal.locally_owned_size() == 0) initialize_dof_vector(diagonal);
    check(exadg_b200_calculate_inverse_diagonal(op, diagonal.data()));
  }

  // Krylov::KrylovSolver::solve with "cg" (iterative_solvers_dealii_wrapper.h:137-221); returns last_step()
  unsigned int solve_cg(VectorType & dst, VectorType const & rhs, int preconditioner, exadg_b200_chebyshev * cheb, double abs_tol, double rel_tol,
                        unsigned int max_iter) const
  {
    int n_iter = 0;
    check(exadg_b200_cg_solve(op, dst.data(), rhs.data(), preconditioner, cheb, abs_tol, rel_tol, (int)max_iter, &n_iter, nullptr));
    return (unsigned int)n_iter;
  }

  void synchronize() const { check(exadg_b200_synchronize(op)); }
  exadg_b200_operator * handle() const { return op; }

private:
  exadg_b200_operator * op = nullptr;
  mutable double time = 0.0; // operator_base.h:473
};

// MultigridPreconditionerBase on DG levels (I/solvers_and_preconditioners/multigrid/multigrid_preconditioner_base.cpp): the level
// operators are created by the caller (coarse -> fine, exadg_b200_multigrid_levels gives the list), vmult = one V-cycle
class MultigridPreconditioner
{
public:
  typedef DeviceVector VectorType;
  explicit MultigridPreconditioner(std::vector<LaplaceOperator const *> const & levels, int smoother_iterations = 5, double smoothing_range = 20.0,
                                   int iterations_eigenvalue_estimation = 20, double coarse_abs_tol = 1e-12, double coarse_rel_tol = 1e-3, int coarse_max_iter = 10000)
  {
    std::vector<exadg_b200_operator *> h;
    for (auto const * l : levels) h.push_back(l->handle());
    check(exadg_b200_multigrid_create((int)h.size(), h.data(), smoother_iterations, smoothing_range, iterations_eigenvalue_estimation, coarse_abs_tol, coarse_rel_tol,
                                      coarse_max_iter, &mg));
  }
  MultigridPreconditioner(MultigridPreconditioner const &) = delete;
  MultigridPreconditioner & operator=(MultigridPreconditioner const &) = delete;
  ~MultigridPreconditioner() { exadg_b200_multigrid_destroy(mg); }
  void vmult(VectorType & dst, VectorType const & src) const { check(exadg_b200_multigrid_vmult(mg, dst.data(), src.data())); }
  // Krylov::KrylovSolver::solve with "cg" and Preconditioner::Multigrid
  unsigned int solve_cg(LaplaceOperator const & A, VectorType & dst, VectorType const & rhs, double abs_tol, double rel_tol, unsigned int max_iter) const
  {
    int n_iter = 0;
    check(exadg_b200_cg_solve_multigrid(A.handle(), dst.data(), rhs.data(), mg, abs_tol, rel_tol, (int)max_iter, &n_iter, nullptr));
    return (unsigned int)n_iter;
  }
  exadg_b200_multigrid * handle() const { return mg; }

private:
  exadg_b200_multigrid * mg = nullptr;
};

} // namespace B200
} // namespace ExaDG
