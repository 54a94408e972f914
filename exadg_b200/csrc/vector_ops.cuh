// Fused vector kernels for the callers of vmult (8a row a12):
//   dealii::SolverCG as driven by Krylov::KrylovSolver::solve
//     (I/solvers_and_preconditioners/solvers/iterative_solvers_dealii_wrapper.h:137-221),
//   JacobiPreconditioner::vmult (preconditioners/jacobi_preconditioner.h:50-62),
//   dealii::PreconditionChebyshev as configured by ChebyshevSmoother
//     (multigrid/smoothers/chebyshev_smoother.h:149-172),
//   invert_diagonal (utilities/invert_diagonal.h:35-46).
// All reductions are deterministic: fixed grid, per-CTA partial sums in a fixed slot, and a final
// single-CTA pass in fixed order (no atomics), so residual histories are reproducible run to run.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace exadg_b200
{
constexpr int RED_BLOCKS = 592; // 4 CTAs per SM on 148 SMs
constexpr int RED_THREADS = 256;

struct Reducer
{
  double * partial = nullptr; // [3][RED_BLOCKS]
  double * result = nullptr;  // [8] device scalars
  double * host = nullptr;    // pinned mirror of result
};

void reducer_init(Reducer & r);
void reducer_free(Reducer & r);

// result[slot] = sum a_i b_i
void dot(const Reducer & r, int slot, const double * a, const double * b, int64_t n, cudaStream_t s);
// x += alpha d ; g += alpha h ; result[slot] = g.g     with alpha = result[num]/result[den] (on device)
void cg_update_x_g(const Reducer & r, int slot, int num, int den, double * x, const double * d, double * g, const double * h, int64_t n, cudaStream_t s);
// z = inv_diag * g ; result[slot] = g.z
void jacobi_dot(const Reducer & r, int slot, double * z, const double * inv_diag, const double * g, int64_t n, cudaStream_t s);
// d = beta d - z,  beta = result[num]/result[den]
void cg_update_d(const Reducer & r, int num, int den, double * d, const double * z, int64_t n, cudaStream_t s);
// y = a x + b y (generic sadd), y = a x
void axpby(double a, const double * x, double b, double * y, int64_t n, cudaStream_t s);
void scale_copy(double a, const double * x, double * y, int64_t n, cudaStream_t s);
void fill(double * x, double v, int64_t n, cudaStream_t s);
// invert_diagonal.h:41-45
void invert_diagonal(double * d, int64_t n, cudaStream_t s);
// Chebyshev: first update  x = x0 + (1/theta) P^-1 (b - r)   [zero start: x0 = 0, r = 0]; xold = x0
void cheb_first(double * x, double * xold, const double * inv_diag, const double * b, const double * r, double inv_theta, bool zero_start, int64_t n, cudaStream_t s);
// x_{j+1} = x_j + f1 (x_j - x_{j-1}) + f2 P^-1 (b - A x_j) ; xold = x_j      (one pass, 5 reads + 2 writes)
void cheb_step(double * x, double * xold, const double * inv_diag, const double * b, const double * r, double f1, double f2, int64_t n, cudaStream_t s);
// start vector of PreconditionChebyshev's eigenvalue estimate: (global index mod 11)
void fill_mod11(double * x, int64_t global_offset, int64_t n, cudaStream_t s);
void add_scalar(double * x, double a, int64_t n, cudaStream_t s);
// two-level transfers of the DG multigrid hierarchy (dealii::MGTwoLevelTransfer as set up by
// I/solvers_and_preconditioners/multigrid/transfer.cpp:28-69), prolongation = embedding, restriction = its transpose, both add into dst:
//   p-transfer (h = 0): FE_DGQ(nc - 1) -> FE_DGQ(nf - 1) on the same cells; I[0] = [nf][nc] coarse Lagrange basis at the fine Gauss-Lobatto nodes;
//   h-transfer (h = 1): parent cell c -> children 8 c + (x + 2 y + 4 z) of a global refinement, nf = nc = n; I[b] = [n][n] parent basis at the
//   nodes of the child covering the lower (b = 0) / upper (b = 1) half of the parent interval.
struct TransferTable { int nf, nc, h; double I[2][8 * 8]; };
void prolongate_add(const TransferTable & t, double * fine, const double * coarse, int64_t n_coarse_cells, cudaStream_t s);
void restrict_add(const TransferTable & t, double * coarse, const double * fine, int64_t n_coarse_cells, cudaStream_t s);
// sum of entries -> result[slot]
void sum(const Reducer & r, int slot, const double * a, int64_t n, cudaStream_t s);

} // namespace exadg_b200
