// Geometry / penalty setup kernels: the device analogue of deal.II's MappingInfo
// (inverse Jacobians, JxW, normals) and of IP::calculate_penalty_parameter
// (I/operators/interior_penalty_parameter.h:43-99).  Runs once per operator.
//
// Stored metrics (what LaplaceKernel::get_mapping_flags asks for, laplace_operator.h:102-126):
//   cells: symmetric G = J^-1 J^-T |J| w_q                      6 doubles / quadrature point
//   faces (unique): a_minus = J_minus^-1 n, a_plus = J_plus^-1 n, JxW   7 doubles / face quadrature point
// so that  grad_xi-test . (G grad_xi u)  is the cell integrand (laplace_operator.cpp:129-137) and
// dn u = a . grad_xi u is get_normal_derivative (laplace_operator.cpp:149-150).
#include <cstdio>
#include <mutex>
#include <set>
#include <stdexcept>
#include <utility>

#include "operator.cuh"

namespace exadg_b200
{
void cuda_check(cudaError_t e, const char * what)
{
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}

bool first_use_on_device(const void * key)
{
  static std::mutex mutex;
  static std::set<std::pair<const void *, int>> seen;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mutex);
  return seen.insert(std::make_pair(key, dev)).second;
}

namespace
{
struct MapTables
{
  int np;          // mapping_degree + 1
  double gl[9];    // Gauss-Lobatto support points of MappingQ(m)
  int nq;          // k+1
  double xq[EXADG_MAX_N], w[EXADG_MAX_N];
};

__device__ inline void lagrange_dev(int n, const double * nodes, double x, double * v, double * d)
{
  for (int j = 0; j < n; ++j) {
    double pv = 1.0, pd = 0.0;
    for (int i = 0; i < n; ++i) if (i != j) pv *= (x - nodes[i]) / (nodes[j] - nodes[i]);
    for (int m = 0; m < n; ++m) if (m != j) {
      double t = 1.0 / (nodes[j] - nodes[m]);
      for (int i = 0; i < n; ++i) if (i != j && i != m) t *= (x - nodes[i]) / (nodes[j] - nodes[i]);
      pd += t;
    }
    v[j] = pv; d[j] = pd;
  }
}

// J[i][j] = d x_i / d xi_j of the MappingQ(m) interpolant of cell `c` at xi
__device__ inline void jacobian_dev(const MapTables & t, const double * __restrict__ xmap, int64_t c, const double xi[3], double J[9])
{
  double v[3][9], d[3][9];
  for (int e = 0; e < 3; ++e) lagrange_dev(t.np, t.gl, xi[e], v[e], d[e]);
  for (int i = 0; i < 9; ++i) J[i] = 0.0;
  const int np = t.np;
  const double * X = xmap + (size_t)c * np * np * np * 3;
  for (int a2 = 0; a2 < np; ++a2) for (int a1 = 0; a1 < np; ++a1) for (int a0 = 0; a0 < np; ++a0) {
    const double * p = X + (a0 + np * (a1 + np * a2)) * 3;
    const double g0 = d[0][a0] * v[1][a1] * v[2][a2], g1 = v[0][a0] * d[1][a1] * v[2][a2], g2 = v[0][a0] * v[1][a1] * d[2][a2];
    for (int i = 0; i < 3; ++i) { J[i * 3 + 0] += p[i] * g0; J[i * 3 + 1] += p[i] * g1; J[i * 3 + 2] += p[i] * g2; }
  }
}

__device__ inline double det3(const double * J)
{
  return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
}
// Ji[e][i] = d xi_e / d x_i
__device__ inline void inv3(const double * J, double det, double * Ji)
{
  const double id = 1.0 / det;
  Ji[0] = (J[4] * J[8] - J[5] * J[7]) * id; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * id; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * id;
  Ji[3] = (J[5] * J[6] - J[3] * J[8]) * id; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * id; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * id;
  Ji[6] = (J[3] * J[7] - J[4] * J[6]) * id; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * id; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * id;
}

__device__ inline void face_xi(const MapTables & t, int f, int qa, int qb, double xi[3])
{
  const int d = f >> 1, s = f & 1, t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
  xi[d] = (double)s; xi[t1] = t.xq[qa]; xi[t2] = t.xq[qb];
}

// tau_K = surface / volume, weights 1/2 on interior (and periodic) faces, 1 on true boundary faces
// (interior_penalty_parameter.h:68-98).  One thread per locally relevant cell, fixed summation order.
__global__ void tau_kernel(MapTables t, const double * __restrict__ xmap, const uint8_t * __restrict__ bt, int64_t n_cells, double * __restrict__ tau)
{
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int nq = t.nq;
  double volume = 0.0, surface = 0.0, J[9], xi[3];
  for (int q2 = 0; q2 < nq; ++q2) for (int q1 = 0; q1 < nq; ++q1) for (int q0 = 0; q0 < nq; ++q0) {
    xi[0] = t.xq[q0]; xi[1] = t.xq[q1]; xi[2] = t.xq[q2];
    jacobian_dev(t, xmap, c, xi, J);
    volume += det3(J) * t.w[q0] * t.w[q1] * t.w[q2];
  }
  for (int f = 0; f < 6; ++f) {
    const double factor = bt[c * 6 + f] != BT_INTERIOR ? 1.0 : 0.5;
    const int d = f >> 1;
    for (int qb = 0; qb < nq; ++qb) for (int qa = 0; qa < nq; ++qa) {
      face_xi(t, f, qa, qb, xi);
      jacobian_dev(t, xmap, c, xi, J);
      const double det = det3(J); double Ji[9]; inv3(J, det, Ji);
      const double len = sqrt(Ji[d * 3] * Ji[d * 3] + Ji[d * 3 + 1] * Ji[d * 3 + 1] + Ji[d * 3 + 2] * Ji[d * 3 + 2]);
      surface += fabs(det) * len * t.w[qa] * t.w[qb] * factor;
    }
  }
  tau[c] = surface / volume;
}

__global__ void cell_metric_kernel(MapTables t, const double * __restrict__ xmap, int64_t n_cells, double * __restrict__ cellG, double * __restrict__ cellJxW)
{
  const int nq = t.nq, nq3 = nq * nq * nq;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_cells * nq3) return;
  const int64_t c = idx / nq3; const int q = (int)(idx % nq3);
  const int q0 = q % nq, q1 = (q / nq) % nq, q2 = q / (nq * nq);
  double xi[3] = {t.xq[q0], t.xq[q1], t.xq[q2]}, J[9], Ji[9];
  jacobian_dev(t, xmap, c, xi, J);
  const double det = det3(J); inv3(J, det, Ji);
  const double jxw = det * t.w[q0] * t.w[q1] * t.w[q2];
  if (cellJxW) cellJxW[idx] = jxw;
  // G[e][g] = sum_i Ji[e][i] Ji[g][i] * JxW
  double * out = cellG + (size_t)c * 6 * nq3 + q;
  const int e1[6] = {0, 1, 2, 0, 0, 1}, e2[6] = {0, 1, 2, 1, 2, 2};
  for (int k = 0; k < 6; ++k) {
    const double * a = Ji + e1[k] * 3, * b = Ji + e2[k] * 3;
    out[(size_t)k * nq3] = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) * jxw;
  }
}

__global__ void face_metric_kernel(MapTables t, const double * __restrict__ xmap, const double * __restrict__ tau_cell,
                                   const int32_t * __restrict__ face_cells, const uint8_t * __restrict__ face_nos, const uint8_t * __restrict__ face_bt,
                                   int64_t n_faces, double penalty_factor, double * __restrict__ faceG, double * __restrict__ tau_f)
{
  const int nq = t.nq, nq2 = nq * nq;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_faces * nq2) return;
  const int64_t F = idx / nq2; const int q = (int)(idx % nq2);
  const int qa = q % nq, qb = q / nq;
  const int64_t cm = face_cells[F * 2], cp = face_cells[F * 2 + 1];
  const int fm = face_nos[F * 2], fp = face_nos[F * 2 + 1];
  double xi[3], J[9], Ji[9];
  face_xi(t, fm, qa, qb, xi);
  jacobian_dev(t, xmap, cm, xi, J);
  const double det = det3(J); inv3(J, det, Ji);
  const int d = fm >> 1; const double sgn = (fm & 1) ? 1.0 : -1.0;
  double nv[3] = {Ji[d * 3], Ji[d * 3 + 1], Ji[d * 3 + 2]};
  const double len = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
  for (int i = 0; i < 3; ++i) nv[i] *= sgn / len; // outward unit normal of the minus side
  double * out = faceG + (size_t)F * 7 * nq2 + q;
  for (int e = 0; e < 3; ++e) out[(size_t)e * nq2] = Ji[e * 3] * nv[0] + Ji[e * 3 + 1] * nv[1] + Ji[e * 3 + 2] * nv[2];
  if (cp >= 0) {
    face_xi(t, fp, qa, qb, xi);
    jacobian_dev(t, xmap, cp, xi, J);
    const double detp = det3(J); inv3(J, detp, Ji);
    for (int e = 0; e < 3; ++e) out[(size_t)(3 + e) * nq2] = Ji[e * 3] * nv[0] + Ji[e * 3 + 1] * nv[1] + Ji[e * 3 + 2] * nv[2];
  } else {
    for (int e = 0; e < 3; ++e) out[(size_t)(3 + e) * nq2] = 0.0;
  }
  out[(size_t)6 * nq2] = fabs(det) * len * t.w[qa] * t.w[qb];
  if (q == 0) {
    // laplace_operator.h:128-151: max of both sides on interior faces, own value on boundary faces
    const double tk = (cp >= 0 && face_bt[F] == BT_INTERIOR) ? fmax(tau_cell[cm], tau_cell[cp]) : tau_cell[cm];
    tau_f[F] = tk * penalty_factor;
  }
}

template<typename T>
T * to_device(const std::vector<T> & v)
{
  T * p = nullptr;
  if (v.empty()) return p;
  CUDA_CHECK(cudaMalloc(&p, v.size() * sizeof(T)));
  CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}
} // namespace

void setup_geometry(DeviceOperator & op, const HostMesh & mesh, double ip_factor, cudaStream_t stream)
{
  const int n = op.n;
  const int64_t nloc = mesh.n_owned + mesh.n_ghost;
  Tables1D tab(op.degree);
  MapTables t;
  t.np = mesh.mapping_degree + 1; t.nq = n;
  std::vector<real_t> gl; lobatto_points(t.np, gl);
  for (int i = 0; i < t.np; ++i) t.gl[i] = (double)gl[i];
  for (int i = 0; i < n; ++i) { t.xq[i] = (double)tab.xq[i]; t.w[i] = (double)tab.w[i]; }
  // interior_penalty_parameter.h:124 (hypercube elements)
  const double penalty_factor = ip_factor * (op.degree + 1.0) * (op.degree + 1.0);

  op.nb = to_device(mesh.nb);
  op.face_id = to_device(mesh.face_id);
  op.face_info = to_device(mesh.face_info);
  if (mesh.n_ghost > 0) CUDA_CHECK(cudaMalloc(&op.ghost, (size_t)mesh.n_ghost * n * n * n * op.n_components * sizeof(double))); // [ghost cell][component][n^3]

  if (op.cartesian) {
    // uniform box: tau_K = sum_d 1/h_d (all faces weighted 1/2), same for every cell
    double tk = 0.0;
    for (int e = 0; e < 3; ++e) tk += 1.0 / mesh.h[e];
    op.tau_hat = tk * penalty_factor; // multiplied by h_d per direction in the kernel parameters
    return;
  }

  if (nloc == 0 || mesh.n_owned == 0) return; // a rank without cells (more ranks than cells): nothing to set up
  double * xmap = to_device(mesh.xmap);
  uint8_t * bt = to_device(mesh.bt);
  int32_t * face_cells = to_device(mesh.face_cells);
  uint8_t * face_nos = to_device(mesh.face_nos);
  uint8_t * face_bt = to_device(mesh.face_bt);
  const int nq2 = n * n, nq3 = nq2 * n;
  CUDA_CHECK(cudaMalloc(&op.tau_cell, (size_t)nloc * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&op.cellG, (size_t)mesh.n_owned * 6 * nq3 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&op.faceG, (size_t)mesh.n_faces * 7 * nq2 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&op.tau_f, (size_t)mesh.n_faces * sizeof(double)));
  const int B = 128;
  tau_kernel<<<(unsigned)((nloc + B - 1) / B), B, 0, stream>>>(t, xmap, bt, nloc, op.tau_cell);
  if (op.helmholtz) CUDA_CHECK(cudaMalloc(&op.cellJxW, (size_t)mesh.n_owned * nq3 * sizeof(double)));
  cell_metric_kernel<<<(unsigned)((mesh.n_owned * nq3 + B - 1) / B), B, 0, stream>>>(t, xmap, mesh.n_owned, op.cellG, op.cellJxW);
  face_metric_kernel<<<(unsigned)((mesh.n_faces * nq2 + B - 1) / B), B, 0, stream>>>(t, xmap, op.tau_cell, face_cells, face_nos, face_bt, mesh.n_faces,
                                                                                     penalty_factor, op.faceG, op.tau_f);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(stream));
  cudaFree(xmap); cudaFree(bt); cudaFree(face_cells); cudaFree(face_nos); cudaFree(face_bt);
}

} // namespace exadg_b200
