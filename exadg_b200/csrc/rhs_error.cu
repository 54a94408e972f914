// Inhomogeneous boundary data, right-hand side and error norms on the GPU:
//   OperatorBase::rhs / rhs_add / evaluate / evaluate_add        I/operators/operator_base.cpp:509-606
//   boundary_face_loop_inhom_operator                             operator_base.cpp:1436-1518
//   calculate_exterior_value / calculate_exterior_normal_gradient I/operators/weak_boundary_conditions.h:72-134, 188-234
//   RHSOperator (volume source term (f, v))                       I/operators/rhs_operator.h, I/poisson/spatial_discretization/operator.cpp:414-423
//   calculate_error (relative L2 norm, Gauss(k+3))                I/postprocessor/error_calculation.cpp:36-115
// The reference evaluates dealii::Function objects at quadrature points; a C ABI cannot take functions, so the library hands out
// the physical coordinates of its quadrature points and takes the function values back as arrays (the reference-side binding
// evaluates its Function objects there).  None of this is on the vmult hot path: the kernels are written for
// clarity (geometry recomputed from the MappingQ(m) support points, fixed summation orders, no atomics) and run once per solve.
#include <stdexcept>

#include "operator.cuh"
#include "tables.hpp"

namespace exadg_b200
{
namespace
{
constexpr int PP_MAXQ = 10; // k + 3 quadrature points for k <= 7

struct PpTables
{
  int np;                   // mapping_degree + 1
  double gl[9];             // Gauss-Lobatto support points of MappingQ(m)
  int n;                    // k + 1 basis functions per direction
  int nq;                   // quadrature points per direction of this evaluation
  double xq[PP_MAXQ], w[PP_MAXQ];
  double S[PP_MAXQ * EXADG_MAX_N], D[PP_MAXQ * EXADG_MAX_N]; // l_j(x_q), l_j'(x_q)
  double fd[2][EXADG_MAX_N];                                 // l_j'(s), s = 0, 1 (l_j(s) is the Kronecker delta with the end node)
};

PpTables make_pp_tables(int degree, int mapping_degree, int nq)
{
  if (nq > PP_MAXQ || degree + 1 > EXADG_MAX_N) throw std::invalid_argument("quadrature too large");
  PpTables t;
  t.np = mapping_degree + 1; t.n = degree + 1; t.nq = nq;
  std::vector<real_t> gl, xn, xq, w, v, d;
  lobatto_points(t.np, gl);
  for (int i = 0; i < t.np; ++i) t.gl[i] = (double)gl[i];
  if (degree == 0) xn.assign(1, 0.5L); else lobatto_points(t.n, xn);
  gauss_points(nq, xq, w);
  for (int q = 0; q < nq; ++q) {
    t.xq[q] = (double)xq[q]; t.w[q] = (double)w[q];
    lagrange_at(xn, xq[q], v, d);
    for (int j = 0; j < t.n; ++j) { t.S[q * t.n + j] = (double)v[j]; t.D[q * t.n + j] = (double)d[j]; }
  }
  for (int s = 0; s < 2; ++s) {
    lagrange_at(xn, (real_t)s, v, d);
    for (int j = 0; j < t.n; ++j) t.fd[s][j] = (double)d[j];
  }
  return t;
}

__device__ inline void pp_lagrange(int n, const double * nodes, double x, double * v, double * d)
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

// position x and Jacobian J[i][j] = d x_i / d xi_j of the MappingQ(m) interpolant of cell c at xi
__device__ inline void pp_map(const PpTables & t, const double * __restrict__ xmap, int64_t c, const double xi[3], double x[3], double J[9])
{
  double v[3][9], d[3][9];
  for (int e = 0; e < 3; ++e) pp_lagrange(t.np, t.gl, xi[e], v[e], d[e]);
  for (int i = 0; i < 9; ++i) J[i] = 0.0;
  x[0] = x[1] = x[2] = 0.0;
  const int np = t.np;
  const double * X = xmap + (size_t)c * np * np * np * 3;
  for (int a2 = 0; a2 < np; ++a2) for (int a1 = 0; a1 < np; ++a1) for (int a0 = 0; a0 < np; ++a0) {
    const double * p = X + (a0 + np * (a1 + np * a2)) * 3;
    const double N = v[0][a0] * v[1][a1] * v[2][a2];
    const double g0 = d[0][a0] * v[1][a1] * v[2][a2], g1 = v[0][a0] * d[1][a1] * v[2][a2], g2 = v[0][a0] * v[1][a1] * d[2][a2];
    for (int i = 0; i < 3; ++i) { x[i] += p[i] * N; J[i * 3 + 0] += p[i] * g0; J[i * 3 + 1] += p[i] * g1; J[i * 3 + 2] += p[i] * g2; }
  }
}
__device__ inline double pp_det3(const double * J)
{
  return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
}
__device__ inline void pp_inv3(const double * J, double det, double * Ji)
{
  const double id = 1.0 / det;
  Ji[0] = (J[4] * J[8] - J[5] * J[7]) * id; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * id; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * id;
  Ji[3] = (J[5] * J[6] - J[3] * J[8]) * id; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * id; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * id;
  Ji[6] = (J[3] * J[7] - J[4] * J[6]) * id; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * id; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * id;
}

// cell quadrature points: xyz[cell][q][3], jxw[cell][q]   (q = q0 + nq (q1 + nq q2))
__global__ void cell_points_kernel(PpTables t, const double * __restrict__ xmap, int64_t n_cells, double * __restrict__ xyz, double * __restrict__ jxw)
{
  const int nq = t.nq, nq3 = nq * nq * nq;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_cells * nq3) return;
  const int64_t c = idx / nq3; const int q = (int)(idx % nq3);
  const int q0 = q % nq, q1 = (q / nq) % nq, q2 = q / (nq * nq);
  const double xi[3] = {t.xq[q0], t.xq[q1], t.xq[q2]};
  double x[3], J[9];
  pp_map(t, xmap, c, xi, x, J);
  if (xyz) for (int i = 0; i < 3; ++i) xyz[idx * 3 + i] = x[i];
  if (jxw) jxw[idx] = pp_det3(J) * t.w[q0] * t.w[q1] * t.w[q2];
}

// boundary-face quadrature points of face b = (cell, face number): position, JxW, and the coefficients of the normal derivative
// in reference coordinates, d_n phi = sum_e cn[e] d phi / d xi_e with cn = J^-1 n (n = outward unit normal)
__global__ void bface_points_kernel(PpTables t, const double * __restrict__ xmap, const int32_t * __restrict__ bf_cell, const uint8_t * __restrict__ bf_face, int64_t n_bf,
                                    double * __restrict__ xyz, double * __restrict__ jxw, double * __restrict__ cn)
{
  const int nq = t.nq, nq2 = nq * nq;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_bf * nq2) return;
  const int64_t b = idx / nq2; const int q = (int)(idx % nq2);
  const int qa = q % nq, qb = q / nq;
  const int f = bf_face[b], d = f >> 1, s = f & 1, t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
  double xi[3]; xi[d] = (double)s; xi[t1] = t.xq[qa]; xi[t2] = t.xq[qb];
  double x[3], J[9], Ji[9];
  pp_map(t, xmap, bf_cell[b], xi, x, J);
  const double det = pp_det3(J); pp_inv3(J, det, Ji);
  const double sgn = s ? 1.0 : -1.0;
  double nv[3] = {Ji[d * 3], Ji[d * 3 + 1], Ji[d * 3 + 2]};
  const double len = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
  for (int i = 0; i < 3; ++i) nv[i] *= sgn / len;
  for (int i = 0; i < 3; ++i) xyz[idx * 3 + i] = x[i];
  jxw[idx] = fabs(det) * len * t.w[qa] * t.w[qb];
  for (int e = 0; e < 3; ++e) cn[idx * 3 + e] = Ji[e * 3] * nv[0] + Ji[e * 3 + 1] * nv[1] + Ji[e * 3 + 2] * nv[2];
}

// inhomogeneous part of the boundary integrals (weak_boundary_conditions.h: interior value and normal gradient are zero for
// OperatorType::inhomogeneous; Dirichlet: u+ = 2 g, d_n u+ = 0; Neumann: u+ = 0, d_n u+ = 2 h), fluxes as in
// laplace_operator.h:180-197, tested like do_boundary_integral (laplace_operator.cpp:221-265):
//   tmp_i = sum_q [ d_n phi_i * gradient_flux - phi_i * value_flux ] JxW,   dst_i += sign * tmp_i
// One thread per (boundary cell, local DoF): the cell's boundary faces are visited in face order (deterministic, no atomics).
__global__ void boundary_inhom_kernel(PpTables t, const int32_t * __restrict__ bc_cell, const int32_t * __restrict__ bc_face_index /*[n_bc][6]*/, int64_t n_bc,
                                      const uint8_t * __restrict__ bf_type, const double * __restrict__ bf_tau, const double * __restrict__ jxw,
                                      const double * __restrict__ cn, const double * __restrict__ values, double sign, double * __restrict__ dst)
{
  const int n = t.n, n3 = n * n * n, nq = t.nq, nq2 = nq * nq;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_bc * n3) return;
  const int64_t bc = idx / n3; const int i = (int)(idx % n3);
  const int ii[3] = {i % n, (i / n) % n, i / (n * n)};
  double acc = 0.0;
  for (int f = 0; f < 6; ++f) {
    const int32_t b = bc_face_index[bc * 6 + f];
    if (b < 0) continue;
    const int d = f >> 1, s = f & 1, t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
    const int end = s ? n - 1 : 0;
    const double phi_d = (ii[d] == end) ? 1.0 : 0.0; // Gauss-Lobatto nodal basis: l_j(s) = delta
    const double dphi_d = t.fd[s][ii[d]];
    const double tau = bf_tau[b];
    const bool dirichlet = bf_type[b] == BT_DIRICHLET;
    for (int qb = 0; qb < nq; ++qb)
      for (int qa = 0; qa < nq; ++qa) {
        const int64_t p = (int64_t)b * nq2 + qa + nq * qb;
        const double g = values[p];
        const double vp = dirichlet ? 2.0 * g : 0.0, dp = dirichlet ? 0.0 : 2.0 * g; // exterior value / normal gradient, interior ones are 0
        const double gradient_flux = 0.5 * vp;                   // -1/2 (u- - u+)
        const double value_flux = 0.5 * dp + tau * vp;           // 1/2 (dn u- + dn u+) - tau (u- - u+)
        const double sa = t.S[qa * n + ii[t1]], sb = t.S[qb * n + ii[t2]];
        const double da = t.D[qa * n + ii[t1]], db = t.D[qb * n + ii[t2]];
        double gref[3];
        gref[d] = dphi_d * sa * sb; gref[t1] = phi_d * da * sb; gref[t2] = phi_d * sa * db;
        const double dn_phi = cn[p * 3] * gref[0] + cn[p * 3 + 1] * gref[1] + cn[p * 3 + 2] * gref[2];
        acc += (dn_phi * gradient_flux - phi_d * sa * sb * value_flux) * jxw[p];
      }
  }
  dst[(int64_t)bc_cell[bc] * n3 + i] += sign * acc;
}

// dst_i += sum_q phi_i(x_q) f_q JxW_q   (RHSOperator); one thread per (cell, local DoF), q in fixed order
__global__ void source_kernel(PpTables t, int64_t n_cells, const double * __restrict__ jxw, const double * __restrict__ f, double * __restrict__ dst)
{
  const int n = t.n, n3 = n * n * n, nq = t.nq;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_cells * n3) return;
  const int64_t c = idx / n3; const int i = (int)(idx % n3);
  const int i0 = i % n, i1 = (i / n) % n, i2 = i / (n * n);
  const int64_t base = c * nq * nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    double a2 = 0.0;
    for (int q1 = 0; q1 < nq; ++q1) {
      double a1 = 0.0;
      for (int q0 = 0; q0 < nq; ++q0) { const int64_t p = base + q0 + nq * (q1 + nq * q2); a1 += t.S[q0 * n + i0] * f[p] * jxw[p]; }
      a2 += t.S[q1 * n + i1] * a1;
    }
    acc += t.S[q2 * n + i2] * a2;
  }
  dst[idx] += acc;
}

// per-cell squared L2 norms of (u_h - u_exact) and u_exact with Gauss(nq): out[cell], out[n_cells + cell]
__global__ void l2_cell_kernel(PpTables t, int64_t n_cells, const double * __restrict__ jxw, const double * __restrict__ u, const double * __restrict__ exact, double * __restrict__ out)
{
  const int n = t.n, n3 = n * n * n, nq = t.nq;
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const double * uc = u + c * n3;
  double e2 = 0.0, n2 = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) for (int q1 = 0; q1 < nq; ++q1) for (int q0 = 0; q0 < nq; ++q0) {
    double v = 0.0;
    for (int i2 = 0; i2 < n; ++i2) {
      double a2 = 0.0;
      for (int i1 = 0; i1 < n; ++i1) {
        double a1 = 0.0;
        for (int i0 = 0; i0 < n; ++i0) a1 += t.S[q0 * n + i0] * uc[i0 + n * (i1 + n * i2)];
        a2 += t.S[q1 * n + i1] * a1;
      }
      v += t.S[q2 * n + i2] * a2;
    }
    const int64_t p = c * nq * nq * nq + q0 + nq * (q1 + nq * q2);
    const double ex = exact[p], df = v - ex, w = jxw[p];
    e2 += df * df * w; n2 += ex * ex * w;
  }
  out[c] = e2; out[n_cells + c] = n2;
}
} // namespace

// ---- host side: state kept per operator ----
struct PostData
{
  double * d_xmap = nullptr;                     // [(owned)][(m+1)^3][3]
  // boundary faces of the owned cells, cell-major / face-minor
  int64_t n_bf = 0, n_bc = 0;
  std::vector<int32_t> bf_cell; std::vector<uint8_t> bf_face, bf_type;
  int32_t * d_bf_cell = nullptr; uint8_t * d_bf_face = nullptr, * d_bf_type = nullptr;
  int32_t * d_bc_cell = nullptr, * d_bc_face_index = nullptr;
  double * d_bf_xyz = nullptr, * d_bf_jxw = nullptr, * d_bf_cn = nullptr, * d_bf_tau = nullptr, * d_bf_values = nullptr;
  bool have_values = false;
};

template<typename T>
static T * pp_upload(const std::vector<T> & v)
{
  T * p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

void post_destroy(void * p)
{
  PostData * P = static_cast<PostData *>(p);
  if (!P) return;
  cudaFree(P->d_xmap); cudaFree(P->d_bf_cell); cudaFree(P->d_bf_face); cudaFree(P->d_bf_type); cudaFree(P->d_bc_cell); cudaFree(P->d_bc_face_index);
  cudaFree(P->d_bf_xyz); cudaFree(P->d_bf_jxw); cudaFree(P->d_bf_cn); cudaFree(P->d_bf_tau); cudaFree(P->d_bf_values);
  delete P;
}

// builds the boundary-face tables and their geometry on first use
void * post_get(void *& slot, const DeviceOperator & op, const HostMesh & mesh, double penalty_factor, cudaStream_t stream)
{
  if (slot) return slot;
  if (mesh.xmap.empty()) throw std::runtime_error("mapping support points are not available for this operator");
  PostData * P = new PostData;
  slot = P;
  const int np3 = (mesh.mapping_degree + 1) * (mesh.mapping_degree + 1) * (mesh.mapping_degree + 1);
  std::vector<double> owned(mesh.xmap.begin(), mesh.xmap.begin() + (size_t)mesh.n_owned * np3 * 3);
  P->d_xmap = pp_upload(owned);
  std::vector<int32_t> bc_cell, bc_face_index;
  for (int64_t c = 0; c < mesh.n_owned; ++c) {
    bool any = false;
    for (int f = 0; f < 6; ++f) any |= (mesh.bt[c * 6 + f] != BT_INTERIOR);
    if (!any) continue;
    bc_cell.push_back((int32_t)c);
    for (int f = 0; f < 6; ++f) {
      if (mesh.bt[c * 6 + f] != BT_INTERIOR) {
        bc_face_index.push_back((int32_t)P->bf_cell.size());
        P->bf_cell.push_back((int32_t)c); P->bf_face.push_back((uint8_t)f); P->bf_type.push_back(mesh.bt[c * 6 + f]);
      } else bc_face_index.push_back(-1);
    }
  }
  P->n_bf = (int64_t)P->bf_cell.size(); P->n_bc = (int64_t)bc_cell.size();
  P->d_bf_cell = pp_upload(P->bf_cell); P->d_bf_face = pp_upload(P->bf_face); P->d_bf_type = pp_upload(P->bf_type);
  P->d_bc_cell = pp_upload(bc_cell); P->d_bc_face_index = pp_upload(bc_face_index);
  const int nq2 = op.n * op.n;
  const size_t npts = (size_t)std::max<int64_t>(P->n_bf, 1) * nq2;
  CUDA_CHECK(cudaMalloc(&P->d_bf_xyz, npts * 3 * sizeof(double))); CUDA_CHECK(cudaMalloc(&P->d_bf_jxw, npts * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&P->d_bf_cn, npts * 3 * sizeof(double))); CUDA_CHECK(cudaMalloc(&P->d_bf_values, npts * sizeof(double)));
  CUDA_CHECK(cudaMemsetAsync(P->d_bf_values, 0, npts * sizeof(double), stream));
  // tau of a boundary face: tau_K (k+1)^2 IP_factor of its cell (laplace_operator.h:142-151); tau_f of the operator holds exactly that
  std::vector<double> tau(std::max<int64_t>(P->n_bf, 1), 0.0);
  if (P->n_bf > 0) {
    if (!op.tau_f) throw std::runtime_error("boundary faces without stored penalty parameters");
    std::vector<double> tau_f(mesh.n_faces);
    CUDA_CHECK(cudaMemcpy(tau_f.data(), op.tau_f, (size_t)mesh.n_faces * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < P->n_bf; ++b) tau[b] = tau_f[mesh.face_id[(int64_t)P->bf_cell[b] * 6 + P->bf_face[b]]];
  }
  (void)penalty_factor;
  P->d_bf_tau = pp_upload(tau);
  if (P->n_bf > 0) {
    const PpTables t = make_pp_tables(op.degree, mesh.mapping_degree, op.n);
    const int64_t total = P->n_bf * nq2;
    bface_points_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(t, P->d_xmap, P->d_bf_cell, P->d_bf_face, P->n_bf, P->d_bf_xyz, P->d_bf_jxw, P->d_bf_cn);
    CUDA_CHECK(cudaGetLastError());
  }
  return slot;
}

int64_t post_n_boundary_faces(void * p) { return static_cast<PostData *>(p)->n_bf; }

void post_boundary_points(void * p, const DeviceOperator & op, double * xyz_host, uint8_t * type_host, cudaStream_t stream)
{
  PostData * P = static_cast<PostData *>(p);
  const size_t npts = (size_t)P->n_bf * op.n * op.n;
  if (xyz_host && npts) { CUDA_CHECK(cudaMemcpyAsync(xyz_host, P->d_bf_xyz, npts * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream)); CUDA_CHECK(cudaStreamSynchronize(stream)); }
  if (type_host) std::copy(P->bf_type.begin(), P->bf_type.end(), type_host);
}

void post_set_boundary_values(void * p, const DeviceOperator & op, const double * values_host, cudaStream_t stream)
{
  PostData * P = static_cast<PostData *>(p);
  const size_t npts = (size_t)P->n_bf * op.n * op.n;
  if (npts) { CUDA_CHECK(cudaMemcpyAsync(P->d_bf_values, values_host, npts * sizeof(double), cudaMemcpyHostToDevice, stream)); CUDA_CHECK(cudaStreamSynchronize(stream)); }
  P->have_values = true;
}

// dst += sign * (inhomogeneous boundary integrals); rhs_add uses sign = -1 (operator_base.cpp:533-535), evaluate_add sign = +1
void post_boundary_inhom_add(void * p, const DeviceOperator & op, const HostMesh & mesh, double sign, double * dst, cudaStream_t stream)
{
  PostData * P = static_cast<PostData *>(p);
  if (P->n_bc == 0) return;
  const PpTables t = make_pp_tables(op.degree, mesh.mapping_degree, op.n);
  const int64_t total = P->n_bc * op.n * op.n * op.n;
  boundary_inhom_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(t, P->d_bc_cell, P->d_bc_face_index, P->n_bc, P->d_bf_type, P->d_bf_tau, P->d_bf_jxw, P->d_bf_cn,
                                                                            P->d_bf_values, sign, dst);
  CUDA_CHECK(cudaGetLastError());
}

void post_cell_points(void * p, const DeviceOperator & op, const HostMesh & mesh, int nq, double * xyz_host, cudaStream_t stream)
{
  PostData * P = static_cast<PostData *>(p);
  const PpTables t = make_pp_tables(op.degree, mesh.mapping_degree, nq);
  const int64_t total = mesh.n_owned * nq * nq * nq;
  if (total == 0) return;
  double * d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, (size_t)total * 3 * sizeof(double)));
  cell_points_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(t, P->d_xmap, mesh.n_owned, d, nullptr);
  CUDA_CHECK(cudaMemcpyAsync(xyz_host, d, (size_t)total * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  cudaFree(d);
}

// dst += (f, phi_i), f given at the Gauss(k+1) points of the owned cells (host array)
void post_source_add(void * p, const DeviceOperator & op, const HostMesh & mesh, const double * f_host, double * dst, cudaStream_t stream)
{
  PostData * P = static_cast<PostData *>(p);
  const int nq = op.n;
  const PpTables t = make_pp_tables(op.degree, mesh.mapping_degree, nq);
  const int64_t npts = mesh.n_owned * nq * nq * nq;
  if (npts == 0) return;
  double * d_f = nullptr, * d_jxw = nullptr;
  CUDA_CHECK(cudaMalloc(&d_f, (size_t)npts * sizeof(double))); CUDA_CHECK(cudaMalloc(&d_jxw, (size_t)npts * sizeof(double)));
  CUDA_CHECK(cudaMemcpyAsync(d_f, f_host, (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, stream));
  cell_points_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, stream>>>(t, P->d_xmap, mesh.n_owned, nullptr, d_jxw);
  const int64_t total = mesh.n_owned * op.n * op.n * op.n;
  source_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(t, mesh.n_owned, d_jxw, d_f, dst);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(stream));
  cudaFree(d_f); cudaFree(d_jxw);
}

// per-cell squared norms -> d_out[0 .. n) (error), [n .. 2n) (exact solution); caller reduces deterministically
void post_l2_cells(void * p, const DeviceOperator & op, const HostMesh & mesh, int nq, const double * u, const double * exact_host, double * d_out, cudaStream_t stream)
{
  PostData * P = static_cast<PostData *>(p);
  const PpTables t = make_pp_tables(op.degree, mesh.mapping_degree, nq);
  const int64_t npts = mesh.n_owned * nq * nq * nq;
  if (npts == 0) return;
  double * d_ex = nullptr, * d_jxw = nullptr;
  CUDA_CHECK(cudaMalloc(&d_ex, (size_t)npts * sizeof(double))); CUDA_CHECK(cudaMalloc(&d_jxw, (size_t)npts * sizeof(double)));
  CUDA_CHECK(cudaMemcpyAsync(d_ex, exact_host, (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, stream));
  cell_points_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, stream>>>(t, P->d_xmap, mesh.n_owned, nullptr, d_jxw);
  l2_cell_kernel<<<(unsigned)((mesh.n_owned + 63) / 64), 64, 0, stream>>>(t, mesh.n_owned, d_jxw, u, d_ex, d_out);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(stream));
  cudaFree(d_ex); cudaFree(d_jxw);
}

} // namespace exadg_b200
