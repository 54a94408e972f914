// Device-side data of the SIPG Laplace operator: the GPU replacement for what
// dealii::MatrixFree / FEEvaluation / FEFaceEvaluation hold for this path (a3, a5, a8, a10, a11).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "mesh.hpp"

namespace exadg_b200
{
#define EXADG_MAX_N 8 // degrees 1..7

struct DeviceOperator
{
  int degree = 0, n = 0;
  int64_t n_owned = 0, n_ghost = 0, n_faces = 0;
  int64_t n_global_dofs = 0;
  // connectivity
  int32_t * nb = nullptr;       // [owned][6]
  int32_t * face_id = nullptr;  // [owned][6]
  uint8_t * face_info = nullptr; // [owned][6]
  // general geometry (null on the Cartesian fast path)
  double * cellG = nullptr;     // [owned][6][n^3]  symmetric J^-1 J^-T JxW: xx,yy,zz,xy,xz,yz
  double * faceG = nullptr;     // [n_faces][7][n^2] a_minus(3), a_plus(3), JxW
  double * tau_f = nullptr;     // [n_faces] penalty incl. (k+1)^2 IP_factor
  double * tau_cell = nullptr;  // [owned+ghost] surface/volume
  double * cellJxW = nullptr;   // [owned][n^3] JxW at the cell quadrature points (mass term / inverse mass; Helmholtz operators only)
  // Helmholtz / viscous operator: mass_coeff (v,u) + laplace_coeff a_SIPG(u,v) on each of n_components components
  bool helmholtz = false; int n_components = 1; double mass_coeff = 0.0, laplace_coeff = 1.0;
  // ghost values of src, filled by the halo exchange
  double * ghost = nullptr;     // [n_ghost][n^3]
  // Cartesian fast path
  bool cartesian = false;       // uniform boxes + all faces interior
  double h[3] = {0, 0, 0};
  double tau_hat = 0;           // tau_K (k+1)^2 IP_factor of the uniform box (times h_d = tau_hat_d)
  void * cart_plan = nullptr;   // batch plan of the Cartesian kernel (vmult_cartesian.cu)
  int cart_variant = -1;        // n = 5 kernel of this operator (-1: the process-wide default, cartesian_kernel_variant)
};

// ---- geometry.cu ----
void setup_geometry(DeviceOperator & op, const HostMesh & mesh, double ip_factor, cudaStream_t stream);

// ---- vmult_general.cu ----
// dst (+)= A src on owned cells listed in cells[0..n_cells) (or all owned cells if cells == nullptr)
void launch_vmult_general(const DeviceOperator & op, double * dst, const double * src, bool add, const int32_t * cells, int64_t n_cells, cudaStream_t stream);
void launch_diagonal_general(const DeviceOperator & op, double * diag, bool add, cudaStream_t stream);
// InverseMassOperator (I/operators/inverse_mass_operator.h; dealii CellwiseInverseMassMatrix): dst = M^-1 src cell by cell,
// M_K = S^T diag(JxW) S with as many Gauss points as basis functions per direction => M_K^-1 = S^-1 diag(1/JxW) S^-T
void launch_inverse_mass(const DeviceOperator & op, double * dst, const double * src, cudaStream_t stream);

// ---- vmult_cartesian.cu ----
bool cartesian_supported(int n);
// selects the n = 5 kernel of the fast path (0 pipelined, 1 (default) / 2 warp-specialised with producer depth 8 / 12, 3 with 4 producer warps; -1 only queries); returns the previous value
int cartesian_kernel_variant(int set);
size_t cartesian_plan_create(DeviceOperator & op, const HostMesh & mesh);
void cartesian_plan_destroy(DeviceOperator & op);
// which: 0 all cell batches, 1 batches touching no ghost cell, 2 batches touching ghost cells
void launch_vmult_cartesian_part(const DeviceOperator & op, double * dst, const double * src, bool add, int which, cudaStream_t stream);
// explicit device list of batch ids (operators without ghost cells): chunks of the pipelined host-buffer vmult
void launch_vmult_cartesian_list(const DeviceOperator & op, double * dst, const double * src, bool add, const int32_t * list, int n_list, cudaStream_t stream);
int cartesian_batch_size(const DeviceOperator & op);
int cartesian_n_batches(const DeviceOperator & op);

// in-kernel ghost hand-over of the single-launch partitioned vmult (NVLink peer-memory halo): this rank's flag slots, indexed by
// peer rank, reach `epoch` once that peer has stored its cells into this rank's ghost buffer
struct GhostSync
{
  const long long * flags = nullptr; long long epoch = 0; int n_peers = 0; int peer_rank[16] = {0};
  // the export of this rank's cells, done by the same launch before its first batch (every CTA copies a slice of every send list
  // straight into the peer's ghost buffer over NVLink; the CTA that completes a peer's list publishes the epoch in the peer's flag
  // slot).  done[p] counts finished CTAs over all launches and is never reset: the launch grid must not change between vmults.
  const int32_t * send_cells[16] = {nullptr}; int64_t n_send[16] = {0}; double * peer_ghost[16] = {nullptr}; long long * peer_flag[16] = {nullptr};
  unsigned long long * done = nullptr;
  long long * put_seq = nullptr; int * put_grid = nullptr; // host-side state of the operator: launches counted by `done`, their grid
  int * counter = nullptr; // device work counter (zeroed per launch) for dynamic item claiming; nullptr: static striding, every CTA exports
};
// one launch over all batches of a partition, batches without ghost neighbours first; the producers of the remaining batches acquire
// the peers' flags inside the kernel.  Returns false (nothing launched) if the operator has no kernel with that capability.
bool launch_vmult_cartesian_fused(const DeviceOperator & op, double * dst, const double * src, bool add, const GhostSync & gs, cudaStream_t stream);

// ---- vmult_cartesian_ws.cu (warp-specialised kernel, n = 5) ----
bool ws_supported(int n);
void * ws_plan_create(const DeviceOperator & op, const HostMesh & mesh);
void ws_plan_destroy(void * plan);
void ws_launch(const DeviceOperator & op, const void * plan, double * dst, const double * src, bool add, const int32_t * batches, int n_items, int n_sm, int depth, bool gh,
               cudaStream_t stream, const GhostSync * gs = nullptr, int first_ghost_item = 0);

// ---- rhs_error.cu (inhomogeneous boundary data, right-hand side, error norms) ----
void * post_get(void *& slot, const DeviceOperator & op, const HostMesh & mesh, double penalty_factor, cudaStream_t stream);
void post_destroy(void * p);
int64_t post_n_boundary_faces(void * p);
void post_boundary_points(void * p, const DeviceOperator & op, double * xyz_host, uint8_t * type_host, cudaStream_t stream);
void post_set_boundary_values(void * p, const DeviceOperator & op, const double * values_host, cudaStream_t stream);
void post_boundary_inhom_add(void * p, const DeviceOperator & op, const HostMesh & mesh, double sign, double * dst, cudaStream_t stream);
void post_cell_points(void * p, const DeviceOperator & op, const HostMesh & mesh, int nq, double * xyz_host, cudaStream_t stream);
void post_source_add(void * p, const DeviceOperator & op, const HostMesh & mesh, const double * f_host, double * dst, cudaStream_t stream);
void post_l2_cells(void * p, const DeviceOperator & op, const HostMesh & mesh, int nq, const double * u, const double * exact_host, double * d_out, cudaStream_t stream);

// ---- microbench.cu ----
void fp64_peak(double * dfma_tflops, double * dmma_tflops);

void cuda_check(cudaError_t e, const char * what);
// true the first time it is called for (key, current device): per-device one-time configuration (cudaFuncSetAttribute applies to
// the current device only)
bool first_use_on_device(const void * key);
#define CUDA_CHECK(x) ::exadg_b200::cuda_check((x), #x)

} // namespace exadg_b200
