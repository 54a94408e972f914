// C ABI (include/exadg_b200.h) and the host-side operator object behind it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "../../include/exadg_b200.h"
#include "host_pipeline.hpp"
#include "operator.cuh"
#include "vector_ops.cuh"

using namespace exadg_b200;

namespace
{
thread_local std::string g_last_error;

// ---- NCCL, resolved at run time (the process usually already holds torch's libnccl.so.2) ----
struct NcclApi
{
  void * lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, char[128], int) = nullptr; // ncclUniqueId passed by value (128 bytes)
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  bool ok = false;
};
struct NcclId { char bytes[128]; };

NcclApi & nccl()
{
  static NcclApi api;
  if (api.lib) return api;
  const char * names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char * nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) return api;
  api.GetUniqueId = (int (*)(void *))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void **, int, char[128], int))dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
  api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
  api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
  api.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))dlsym(api.lib, "ncclSend");
  api.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))dlsym(api.lib, "ncclRecv");
  api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
  api.ok = api.GetUniqueId && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.AllReduce;
  return api;
}
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

__global__ void pack_cells_kernel(const double * __restrict__ src, const int32_t * __restrict__ cells, int64_t n_cells, int n3, double * __restrict__ out)
{
  const int64_t total = n_cells * n3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / n3; const int k = (int)(i % n3);
    out[i] = src[(int64_t)cells[c] * n3 + k];
  }
}

// ---- NVLink peer-memory halo exchange (no NCCL in the data path) ------------------------------------------
// Every rank maps its peers' ghost buffers (CUDA IPC).  One kernel gathers the owned cells each peer needs and
// stores them straight into that peer's ghost range over NVLink (pack + put fused); a signal kernel then
// publishes the epoch in the peer's flag slot; the consumer's wait kernel (stream-ordered before the cells that
// touch ghosts) spins on its own flags.  Ghost buffers are double-buffered by epoch parity: a peer can be at most
// one vmult ahead (it needs our data of that vmult to finish it), so no back-pressure handshake is needed.
constexpr int MAX_PEERS = 16;
struct PutArgs
{
  const int32_t * cells[MAX_PEERS]; int64_t n_cells[MAX_PEERS]; double * dst[MAX_PEERS];
  long long * peer_flag[MAX_PEERS]; // slot of this rank in the peer's flag array
  int peer_rank[MAX_PEERS];
  int n_peers;
};
__global__ void put_cells_kernel(const PutArgs a, const double * __restrict__ src, int n3)
{
  const int p = blockIdx.y;
  const int32_t * __restrict__ cells = a.cells[p];
  double * __restrict__ dst = a.dst[p]; // peer memory
  const int64_t total = a.n_cells[p] * n3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / n3; const int k = (int)(i % n3);
    dst[i] = src[(int64_t)cells[c] * n3 + k];
  }
}
__global__ void signal_peers_kernel(const PutArgs a, long long epoch)
{
  const int p = threadIdx.x;
  if (p < a.n_peers) {
    __threadfence_system(); // the stores of put_cells_kernel (previous kernel on this stream) are performed
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(a.peer_flag[p]), "l"(epoch) : "memory");
  }
}
__global__ void wait_peers_kernel(const PutArgs a, const long long * my_flags, long long epoch)
{
  const int p = threadIdx.x;
  if (p < a.n_peers) {
    long long v;
    do { asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(my_flags + a.peer_rank[p]) : "memory"); } while (v < epoch);
  }
}

// closed-form diagonal on the uniform periodic Cartesian box: A = sum_d c_d (M x M x L_d)  =>
// A_ii = sum_d c_d M_aa M_bb (L_d)_cc, identical for every cell
__global__ void cartesian_diagonal_kernel(double * __restrict__ diag, int64_t n_dofs, int n, const double * __restrict__ cell_diag, int add)
{
  const int n3 = n * n * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_dofs; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = cell_diag[i % n3];
    diag[i] = add ? diag[i] + v : v;
  }
}
} // namespace

struct exadg_b200_operator
{
  DeviceOperator dev;
  HostMesh mesh; // connectivity kept for the halo plan (mapping points dropped after setup)
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev_packed = nullptr, ev_halo = nullptr, ev_order = nullptr;
  Reducer red;
  int64_t launches = 0;
  int64_t n_local = 0; // locally owned DoFs
  // halo
  void * comm = nullptr; bool own_comm = false;
  std::vector<int32_t *> d_send_lists; std::vector<double *> d_send_bufs;
  int32_t * d_interior = nullptr, * d_boundary = nullptr; int64_t n_interior = 0, n_boundary = 0;
  // uniform box with Dirichlet / Neumann faces: batches of cells that see the interior penalty on all their faces run the affine fast
  // kernel, the two cell layers next to the boundary (and the batches they share) the general kernel
  bool hybrid = false; int32_t * d_hyb_batches = nullptr, * d_hyb_cells = nullptr; int64_t n_hyb_batches = 0, n_hyb_cells = 0;
  // Helmholtz / viscous operator on a uniform periodic box (the Taylor-Green setup): vmult on the affine fast kernels, one component of a
  // cell batch per CTA (CartArgs::ncomp); dev_helm holds the batch plan.  The diagonal and everything else stay with dev (general kernel).
  bool helm_fast = false; DeviceOperator dev_helm;
  // peer-memory halo (NVLink): one region [ghost A | ghost B | flags[world]] mapped by all peers
  bool p2p = false; char * p2p_region = nullptr; size_t p2p_ghost_bytes = 0; long long p2p_epoch = 0;
  unsigned long long * d_put_done = nullptr; long long put_seq = 0; int put_grid = 0; int * d_work_counter = nullptr; // ticket counters of the in-launch export (GhostSync)
  std::vector<void *> p2p_peer_regions; // opened IPC mappings, indexed like mesh.peers
  std::vector<int64_t> p2p_peer_recv_begin, p2p_peer_ghost_bytes;
  double * ghost_alloc = nullptr; // the ghost buffer of the NCCL path (dev.ghost points into p2p_region once p2p is on)
  // work vectors
  double * w[4] = {nullptr, nullptr, nullptr, nullptr};
  double * d_cell_diag = nullptr;
  void * post = nullptr; // boundary data / right-hand side / error norms (rhs_error.cu), built on first use
  double * d_stage_src = nullptr, * d_stage_dst = nullptr;
  // pipelined host-buffer vmult (exadg_b200_vmult_host_pipelined)
  HostPipelinePlan hp; bool hp_built = false;
  cudaStream_t hp_in = nullptr, hp_out = nullptr; cudaEvent_t hp_start = nullptr;
  std::vector<cudaEvent_t> hp_ev_in, hp_ev_cmp; int32_t * d_iota = nullptr;
  // second plan: pieces in address order, per-unit readiness, dst stored by the kernels into the device-mapped host buffer
  HostStreamPlan hs; bool hs_built = false; int32_t * d_hs_units = nullptr; std::vector<cudaEvent_t> hs_ev;
  int hp_mode = 0; // 0 auto (direct on the affine path if dst_host is device-accessible), 1 staged (chunk plan, copy-engine download), 2 direct

  double * work(int i)
  {
    if (!w[i]) { CUDA_CHECK(cudaMalloc(&w[i], (size_t)std::max<int64_t>(n_local, 1) * sizeof(double))); }
    return w[i];
  }
};

struct exadg_b200_chebyshev
{
  exadg_b200_operator * op = nullptr;
  int degree = 5; double smoothing_range = 20; int eig_cg_n_iterations = 20;
  double lambda_min_est = 0, lambda_max_est = 0, theta = 1, delta = 0;
  double * inv_diag = nullptr, * xold = nullptr, * r = nullptr;
};

// MultigridPreconditionerBase / MultigridAlgorithm on a hierarchy of DG level operators; see the comments at
// exadg_b200_multigrid_create.  Levels are ordered coarse -> fine like level_info of the reference.
struct exadg_b200_multigrid
{
  std::vector<exadg_b200_operator *> ops;
  std::vector<exadg_b200_chebyshev *> smoothers;      // [level], nullptr on level 0
  std::vector<double *> defect, solution, t;          // MultigridAlgorithm::defect / solution / t (multigrid_algorithm.h:64-66)
  std::vector<TransferTable> transfer;                // [level]: between level - 1 and level
  double * coarse_inv_diag = nullptr, * coarse_rhs = nullptr;
  double coarse_abs_tol = 1e-12, coarse_rel_tol = 1e-3; int coarse_max_iter = 10000;
  int64_t coarse_iterations = 0, cycles = 0;
};

namespace
{
template<typename F>
int guarded(F && f)
{
  try { return f(); }
  catch (const std::invalid_argument & e) { g_last_error = e.what(); return EXADG_B200_ERR_ARG; }
  catch (const std::exception & e) { g_last_error = e.what(); return std::string(e.what()).find("CUDA") != std::string::npos ? EXADG_B200_ERR_CUDA : EXADG_B200_ERR_UNSUPPORTED; }
}

// exadg_b200_create with n_cells_ghost > 0: the partition (ghost import, global dot products) is owned by the caller; the solver
// and smoother entry points of the library cannot provide either and refuse instead of returning rank-local results
void require_self_contained(const exadg_b200_operator * op, const char * what);

void check_ptr(const void * p, const char * name)
{
  if (!p) throw std::invalid_argument(std::string(name) + " is null");
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) throw std::invalid_argument(std::string(name) + " must be 16-byte aligned");
}

void require_self_contained(const exadg_b200_operator * op, const char * what)
{
  if (op->dev.n_ghost > 0 && op->mesh.peers.empty())
    throw std::runtime_error(std::string(what) + ": this operator has ghost cells filled by the caller (exadg_b200_create with n_cells_ghost > 0); "
                             "the library cannot refresh them or reduce over the caller's ranks - drive vmult from the caller's solver instead");
}

void allreduce(exadg_b200_operator * op, double * dev_scalars, int count)
{
  if (op->mesh.world <= 1) return;
  if (!op->comm) throw std::runtime_error("world > 1 requires a communicator (exadg_b200_set_nccl_comm / exadg_b200_nccl_init)");
  if (nccl().AllReduce(dev_scalars, dev_scalars, (size_t)count, NCCL_FLOAT64, NCCL_SUM, op->comm, op->stream) != 0) throw std::runtime_error("ncclAllReduce failed");
}

void finish_setup(exadg_b200_operator * op, double ip_factor, bool force_general)
{
  HostMesh & M = op->mesh;
  DeviceOperator & D = op->dev;
  if (!M.standard_orientation()) throw std::runtime_error("only meshes in standard orientation are supported");
  D.n = D.degree + 1;
  D.n_owned = M.n_owned; D.n_ghost = M.n_ghost; D.n_faces = M.n_faces;
  const int64_t n3 = (int64_t)D.n * D.n * D.n;
  D.n_global_dofs = M.n_global_cells * n3 * D.n_components;
  op->n_local = M.n_owned * n3 * D.n_components;
  for (int e = 0; e < 3; ++e) D.h[e] = M.h[e];
  // the mass term and the component blocks live in the general kernel only
  D.cartesian = M.n_owned > 0 && M.cartesian_uniform && M.all_interior() && !force_general && !D.helmholtz && cartesian_supported(D.n);
  if (D.cartesian) {
    // tau_hat is needed by the kernel tables: uniform box => tau_K = sum_d 1/h_d (interior_penalty_parameter.h:68-98)
    double tk = 0.0;
    for (int e = 0; e < 3; ++e) tk += 1.0 / M.h[e];
    D.tau_hat = tk * ip_factor * (D.degree + 1.0) * (D.degree + 1.0);
    if (cartesian_plan_create(D, M) == 0) D.cartesian = false;
  }
  if (!D.cartesian && M.n_owned > 0 && M.cartesian_uniform && !M.all_interior() && !force_general && !D.helmholtz && M.world <= 1 && M.n_ghost == 0 && cartesian_supported(D.n)
      && !getenv("EXADG_B200_NO_HYBRID")) {
    // tau_K counts true boundary faces with weight 1 (interior_penalty_parameter.h:88-89), and an interior face takes the larger of its two
    // cells' values (laplace_operator.h:128-140): a cell sees the uniform interior penalty on all its faces iff neither it nor any of its
    // neighbours has a boundary face.  Batches made of such cells only go to the fast kernel (they read, but never write, the others).
    std::vector<uint8_t> has_bf(M.n_owned, 0);
    for (int64_t c = 0; c < M.n_owned; ++c) for (int f = 0; f < 6; ++f) if (M.nb[c * 6 + f] < 0) has_bf[c] = 1;
    std::vector<uint8_t> regular(M.n_owned, 1);
    for (int64_t c = 0; c < M.n_owned; ++c) {
      if (has_bf[c]) { regular[c] = 0; continue; }
      for (int f = 0; f < 6; ++f) if (has_bf[M.nb[c * 6 + f]]) regular[c] = 0;
    }
    double tk = 0.0;
    for (int e = 0; e < 3; ++e) tk += 1.0 / M.h[e];
    D.tau_hat = tk * ip_factor * (D.degree + 1.0) * (D.degree + 1.0);
    const std::vector<int32_t> nb_saved = M.nb;
    for (int64_t i = 0; i < M.n_owned * 6; ++i) if (M.nb[i] < 0) M.nb[i] = (int32_t)(i / 6); // the plan of such batches is never launched
    const size_t ok = cartesian_plan_create(D, M);
    M.nb = nb_saved;
    if (ok != 0) {
      const int B = cartesian_batch_size(D), nbatch = cartesian_n_batches(D);
      std::vector<int32_t> fast, slow;
      for (int b = 0; b < nbatch; ++b) {
        const int64_t b0 = (int64_t)b * B, b1 = std::min<int64_t>(b0 + B, M.n_owned);
        bool reg = true;
        for (int64_t c = b0; c < b1; ++c) reg &= (regular[c] != 0);
        if (reg) fast.push_back(b); else for (int64_t c = b0; c < b1; ++c) slow.push_back((int32_t)c);
      }
      if (!fast.empty()) {
        op->hybrid = true; op->n_hyb_batches = (int64_t)fast.size(); op->n_hyb_cells = (int64_t)slow.size();
        CUDA_CHECK(cudaMalloc(&op->d_hyb_batches, fast.size() * 4)); CUDA_CHECK(cudaMemcpy(op->d_hyb_batches, fast.data(), fast.size() * 4, cudaMemcpyHostToDevice));
        if (!slow.empty()) { CUDA_CHECK(cudaMalloc(&op->d_hyb_cells, slow.size() * 4)); CUDA_CHECK(cudaMemcpy(op->d_hyb_cells, slow.data(), slow.size() * 4, cudaMemcpyHostToDevice)); }
      } else cartesian_plan_destroy(D);
    }
  }
  CUDA_CHECK(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
  {
    // the halo exchange must not queue behind the interior-cell kernel: highest priority for its stream
    int lo = 0, hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_CHECK(cudaStreamCreateWithPriority(&op->comm_stream, cudaStreamNonBlocking, hi));
  }
  CUDA_CHECK(cudaEventCreateWithFlags(&op->ev_packed, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&op->ev_halo, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&op->ev_order, cudaEventDisableTiming));
  reducer_init(op->red);
  setup_geometry(D, M, ip_factor, op->stream);
  if (D.helmholtz && M.n_owned > 0 && M.cartesian_uniform && M.all_interior() && !force_general && M.world <= 1 && M.n_ghost == 0 && cartesian_supported(D.n)
      && !getenv("EXADG_B200_NO_HELMHOLTZ_FAST")) {
    // Helmholtz / viscous operator on a uniform periodic box: the affine fast kernels, one component of a cell batch per CTA.  The batch plan
    // is that of the scalar operator on this mesh (dev_helm shares the neighbour table of dev and owns only the plan)
    DeviceOperator & H = op->dev_helm;
    H = DeviceOperator();
    H.degree = D.degree; H.n = D.n; H.n_owned = D.n_owned; H.n_global_dofs = D.n_global_dofs; H.nb = D.nb;
    H.helmholtz = true; H.n_components = D.n_components; H.mass_coeff = D.mass_coeff; H.laplace_coeff = D.laplace_coeff;
    for (int e = 0; e < 3; ++e) H.h[e] = M.h[e];
    double tk = 0.0;
    for (int e = 0; e < 3; ++e) tk += 1.0 / M.h[e];
    H.tau_hat = tk * ip_factor * (D.degree + 1.0) * (D.degree + 1.0);
    H.cartesian = true;
    if (cartesian_plan_create(H, M) != 0) op->helm_fast = true;
  }
  // halo plan on the device + interior/boundary split for overlap
  if (M.world > 1) {
    for (auto & p : M.peers) {
      int32_t * l = nullptr; double * b = nullptr;
      CUDA_CHECK(cudaMalloc(&l, p.send_cells.size() * sizeof(int32_t)));
      CUDA_CHECK(cudaMemcpy(l, p.send_cells.data(), p.send_cells.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMalloc(&b, p.send_cells.size() * n3 * D.n_components * sizeof(double))); // whole cell blocks: all components
      op->d_send_lists.push_back(l); op->d_send_bufs.push_back(b);
    }
    std::vector<int32_t> interior, boundary;
    for (int64_t c = 0; c < M.n_owned; ++c) {
      bool touches = false;
      for (int f = 0; f < 6; ++f) touches |= (M.nb[c * 6 + f] >= M.n_owned);
      (touches ? boundary : interior).push_back((int32_t)c);
    }
    if (D.n_components > 1) { // the general kernel's items are (cell, component) blocks
      auto blocks = [&](std::vector<int32_t> & v) { std::vector<int32_t> b; b.reserve(v.size() * D.n_components); for (int32_t c : v) for (int k = 0; k < D.n_components; ++k) b.push_back(c * D.n_components + k); v.swap(b); };
      blocks(interior); blocks(boundary);
    }
    op->n_interior = (int64_t)interior.size(); op->n_boundary = (int64_t)boundary.size();
    if (!interior.empty()) { CUDA_CHECK(cudaMalloc(&op->d_interior, interior.size() * 4)); CUDA_CHECK(cudaMemcpy(op->d_interior, interior.data(), interior.size() * 4, cudaMemcpyHostToDevice)); }
    if (!boundary.empty()) { CUDA_CHECK(cudaMalloc(&op->d_boundary, boundary.size() * 4)); CUDA_CHECK(cudaMemcpy(op->d_boundary, boundary.data(), boundary.size() * 4, cudaMemcpyHostToDevice)); }
  }
  // the mapping support points stay on the host: rhs / evaluate / error norms rebuild their geometry from them on first use
}

// which: 0 all owned cells, 1 cells (batches) that touch no ghost, 2 those that do
void launch_vmult(exadg_b200_operator * op, double * dst, const double * src, bool add, int which, cudaStream_t stream = nullptr)
{
  if (!stream) stream = op->stream;
  if (op->helm_fast && which == 0) launch_vmult_cartesian_part(op->dev_helm, dst, src, add, 0, stream);
  else if (op->dev.cartesian) launch_vmult_cartesian_part(op->dev, dst, src, add, which, stream);
  else if (op->hybrid && which == 0) {
    launch_vmult_cartesian_list(op->dev, dst, src, add, op->d_hyb_batches, (int)op->n_hyb_batches, stream);
    if (op->n_hyb_cells > 0) { launch_vmult_general(op->dev, dst, src, add, op->d_hyb_cells, op->n_hyb_cells, stream); op->launches++; }
  }
  else if (which == 0) launch_vmult_general(op->dev, dst, src, add, nullptr, 0, stream);
  else {
    const int64_t nc = which == 1 ? op->n_interior : op->n_boundary;
    if (nc == 0) return;
    launch_vmult_general(op->dev, dst, src, add, which == 1 ? op->d_interior : op->d_boundary, nc, stream);
  }
  op->launches++;
}

// dst (+)= A src including the ghost import of src (MatrixFree::loop's update_ghost_values, overlapped
// with the cells that touch no ghost)
// boundary_only (partitioned operators): ghost import + the cells (batches) that touch ghost cells; the others were applied by the caller
void apply(exadg_b200_operator * op, double * dst, const double * src, bool add, bool boundary_only = false)
{
  if (op->n_local == 0 && op->mesh.peers.empty()) return; // is_empty_locally: a rank without cells (and without neighbours) has nothing to do
  check_ptr(dst, "dst"); check_ptr(src, "src");
  if (dst == src) throw std::invalid_argument("dst and src must not alias");
  HostMesh & M = op->mesh;
  if (M.world <= 1 || M.peers.empty()) { if (!boundary_only) launch_vmult(op, dst, src, add, 0); return; }
  const int n3 = op->dev.n * op->dev.n * op->dev.n * op->dev.n_components; // doubles per cell block (all components of a cell travel together)
  if (op->p2p) {
    const long long epoch = ++op->p2p_epoch;
    const int buf = (int)(epoch & 1);
    if (M.peers.size() > (size_t)MAX_PEERS) throw std::runtime_error("too many halo peers");
    PutArgs a; a.n_peers = (int)M.peers.size();
    int64_t max_total = 1;
    for (size_t i = 0; i < M.peers.size(); ++i) {
      a.cells[i] = op->d_send_lists[i]; a.n_cells[i] = (int64_t)M.peers[i].send_cells.size();
      char * peer = static_cast<char *>(op->p2p_peer_regions[i]);
      const size_t pgb = (size_t)op->p2p_peer_ghost_bytes[i]; // the PEER's buffer size fixes the layout of its region
      a.dst[i] = reinterpret_cast<double *>(peer + (size_t)buf * pgb) + op->p2p_peer_recv_begin[i] * n3;
      a.peer_flag[i] = reinterpret_cast<long long *>(peer + 2 * pgb) + M.rank;
      a.peer_rank[i] = M.peers[i].rank;
      max_total = std::max<int64_t>(max_total, a.n_cells[i] * n3);
    }
    op->dev.ghost = reinterpret_cast<double *>(op->p2p_region + (size_t)buf * op->p2p_ghost_bytes);
    // src must be complete (work queued on the compute stream) before it is read on the communication stream
    CUDA_CHECK(cudaEventRecord(op->ev_packed, op->stream));
    CUDA_CHECK(cudaStreamWaitEvent(op->comm_stream, op->ev_packed, 0));
    if (!boundary_only && op->dev.cartesian && op->d_put_done) {
      // single launch: the operator kernel exports this rank's cells before its first batch (every CTA a slice, the last one per
      // peer publishes the epoch), runs the batches without ghost neighbours, and its producers acquire the peers' flags before
      // the first batch that reads ghost cells.  No communication stream, no events.
      GhostSync gs;
      gs.flags = reinterpret_cast<const long long *>(op->p2p_region + 2 * op->p2p_ghost_bytes); gs.epoch = epoch; gs.n_peers = a.n_peers;
      gs.done = op->d_put_done; gs.put_seq = &op->put_seq; gs.put_grid = &op->put_grid;
      static const bool static_items = getenv("EXADG_B200_STATIC_ITEMS") != nullptr; // measurement switch: every CTA exports, static striding
      gs.counter = static_items ? nullptr : op->d_work_counter;
      for (int i = 0; i < a.n_peers; ++i) {
        gs.peer_rank[i] = a.peer_rank[i]; gs.send_cells[i] = a.cells[i]; gs.n_send[i] = a.n_cells[i]; gs.peer_ghost[i] = a.dst[i]; gs.peer_flag[i] = a.peer_flag[i];
      }
      if (launch_vmult_cartesian_fused(op->dev, dst, src, add, gs, op->stream)) {
        op->launches += 1;
        CUDA_CHECK(cudaGetLastError());
        return;
      }
    }
    const dim3 grid((unsigned)std::min<int64_t>((max_total + 255) / 256, 64), (unsigned)a.n_peers);
    put_cells_kernel<<<grid, 256, 0, op->comm_stream>>>(a, src, n3);
    signal_peers_kernel<<<1, 32, 0, op->comm_stream>>>(a, epoch);
    wait_peers_kernel<<<1, 32, 0, op->comm_stream>>>(a, reinterpret_cast<const long long *>(op->p2p_region + 2 * op->p2p_ghost_bytes), epoch);
    op->launches += 3;
    // cells that touch ghosts run on the (high-priority) communication stream right behind the wait kernel and
    // overlap with the tail of the interior-cell kernel on the compute stream; the streams join at the end
    launch_vmult(op, dst, src, add, 2, op->comm_stream);
    CUDA_CHECK(cudaEventRecord(op->ev_halo, op->comm_stream));
    if (!boundary_only) launch_vmult(op, dst, src, add, 1);
    CUDA_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (!op->comm) throw std::runtime_error("world > 1 requires a communicator (exadg_b200_set_nccl_comm / exadg_b200_nccl_init)");
  for (size_t i = 0; i < M.peers.size(); ++i) {
    const int64_t nc = (int64_t)M.peers[i].send_cells.size();
    const int64_t total = nc * n3;
    pack_cells_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, op->stream>>>(src, op->d_send_lists[i], nc, n3, op->d_send_bufs[i]);
    op->launches++;
  }
  CUDA_CHECK(cudaEventRecord(op->ev_packed, op->stream));
  CUDA_CHECK(cudaStreamWaitEvent(op->comm_stream, op->ev_packed, 0));
  static const bool skip_exchange = getenv("EXADG_B200_SKIP_EXCHANGE") != nullptr; // timing experiments only (results wrong)
  NcclApi & api = nccl();
  if (!skip_exchange) {
  api.GroupStart();
  for (size_t i = 0; i < M.peers.size(); ++i) {
    const PeerPlan & p = M.peers[i];
    api.Send(op->d_send_bufs[i], p.send_cells.size() * (size_t)n3, NCCL_FLOAT64, p.rank, op->comm, op->comm_stream);
    api.Recv(op->dev.ghost + p.recv_begin * n3, (size_t)p.recv_count * n3, NCCL_FLOAT64, p.rank, op->comm, op->comm_stream);
  }
  if (api.GroupEnd() != 0) throw std::runtime_error("NCCL halo exchange failed");
  }
  launch_vmult(op, dst, src, add, 2, op->comm_stream);
  CUDA_CHECK(cudaEventRecord(op->ev_halo, op->comm_stream));
  if (!boundary_only) launch_vmult(op, dst, src, add, 1);
  CUDA_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
}

void diagonal(exadg_b200_operator * op, double * diag, bool add)
{
  if (op->n_local == 0) return;
  check_ptr(diag, "diagonal");
  DeviceOperator & D = op->dev;
  if (D.cartesian) {
    const int n = D.n, n3 = n * n * n;
    if (!op->d_cell_diag) {
      // 1-D own-side operator L_d = K/h_d-scaled + face terms; see vmult_cartesian.cu for the derivation
      Tables1D tab(D.degree);
      std::vector<double> cd(n3, 0.0);
      for (int d = 0; d < 3; ++d) {
        const int e = (d + 1) % 3, f = (d + 2) % 3;
        const double cdir = D.h[e] * D.h[f] / D.h[d];
        const real_t tau_hat = (real_t)D.tau_hat * D.h[d];
        std::vector<real_t> Ld(n);
        for (int i = 0; i < n; ++i) {
          real_t v = tab.K[i * n + i];
          for (int s = 0; s < 2; ++s) {
            const real_t sig = s ? 1 : -1;
            const real_t es = (i == (s ? n - 1 : 0)) ? 1 : 0;
            v += -sig * tab.fd[s][i] * es + tau_hat * es * es; // -1/2 sig (d e^T + e d^T)_ii + tau e e^T
          }
          Ld[i] = v;
        }
        for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
          const int idx[3] = {i, j, k};
          cd[i + n * (j + n * k)] += (double)(cdir * Ld[idx[d]] * tab.M[idx[e] * n + idx[e]] * tab.M[idx[f] * n + idx[f]]);
        }
      }
      CUDA_CHECK(cudaMalloc(&op->d_cell_diag, n3 * sizeof(double)));
      CUDA_CHECK(cudaMemcpy(op->d_cell_diag, cd.data(), n3 * sizeof(double), cudaMemcpyHostToDevice));
    }
    cartesian_diagonal_kernel<<<148 * 8, 256, 0, op->stream>>>(diag, op->n_local, n, op->d_cell_diag, add ? 1 : 0);
    op->launches++;
  } else {
    launch_diagonal_general(D, diag, add, op->stream);
    op->launches++;
  }
  CUDA_CHECK(cudaGetLastError());
}

double read_scalar(exadg_b200_operator * op, int slot)
{
  CUDA_CHECK(cudaMemcpyAsync(op->red.host + slot, op->red.result + slot, sizeof(double), cudaMemcpyDeviceToHost, op->stream));
  CUDA_CHECK(cudaStreamSynchronize(op->stream));
  return op->red.host[slot];
}

// all reduction slots -> pinned host mirror, ordered behind the kernels queued on the operator's stream
void read_all_scalars(exadg_b200_operator * op)
{
  CUDA_CHECK(cudaMemcpyAsync(op->red.host, op->red.result, 8 * sizeof(double), cudaMemcpyDeviceToHost, op->stream));
  CUDA_CHECK(cudaStreamSynchronize(op->stream));
}

void cheb_run(exadg_b200_chebyshev * ch, double * x, const double * b, bool zero_start);
void mg_vmult(exadg_b200_multigrid * mg, double * dst, const double * src);
thread_local exadg_b200_multigrid * g_cg_multigrid = nullptr; // preconditioner object of EXADG_B200_PRECOND_MULTIGRID for the running cg()

void precondition(exadg_b200_operator * op, int precond, const double * inv_diag, exadg_b200_chebyshev * cheb, int slot, double * z, const double * g)
{
  const int64_t n = op->n_local;
  if (precond == EXADG_B200_PRECOND_POINT_JACOBI) { jacobi_dot(op->red, slot, z, inv_diag, g, n, op->stream); op->launches++; }
  else if (precond == EXADG_B200_PRECOND_MULTIGRID) { mg_vmult(g_cg_multigrid, z, g); dot(op->red, slot, g, z, n, op->stream); op->launches++; }
  else { cheb_run(cheb, z, g, true); dot(op->red, slot, g, z, n, op->stream); op->launches++; }
  allreduce(op, op->red.result + slot, 1);
}

// dealii::SolverCG restated; scalars alpha/beta never leave the device, the host
// only reads the residual norm once per iteration for ReductionControl::check.
int cg(exadg_b200_operator * op, double * x, const double * b, int precond, const double * inv_diag, exadg_b200_chebyshev * cheb,
       double abs_tol, double rel_tol, int max_iter, int * n_iter, double * residuals, std::vector<double> * alphas, std::vector<double> * betas)
{
  const int64_t n = op->n_local;
  cudaStream_t s = op->stream;
  double * g = op->work(0), * d = op->work(1), * h = op->work(2);
  int S_GH = 0, S_DH = 1, S_NEW = 2, S_RES = 3;
  // g = A x - b
  apply(op, g, x, false);
  axpby(-1.0, b, 1.0, g, n, s); op->launches++;
  dot(op->red, S_RES, g, g, n, s); op->launches++;
  allreduce(op, op->red.result + S_RES, 1);
  double res = std::sqrt(read_scalar(op, S_RES));
  const double res0 = res, reduced_tol = rel_tol * res0;
  if (residuals) residuals[0] = res;
  int it = 0, state = 0;
  auto check = [&](int step, double r) {
    if (r < reduced_tol) return 1;   // ReductionControl::check (strict)
    if (r <= abs_tol) return 1;      // SolverControl::check
    if (step >= max_iter || std::isnan(r)) return 2;
    return 0;
  };
  state = check(0, res);
  if (state == 0) {
    if (precond != EXADG_B200_PRECOND_NONE) { precondition(op, precond, inv_diag, cheb, S_GH, h, g); scale_copy(-1.0, h, d, n, s); op->launches++; }
    else { scale_copy(-1.0, g, d, n, s); op->launches++; std::swap(S_GH, S_RES); }
  }
  while (state == 0) {
    ++it;
    apply(op, h, d, false);
    dot(op->red, S_DH, d, h, n, s); op->launches++;
    allreduce(op, op->red.result + S_DH, 1);
    cg_update_x_g(op->red, S_RES, S_GH, S_DH, x, d, g, h, n, s); op->launches++;
    allreduce(op, op->red.result + S_RES, 1);
    res = std::sqrt(read_scalar(op, S_RES));
    if (residuals) residuals[it] = res;
    if (alphas) { // Lanczos coefficients for the eigenvalue estimate
      read_all_scalars(op);
      alphas->push_back(op->red.host[S_GH] / op->red.host[S_DH]);
    }
    state = check(it, res);
    if (state != 0) break;
    if (precond != EXADG_B200_PRECOND_NONE) {
      precondition(op, precond, inv_diag, cheb, S_NEW, h, g);
      cg_update_d(op->red, S_NEW, S_GH, d, h, n, s); op->launches++;
      if (betas) { read_all_scalars(op); betas->push_back(op->red.host[S_NEW] / op->red.host[S_GH]); }
      std::swap(S_GH, S_NEW);
    } else {
      cg_update_d(op->red, S_RES, S_GH, d, g, n, s); op->launches++;
      if (betas) { read_all_scalars(op); betas->push_back(op->red.host[S_RES] / op->red.host[S_GH]); }
      std::swap(S_GH, S_RES);
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (n_iter) *n_iter = it;
  return state == 1 ? EXADG_B200_OK : EXADG_B200_ERR_NOT_CONVERGED;
}

void tridiag_extreme_eigs(const std::vector<double> & a, const std::vector<double> & b, double & emin, double & emax)
{
  const int n = (int)a.size();
  double lo = a[0], hi = a[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(b[i - 1]) : 0.0) + (i < n - 1 ? std::fabs(b[i]) : 0.0);
    lo = std::min(lo, a[i] - r); hi = std::max(hi, a[i] + r);
  }
  for (int which = 0; which < 2; ++which) {
    const int target = which == 0 ? 1 : n;
    double l = lo, u = hi;
    for (int it = 0; it < 200; ++it) {
      const double x = 0.5 * (l + u);
      int cnt = 0; double q = 1.0;
      for (int i = 0; i < n; ++i) {
        q = a[i] - x - (i > 0 ? b[i - 1] * b[i - 1] / q : 0.0);
        if (q == 0.0) q = 1e-300;
        if (q < 0.0) ++cnt;
      }
      if (cnt >= target) u = x; else l = x;
    }
    (which == 0 ? emin : emax) = 0.5 * (l + u);
  }
}

// dealii::PreconditionChebyshev::vmult / step with point-Jacobi (chebyshev_smoother.h:79-119)
void cheb_run(exadg_b200_chebyshev * ch, double * x, const double * b, bool zero_start)
{
  exadg_b200_operator * op = ch->op;
  const int64_t n = op->n_local; cudaStream_t s = op->stream;
  const double theta = ch->theta, delta = ch->delta;
  if (!zero_start) apply(op, ch->r, x, false);
  cheb_first(x, ch->xold, ch->inv_diag, b, ch->r, 1.0 / theta, zero_start, n, s); op->launches++;
  if (ch->degree < 2 || std::fabs(delta) < 1e-40) return;
  double rhok = delta / theta; const double sigma = theta / delta;
  for (int k = 0; k < ch->degree - 1; ++k) {
    apply(op, ch->r, x, false);
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / delta;
    rhok = rhokp;
    cheb_step(x, ch->xold, ch->inv_diag, b, ch->r, f1, f2, n, s); op->launches++;
  }
}

// ---- multigrid ----
// 1-D embedding matrix: Lagrange basis on the Gauss-Lobatto nodes of FE_DGQ(kc) evaluated at x = a + b xi_i, xi_i the nodes of FE_DGQ(kf)
void embedding_1d(int kf, int kc, real_t a, real_t b, double * I /*[kf+1][kc+1]*/)
{
  Tables1D f(kf), c(kc);
  std::vector<real_t> v, d;
  for (int i = 0; i <= kf; ++i) {
    lagrange_at(c.xn, a + b * f.xn[i], v, d);
    for (int j = 0; j <= kc; ++j) I[i * (kc + 1) + j] = (double)v[j];
  }
}

// MultigridAlgorithm::v_cycle (I/solvers_and_preconditioners/multigrid/multigrid_algorithm.h:173-243), preconditioner mode
void mg_v_cycle(exadg_b200_multigrid * mg, int level)
{
  exadg_b200_operator * op = mg->ops[level];
  const int64_t n = op->n_local; cudaStream_t s = op->stream;
  if (level == 0) {
    // MGCoarseKrylov::operator() (coarse_grid_solvers.h:166-232): r = src (minus its mean for singular operators), CG with point Jacobi
    // from the solution vector as it stands (zero before the first cycle, the previous coarse solution afterwards, as in the reference)
    CUDA_CHECK(cudaMemcpyAsync(mg->coarse_rhs, mg->defect[0], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (op->mesh.pure_neumann_or_periodic()) {
      sum(op->red, 4, mg->coarse_rhs, n, s); op->launches++;
      allreduce(op, op->red.result + 4, 1);
      const double mean = read_scalar(op, 4) / (double)op->dev.n_global_dofs;
      add_scalar(mg->coarse_rhs, -mean, n, s); op->launches++;
    }
    int its = 0;
    const int st = cg(op, mg->solution[0], mg->coarse_rhs, EXADG_B200_PRECOND_POINT_JACOBI, mg->coarse_inv_diag, nullptr, mg->coarse_abs_tol, mg->coarse_rel_tol,
                      mg->coarse_max_iter, &its, nullptr, nullptr, nullptr);
    mg->coarse_iterations += its;
    if (st != EXADG_B200_OK) throw std::runtime_error("multigrid: the coarse-grid CG solver did not converge (SolverControl::NoConvergence)");
    return;
  }
  exadg_b200_operator * opc = mg->ops[level - 1];
  cheb_run(mg->smoothers[level], mg->solution[level], mg->defect[level], true);      // pre-smoothing, zero initial guess
  apply(op, mg->t[level], mg->solution[level], false);                                // vmult_interface_down
  axpby(1.0, mg->defect[level], -1.0, mg->t[level], n, s); op->launches++;            // t = defect - A solution
  const int64_t n_coarse_blocks = opc->dev.n_owned * opc->dev.n_components; // p-transfer of a vector-valued field: block by block
  restrict_add(mg->transfer[level], mg->defect[level - 1], mg->t[level], n_coarse_blocks, s); op->launches++;
  mg_v_cycle(mg, level - 1);
  prolongate_add(mg->transfer[level], mg->solution[level], mg->solution[level - 1], n_coarse_blocks, s); op->launches++;
  cheb_run(mg->smoothers[level], mg->solution[level], mg->defect[level], false);     // post-smoothing
}

// MultigridAlgorithm::vmult (multigrid_algorithm.h:88-109)
void mg_vmult(exadg_b200_multigrid * mg, double * dst, const double * src)
{
  if (!mg) throw std::invalid_argument("null multigrid");
  const int L = (int)mg->ops.size() - 1;
  cudaStream_t s = mg->ops[L]->stream;
  for (int l = 0; l < L; ++l) mg->ops[l]->stream = s; // the coarser levels follow the stream of the finest operator (they do not own one)
  for (int l = 0; l < L; ++l) { fill(mg->defect[l], 0.0, mg->ops[l]->n_local, s); mg->ops[l]->launches++; }
  CUDA_CHECK(cudaMemcpyAsync(mg->defect[L], src, (size_t)mg->ops[L]->n_local * sizeof(double), cudaMemcpyDeviceToDevice, s));
  mg_v_cycle(mg, L);
  CUDA_CHECK(cudaMemcpyAsync(dst, mg->solution[L], (size_t)mg->ops[L]->n_local * sizeof(double), cudaMemcpyDeviceToDevice, s));
  mg->cycles++;
}
} // namespace

extern "C" {

const char * exadg_b200_last_error(void) { return g_last_error.c_str(); }
int exadg_b200_version(void) { return 100; }

static void set_helmholtz(exadg_b200_operator * op, const exadg_b200_helmholtz_data * hd)
{
  if (!hd) return;
  if (hd->n_components < 1 || hd->n_components > 3) throw std::invalid_argument("n_components must be 1..3");
  op->dev.helmholtz = true; op->dev.n_components = hd->n_components; op->dev.mass_coeff = hd->scaling_factor_mass; op->dev.laplace_coeff = hd->viscosity;
}

static int create_hypercube(const exadg_b200_hypercube_desc * desc, const exadg_b200_helmholtz_data * hdata, exadg_b200_operator ** out)
{
  return guarded([&]() {
    if (!desc || !out) throw std::invalid_argument("null argument");
    if (desc->degree < 1 || desc->degree > 7) throw std::invalid_argument("degree must be in 1..7");
    HypercubeDesc hd;
    hd.n_sub = desc->n_subdivisions; hd.refine = desc->n_refinements; hd.mapping_degree = desc->mapping_degree;
    hd.deformation = desc->deformation; hd.frequency = desc->frequency;
    for (int f = 0; f < 6; ++f) { if (desc->boundary[f] < 0 || desc->boundary[f] > 2) throw std::invalid_argument("bad boundary type"); hd.bc[f] = desc->boundary[f]; }
    hd.rank = desc->rank; hd.world = desc->world < 1 ? 1 : desc->world;
    std::unique_ptr<exadg_b200_operator> op(new exadg_b200_operator);
    op->dev.degree = desc->degree;
    set_helmholtz(op.get(), hdata);
    op->mesh = make_hypercube(hd);
    finish_setup(op.get(), desc->ip_factor, desc->force_general != 0);
    *out = op.release();
    return EXADG_B200_OK;
  });
}

static int create_from_mesh(const exadg_b200_mesh_desc * desc, const exadg_b200_helmholtz_data * hdata, exadg_b200_operator ** out)
{
  return guarded([&]() {
    if (!desc || !out || !desc->mapping_points || !desc->neighbors || !desc->neighbor_face || !desc->boundary_type) throw std::invalid_argument("null argument");
    if (desc->degree < 1 || desc->degree > 7) throw std::invalid_argument("degree must be in 1..7");
    if (desc->mapping_degree < 1 || desc->mapping_degree > 8) throw std::invalid_argument("mapping_degree must be in 1..8");
    std::unique_ptr<exadg_b200_operator> op(new exadg_b200_operator);
    op->dev.degree = desc->degree;
    set_helmholtz(op.get(), hdata);
    HostMesh & M = op->mesh;
    M.mapping_degree = desc->mapping_degree; M.n_owned = desc->n_cells_owned; M.n_ghost = desc->n_cells_ghost;
    M.n_global_cells = desc->n_global_cells > 0 ? desc->n_global_cells : desc->n_cells_owned;
    M.global_offset = desc->global_cell_offset;
    M.singular = desc->operator_is_singular < 0 ? -1 : (desc->operator_is_singular ? 1 : 0);
    const int64_t nloc = M.n_owned + M.n_ghost;
    const int np3 = (M.mapping_degree + 1) * (M.mapping_degree + 1) * (M.mapping_degree + 1);
    M.xmap.assign(desc->mapping_points, desc->mapping_points + (size_t)nloc * np3 * 3);
    M.nb.assign(desc->neighbors, desc->neighbors + M.n_owned * 6);
    M.nbface.assign(desc->neighbor_face, desc->neighbor_face + M.n_owned * 6);
    M.bt.assign(desc->boundary_type, desc->boundary_type + nloc * 6);
    for (int64_t i = 0; i < M.n_owned * 6; ++i) {
      if (M.nb[i] >= nloc) throw std::invalid_argument("neighbour index out of range");
      if ((M.nb[i] < 0) != (M.bt[i] != BT_INTERIOR)) throw std::invalid_argument("boundary_type inconsistent with neighbors");
    }
    double h[3];
    M.cartesian_uniform = detect_cartesian_uniform(M, h);
    if (M.cartesian_uniform) for (int e = 0; e < 3; ++e) M.h[e] = h[e];
    M.build_faces();
    finish_setup(op.get(), desc->ip_factor, desc->force_general != 0);
    *out = op.release();
    return EXADG_B200_OK;
  });
}

int exadg_b200_create_hypercube(const exadg_b200_hypercube_desc * desc, exadg_b200_operator ** out) { return create_hypercube(desc, nullptr, out); }
int exadg_b200_create(const exadg_b200_mesh_desc * desc, exadg_b200_operator ** out) { return create_from_mesh(desc, nullptr, out); }
int exadg_b200_create_hypercube_helmholtz(const exadg_b200_hypercube_desc * desc, const exadg_b200_helmholtz_data * data, exadg_b200_operator ** out)
{ if (!data) { g_last_error = "null helmholtz data"; return EXADG_B200_ERR_ARG; } return create_hypercube(desc, data, out); }
int exadg_b200_create_helmholtz(const exadg_b200_mesh_desc * desc, const exadg_b200_helmholtz_data * data, exadg_b200_operator ** out)
{ if (!data) { g_last_error = "null helmholtz data"; return EXADG_B200_ERR_ARG; } return create_from_mesh(desc, data, out); }

int exadg_b200_n_components(const exadg_b200_operator * op) { return op ? op->dev.n_components : -1; }
/* MomentumOperator::set_scaling_factor_mass_operator (momentum_operator.cpp:147-153): gamma_0 / dt changes with the time step */
int exadg_b200_set_scaling_factor_mass(exadg_b200_operator * op, double scaling_factor_mass)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    if (!op->dev.helmholtz) throw std::runtime_error("exadg_b200_set_scaling_factor_mass: not a Helmholtz operator (create it with exadg_b200_create_*_helmholtz)");
    op->dev.mass_coeff = scaling_factor_mass; op->dev_helm.mass_coeff = scaling_factor_mass;
    return EXADG_B200_OK;
  });
}
/* InverseMassOperator::apply (I/operators/inverse_mass_operator.h) */
int exadg_b200_inverse_mass_vmult(exadg_b200_operator * op, double * dst, const double * src)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    check_ptr(dst, "dst"); check_ptr(src, "src");
    launch_inverse_mass(op->dev, dst, src, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}

int exadg_b200_destroy(exadg_b200_operator * op)
{
  if (!op) return EXADG_B200_OK;
  cudaDeviceSynchronize();
  DeviceOperator & D = op->dev;
  cartesian_plan_destroy(D);
  cartesian_plan_destroy(op->dev_helm); // (its neighbour table is dev's)
  if (op->p2p) D.ghost = op->ghost_alloc;
  cudaFree(D.cellJxW); cudaFree(op->d_hyb_batches); cudaFree(op->d_hyb_cells);
  cudaFree(D.nb); cudaFree(D.face_id); cudaFree(D.face_info); cudaFree(D.cellG); cudaFree(D.faceG); cudaFree(D.tau_f); cudaFree(D.tau_cell); cudaFree(D.ghost);
  for (auto p : op->p2p_peer_regions) if (p) cudaIpcCloseMemHandle(p);
  if (op->p2p_region) cudaFree(op->p2p_region);
  cudaFree(op->d_put_done); cudaFree(op->d_work_counter);
  for (auto p : op->d_send_lists) cudaFree(p);
  for (auto p : op->d_send_bufs) cudaFree(p);
  cudaFree(op->d_interior); cudaFree(op->d_boundary); cudaFree(op->d_cell_diag);
  post_destroy(op->post);
  for (int i = 0; i < 4; ++i) cudaFree(op->w[i]);
  cudaFree(op->d_stage_src); cudaFree(op->d_stage_dst);
  cudaFree(op->d_iota);
  cudaFree(op->d_hs_units);
  for (auto e : op->hs_ev) cudaEventDestroy(e);
  for (auto e : op->hp_ev_in) cudaEventDestroy(e);
  for (auto e : op->hp_ev_cmp) cudaEventDestroy(e);
  if (op->hp_start) cudaEventDestroy(op->hp_start);
  if (op->hp_in) cudaStreamDestroy(op->hp_in);
  if (op->hp_out) cudaStreamDestroy(op->hp_out);
  reducer_free(op->red);
  if (op->own_comm && op->comm) nccl().CommDestroy(op->comm);
  if (op->ev_packed) cudaEventDestroy(op->ev_packed);
  if (op->ev_halo) cudaEventDestroy(op->ev_halo);
  if (op->ev_order) cudaEventDestroy(op->ev_order);
  if (op->own_stream && op->stream) cudaStreamDestroy(op->stream);
  if (op->comm_stream) cudaStreamDestroy(op->comm_stream);
  delete op;
  return EXADG_B200_OK;
}

int exadg_b200_set_stream(exadg_b200_operator * op, void * cuda_stream)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    if (op->own_stream && op->stream) { CUDA_CHECK(cudaStreamSynchronize(op->stream)); cudaStreamDestroy(op->stream); }
    op->stream = (cudaStream_t)cuda_stream; op->own_stream = false;
    return EXADG_B200_OK;
  });
}
int exadg_b200_synchronize(exadg_b200_operator * op) { return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); CUDA_CHECK(cudaStreamSynchronize(op->stream)); return EXADG_B200_OK; }); }
int exadg_b200_wait_stream(exadg_b200_operator * op, void * cuda_stream)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    cudaStream_t other = (cudaStream_t)cuda_stream;
    if (other == op->stream) return EXADG_B200_OK;
    CUDA_CHECK(cudaEventRecord(op->ev_order, other));
    CUDA_CHECK(cudaStreamWaitEvent(op->stream, op->ev_order, 0));
    return EXADG_B200_OK;
  });
}
int exadg_b200_stream_wait_operator(exadg_b200_operator * op, void * cuda_stream)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    cudaStream_t other = (cudaStream_t)cuda_stream;
    if (other == op->stream) return EXADG_B200_OK;
    CUDA_CHECK(cudaEventRecord(op->ev_order, op->stream));
    CUDA_CHECK(cudaStreamWaitEvent(other, op->ev_order, 0));
    return EXADG_B200_OK;
  });
}
int exadg_b200_operator_is_singular(const exadg_b200_operator * op) { return (op && op->mesh.pure_neumann_or_periodic()) ? 1 : 0; }
int exadg_b200_degree(const exadg_b200_operator * op) { return op ? op->dev.degree : -1; }
int exadg_b200_set_kernel_variant(exadg_b200_operator * op, int variant)
{
  if (!op || variant < -1 || variant > 15) return EXADG_B200_ERR_ARG;
  op->dev.cart_variant = variant;
  return EXADG_B200_OK;
}
int exadg_b200_get_kernel_variant(const exadg_b200_operator * op) { return op ? (op->dev.cart_variant >= 0 ? op->dev.cart_variant : cartesian_kernel_variant(-1)) : -1; }

int64_t exadg_b200_n(const exadg_b200_operator * op) { return op ? op->dev.n_global_dofs : -1; }
int64_t exadg_b200_local_size(const exadg_b200_operator * op) { return op ? op->n_local : -1; }
int64_t exadg_b200_n_cells_owned(const exadg_b200_operator * op) { return op ? op->dev.n_owned : -1; }
int64_t exadg_b200_n_cells_ghost(const exadg_b200_operator * op) { return op ? op->dev.n_ghost : -1; }
int exadg_b200_is_cartesian_path(const exadg_b200_operator * op) { return op ? (op->dev.cartesian ? 1 : (op->hybrid ? 2 : 0)) : 0; }
int exadg_b200_kernel_launches(const exadg_b200_operator * op, int64_t * count) { if (!op || !count) return EXADG_B200_ERR_ARG; *count = op->launches; return EXADG_B200_OK; }

int exadg_b200_initialize_dof_vector(const exadg_b200_operator * op, double ** vec)
{
  return guarded([&]() {
    if (!op || !vec) throw std::invalid_argument("null argument");
    const size_t bytes = (size_t)std::max<int64_t>(op->n_local, 1) * sizeof(double);
    CUDA_CHECK(cudaMalloc(vec, bytes));
    CUDA_CHECK(cudaMemsetAsync(*vec, 0, bytes, op->stream)); // ordered before the kernels of this operator that write the vector
    return EXADG_B200_OK;
  });
}
int exadg_b200_free_dof_vector(double * vec) { return guarded([&]() { CUDA_CHECK(cudaFree(vec)); return EXADG_B200_OK; }); }

int exadg_b200_vmult(exadg_b200_operator * op, double * dst, const double * src)
{ return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); apply(op, dst, src, false); return EXADG_B200_OK; }); }
int exadg_b200_vmult_add(exadg_b200_operator * op, double * dst, const double * src)
{ return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); apply(op, dst, src, true); return EXADG_B200_OK; }); }

int exadg_b200_vmult_host(exadg_b200_operator * op, double * dst_host, const double * src_host)
{
  return guarded([&]() {
    if (!op || !dst_host || !src_host) throw std::invalid_argument("null argument");
    const size_t bytes = (size_t)op->n_local * sizeof(double);
    if (!op->d_stage_src) CUDA_CHECK(cudaMalloc(&op->d_stage_src, std::max<size_t>(bytes, 16)));
    if (!op->d_stage_dst) CUDA_CHECK(cudaMalloc(&op->d_stage_dst, std::max<size_t>(bytes, 16)));
    CUDA_CHECK(cudaMemcpyAsync(op->d_stage_src, src_host, bytes, cudaMemcpyHostToDevice, op->stream));
    apply(op, op->d_stage_dst, op->d_stage_src, false);
    CUDA_CHECK(cudaMemcpyAsync(dst_host, op->d_stage_dst, bytes, cudaMemcpyDeviceToHost, op->stream));
    CUDA_CHECK(cudaStreamSynchronize(op->stream));
    return EXADG_B200_OK;
  });
}

// Same result as exadg_b200_vmult_host, but the upload of src, the operator and the download of dst overlap.  Staged variant: the vector is cut
// into contiguous chunks (host_pipeline.hpp), a chunk is applied as soon as the chunks holding its face neighbours have arrived and is
// downloaded right behind its kernel on a third stream (unpartitioned operators only).  Direct variant: vmult_host_direct below.  Host
// buffers should be pinned.
static int64_t host_pipeline_cells_per_chunk(int batch)
{
  // about 12288 cells (24 blocks of 8^3 cells of the Morton curve on refined hypercubes; 12 MB per copy at k = 4), a multiple of the kernels'
  // batch size.  Measured on the 96^3 box (k = 4, scripts/r02_shot24.sh): 1536 / 4608 / 12288 / 32768 cells per chunk -> 3.76 / 4.21 / 4.34 / 3.97
  // GDoF/s end to end: small chunks overlap best on paper but pay per-chunk launch, event and copy-setup costs.
  static const int64_t target = []() { const char * e = getenv("EXADG_B200_HP_CELLS"); const long v = e ? std::atol(e) : 12288; return (int64_t)(v > 0 ? v : 12288); }();
  const int64_t per = std::max<int64_t>(1, (target + batch / 2) / batch);
  return per * batch;
}

// Direct variant (HostStreamPlan, host_pipeline.hpp): src is uploaded in pieces in address order; behind every piece ONE launch applies
// the units (batches / cells) whose own cells and face neighbours are complete with it, and the kernels store dst straight into the
// caller's pinned host buffer through its device mapping (bulk stores / coalesced stores over PCIe; every DoF of dst is written exactly
// once, as on the device path).  No staging vector and no copy-engine download: the download has the granularity of a kernel unit.
static int64_t host_stream_cells_per_piece(int unit)
{
  // measured on the 96^3 box (k = 4, scripts/r02_shot40.sh): 3072 / 6144 / 12288 / 24576 / 49152 cells per piece -> 4.47 / 4.79 / 5.03 / 5.19 / 5.18 GDoF/s
  // end to end (staged variant 4.34, sequential entry 3.25)
  static const int64_t target = []() { const char * e = getenv("EXADG_B200_HS_CELLS"); const long v = e ? std::atol(e) : 24576; return (int64_t)(v > 0 ? v : 24576); }();
  const int64_t per = std::max<int64_t>(1, (target + unit / 2) / unit);
  return per * unit;
}

// device address of a host buffer the GPU can write (cudaHostAlloc / cudaHostRegister memory under unified addressing), else nullptr
static double * device_view_of_host(double * host)
{
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return static_cast<double *>(at.devicePointer);
}

static int vmult_host_direct(exadg_b200_operator * op, double * dst_dev, const double * src_host)
{
  HostMesh & M = op->mesh;
  const int n3 = op->dev.n * op->dev.n * op->dev.n;
  const bool cart = op->dev.cartesian;
  const int B = cart ? cartesian_batch_size(op->dev) : 1;
  const size_t bytes = (size_t)op->n_local * sizeof(double);
  if (!op->hs_built) {
    op->hs = build_host_stream_plan(M.nb.data(), M.n_owned, B, host_stream_cells_per_piece(B), /*allow_ghosts=*/M.world > 1 && !M.peers.empty());
    const HostStreamPlan & P = op->hs;
    if (P.n_steps > 0) {
      CUDA_CHECK(cudaMalloc(&op->d_hs_units, std::max<size_t>(P.units.size(), 1) * sizeof(int32_t)));
      CUDA_CHECK(cudaMemcpy(op->d_hs_units, P.units.data(), P.units.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      if (!op->hp_in) CUDA_CHECK(cudaStreamCreateWithFlags(&op->hp_in, cudaStreamNonBlocking));
      if (!op->hp_start) CUDA_CHECK(cudaEventCreateWithFlags(&op->hp_start, cudaEventDisableTiming));
      op->hs_ev.assign(P.n_steps + 1, nullptr); // the last one: all uploads complete
      for (auto & e : op->hs_ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    op->hs_built = true;
  }
  const HostStreamPlan & P = op->hs;
  const bool import_ghosts = M.world > 1 && !M.peers.empty();
  // (a rank of a partition that owns no cell has no plan but still takes part in the ghost import below)
  if (P.n_steps == 0 && !(import_ghosts && M.n_owned == 0)) { g_last_error = "the pipelined host-buffer vmult does not apply to this operator"; return (int)EXADG_B200_ERR_UNSUPPORTED; }
  if (!op->d_stage_src) CUDA_CHECK(cudaMalloc(&op->d_stage_src, std::max<size_t>(bytes, 16)));
  if (P.n_steps > 0) {
    // the upload stream may not overtake earlier work of the operator's stream on the staging vector
    CUDA_CHECK(cudaEventRecord(op->hp_start, op->stream));
    CUDA_CHECK(cudaStreamWaitEvent(op->hp_in, op->hp_start, 0));
  }
  for (int i = 0; i < P.n_steps; ++i) {
    const int64_t c0 = P.piece_begin[i], c1 = P.piece_begin[i + 1];
    CUDA_CHECK(cudaMemcpyAsync(op->d_stage_src + c0 * n3, src_host + c0 * n3, (size_t)(c1 - c0) * n3 * sizeof(double), cudaMemcpyHostToDevice, op->hp_in));
    const int64_t u0 = P.step_begin[i], n_units = P.step_begin[i + 1] - u0;
    if (n_units == 0) continue; // uploads complete in order: a later event covers this piece as well
    CUDA_CHECK(cudaEventRecord(op->hs_ev[i], op->hp_in));
    CUDA_CHECK(cudaStreamWaitEvent(op->stream, op->hs_ev[i], 0));
    if (cart) launch_vmult_cartesian_list(op->dev, dst_dev, op->d_stage_src, false, op->d_hs_units + u0, (int)n_units, op->stream);
    else launch_vmult_general(op->dev, dst_dev, op->d_stage_src, false, op->d_hs_units + u0, n_units, op->stream);
    op->launches++;
  }
  if (import_ghosts) {
    // partitioned operator: the units above touch no ghost cell; the others follow the ghost import of src (all ranks enter this call), which
    // needs this rank's complete src on the device
    if (P.n_steps > 0) {
      CUDA_CHECK(cudaEventRecord(op->hs_ev[P.n_steps], op->hp_in));
      CUDA_CHECK(cudaStreamWaitEvent(op->stream, op->hs_ev[P.n_steps], 0));
    }
    apply(op, dst_dev, op->d_stage_src, false, /*boundary_only=*/true);
  }
  CUDA_CHECK(cudaStreamSynchronize(op->stream)); // kernel completion makes the stores to host memory visible to the caller
  return (int)EXADG_B200_OK;
}

int exadg_b200_set_host_pipeline_mode(exadg_b200_operator * op, int mode)
{
  if (!op) return -1;
  const int prev = op->hp_mode;
  if (mode >= 0 && mode <= 2) op->hp_mode = mode;
  return prev;
}

int exadg_b200_vmult_host_pipelined(exadg_b200_operator * op, double * dst_host, const double * src_host)
{
  return guarded([&]() {
    if (!op || !dst_host || !src_host) throw std::invalid_argument("null argument");
    HostMesh & M = op->mesh;
    const bool partitioned = M.world > 1 || M.n_ghost > 0;
    if (op->dev.helmholtz) { g_last_error = "the pipelined host-buffer vmult is implemented for the scalar Laplace operator (use exadg_b200_vmult_host)"; return (int)EXADG_B200_ERR_UNSUPPORTED; }
    if (partitioned && M.peers.empty() && M.n_ghost > 0) { g_last_error = "the pipelined host-buffer vmult cannot import ghost cells the caller owns"; return (int)EXADG_B200_ERR_UNSUPPORTED; }
    {
      static const int env_mode = []() { const char * e = getenv("EXADG_B200_HOST_PIPELINE"); return !e ? 0 : (!strcmp(e, "staged") ? 1 : (!strcmp(e, "direct") ? 2 : 0)); }();
      const int mode = op->hp_mode ? op->hp_mode : env_mode;
      // auto: the affine kernels store whole batches (bulk / coalesced stores), which PCIe takes at full rate; the general kernel stores line by
      // line and keeps the staged download
      if (mode == 2 || (mode == 0 && op->dev.cartesian)) {
        double * const dst_dev = device_view_of_host(dst_host);
        if (dst_dev) return vmult_host_direct(op, dst_dev, src_host);
        if (mode == 2) throw std::invalid_argument("direct mode needs a dst_host the GPU can address (cudaHostAlloc / cudaHostRegister memory)");
      }
    }
    if (partitioned) { g_last_error = "the staged variant of the pipelined host-buffer vmult is for unpartitioned operators"; return (int)EXADG_B200_ERR_UNSUPPORTED; }
    const int n3 = op->dev.n * op->dev.n * op->dev.n;
    const bool cart = op->dev.cartesian;
    const int B = cart ? cartesian_batch_size(op->dev) : 1;
    const size_t bytes = (size_t)op->n_local * sizeof(double);
    if (!op->hp_built) {
      op->hp = build_host_pipeline(M.nb.data(), M.n_owned, host_pipeline_cells_per_chunk(B));
      const int64_t n_units = cart ? (int64_t)cartesian_n_batches(op->dev) : M.n_owned;
      std::vector<int32_t> iota((size_t)std::max<int64_t>(n_units, 1));
      for (size_t i = 0; i < iota.size(); ++i) iota[i] = (int32_t)i;
      CUDA_CHECK(cudaMalloc(&op->d_iota, iota.size() * sizeof(int32_t)));
      CUDA_CHECK(cudaMemcpy(op->d_iota, iota.data(), iota.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      if (!op->hp_in) CUDA_CHECK(cudaStreamCreateWithFlags(&op->hp_in, cudaStreamNonBlocking)); // (shared with the direct variant)
      CUDA_CHECK(cudaStreamCreateWithFlags(&op->hp_out, cudaStreamNonBlocking));
      if (!op->hp_start) CUDA_CHECK(cudaEventCreateWithFlags(&op->hp_start, cudaEventDisableTiming));
      op->hp_ev_in.assign(op->hp.n_chunks, nullptr); op->hp_ev_cmp.assign(op->hp.n_chunks, nullptr);
      for (int c = 0; c < op->hp.n_chunks; ++c) {
        CUDA_CHECK(cudaEventCreateWithFlags(&op->hp_ev_in[c], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&op->hp_ev_cmp[c], cudaEventDisableTiming));
      }
      op->hp_built = true;
    }
    const HostPipelinePlan & P = op->hp;
    if (P.n_chunks == 0) { g_last_error = "the pipelined host-buffer vmult is for unpartitioned operators"; return (int)EXADG_B200_ERR_UNSUPPORTED; }
    if (!op->d_stage_src) CUDA_CHECK(cudaMalloc(&op->d_stage_src, std::max<size_t>(bytes, 16)));
    if (!op->d_stage_dst) CUDA_CHECK(cudaMalloc(&op->d_stage_dst, std::max<size_t>(bytes, 16)));
    // neither copy stream may overtake earlier work of the operator's stream on the staging vectors
    CUDA_CHECK(cudaEventRecord(op->hp_start, op->stream));
    CUDA_CHECK(cudaStreamWaitEvent(op->hp_in, op->hp_start, 0));
    CUDA_CHECK(cudaStreamWaitEvent(op->hp_out, op->hp_start, 0));
    auto range = [&](int c, int64_t & c0, int64_t & c1) { c0 = (int64_t)c * P.cells_per_chunk; c1 = std::min<int64_t>(M.n_owned, c0 + P.cells_per_chunk); };
    size_t next = 0; // next entry of compute_order (its chunks become computable in upload order)
    for (int32_t b : P.upload_order) {
      int64_t c0, c1; range(b, c0, c1);
      CUDA_CHECK(cudaMemcpyAsync(op->d_stage_src + c0 * n3, src_host + c0 * n3, (size_t)(c1 - c0) * n3 * sizeof(double), cudaMemcpyHostToDevice, op->hp_in));
      CUDA_CHECK(cudaEventRecord(op->hp_ev_in[b], op->hp_in));
      while (next < P.compute_order.size() && P.ready_chunk[P.compute_order[next]] == b) {
        const int c = P.compute_order[next++];
        range(c, c0, c1);
        CUDA_CHECK(cudaStreamWaitEvent(op->stream, op->hp_ev_in[b], 0)); // uploads complete in order: the last dependency covers the others
        if (cart) launch_vmult_cartesian_list(op->dev, op->d_stage_dst, op->d_stage_src, false, op->d_iota + c0 / B, (int)((c1 - c0 + B - 1) / B), op->stream);
        else launch_vmult_general(op->dev, op->d_stage_dst, op->d_stage_src, false, op->d_iota + c0, c1 - c0, op->stream);
        op->launches++;
        CUDA_CHECK(cudaEventRecord(op->hp_ev_cmp[c], op->stream));
        CUDA_CHECK(cudaStreamWaitEvent(op->hp_out, op->hp_ev_cmp[c], 0));
        CUDA_CHECK(cudaMemcpyAsync(dst_host + c0 * n3, op->d_stage_dst + c0 * n3, (size_t)(c1 - c0) * n3 * sizeof(double), cudaMemcpyDeviceToHost, op->hp_out));
      }
    }
    if (next != P.compute_order.size()) throw std::runtime_error("host pipeline plan incomplete");
    CUDA_CHECK(cudaStreamSynchronize(op->hp_out));
    CUDA_CHECK(cudaStreamSynchronize(op->stream));
    return (int)EXADG_B200_OK;
  });
}

// host-only view of that plan (no CUDA call), for the CPU tests: sizes with null arrays, then the three tables of n_chunks entries;
// model = duration of one call in units of a one-direction transfer (2 = no overlap)
int exadg_b200_host_pipeline_plan(const exadg_b200_hypercube_desc * desc, int64_t cells_per_chunk, int32_t * n_chunks, int32_t * upload_order, int32_t * compute_order,
                                  int32_t * ready_chunk, double * model)
{
  return guarded([&]() {
    if (!desc || !n_chunks) throw std::invalid_argument("null argument");
    HypercubeDesc hd;
    hd.n_sub = desc->n_subdivisions; hd.refine = desc->n_refinements; hd.mapping_degree = 1;
    hd.deformation = 0.0; hd.frequency = desc->frequency;
    for (int f = 0; f < 6; ++f) hd.bc[f] = desc->boundary[f];
    hd.rank = desc->rank; hd.world = desc->world < 1 ? 1 : desc->world;
    const HostMesh mesh = make_hypercube(hd);
    const HostPipelinePlan P = build_host_pipeline(mesh.nb.data(), mesh.n_owned, cells_per_chunk > 0 ? cells_per_chunk : host_pipeline_cells_per_chunk(24));
    *n_chunks = P.n_chunks;
    if (upload_order) std::copy(P.upload_order.begin(), P.upload_order.end(), upload_order);
    if (compute_order) std::copy(P.compute_order.begin(), P.compute_order.end(), compute_order);
    if (ready_chunk) std::copy(P.ready_chunk.begin(), P.ready_chunk.end(), ready_chunk);
    if (model) *model = host_pipeline_model(P);
    return (int)EXADG_B200_OK;
  });
}

// host-only view of the stream plan of the direct variant (no CUDA call; CPU tests): sizes with null arrays, then the tables
int exadg_b200_host_stream_plan(const exadg_b200_hypercube_desc * desc, int unit, int64_t cells_per_piece, int32_t * n_steps, int64_t * n_units, int64_t * piece_begin,
                                int64_t * step_begin, int32_t * units, double * model)
{
  return guarded([&]() {
    if (!desc || !n_steps || !n_units) throw std::invalid_argument("null argument");
    if (unit <= 0) throw std::invalid_argument("unit must be positive");
    HypercubeDesc hd;
    hd.n_sub = desc->n_subdivisions; hd.refine = desc->n_refinements; hd.mapping_degree = 1;
    hd.deformation = 0.0; hd.frequency = desc->frequency;
    for (int f = 0; f < 6; ++f) hd.bc[f] = desc->boundary[f];
    hd.rank = desc->rank; hd.world = desc->world < 1 ? 1 : desc->world;
    const HostMesh mesh = make_hypercube(hd);
    const HostStreamPlan P = build_host_stream_plan(mesh.nb.data(), mesh.n_owned, unit, cells_per_piece > 0 ? cells_per_piece : host_stream_cells_per_piece(unit), hd.world > 1);
    *n_steps = P.n_steps; *n_units = (int64_t)P.units.size();
    if (piece_begin) std::copy(P.piece_begin.begin(), P.piece_begin.end(), piece_begin);
    if (step_begin) std::copy(P.step_begin.begin(), P.step_begin.end(), step_begin);
    if (units) std::copy(P.units.begin(), P.units.end(), units);
    if (model) *model = host_stream_model(P);
    return (int)EXADG_B200_OK;
  });
}

int exadg_b200_calculate_diagonal(exadg_b200_operator * op, double * d) { return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); diagonal(op, d, false); return EXADG_B200_OK; }); }
int exadg_b200_add_diagonal(exadg_b200_operator * op, double * d) { return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); diagonal(op, d, true); return EXADG_B200_OK; }); }
int exadg_b200_calculate_inverse_diagonal(exadg_b200_operator * op, double * d)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    diagonal(op, d, false);
    invert_diagonal(d, op->n_local, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}


/* dealii::VectorTools::subtract_mean_value as used for the singular pressure-Poisson system (SURVEY 8 f-2;
 * incompressible_navier_stokes/time_integration/time_int_bdf_dual_splitting.cpp:655-656, compute_eigenvalues.h:52-53): global mean over all ranks */
int exadg_b200_subtract_mean_value(exadg_b200_operator * op, double * vec)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    require_self_contained(op, "exadg_b200_subtract_mean_value");
    check_ptr(vec, "vec");
    sum(op->red, 4, vec, op->n_local, op->stream); op->launches++;
    allreduce(op, op->red.result + 4, 1);
    const double mean = read_scalar(op, 4) / (double)op->dev.n_global_dofs;
    add_scalar(vec, -mean, op->n_local, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}

/* ---- inhomogeneous boundary data, right-hand side, error norms ---- */
static void * post_of(exadg_b200_operator * op)
{
  if (op->dev.helmholtz) throw std::runtime_error("rhs / evaluate / error norms are implemented for the scalar Laplace operator only");
  return post_get(op->post, op->dev, op->mesh, 0.0, op->stream);
}

int exadg_b200_n_boundary_faces(exadg_b200_operator * op, int64_t * n_faces)
{
  return guarded([&]() { if (!op || !n_faces) throw std::invalid_argument("null argument"); *n_faces = post_n_boundary_faces(post_of(op)); return EXADG_B200_OK; });
}
int exadg_b200_boundary_quadrature_points(exadg_b200_operator * op, double * xyz_host, uint8_t * type_host)
{
  return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); post_boundary_points(post_of(op), op->dev, xyz_host, type_host, op->stream); return EXADG_B200_OK; });
}
int exadg_b200_set_boundary_values(exadg_b200_operator * op, const double * values_host)
{
  return guarded([&]() { if (!op || !values_host) throw std::invalid_argument("null argument"); post_set_boundary_values(post_of(op), op->dev, values_host, op->stream); return EXADG_B200_OK; });
}
int exadg_b200_rhs_add(exadg_b200_operator * op, double * dst)
{
  return guarded([&]() { if (!op) throw std::invalid_argument("null operator"); check_ptr(dst, "dst"); post_boundary_inhom_add(post_of(op), op->dev, op->mesh, -1.0, dst, op->stream); op->launches++; return EXADG_B200_OK; });
}
int exadg_b200_rhs(exadg_b200_operator * op, double * dst)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    check_ptr(dst, "dst");
    CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)op->n_local * sizeof(double), op->stream));
    post_boundary_inhom_add(post_of(op), op->dev, op->mesh, -1.0, dst, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}
int exadg_b200_evaluate_add(exadg_b200_operator * op, double * dst, const double * src)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    apply(op, dst, src, true);
    post_boundary_inhom_add(post_of(op), op->dev, op->mesh, +1.0, dst, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}
int exadg_b200_evaluate(exadg_b200_operator * op, double * dst, const double * src)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    apply(op, dst, src, false);
    post_boundary_inhom_add(post_of(op), op->dev, op->mesh, +1.0, dst, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}
int exadg_b200_cell_quadrature_points(exadg_b200_operator * op, int n_q_points_1d, double * xyz_host)
{
  return guarded([&]() {
    if (!op || !xyz_host) throw std::invalid_argument("null argument");
    if (n_q_points_1d < 1 || n_q_points_1d > 10) throw std::invalid_argument("n_q_points_1d must be in 1..10");
    post_cell_points(post_of(op), op->dev, op->mesh, n_q_points_1d, xyz_host, op->stream);
    return EXADG_B200_OK;
  });
}
int exadg_b200_integrate_source_add(exadg_b200_operator * op, double * dst, const double * f_host)
{
  return guarded([&]() {
    if (!op || !f_host) throw std::invalid_argument("null argument");
    check_ptr(dst, "dst");
    post_source_add(post_of(op), op->dev, op->mesh, f_host, dst, op->stream); op->launches += 2;
    return EXADG_B200_OK;
  });
}
int exadg_b200_l2_error(exadg_b200_operator * op, const double * u, const double * exact_host, int relative, double * error)
{
  return guarded([&]() {
    if (!op || !exact_host || !error) throw std::invalid_argument("null argument");
    check_ptr(u, "u");
    const int nq = op->dev.degree + 3; // additional_quadrature_points = 3 (error_calculation.cpp:46)
    const int64_t nc = op->dev.n_owned;
    double * d_out = nullptr;
    CUDA_CHECK(cudaMalloc(&d_out, (size_t)std::max<int64_t>(2 * nc, 1) * sizeof(double)));
    CUDA_CHECK(cudaMemsetAsync(d_out, 0, (size_t)std::max<int64_t>(2 * nc, 1) * sizeof(double), op->stream));
    post_l2_cells(post_of(op), op->dev, op->mesh, nq, u, exact_host, d_out, op->stream); op->launches += 2;
    sum(op->red, 5, d_out, nc, op->stream); sum(op->red, 6, d_out + nc, nc, op->stream); op->launches += 2;
    allreduce(op, op->red.result + 5, 2);
    read_all_scalars(op);
    cudaFree(d_out);
    const double e2 = op->red.host[5], n2 = op->red.host[6];
    if (relative) {
      if (!(std::sqrt(n2) > 1e-15)) throw std::runtime_error("Cannot compute relative error since norm of solution tends to zero.");
      *error = std::sqrt(e2) / std::sqrt(n2);
    } else *error = std::sqrt(e2);
    return EXADG_B200_OK;
  });
}

int exadg_b200_jacobi_vmult(exadg_b200_operator * op, double * dst, const double * src, const double * inv_diag)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    check_ptr(dst, "dst"); check_ptr(src, "src"); check_ptr(inv_diag, "inverse_diagonal");
    jacobi_dot(op->red, 7, dst, inv_diag, src, op->n_local, op->stream); op->launches++;
    return EXADG_B200_OK;
  });
}

int exadg_b200_cg_solve(exadg_b200_operator * op, double * x, const double * b, int preconditioner, exadg_b200_chebyshev * cheb,
                        double abs_tol, double rel_tol, int max_iter, int * n_iter, double * residuals)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    require_self_contained(op, "exadg_b200_cg_solve");
    check_ptr(x, "x"); check_ptr(b, "b");
    double * inv_diag = nullptr;
    if (preconditioner == EXADG_B200_PRECOND_POINT_JACOBI) {
      inv_diag = op->work(3);
      diagonal(op, inv_diag, false);
      invert_diagonal(inv_diag, op->n_local, op->stream); op->launches++;
    } else if (preconditioner == EXADG_B200_PRECOND_CHEBYSHEV) {
      if (!cheb || cheb->op != op) throw std::invalid_argument("Chebyshev preconditioner missing or built for another operator");
    } else if (preconditioner != EXADG_B200_PRECOND_NONE) throw std::invalid_argument("unknown preconditioner");
    return cg(op, x, b, preconditioner, inv_diag, cheb, abs_tol, rel_tol, max_iter, n_iter, residuals, nullptr, nullptr);
  });
}

int exadg_b200_chebyshev_create(exadg_b200_operator * op, int degree, double smoothing_range, int eig_cg_n_iterations, exadg_b200_chebyshev ** out)
{
  return guarded([&]() {
    if (!op || !out) throw std::invalid_argument("null argument");
    require_self_contained(op, "exadg_b200_chebyshev_create");
    std::unique_ptr<exadg_b200_chebyshev> ch(new exadg_b200_chebyshev);
    ch->op = op; ch->degree = degree; ch->smoothing_range = smoothing_range; ch->eig_cg_n_iterations = eig_cg_n_iterations;
    const int64_t n = op->n_local; const size_t bytes = (size_t)std::max<int64_t>(n, 1) * sizeof(double);
    CUDA_CHECK(cudaMalloc(&ch->inv_diag, bytes)); CUDA_CHECK(cudaMalloc(&ch->xold, bytes)); CUDA_CHECK(cudaMalloc(&ch->r, bytes));
    diagonal(op, ch->inv_diag, false);
    invert_diagonal(ch->inv_diag, n, op->stream); op->launches++;
    // eigenvalue estimate: <= eig_cg_n_iterations Jacobi-preconditioned CG steps on (global index mod 11) - mean
    double * rhs = ch->r, * sol = ch->xold;
    fill_mod11(rhs, op->mesh.global_offset * (int64_t)(op->dev.n * op->dev.n * op->dev.n), n, op->stream); op->launches++;
    sum(op->red, 6, rhs, n, op->stream); op->launches++;
    allreduce(op, op->red.result + 6, 1);
    const double mean = read_scalar(op, 6) / (double)op->dev.n_global_dofs;
    add_scalar(rhs, -mean, n, op->stream); op->launches++;
    fill(sol, 0.0, n, op->stream); op->launches++;
    std::vector<double> alphas, betas; int its = 0;
    cg(op, sol, rhs, EXADG_B200_PRECOND_POINT_JACOBI, ch->inv_diag, nullptr, 1.4901161193847656e-08, 1e-2, eig_cg_n_iterations, &its, nullptr, &alphas, &betas);
    if (!alphas.empty()) {
      std::vector<double> a(alphas.size()), b(alphas.size() > 1 ? alphas.size() - 1 : 0);
      for (size_t i = 0; i < alphas.size(); ++i) {
        a[i] = 1.0 / alphas[i] + (i > 0 ? betas[i - 1] / alphas[i - 1] : 0.0);
        if (i + 1 < alphas.size()) b[i] = std::sqrt(betas[i]) / alphas[i];
      }
      tridiag_extreme_eigs(a, b, ch->lambda_min_est, ch->lambda_max_est);
    } else { ch->lambda_min_est = ch->lambda_max_est = 1.0; }
    const double max_ev = 1.2 * ch->lambda_max_est;
    const double alpha = smoothing_range > 1.0 ? max_ev / smoothing_range : std::min(0.9 * max_ev, ch->lambda_min_est);
    ch->delta = 0.5 * (max_ev - alpha); ch->theta = 0.5 * (max_ev + alpha);
    *out = ch.release();
    return EXADG_B200_OK;
  });
}
int exadg_b200_chebyshev_destroy(exadg_b200_chebyshev * ch)
{
  if (!ch) return EXADG_B200_OK;
  cudaFree(ch->inv_diag); cudaFree(ch->xold); cudaFree(ch->r);
  delete ch;
  return EXADG_B200_OK;
}
int exadg_b200_chebyshev_get(const exadg_b200_chebyshev * ch, double * lmin, double * lmax, double * theta, double * delta)
{
  if (!ch) return EXADG_B200_ERR_ARG;
  if (lmin) *lmin = ch->lambda_min_est; if (lmax) *lmax = ch->lambda_max_est; if (theta) *theta = ch->theta; if (delta) *delta = ch->delta;
  return EXADG_B200_OK;
}
int exadg_b200_chebyshev_set_interval(exadg_b200_chebyshev * ch, double theta, double delta)
{ if (!ch) return EXADG_B200_ERR_ARG; ch->theta = theta; ch->delta = delta; return EXADG_B200_OK; }
int exadg_b200_chebyshev_vmult(exadg_b200_chebyshev * ch, double * dst, const double * src)
{ return guarded([&]() { if (!ch) throw std::invalid_argument("null smoother"); check_ptr(dst, "dst"); check_ptr(src, "src"); cheb_run(ch, dst, src, true); return EXADG_B200_OK; }); }
int exadg_b200_chebyshev_step(exadg_b200_chebyshev * ch, double * dst, const double * src)
{ return guarded([&]() { if (!ch) throw std::invalid_argument("null smoother"); check_ptr(dst, "dst"); check_ptr(src, "src"); cheb_run(ch, dst, src, false); return EXADG_B200_OK; }); }

/* ---- multigrid ---- */
int exadg_b200_multigrid_levels(int mg_type, int p_sequence, int degree, int n_h_levels, int max_levels, int * n_levels, int * h_level, int * level_degree)
{
  return guarded([&]() {
    if (!n_levels || degree < 1 || n_h_levels < 1) throw std::invalid_argument("bad argument");
    // MultigridPreconditionerBase::initialize_levels (multigrid_preconditioner_base.cpp:97-323) for the DG-only types
    std::vector<int> p_levels;
    if (mg_type == EXADG_B200_MG_H) p_levels.push_back(degree);
    else {
      int p = degree;
      do {
        p_levels.push_back(p);
        switch (p_sequence) {
          case EXADG_B200_PSEQ_GO_TO_ONE: p = 1; break;
          case EXADG_B200_PSEQ_DECREASE_BY_ONE: p = std::max(p - 1, 1); break;
          case EXADG_B200_PSEQ_BISECT: p = std::max(p / 2, 1); break;
          default: throw std::invalid_argument("No valid p-sequence selected!");
        }
      } while (p != p_levels.back());
      std::reverse(p_levels.begin(), p_levels.end());
    }
    std::vector<std::pair<int, int>> info; // (h_level, degree), coarse -> fine
    const bool use_h = (mg_type == EXADG_B200_MG_H || mg_type == EXADG_B200_MG_HP || mg_type == EXADG_B200_MG_PH);
    const int nh = use_h ? n_h_levels : 1;
    const int h_fine = n_h_levels - 1, h_first = use_h ? 0 : h_fine;
    if (mg_type == EXADG_B200_MG_H) for (int h = 0; h < nh; ++h) info.push_back({h, p_levels.front()});
    else if (mg_type == EXADG_B200_MG_P) for (int p : p_levels) info.push_back({h_fine, p});
    else if (mg_type == EXADG_B200_MG_PH) {
      for (int h = 0; h < nh - 1; ++h) info.push_back({h_first + h, p_levels.front()});
      for (int p : p_levels) info.push_back({h_fine, p});
    } else if (mg_type == EXADG_B200_MG_HP) {
      for (size_t p = 0; p + 1 < p_levels.size(); ++p) info.push_back({h_first, p_levels[p]});
      for (int h = 0; h < nh; ++h) info.push_back({h_first + h, p_levels.back()});
    } else throw std::invalid_argument("This multigrid type is not implemented! (DG-only types: hMG, pMG, hpMG, phMG)");
    *n_levels = (int)info.size();
    if (h_level && level_degree) {
      if ((int)info.size() > max_levels) throw std::invalid_argument("max_levels too small");
      for (size_t l = 0; l < info.size(); ++l) { h_level[l] = info[l].first; level_degree[l] = info[l].second; }
    }
    return EXADG_B200_OK;
  });
}

int exadg_b200_multigrid_destroy(exadg_b200_multigrid * mg)
{
  if (!mg) return EXADG_B200_OK;
  cudaDeviceSynchronize();
  for (auto * c : mg->smoothers) exadg_b200_chebyshev_destroy(c);
  for (auto * v : mg->defect) cudaFree(v);
  for (auto * v : mg->solution) cudaFree(v);
  for (auto * v : mg->t) cudaFree(v);
  cudaFree(mg->coarse_inv_diag); cudaFree(mg->coarse_rhs);
  delete mg;
  return EXADG_B200_OK;
}

int exadg_b200_multigrid_create(int n_levels, exadg_b200_operator * const * ops, int smoother_degree, double smoothing_range, int eig_cg_n_iterations, double coarse_abs_tol,
                                double coarse_rel_tol, int coarse_max_iter, exadg_b200_multigrid ** out)
{
  return guarded([&]() {
    if (n_levels < 1 || !ops || !out) throw std::invalid_argument("bad argument");
    std::unique_ptr<exadg_b200_multigrid, int (*)(exadg_b200_multigrid *)> mg(new exadg_b200_multigrid, exadg_b200_multigrid_destroy);
    mg->coarse_abs_tol = coarse_abs_tol; mg->coarse_rel_tol = coarse_rel_tol; mg->coarse_max_iter = coarse_max_iter;
    exadg_b200_operator * fine = ops[n_levels - 1];
    if (!fine) throw std::invalid_argument("null level operator");
    mg->smoothers.assign(n_levels, nullptr); mg->transfer.resize(n_levels);
    for (int l = 0; l < n_levels; ++l) {
      exadg_b200_operator * op = ops[l];
      if (!op) throw std::invalid_argument("null level operator");
      require_self_contained(op, "exadg_b200_multigrid_create");
      // all levels run on the stream of the finest operator (one V-cycle is one ordered sequence of launches)
      if (op != fine) { CUDA_CHECK(cudaStreamSynchronize(op->stream)); if (op->own_stream && op->stream) { CUDA_CHECK(cudaStreamDestroy(op->stream)); } op->stream = fine->stream; op->own_stream = false; }
      mg->ops.push_back(op);
      const size_t bytes = (size_t)std::max<int64_t>(op->n_local, 1) * sizeof(double);
      double * v[3];
      for (auto & x : v) { CUDA_CHECK(cudaMalloc(&x, bytes)); CUDA_CHECK(cudaMemsetAsync(x, 0, bytes, fine->stream)); }
      mg->defect.push_back(v[0]); mg->solution.push_back(v[1]); mg->t.push_back(v[2]);
      if (l > 0) {
        // only one type of transfer between two consecutive levels (multigrid_preconditioner_base.cpp:309-318)
        exadg_b200_operator * c = ops[l - 1];
        TransferTable & T = mg->transfer[l];
        std::memset(&T, 0, sizeof(T));
        T.nf = op->dev.n; T.nc = c->dev.n;
        if (op->dev.n_components != c->dev.n_components) throw std::invalid_argument("multigrid levels with different numbers of components");
        if (op->dev.n_components > 1 && op->dev.degree == c->dev.degree) throw std::runtime_error("h-transfer of vector-valued fields is not implemented");
        if (op->dev.n_owned == c->dev.n_owned && op->dev.degree > c->dev.degree && op->mesh.global_offset == c->mesh.global_offset) {
          T.h = 0; embedding_1d(op->dev.degree, c->dev.degree, 0, 1, T.I[0]);
        } else if (op->dev.degree == c->dev.degree && op->dev.n_owned == 8 * c->dev.n_owned && op->mesh.global_offset == 8 * c->mesh.global_offset) {
          T.h = 1; embedding_1d(op->dev.degree, op->dev.degree, 0, 0.5L, T.I[0]); embedding_1d(op->dev.degree, op->dev.degree, 0.5L, 0.5L, T.I[1]);
        } else
          throw std::invalid_argument("Between two consecutive multigrid levels, only one type of transfer is allowed: either the degree decreases on the same cells, or "
                                      "every coarse cell c has the children 8 c .. 8 c + 7 on this rank (global refinement, aligned partition)");
        exadg_b200_chebyshev * ch = nullptr;
        const int st = exadg_b200_chebyshev_create(op, smoother_degree, smoothing_range, eig_cg_n_iterations, &ch);
        if (st != EXADG_B200_OK) throw std::runtime_error(g_last_error);
        mg->smoothers[l] = ch;
      }
    }
    exadg_b200_operator * c0 = mg->ops[0];
    const size_t b0 = (size_t)std::max<int64_t>(c0->n_local, 1) * sizeof(double);
    CUDA_CHECK(cudaMalloc(&mg->coarse_inv_diag, b0)); CUDA_CHECK(cudaMalloc(&mg->coarse_rhs, b0));
    diagonal(c0, mg->coarse_inv_diag, false);
    invert_diagonal(mg->coarse_inv_diag, c0->n_local, c0->stream); c0->launches++;
    CUDA_CHECK(cudaStreamSynchronize(fine->stream));
    *out = mg.release();
    return EXADG_B200_OK;
  });
}

int exadg_b200_multigrid_vmult(exadg_b200_multigrid * mg, double * dst, const double * src)
{ return guarded([&]() { check_ptr(dst, "dst"); check_ptr(src, "src"); mg_vmult(mg, dst, src); return EXADG_B200_OK; }); }

int exadg_b200_multigrid_info(const exadg_b200_multigrid * mg, int * n_levels, int64_t * coarse_iterations, int64_t * cycles)
{
  if (!mg) return EXADG_B200_ERR_ARG;
  if (n_levels) *n_levels = (int)mg->ops.size();
  if (coarse_iterations) *coarse_iterations = mg->coarse_iterations;
  if (cycles) *cycles = mg->cycles;
  return EXADG_B200_OK;
}

int exadg_b200_multigrid_smoother(const exadg_b200_multigrid * mg, int level, exadg_b200_chebyshev ** smoother)
{
  if (!mg || !smoother || level < 1 || level >= (int)mg->ops.size()) return EXADG_B200_ERR_ARG;
  *smoother = mg->smoothers[level];
  return EXADG_B200_OK;
}

int exadg_b200_cg_solve_multigrid(exadg_b200_operator * op, double * x, const double * b, exadg_b200_multigrid * mg, double abs_tol, double rel_tol, int max_iter,
                                  int * n_iter, double * residuals)
{
  return guarded([&]() {
    if (!op || !mg) throw std::invalid_argument("null argument");
    require_self_contained(op, "exadg_b200_cg_solve_multigrid");
    check_ptr(x, "x"); check_ptr(b, "b");
    if (mg->ops.back()->n_local != op->n_local) throw std::invalid_argument("the finest multigrid level does not match the operator");
    if (mg->ops.back()->stream != op->stream) throw std::invalid_argument("the finest multigrid level must run on the operator's stream (exadg_b200_set_stream)");
    g_cg_multigrid = mg;
    struct Reset { ~Reset() { g_cg_multigrid = nullptr; } } reset;
    return cg(op, x, b, EXADG_B200_PRECOND_MULTIGRID, nullptr, nullptr, abs_tol, rel_tol, max_iter, n_iter, residuals, nullptr, nullptr);
  });
}

int exadg_b200_set_nccl_comm(exadg_b200_operator * op, void * comm)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    if (!nccl().ok) throw std::runtime_error("libnccl.so.2 could not be loaded");
    op->comm = comm; op->own_comm = false;
    return EXADG_B200_OK;
  });
}
int exadg_b200_nccl_unique_id(char * id128)
{
  return guarded([&]() {
    if (!nccl().ok) throw std::runtime_error("libnccl.so.2 could not be loaded");
    if (nccl().GetUniqueId(id128) != 0) throw std::runtime_error("ncclGetUniqueId failed");
    return EXADG_B200_OK;
  });
}
int exadg_b200_nccl_init(exadg_b200_operator * op, const char * id128)
{
  return guarded([&]() {
    if (!op) throw std::invalid_argument("null operator");
    if (!nccl().ok) throw std::runtime_error("libnccl.so.2 could not be loaded");
    NcclId id; std::memcpy(id.bytes, id128, 128);
    typedef int (*init_fn)(void **, int, NcclId, int);
    init_fn f = (init_fn)dlsym(nccl().lib, "ncclCommInitRank");
    if (!f || f(&op->comm, op->mesh.world, id, op->mesh.rank) != 0) throw std::runtime_error("ncclCommInitRank failed");
    op->own_comm = true;
    return EXADG_B200_OK;
  });
}

// ---- host-only view of the partition / halo plan (no CUDA call): used by the world_size-2 gloo tests ----
struct exadg_b200_plan { HostMesh mesh; };

int exadg_b200_plan_create(const exadg_b200_hypercube_desc * desc, exadg_b200_plan ** out)
{
  return guarded([&]() {
    if (!desc || !out) throw std::invalid_argument("null argument");
    HypercubeDesc hd;
    hd.n_sub = desc->n_subdivisions; hd.refine = desc->n_refinements; hd.mapping_degree = desc->mapping_degree < 1 ? 1 : desc->mapping_degree;
    hd.deformation = desc->deformation; hd.frequency = desc->frequency;
    for (int f = 0; f < 6; ++f) hd.bc[f] = desc->boundary[f];
    hd.rank = desc->rank; hd.world = desc->world < 1 ? 1 : desc->world;
    std::unique_ptr<exadg_b200_plan> p(new exadg_b200_plan);
    p->mesh = make_hypercube(hd);
    *out = p.release();
    return EXADG_B200_OK;
  });
}
int exadg_b200_plan_destroy(exadg_b200_plan * p) { delete p; return EXADG_B200_OK; }
int exadg_b200_plan_sizes(const exadg_b200_plan * p, int64_t * n_owned, int64_t * n_ghost, int64_t * global_offset, int * n_peers)
{
  if (!p) return EXADG_B200_ERR_ARG;
  if (n_owned) *n_owned = p->mesh.n_owned; if (n_ghost) *n_ghost = p->mesh.n_ghost;
  if (global_offset) *global_offset = p->mesh.global_offset; if (n_peers) *n_peers = (int)p->mesh.peers.size();
  return EXADG_B200_OK;
}
int exadg_b200_plan_peer(const exadg_b200_plan * p, int i, int * peer_rank, int64_t * n_send, int64_t * recv_begin, int64_t * recv_count, int32_t * send_cells)
{
  if (!p || i < 0 || i >= (int)p->mesh.peers.size()) return EXADG_B200_ERR_ARG;
  const PeerPlan & pp = p->mesh.peers[i];
  if (peer_rank) *peer_rank = pp.rank; if (n_send) *n_send = (int64_t)pp.send_cells.size();
  if (recv_begin) *recv_begin = pp.recv_begin; if (recv_count) *recv_count = pp.recv_count;
  if (send_cells) std::memcpy(send_cells, pp.send_cells.data(), pp.send_cells.size() * sizeof(int32_t));
  return EXADG_B200_OK;
}
int exadg_b200_plan_tables(const exadg_b200_plan * p, int32_t * neighbors, int64_t * ghost_global_ids)
{
  if (!p) return EXADG_B200_ERR_ARG;
  if (neighbors) std::memcpy(neighbors, p->mesh.nb.data(), p->mesh.nb.size() * sizeof(int32_t));
  if (ghost_global_ids) std::memcpy(ghost_global_ids, p->mesh.ghost_global.data(), p->mesh.ghost_global.size() * sizeof(int64_t));
  return EXADG_B200_OK;
}

// Peer-memory halo: step 1, allocate the shared region and export its IPC handle and this rank's receive offsets
int exadg_b200_p2p_export(exadg_b200_operator * op, char * handle64, int64_t * recv_begin_by_rank)
{
  return guarded([&]() {
    if (!op || !handle64 || !recv_begin_by_rank) throw std::invalid_argument("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const HostMesh & M = op->mesh;
    const size_t n3 = (size_t)op->dev.n * op->dev.n * op->dev.n * op->dev.n_components;
    if (!op->p2p_region) {
      op->p2p_ghost_bytes = ((size_t)std::max<int64_t>(M.n_ghost, 1) * n3 * sizeof(double) + 255) / 256 * 256;
      const size_t bytes = 2 * op->p2p_ghost_bytes + (size_t)M.world * sizeof(long long);
      CUDA_CHECK(cudaMalloc(&op->p2p_region, bytes));
      CUDA_CHECK(cudaMemset(op->p2p_region, 0, bytes));
      CUDA_CHECK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, op->p2p_region));
    std::memcpy(handle64, &h, 64);
    for (int r = 0; r < M.world; ++r) recv_begin_by_rank[r] = -1;
    for (auto & p : M.peers) recv_begin_by_rank[p.rank] = p.recv_begin;
    recv_begin_by_rank[M.world] = (int64_t)op->p2p_ghost_bytes; // entry [world]: size of one ghost buffer of this rank
    return EXADG_B200_OK;
  });
}
// step 2 (after an all-gather of step 1's outputs): map the peers' regions; from now on vmult uses NVLink stores
int exadg_b200_p2p_connect(exadg_b200_operator * op, const char * handles /*[world][64]*/, const int64_t * recv_begin_table /*[world][world+1]*/)
{
  return guarded([&]() {
    if (!op || !handles || !recv_begin_table) throw std::invalid_argument("null argument");
    if (!op->p2p_region) throw std::runtime_error("call exadg_b200_p2p_export first");
    const HostMesh & M = op->mesh;
    op->p2p_peer_regions.assign(M.peers.size(), nullptr);
    op->p2p_peer_recv_begin.assign(M.peers.size(), 0);
    op->p2p_peer_ghost_bytes.assign(M.peers.size(), 0);
    for (size_t i = 0; i < M.peers.size(); ++i) {
      const int r = M.peers[i].rank;
      cudaIpcMemHandle_t h; std::memcpy(&h, handles + (size_t)r * 64, 64);
      CUDA_CHECK(cudaIpcOpenMemHandle(&op->p2p_peer_regions[i], h, cudaIpcMemLazyEnablePeerAccess));
      const int64_t rb = recv_begin_table[(size_t)r * (M.world + 1) + M.rank]; // where rank r stores the cells it receives from us
      op->p2p_peer_ghost_bytes[i] = recv_begin_table[(size_t)r * (M.world + 1) + M.world];
      if (rb < 0) throw std::runtime_error("asymmetric halo plan (peer does not expect our cells)");
      op->p2p_peer_recv_begin[i] = rb;
    }
    op->ghost_alloc = op->dev.ghost;
    // ticket counters of the fused put + signal kernel; the grid is fixed from here on (the counters accumulate over the vmults)
    if (!op->d_put_done) {
      CUDA_CHECK(cudaMalloc(&op->d_work_counter, sizeof(int)));
      CUDA_CHECK(cudaMalloc(&op->d_put_done, MAX_PEERS * sizeof(unsigned long long)));
      CUDA_CHECK(cudaMemset(op->d_put_done, 0, MAX_PEERS * sizeof(unsigned long long)));
      CUDA_CHECK(cudaDeviceSynchronize());
    }
    op->p2p = true;
    return EXADG_B200_OK;
  });
}

int exadg_b200_halo_n_peers(const exadg_b200_operator * op) { return op ? (int)op->mesh.peers.size() : -1; }
int exadg_b200_halo_peer(const exadg_b200_operator * op, int i, int * peer_rank, int64_t * send_cells, int64_t * recv_begin, int64_t * recv_cells)
{
  if (!op || i < 0 || i >= (int)op->mesh.peers.size()) return EXADG_B200_ERR_ARG;
  const PeerPlan & p = op->mesh.peers[i];
  if (peer_rank) *peer_rank = p.rank; if (send_cells) *send_cells = (int64_t)p.send_cells.size();
  if (recv_begin) *recv_begin = p.recv_begin; if (recv_cells) *recv_cells = p.recv_count;
  return EXADG_B200_OK;
}
int exadg_b200_halo_send_list(const exadg_b200_operator * op, int i, int32_t * cells)
{
  if (!op || !cells || i < 0 || i >= (int)op->mesh.peers.size()) return EXADG_B200_ERR_ARG;
  std::memcpy(cells, op->mesh.peers[i].send_cells.data(), op->mesh.peers[i].send_cells.size() * sizeof(int32_t));
  return EXADG_B200_OK;
}
int exadg_b200_ghost_global_ids(const exadg_b200_operator * op, int64_t * ids)
{
  if (!op || !ids) return EXADG_B200_ERR_ARG;
  std::memcpy(ids, op->mesh.ghost_global.data(), op->mesh.ghost_global.size() * sizeof(int64_t));
  return EXADG_B200_OK;
}
int exadg_b200_cartesian_kernel(int variant) { return cartesian_kernel_variant(variant); }

int exadg_b200_fp64_peak(double * dfma_tflops, double * dmma_tflops)
{ return guarded([&]() { fp64_peak(dfma_tflops, dmma_tflops); return EXADG_B200_OK; }); }
double * exadg_b200_ghost_buffer(exadg_b200_operator * op) { return op ? op->dev.ghost : nullptr; }
int exadg_b200_halo_pack(exadg_b200_operator * op, int i, const double * src, double * send_buffer)
{
  return guarded([&]() {
    if (!op || i < 0 || i >= (int)op->mesh.peers.size()) throw std::invalid_argument("bad peer index");
    const int n3 = op->dev.n * op->dev.n * op->dev.n * op->dev.n_components;
    const int64_t nc = (int64_t)op->mesh.peers[i].send_cells.size();
    pack_cells_kernel<<<(unsigned)std::min<int64_t>((nc * n3 + 255) / 256, 148 * 8), 256, 0, op->stream>>>(src, op->d_send_lists[i], nc, n3, send_buffer);
    op->launches++;
    CUDA_CHECK(cudaGetLastError());
    return EXADG_B200_OK;
  });
}

} // extern "C"
