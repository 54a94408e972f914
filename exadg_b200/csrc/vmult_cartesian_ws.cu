// Warp-specialised Cartesian vmult for n = 5 (k = 4): device side of cart_ws.hpp - the PTX run-time interface (named
// barriers, mbarrier + 1-D bulk copies), the kernel wrapper, the device copy of the batch plan and the launch.
// The CTA body itself lives in cart_ws.hpp and is also compiled for the CPU emulation (tests/cpp/ws_emulate.cpp).
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "cart_ws.hpp"
#include "cart_wp.hpp"
#include "operator.cuh"

namespace exadg_b200
{
namespace
{
using namespace ws;

__device__ __forceinline__ uint32_t s32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }

// export of this rank's boundary cells by the operator launch itself (single-launch partitioned vmult): see GhostSync
struct PutDesc
{
  const int32_t * cells[16]; long long n_cells[16]; double * dst[16]; long long * flag[16];
  unsigned long long * done; long long epoch; long long seq; int n_peers; int n_export; // CTAs 0 .. n_export - 1 share the export
};

__device__ __forceinline__ void put_slice(const PutDesc & P, const double * __restrict__ src, int n3)
{
  const long long stride = (long long)P.n_export * blockDim.x;
  for (int p = 0; p < P.n_peers; ++p) {
    const int32_t * __restrict__ cells = P.cells[p];
    double * __restrict__ dst = P.dst[p]; // peer memory
    const long long total = P.n_cells[p] * n3;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
      double v[4]; // four independent loads in flight per thread
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long i = i0 + q * stride;
        if (i < total) { const long long c = i / n3; const int k = (int)(i - c * n3); v[q] = src[(long long)cells[c] * n3 + k]; }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { const long long i = i0 + q * stride; if (i < total) dst[i] = v[q]; }
    }
  }
  __threadfence_system(); // this thread's peer stores are performed before the ticket below
  __syncthreads();
  if (threadIdx.x < P.n_peers) {
    const int p = threadIdx.x;
    const unsigned long long ticket = atomicAdd(P.done + p, 1ull) + 1ull;
    if (ticket == (unsigned long long)P.seq * P.n_export) {
      __threadfence_system(); // the other CTAs' stores (ordered before their tickets) are performed before the flag
      asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(P.flag[p]), "l"(P.epoch) : "memory");
    }
  }
}

// NT: threads per CTA (CTA-wide barrier); REG: re-allocate registers between the roles (4 compute + 4 producer warps launched
// with 128 registers per thread: the compute warpgroup grows to 160, the producer warpgroup shrinks to 96)
template<int NT, bool REG>
struct DeviceRT
{
  double * base;
  uint32_t parity;
  __device__ __forceinline__ double * smem() const { return base; }
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ int cta() const { return (int)blockIdx.x; }
  __device__ __forceinline__ int ncta() const { return (int)gridDim.x; }
  __device__ __forceinline__ void bar_init(void * bar)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // reached from both roles' loops, i.e. from two different barrier instructions: barrier.sync WITHOUT .aligned (bar.sync = barrier.sync.aligned
  // is meant for one instruction executed by all participating threads; compute-sanitizer synccheck reports the aligned form as divergent)
  __device__ __forceinline__ void sync_all() { asm volatile("barrier.sync 2, %0;" ::"n"(NT) : "memory"); }
  __device__ __forceinline__ void role_compute() { if (REG) asm volatile("setmaxnreg.inc.sync.aligned.u32 160;"); }
  __device__ __forceinline__ void role_producer() { if (REG) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;"); }
  __device__ __forceinline__ void sync_compute() { asm volatile("bar.sync 1, %0;" ::"n"(WsCfg<5>::NC) : "memory"); }
  __device__ __forceinline__ void sync_producer(int) { __syncwarp(); }
  // global -> shared bulk copy, completion on the mbarrier (SASS: UBLKCP)
  __device__ __forceinline__ void load_issue(void * bar, double * dst, const double * src, uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
  }
  __device__ __forceinline__ void load_wait(void * bar)
  {
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
    } while (!done);
    parity ^= 1;
  }
  __device__ __forceinline__ int claim(int * counter) { return atomicAdd(counter, 1); }
  __device__ __forceinline__ void flag_wait(const long long * p, long long epoch)
  {
    long long v;
    do { asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); } while (v < epoch);
  }
  // request the 128-byte lines of [p, p + bytes) into L2 (no register, no scoreboard entry)
  __device__ __forceinline__ void prefetch_l2(const double * p, int bytes)
  {
    const char * c = reinterpret_cast<const char *>(p);
#pragma unroll
    for (int i = 0; i < 1024; i += 128)
      if (i < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 8));
  }
  __device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  // shared -> global bulk copy (plain store, or FP64 add-reduction for vmult_add)
  __device__ __forceinline__ void store_issue(double * g, const double * s, uint32_t bytes, bool add)
  {
    if (add) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(g), "r"(s32(s)), "r"(bytes) : "memory");
    else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s32(s)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
  __device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
  // staged variant: a warp copies one neighbour cell (1008 bytes = 63 chunks of 16 bytes) into the staging ring, asynchronously
  __device__ __forceinline__ void stage_cell(double * dst, const double * src, int lane)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + 2 * lane)), "l"(src + 2 * lane) : "memory");
    if (lane < 31) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + 64 + 2 * lane)), "l"(src + 64 + 2 * lane) : "memory");
  }
  __device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void stage_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
  __device__ __forceinline__ void stage_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  // "A free" (compute -> producers) and "B ready" (producers -> compute): named barriers 3 and 4, one side arrives, the other waits
  __device__ __forceinline__ void arrive_a() { asm volatile("bar.arrive 3, %0;" ::"n"(NT) : "memory"); }
  __device__ __forceinline__ void wait_a() { asm volatile("bar.sync 3, %0;" ::"n"(NT) : "memory"); }
  __device__ __forceinline__ void arrive_b() { asm volatile("bar.arrive 4, %0;" ::"n"(NT) : "memory"); }
  __device__ __forceinline__ void wait_b() { asm volatile("bar.sync 4, %0;" ::"n"(NT) : "memory"); }
};

// R: neighbour cells per producer round; GH: the partition has ghost cells (src of the neighbours may live in the ghost buffer);
// NP: producer warps (2, or 4 with register re-allocation between the roles - not measured yet)
template<int N, int R, bool GH, int NP>
__global__ void __launch_bounds__(WsCfg<N, NP>::NT, 2) vmult_cartesian_ws_kernel(const __grid_constant__ WsTables<N> T, const WsArgs A, const __grid_constant__ PutDesc put)
{
  extern __shared__ __align__(128) double ws_shared[];
  if (GH && put.n_peers > 0 && (int)blockIdx.x < put.n_export) put_slice(put, A.src, N * N * N);
  DeviceRT<WsCfg<N, NP>::NT, (NP == 4)> rt{ws_shared, 0u};
  ws_cta<N, R, GH, NP>(rt, T, A);
}

// run-time interface of the warp-private kernel (cart_wp.hpp): mbarriers with one arrival per warp, per-warp bulk stores
template<int NT>
struct DeviceRTwp
{
  double * base;
  __device__ __forceinline__ double * smem() const { return base; }
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ int cta() const { return (int)blockIdx.x; }
  __device__ __forceinline__ int ncta() const { return (int)gridDim.x; }
  __device__ __forceinline__ void sync_all() { asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); }
  __device__ __forceinline__ void sync_warp() { __syncwarp(); }
  __device__ __forceinline__ void role_compute() {}
  __device__ __forceinline__ void role_producer() {}
  __device__ __forceinline__ void mbar_init(void * bar, int count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __device__ __forceinline__ void mbar_arrive(void * bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory"); }
  __device__ __forceinline__ void mbar_wait(void * bar, uint32_t parity)
  {
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
    } while (!done);
  }
  __device__ __forceinline__ void load_issue(void * bar, double * dst, const double * src, uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
  }
  __device__ __forceinline__ int claim(int * counter) { return atomicAdd(counter, 1); }
  __device__ __forceinline__ void flag_wait(const long long * p, long long epoch)
  {
    long long v;
    do { asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); } while (v < epoch);
  }
  __device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __device__ __forceinline__ void store_issue(double * g, const double * s, uint32_t bytes, bool add)
  {
    if (add) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(g), "r"(s32(s)), "r"(bytes) : "memory");
    else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s32(s)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
  __device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
};

// warp-private kernel (cart_wp.hpp): four compute warps that own six cells each + NP producer warps
template<int N, int R, bool GH, int NP>
__global__ void __launch_bounds__(wp::WpCfg<N, NP>::NT, 2) vmult_cartesian_wp_kernel(const __grid_constant__ WsTables<N> T, const WsArgs A, const __grid_constant__ PutDesc put)
{
  extern __shared__ __align__(128) double wp_shared[];
  if (GH && put.n_peers > 0 && (int)blockIdx.x < put.n_export) put_slice(put, A.src, N * N * N);
  DeviceRTwp<wp::WpCfg<N, NP>::NT> rt{wp_shared};
  wp::wp_cta<N, R, GH, NP>(rt, T, A);
}

constexpr size_t WS_MAX_SMEM = 228 * 1024 / 2 - 1024; // half of an SM's shared memory minus the per-CTA reservation

struct WsDevPlan
{
  i2 * d_halo = nullptr; int32_t * d_cnt = nullptr, * d_nloc = nullptr; int64_t * d_nloc8 = nullptr;
  int HL = 0, n_batches = 0, ctas_per_sm = 0;
  int HA = 0, HB = 0;         // staged variant: slots of the x / y traces and of the z traces
  size_t smem = 0, smem4 = 0; // smem4: with 4 producer warps (0 if it does not fit)
  size_t smem_st = 0;         // staged variant (0: not applicable to this mesh)
  size_t smem_wp = 0;         // warp-private kernel (0: not applicable to this mesh)
  WsTables<5> T;
};
} // namespace

bool ws_supported(int n) { return n == 5; }

// device copy of the batch plan; nullptr if the batches are too irregular for the double-buffered trace area
void * ws_plan_create(const DeviceOperator & op, const HostMesh & mesh)
{
  if (!ws_supported(op.n)) return nullptr;
  const WsHostPlan H = ws_build_plan(mesh.nb.data(), mesh.n_owned, WsCfg<5>::B);
  const size_t smem = ws_smem_bytes<5>(H.HL);
  if (smem > WS_MAX_SMEM || H.HL > WsCfg<5>::HLMAX) return nullptr; // two CTAs per SM are what the kernel is built for
  WsDevPlan * P = new WsDevPlan;
  try {
  P->HL = H.HL; P->n_batches = H.n_batches; P->smem = smem;
  P->T = make_ws_tables<5>(op.h, op.tau_hat);
  CUDA_CHECK(cudaMalloc(&P->d_halo, H.halo.size() * sizeof(i2)));
  CUDA_CHECK(cudaMemcpy(P->d_halo, H.halo.data(), H.halo.size() * sizeof(i2), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&P->d_cnt, H.cnt.size() * sizeof(int32_t)));
  CUDA_CHECK(cudaMemcpy(P->d_cnt, H.cnt.data(), H.cnt.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&P->d_nloc, H.nloc.size() * sizeof(int32_t)));
  CUDA_CHECK(cudaMemcpy(P->d_nloc, H.nloc.data(), H.nloc.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (!H.nloc8.empty()) {
    CUDA_CHECK(cudaMalloc(&P->d_nloc8, H.nloc8.size() * sizeof(int64_t)));
    CUDA_CHECK(cudaMemcpy(P->d_nloc8, H.nloc8.data(), H.nloc8.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  }
  P->ctas_per_sm = 2;
  auto configure = [&](auto kernel, int threads, size_t bytes) {
    int occ = 0;
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_MAX_SMEM)); // same for every operator
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, bytes));
    return occ;
  };
  P->ctas_per_sm = std::min(P->ctas_per_sm, configure(vmult_cartesian_ws_kernel<5, 8, false, 2>, WsCfg<5>::NT, smem));
  P->ctas_per_sm = std::min(P->ctas_per_sm, configure(vmult_cartesian_ws_kernel<5, 12, false, 2>, WsCfg<5>::NT, smem));
  P->ctas_per_sm = std::min(P->ctas_per_sm, configure(vmult_cartesian_ws_kernel<5, 8, true, 2>, WsCfg<5>::NT, smem));
  P->ctas_per_sm = std::min(P->ctas_per_sm, configure(vmult_cartesian_ws_kernel<5, 12, true, 2>, WsCfg<5>::NT, smem));
  // experimental variant with 4 producer warps: only if it fits two CTAs per SM as well
  // (a failure here must not take the default kernel down with it)
  try {
    P->smem4 = ws_smem_bytes<5, 4>(H.HL);
    if (P->smem4 > WS_MAX_SMEM || configure(vmult_cartesian_ws_kernel<5, 4, false, 4>, WsCfg<5, 4>::NT, P->smem4) < 2
        || configure(vmult_cartesian_ws_kernel<5, 4, true, 4>, WsCfg<5, 4>::NT, P->smem4) < 2)
      P->smem4 = 0;
  } catch (const std::exception &) { P->smem4 = 0; cudaGetLastError(); }
  // staged variant: whole neighbour cells by bulk copies of 1008 bytes from (cell address rounded down to 16 bytes) - with an even number of
  // owned / ghost cells no copy reaches beyond the end of the vector / ghost buffer
  try {
    P->HA = H.HA; P->HB = H.HB;
    P->smem_st = ws_smem_bytes_staged<5, 4, 3>(H.HA + H.HB);
    if (mesh.n_owned % 2 != 0 || mesh.n_ghost % 2 != 0 || H.HL > WsCfg<5>::HLMAX || P->smem_st > WS_MAX_SMEM
        || configure(vmult_cartesian_ws_kernel<5, 3, false, 4>, WsCfg<5, 4>::NT, P->smem_st) < 2
        || configure(vmult_cartesian_ws_kernel<5, 3, true, 4>, WsCfg<5, 4>::NT, P->smem_st) < 2)
      P->smem_st = 0;
  } catch (const std::exception &) { P->smem_st = 0; cudaGetLastError(); }
  // warp-private kernel: same batch plan; needs an even number of owned cells (every bulk copy a multiple of 16 bytes)
  try {
    P->smem_wp = wp::wp_smem_bytes<5, 2>();
    if (mesh.n_owned % 2 != 0 || !P->d_nloc8 || H.HL > wp::WpCfg<5>::HLMAX || P->smem_wp > WS_MAX_SMEM
        || configure(vmult_cartesian_wp_kernel<5, 8, false, 2>, wp::WpCfg<5, 2>::NT, P->smem_wp) < 2
        || configure(vmult_cartesian_wp_kernel<5, 8, true, 2>, wp::WpCfg<5, 2>::NT, P->smem_wp) < 2
        || configure(vmult_cartesian_wp_kernel<5, 12, false, 2>, wp::WpCfg<5, 2>::NT, P->smem_wp) < 2
        || configure(vmult_cartesian_wp_kernel<5, 12, true, 2>, wp::WpCfg<5, 2>::NT, P->smem_wp) < 2)
      P->smem_wp = 0;
  } catch (const std::exception &) { P->smem_wp = 0; cudaGetLastError(); }
  } catch (...) { ws_plan_destroy(P); throw; }
  if (P->ctas_per_sm < 1) { ws_plan_destroy(P); return nullptr; }
  return P;
}

void ws_plan_destroy(void * p)
{
  WsDevPlan * P = static_cast<WsDevPlan *>(p);
  if (!P) return;
  cudaFree(P->d_halo); cudaFree(P->d_cnt); cudaFree(P->d_nloc); cudaFree(P->d_nloc8);
  delete P;
}

// batches: optional list of batch ids (interior / boundary launches of the multi-GPU path), n_items its length (or all batches)
// depth: neighbour cells per producer round (8 or 12; 4 selects the variant with 4 producer warps); gh: the selected batches may
// have neighbours in the ghost buffer
void ws_launch(const DeviceOperator & op, const void * p, double * dst, const double * src, bool add, const int32_t * batches, int n_items, int n_sm, int depth, bool gh,
               cudaStream_t stream, const GhostSync * gs, int first_ghost_item)
{
  const WsDevPlan * P = static_cast<const WsDevPlan *>(p);
  if (n_items == 0) return;
  WsArgs A;
  A.halo = P->d_halo; A.cnt = P->d_cnt; A.nloc = P->d_nloc; A.nloc8 = P->d_nloc8; A.batches = batches;
  A.src = src; A.ghost = op.ghost; A.dst = dst; A.n_owned = op.n_owned; A.n_items = n_items; A.HL = P->HL; A.add = add ? 1 : 0;
  A.flags = nullptr; A.epoch = 0; A.first_ghost_item = 0; A.n_peers = 0; A.counter = nullptr; A.HA = 0; A.HT = 0;
  for (int i = 0; i < 16; ++i) A.peer_rank[i] = 0;
  static const int l2pf = []() { const char * e = std::getenv("EXADG_B200_WS_L2PF"); return e ? std::atoi(e) : 0; }(); // measurement switch
  A.l2pf = l2pf;
  if (depth == 3 && P->smem_st == 0) depth = 4; // staged variant not applicable to this mesh: the same kernel with strided loads
  if (depth == 3) { A.HA = P->HA; A.HT = P->HA + P->HB; }
  PutDesc put;
  std::memset(&put, 0, sizeof(put));
  if (gs && gs->flags) {
    A.flags = gs->flags; A.epoch = gs->epoch; A.first_ghost_item = first_ghost_item; A.n_peers = gs->n_peers;
    for (int i = 0; i < gs->n_peers && i < 16; ++i) A.peer_rank[i] = gs->peer_rank[i];
    if (gs->done) {
      put.done = gs->done; put.epoch = gs->epoch; put.n_peers = gs->n_peers;
      for (int i = 0; i < gs->n_peers && i < 16; ++i) { put.cells[i] = gs->send_cells[i]; put.n_cells[i] = gs->n_send[i]; put.dst[i] = gs->peer_ghost[i]; put.flag[i] = gs->peer_flag[i]; }
    }
  }
  int grid = std::min(n_items, n_sm * P->ctas_per_sm);
  // in-kernel ghost hand-over: CTAs spin until the PEERS' put kernels have run, and those wait for this rank's previous launch -
  // this rank's own put kernel must therefore always find room next to the persistent CTAs (otherwise two ranks whose operator
  // kernels got all SMs before their put kernels became runnable wait for each other forever): leave four SMs' worth of CTAs out
  if (A.flags && !put.done) grid = std::max(1, std::min(grid, n_sm * P->ctas_per_sm - 4 * P->ctas_per_sm));
  // with the export inside the launch every CTA must be resident at once (a CTA that has not started yet holds back the peers'
  // flags while the resident ones spin on theirs): the grid never exceeds the occupancy computed for this kernel
  if (put.done) { // tickets are counted per launch of a fixed grid
    // with a work counter the first 64 CTAs export and join the batches late (they claim fewer items); otherwise every CTA does
    int * const counter = depth < 100 ? gs->counter : nullptr; // the warp-private kernel strides statically
    static const int n_export_ctas = []() { const char * e = std::getenv("EXADG_B200_EXPORT_CTAS"); const int v = e ? std::atoi(e) : 64; return v > 0 ? v : 64; }();
    put.n_export = counter ? std::min(grid, n_export_ctas) : grid;
    A.counter = counter;
    if (counter) CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    if (*gs->put_grid != grid) { CUDA_CHECK(cudaMemsetAsync(put.done, 0, 16 * sizeof(unsigned long long), stream)); *gs->put_seq = 0; *gs->put_grid = grid; }
    put.seq = ++*gs->put_seq;
  }
  if (depth == 100 && P->smem_wp > 0) { // warp-private kernel
    if (gh) vmult_cartesian_wp_kernel<5, 8, true, 2><<<grid, wp::WpCfg<5, 2>::NT, P->smem_wp, stream>>>(P->T, A, put);
    else vmult_cartesian_wp_kernel<5, 8, false, 2><<<grid, wp::WpCfg<5, 2>::NT, P->smem_wp, stream>>>(P->T, A, put);
  } else if (depth == 101 && P->smem_wp > 0) { // warp-private kernel, 12 neighbour cells per producer round
    if (gh) vmult_cartesian_wp_kernel<5, 12, true, 2><<<grid, wp::WpCfg<5, 2>::NT, P->smem_wp, stream>>>(P->T, A, put);
    else vmult_cartesian_wp_kernel<5, 12, false, 2><<<grid, wp::WpCfg<5, 2>::NT, P->smem_wp, stream>>>(P->T, A, put);
  } else if (depth == 3) { // staged variant (bulk-copied neighbour cells)
    if (gh) vmult_cartesian_ws_kernel<5, 3, true, 4><<<grid, WsCfg<5, 4>::NT, P->smem_st, stream>>>(P->T, A, put);
    else vmult_cartesian_ws_kernel<5, 3, false, 4><<<grid, WsCfg<5, 4>::NT, P->smem_st, stream>>>(P->T, A, put);
  } else if (depth == 4 && P->smem4 > 0) {
    if (gh) vmult_cartesian_ws_kernel<5, 4, true, 4><<<grid, WsCfg<5, 4>::NT, P->smem4, stream>>>(P->T, A, put);
    else vmult_cartesian_ws_kernel<5, 4, false, 4><<<grid, WsCfg<5, 4>::NT, P->smem4, stream>>>(P->T, A, put);
  } else if (depth == 12) {
    if (gh) vmult_cartesian_ws_kernel<5, 12, true, 2><<<grid, WsCfg<5>::NT, P->smem, stream>>>(P->T, A, put);
    else vmult_cartesian_ws_kernel<5, 12, false, 2><<<grid, WsCfg<5>::NT, P->smem, stream>>>(P->T, A, put);
  } else {
    if (gh) vmult_cartesian_ws_kernel<5, 8, true, 2><<<grid, WsCfg<5>::NT, P->smem, stream>>>(P->T, A, put);
    else vmult_cartesian_ws_kernel<5, 8, false, 2><<<grid, WsCfg<5>::NT, P->smem, stream>>>(P->T, A, put);
  }
  CUDA_CHECK(cudaGetLastError());
}

} // namespace exadg_b200
