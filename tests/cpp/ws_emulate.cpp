// CPU emulation of the warp-specialised Cartesian kernel (exadg_b200/csrc/cart_ws.hpp): the very same CTA body, compiled by g++
// against a run-time interface made of OS threads, pthread barriers and synchronous copies.  One OS thread per CUDA thread
// (192 or 256 per CTA), CTAs one after the other.  Test infrastructure only (tests/test_ws_emulation.py): it checks indexing,
// the barrier protocol (under -fsanitize=thread every unordered shared-memory access is reported) and the results against the
// CPU oracle, on a machine without a GPU.
//
// Bulk-copy emulation: a load is performed at issue time (the earliest moment the hardware may write) and becomes visible
// through an acquire/release counter that stands in for the mbarrier phase; a store is performed at issue time and its source
// is compared again when the kernel waits for the read (the latest moment the hardware may read) - a source modified in between
// is an error.
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

#include "../../exadg_b200/csrc/cart_ws.hpp"
#include "../../exadg_b200/csrc/mesh.hpp"

using namespace exadg_b200;
using namespace exadg_b200::ws;

namespace
{
constexpr int N = 5;
#ifndef WSE_R
#define WSE_R 8
#endif
#ifndef WSE_NP
#define WSE_NP 2
#endif
constexpr int NP = WSE_NP;

// named barrier where some threads only arrive (bar.arrive) and the others wait (bar.sync): `total` participants per generation
struct NamedBar
{
  std::mutex m; std::condition_variable cv; int count = 0, gen = 0, total = 0;
  void arrive() { std::unique_lock<std::mutex> l(m); if (++count == total) { count = 0; ++gen; cv.notify_all(); } }
  void sync()
  {
    std::unique_lock<std::mutex> l(m);
    const int g = gen;
    if (++count == total) { count = 0; ++gen; cv.notify_all(); }
    else cv.wait(l, [&] { return gen != g; });
  }
};

struct HostCta
{
  double * smem = nullptr;
  pthread_barrier_t ba, bc, bp[NP];
  NamedBar na, nb;
  std::atomic<int> loads{0};
  const double * st_src = nullptr; size_t st_bytes = 0; std::vector<char> snap; bool st_pending = false;
  std::atomic<int> errors{0};
  int cta = 0, ncta = 1;
};

struct HostRT
{
  HostCta * c; int t; int waits = 0;
  double * smem() { return c->smem; }
  int tid() const { return t; }
  int cta() const { return c->cta; }
  int ncta() const { return c->ncta; }
  void bar_init(void *) {}
  // staged variant: every lane copies its 16-byte chunks at issue time; the warp barrier behind the wait publishes them
  void stage_cell(double * dst, const double * src, int lane)
  {
    if ((reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0) c->errors++;
    std::memcpy(dst + 2 * lane, src + 2 * lane, 16);
    if (lane < 31) std::memcpy(dst + 64 + 2 * lane, src + 64 + 2 * lane, 16);
  }
  void stage_commit() {}
  void stage_wait_prev() {}
  void stage_wait_all() {}
  void arrive_a() { c->na.arrive(); }
  void wait_a() { c->na.sync(); }
  void arrive_b() { c->nb.arrive(); }
  void wait_b() { c->nb.sync(); }
  void sync_all() { pthread_barrier_wait(&c->ba); }
  void sync_compute() { pthread_barrier_wait(&c->bc); }
  void sync_producer(int pw) { pthread_barrier_wait(&c->bp[pw]); }
  void role_compute() {}
  void role_producer() {}
  void load_issue(void *, double * dst, const double * src, uint32_t bytes)
  {
    if (bytes % 16 != 0) c->errors++;
    std::memcpy(dst, src, bytes);
    c->loads.fetch_add(1, std::memory_order_release);
  }
  void load_wait(void *)
  {
    ++waits;
    while (c->loads.load(std::memory_order_acquire) < waits) sched_yield();
  }
  void flag_wait(const long long * p, long long epoch)
  {
    while (__atomic_load_n(p, __ATOMIC_ACQUIRE) < epoch) sched_yield();
  }
  int claim(int * counter) { return __atomic_fetch_add(counter, 1, __ATOMIC_RELAXED); }
  void prefetch_l2(const double *, int) {}
  void fence_async() {}
  void check_store()
  {
    if (!c->st_pending) return;
    if (std::memcmp(c->snap.data(), c->st_src, c->st_bytes) != 0) c->errors++; // the source changed before the read was awaited
    c->st_pending = false;
  }
  void store_issue(double * g, const double * s, uint32_t bytes, bool add)
  {
    if (c->st_pending || bytes % 16 != 0) c->errors++; // protocol: the previous read is awaited before the next store
    const size_t n = bytes / sizeof(double);
    if (add) for (size_t i = 0; i < n; ++i) g[i] += s[i]; else std::memcpy(g, s, bytes);
    c->snap.assign(reinterpret_cast<const char *>(s), reinterpret_cast<const char *>(s) + bytes);
    c->st_src = s; c->st_bytes = bytes; c->st_pending = true;
  }
  void store_wait_read() { check_store(); }
  void store_wait_all() { check_store(); }
};

static int g_dynamic = 0, g_counter = 0;

struct Emu
{
  HostMesh mesh;
  WsHostPlan plan;
  WsTables<N> T;
  std::vector<int32_t> interior, boundary;
};
} // namespace

extern "C" {

void * wse_create(int n_sub, int refine, int rank, int world, double ip_factor)
{
  HypercubeDesc d;
  d.n_sub = n_sub; d.refine = refine; d.mapping_degree = 1; d.rank = rank; d.world = world;
  for (int f = 0; f < 6; ++f) d.bc[f] = 0;
  d.left = -1.0; d.right = 1.0; d.deformation = 0.0; d.frequency = 2;
  Emu * E = new Emu;
  E->mesh = make_hypercube(d);
  E->plan = ws_build_plan(E->mesh.nb.data(), E->mesh.n_owned, WsCfg<N>::B);
  // as finish_setup (csrc/c_api.cu): tau_K = sum_d 1/h_d on the uniform box, times (k+1)^2 IP_factor
  double tk = 0.0;
  for (int e = 0; e < 3; ++e) tk += 1.0 / E->mesh.h[e];
  E->T = make_ws_tables<N>(E->mesh.h, tk * ip_factor * N * N);
  for (int b = 0; b < E->plan.n_batches; ++b) {
    bool ghost = false;
    for (int e = 0; e < ws_count_total(E->plan.cnt[b]); ++e) ghost |= (E->plan.halo[(size_t)b * E->plan.HL + e].y >= E->mesh.n_owned);
    (ghost ? E->boundary : E->interior).push_back(b);
  }
  return E;
}
void wse_destroy(void * h) { delete static_cast<Emu *>(h); }
int64_t wse_n_owned(void * h) { return static_cast<Emu *>(h)->mesh.n_owned; }
int64_t wse_n_ghost(void * h) { return static_cast<Emu *>(h)->mesh.n_ghost; }
int64_t wse_global_offset(void * h) { return static_cast<Emu *>(h)->mesh.global_offset; }
void wse_ghost_global(void * h, int64_t * out) { Emu * E = static_cast<Emu *>(h); std::copy(E->mesh.ghost_global.begin(), E->mesh.ghost_global.end(), out); }
int wse_halo_max(void * h) { return static_cast<Emu *>(h)->plan.HL; }
constexpr bool STAGED = WsStaged<N, WSE_R, NP>::value;
static size_t emu_smem_bytes(const Emu * E)
{
  return STAGED ? ws_smem_bytes_staged<N, NP, WSE_R>(E->plan.HA + E->plan.HB) : ws_smem_bytes<N, NP>(E->plan.HL);
}
int64_t wse_smem_bytes(void * h) { return (int64_t)emu_smem_bytes(static_cast<Emu *>(h)); }
int wse_n_batches(void * h, int which) { Emu * E = static_cast<Emu *>(h); return which == 0 ? E->plan.n_batches : (which == 1 ? (int)E->interior.size() : (int)E->boundary.size()); }

// 1: the CTAs claim their items from a work counter (the scheduling of the single-launch partitioned vmult)
void wse_set_dynamic(int on) { g_dynamic = on; }

// dst (+)= A src on the batches selected by `which` (0 all, 1 batches without ghost neighbours, 2 batches with), n_ctas persistent CTAs;
// returns the number of protocol errors
int wse_vmult(void * h, const double * src, const double * ghost, double * dst, int add, int n_ctas, int which)
{
  Emu * E = static_cast<Emu *>(h);
  WsArgs A;
  A.halo = E->plan.halo.data(); A.cnt = E->plan.cnt.data(); A.nloc = E->plan.nloc.data(); A.nloc8 = E->plan.nloc8.data();
  A.batches = which == 0 ? nullptr : (which == 1 ? E->interior.data() : E->boundary.data());
  A.n_items = wse_n_batches(h, which);
  A.src = src; A.ghost = ghost; A.dst = dst; A.n_owned = E->mesh.n_owned; A.HL = E->plan.HL; A.add = add;
  A.flags = nullptr; A.epoch = 0; A.first_ghost_item = 0; A.n_peers = 0; A.counter = nullptr; A.HA = 0; A.HT = 0;
  A.l2pf = 1; // exercises the (no-op here) prefetch path: indexing only
  if (STAGED) { A.HA = E->plan.HA; A.HT = E->plan.HA + E->plan.HB; }
  if (g_dynamic) { g_counter = 0; A.counter = &g_counter; }
  if (A.n_items == 0) return 0;
  if (E->plan.HL > WsCfg<N>::HLMAX) return -1; // the library falls back to the pipelined kernel
  n_ctas = std::min(n_ctas, A.n_items);
  int errors = 0;
  for (int cta = 0; cta < n_ctas; ++cta) {
    HostCta C;
    std::vector<double> smem_store(emu_smem_bytes(E) / sizeof(double) + 4, -777.0);
    double * smem_aligned = smem_store.data();
    if (reinterpret_cast<uintptr_t>(smem_aligned) & 15) ++smem_aligned; // 16-byte aligned like dynamic shared memory
    C.smem = smem_aligned; C.cta = cta; C.ncta = n_ctas;
    C.na.total = C.nb.total = WsCfg<N, NP>::NT;
    pthread_barrier_init(&C.ba, nullptr, WsCfg<N, NP>::NT);
    pthread_barrier_init(&C.bc, nullptr, WsCfg<N>::NC);
    for (int p = 0; p < NP; ++p) pthread_barrier_init(&C.bp[p], nullptr, 32);
    std::vector<std::thread> threads;
    for (int t = 0; t < WsCfg<N, NP>::NT; ++t)
      threads.emplace_back([&, t]() {
        HostRT rt{&C, t};
        if (E->mesh.n_ghost > 0 || g_dynamic) ws_cta<N, WSE_R, true, NP>(rt, E->T, A); else ws_cta<N, WSE_R, false, NP>(rt, E->T, A);
      });
    for (auto & th : threads) th.join();
    pthread_barrier_destroy(&C.ba); pthread_barrier_destroy(&C.bc); for (int p = 0; p < NP; ++p) pthread_barrier_destroy(&C.bp[p]);
    errors += C.errors.load() + (C.st_pending ? 1 : 0);
  }
  return errors;
}
}
