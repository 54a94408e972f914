// CPU emulation of the warp-private Cartesian kernel (exadg_b200/csrc/cart_wp.hpp): the very same CTA body, compiled by g++
// against a run-time interface made of OS threads, pthread barriers (CTA-wide and per warp), an mbarrier emulation (arrival
// counter + phase counter with acquire/release ordering) and synchronous bulk copies.  One OS thread per CUDA thread, CTAs one
// after the other.  Test infrastructure only (tests/test_wp_emulation.py): it checks indexing, the barrier protocol (under
// -fsanitize=thread every unordered shared-memory access is reported) and the results against the CPU oracle without a GPU.
//
// Bulk-copy emulation: a load is performed at issue time (the earliest moment the hardware may write) and completes the
// mbarrier phase; a store is performed at issue time and its source is compared again when the issuing thread waits for the
// read (the latest moment the hardware may read) - a source modified in between is an error.  The exported interface is the one
// of ws_emulate.cpp, so that the same tests drive both kernels.
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

#include "../../exadg_b200/csrc/cart_wp.hpp"
#include "../../exadg_b200/csrc/mesh.hpp"

using namespace exadg_b200;
using namespace exadg_b200::ws;
using namespace exadg_b200::wp;

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
using Cfg = WpCfg<N, NP>;
constexpr int NW = Cfg::NT / 32;

struct MBar { std::atomic<int> pending{0}; std::atomic<uint32_t> completed{0}; int count = 0; };

struct HostCta
{
  double * smem = nullptr;
  char * bar_base = nullptr; // address of the first mbarrier inside smem
  MBar bars[4];
  pthread_barrier_t ba, bw[NW];
  std::atomic<int> errors{0};
  int cta = 0, ncta = 1;
  MBar & bar(void * p) { return bars[(reinterpret_cast<char *>(p) - bar_base) / 8]; }
};

struct HostRT
{
  HostCta * c; int t;
  const double * st_src = nullptr; size_t st_bytes = 0; std::vector<char> snap; bool st_pending = false;
  double * smem() { return c->smem; }
  int tid() const { return t; }
  int cta() const { return c->cta; }
  int ncta() const { return c->ncta; }
  void sync_all() { pthread_barrier_wait(&c->ba); }
  void sync_warp() { pthread_barrier_wait(&c->bw[t / 32]); }
  void role_compute() {}
  void role_producer() {}
  void mbar_init(void * b, int count)
  {
    if (!c->bar_base) c->bar_base = reinterpret_cast<char *>(b);
    MBar & m = c->bar(b); m.count = count; m.pending.store(count); m.completed.store(0);
  }
  void mbar_arrive(void * b)
  {
    MBar & m = c->bar(b);
    if (m.pending.fetch_sub(1, std::memory_order_acq_rel) == 1) { m.pending.store(m.count, std::memory_order_relaxed); m.completed.fetch_add(1, std::memory_order_release); }
  }
  void mbar_wait(void * b, uint32_t parity)
  {
    MBar & m = c->bar(b);
    long spins = 0;
    while ((m.completed.load(std::memory_order_acquire) & 1u) == parity) { sched_yield(); if (++spins > 200000000L) { c->errors++; std::fprintf(stderr, "mbarrier wait timed out (thread %d)\n", t); return; } }
  }
  void load_issue(void * b, double * dst, const double * src, uint32_t bytes)
  {
    if (bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0) c->errors++;
    std::memcpy(dst, src, bytes);
    mbar_arrive(b);
  }
  void flag_wait(const long long * p, long long epoch)
  {
    while (__atomic_load_n(p, __ATOMIC_ACQUIRE) < epoch) sched_yield();
  }
  int claim(int * counter) { return __atomic_fetch_add(counter, 1, __ATOMIC_RELAXED); }
  void fence_async() {}
  void check_store()
  {
    if (!st_pending) return;
    if (std::memcmp(snap.data(), st_src, st_bytes) != 0) c->errors++; // the source changed before the read was awaited
    st_pending = false;
  }
  void store_issue(double * g, const double * s, uint32_t bytes, bool add)
  {
    if (st_pending || bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(g) & 15) != 0) c->errors++; // protocol: the previous read is awaited before the next store
    const size_t n = bytes / sizeof(double);
    if (add) for (size_t i = 0; i < n; ++i) g[i] += s[i]; else std::memcpy(g, s, bytes);
    snap.assign(reinterpret_cast<const char *>(s), reinterpret_cast<const char *>(s) + bytes);
    st_src = s; st_bytes = bytes; st_pending = true;
  }
  void store_wait_read() { check_store(); }
  void store_wait_all() { check_store(); }
};

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
  E->plan = ws_build_plan(E->mesh.nb.data(), E->mesh.n_owned, Cfg::B);
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
int64_t wse_smem_bytes(void *) { return (int64_t)wp_smem_bytes<N, NP>(); }
int wse_n_batches(void * h, int which) { Emu * E = static_cast<Emu *>(h); return which == 0 ? E->plan.n_batches : (which == 1 ? (int)E->interior.size() : (int)E->boundary.size()); }

// dst (+)= A src on the batches selected by `which` (0 all, 1 batches without ghost neighbours, 2 batches with), n_ctas persistent CTAs;
// returns the number of protocol errors, -1 if the library would not use this kernel for the mesh
int wse_vmult(void * h, const double * src, const double * ghost, double * dst, int add, int n_ctas, int which)
{
  Emu * E = static_cast<Emu *>(h);
  WsArgs A;
  A.halo = E->plan.halo.data(); A.cnt = E->plan.cnt.data(); A.nloc = E->plan.nloc.data(); A.nloc8 = E->plan.nloc8.data();
  A.batches = which == 0 ? nullptr : (which == 1 ? E->interior.data() : E->boundary.data());
  A.n_items = wse_n_batches(h, which);
  A.src = src; A.ghost = ghost; A.dst = dst; A.n_owned = E->mesh.n_owned; A.HL = E->plan.HL; A.add = add;
  A.flags = nullptr; A.epoch = 0; A.first_ghost_item = 0; A.n_peers = 0; A.counter = nullptr; A.HA = 0; A.HT = 0; A.l2pf = 0;
  if (E->plan.HL > Cfg::HLMAX || E->mesh.n_owned % 2 != 0) return -1; // the library falls back to the pipelined kernel
  if (A.n_items == 0) return 0;
  n_ctas = std::min(n_ctas, A.n_items);
  int errors = 0;
  for (int cta = 0; cta < n_ctas; ++cta) {
    HostCta C;
    std::vector<double> smem(wp_smem_bytes<N, NP>() / sizeof(double) + 2, -777.0);
    C.smem = smem.data(); C.cta = cta; C.ncta = n_ctas;
    pthread_barrier_init(&C.ba, nullptr, Cfg::NT);
    for (int w = 0; w < NW; ++w) pthread_barrier_init(&C.bw[w], nullptr, 32);
    std::vector<std::thread> threads;
    std::vector<HostRT> rts(Cfg::NT);
    for (int t = 0; t < Cfg::NT; ++t) {
      rts[t].c = &C; rts[t].t = t;
      threads.emplace_back([&, t]() {
        if (E->mesh.n_ghost > 0) wp_cta<N, WSE_R, true, NP>(rts[t], E->T, A); else wp_cta<N, WSE_R, false, NP>(rts[t], E->T, A);
      });
    }
    for (auto & th : threads) th.join();
    pthread_barrier_destroy(&C.ba); for (int w = 0; w < NW; ++w) pthread_barrier_destroy(&C.bw[w]);
    errors += C.errors.load();
    for (int t = 0; t < Cfg::NT; ++t) errors += rts[t].st_pending ? 1 : 0;
  }
  return errors;
}
}
