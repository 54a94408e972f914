// Warp-specialised Cartesian vmult (n = 5): four compute warps + two producer warps per CTA.
//
// Same operator and the same Kronecker formulation as vmult_cartesian.cu (cell_loop + face_loop of
// I/operators/operator_base.cpp:1349-1397, fluxes I/poisson/spatial_discretization/laplace_operator.h:180-197):
//     y = (M x M x M) sum_d c_d (Minv L_d) u
// What changes is who does what inside the CTA:
//  * warps 0-3 (compute) own the 24-cell batch exactly like the pipelined kernel (one xy-plane of a cell per thread in
//    registers for the x/y sweeps and the final M_x M_y, n z-lines per thread for the z sweep), but they no longer touch
//    global memory: neighbour traces are read branch-free from shared memory through one precomputed index per
//    (cell, face) - a non-negative value is the in-batch neighbour, a negative one the slot of the trace the producer
//    prepared;
//  * warps 4-5 (producers) run one batch ahead: they copy the next batch's index table, fetch the lines of all
//    out-of-batch neighbour cells (a lane per line, R cells = 5 R loads in flight per lane; the halo list is sorted by
//    direction so that the strides are compile-time constants) and reduce them to the end value / end derivative
//    traces in the other half of a double-buffered trace area.  One CTA-wide barrier per batch
//    hands the buffers over; the compute warps synchronise among themselves on a named barrier.
//
// This header is written against a small run-time interface RT (thread ids, barriers, bulk copies), so that the very
// same body is compiled twice: by nvcc with the PTX implementation (vmult_cartesian_ws.cu) and by g++ with an emulation
// on OS threads (tests/cpp/ws_emulate.cpp), which checks indexing, barrier placement (ThreadSanitizer) and results
// against the CPU oracle without a GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "tables.hpp"

#if defined(__CUDACC__)
#define WS_FN __device__ __forceinline__
#define WS_UNROLL _Pragma("unroll")
#else
#define WS_FN inline
#define WS_UNROLL
#endif

namespace exadg_b200
{
namespace ws
{
struct i2 { int x, y; };

template<int N>
struct WsTables
{
  double G[3][N * N];  // c_d Minv (K + own-side face terms)
  double Pf[3][2][N];  // c_d Minv (1/2 sigma_s l'(s) - tau_hat_d e_s)   times the neighbour's end value
  double Qh[3][2][N];  // -c_d 1/2 sigma_s Minv e_s                       times the neighbour's end derivative
  double M[N * N];
  double fd[2][N];     // l_j'(s)
};

template<int N, int NP_ = 2>
struct WsCfg
{
  static constexpr int B = 24;    // cells per batch: 3 octets of the Morton curve, 24 x 5 planes = 120 of 128 compute threads
  static constexpr int NC = 128;  // compute threads (named barrier 1)
  static constexpr int NP = NP_;  // producer warps: 2 by default (5 or 6 warps per CTA cost the same register allocation)
  static constexpr int NT = NC + 32 * NP;
  static constexpr int HLMAX = 64; // halo entries per batch the producers can stage (two per lane)
};

// Staged variant (four producer warps, R = 3): the out-of-batch neighbour cells are not read with strided 8-byte loads into registers but
// copied whole and asynchronously (cp.async, 16 bytes per lane and instruction, 1008 bytes per cell) into a per-warp staging ring in
// shared memory, two rounds of R cells deep, and reduced to traces from there.  (cp.async.bulk was measured first: one bulk copy of 1 KB
// costs its issuing thread ~670 cycles, scripts/tma_small_copy_bench.cu - 64 of them per batch would keep the producers as busy as today.)  The room for the ring comes from a SINGLE trace buffer: the traces of the x and y faces (slots [0, HA)) are
// produced one batch ahead, once the compute warps have left the y phase of the current batch ("A free"), those of the z faces
// (slots [HA, HA + HB)) at the start of their own batch, before its z phase ("B ready").
template<int N, int R, int NP> struct WsStaged { static constexpr bool value = (NP == 4 && R == 3); };
constexpr int WS_STAGE_SLOT = 128;       // doubles per staged cell (1024 bytes: 1000 bytes of the cell + 8 bytes of alignment slack on either side)
constexpr unsigned WS_STAGE_BYTES = 1008; // bytes copied per cell: the cell, extended to 16-byte granularity (63 chunks of 16 bytes)

struct WsArgs
{
  const i2 * halo;          // [n_batches][HL]: (local cell << 3 | face, neighbour cell), out-of-batch faces of the batch, x faces first, then y, z
  const int32_t * cnt;      // [n_batches] numbers of x, y, z entries packed as x | y << 10 | z << 20
  const int32_t * nloc;     // [n_batches][B * 6]: >= 0 in-batch neighbour (local index), < 0: -1 - (entry of the halo list)
  const int64_t * nloc8;    // [n_batches][B]: the six values of a cell packed as signed bytes (face f in bits 8 f .. 8 f + 7)
  const int32_t * batches;  // optional list of batch ids
  const double * src; const double * ghost; double * dst;
  int64_t n_owned; int n_items; int HL; int add;
  // single-launch partitioned vmult (NVLink peer-memory halo, csrc/c_api.cu): items [first_ghost_item, n_items) read ghost cells that
  // the peers store into this rank's ghost buffer during the launch; a producer warp acquires the peers' flags (>= epoch) before
  // its first such item.  flags == nullptr: the ghost buffer is complete before the launch.
  const long long * flags; long long epoch; int first_ghost_item; int n_peers; int peer_rank[16];
  // optional work counter (zeroed before the launch): the CTAs claim their items dynamically instead of striding through them -
  // CTAs that start late (those that export this rank's cells first) then simply process fewer items
  int * counter;
  int HA, HT; // staged variant: first trace slot of the z faces = maximum number of x and y entries of a batch; HT = HA + HB trace slots
  int l2pf;   // producers prefetch the neighbour cells of the batch after next into L2 (measurement switch EXADG_B200_WS_L2PF; 0 = off)
};

// all peers have stored this vmult's ghost cells (every lane acquires every flag: its later loads are ordered behind them)
template<class RT>
WS_FN void ws_acquire_ghosts(RT & rt, const WsArgs & A)
{
  for (int p = 0; p < A.n_peers; ++p) rt.flag_wait(A.flags + A.peer_rank[p], A.epoch);
}

// shared memory of one CTA in bytes (doubles first, then the int tables, then the mbarrier)
template<int N, int NP = 2>
inline size_t ws_smem_bytes(int HL)
{
  constexpr int B = WsCfg<N>::B, N2 = N * N, N3 = N2 * N;
  return ((size_t)2 * B * N3 + (size_t)2 * B * N2 + (size_t)4 * HL * N2) * sizeof(double) + (size_t)2 * B * 6 * sizeof(int)
         + (size_t)NP * WsCfg<N>::HLMAX * sizeof(i2) + 16 + 16 /* item ring */;
}

// staged variant: single trace buffer of HL = HA + HB slots (rounded to 16 bytes), two index tables, two halo lists per producer warp,
// the staging ring and two mbarriers per producer warp
template<int N, int NP, int R>
inline size_t ws_smem_bytes_staged(int HL)
{
  constexpr int B = WsCfg<N>::B, N2 = N * N, N3 = N2 * N;
  const size_t trs = ((size_t)HL * N2 + 1) & ~(size_t)1;
  return ((size_t)2 * B * N3 + (size_t)2 * B * N2 + 2 * trs + (size_t)NP * 2 * R * WS_STAGE_SLOT) * sizeof(double) + (size_t)2 * B * 6 * sizeof(int)
         + (size_t)NP * 2 * WsCfg<N>::HLMAX * sizeof(i2) + 16 + 16 + (size_t)NP * 2 * 8;
}

// ---------------------------------------------------------------------------------------------------------------
// host side: 1-D tables and the batch plan
// ---------------------------------------------------------------------------------------------------------------
template<int N>
inline WsTables<N> make_ws_tables(const double h[3], double tau_op)
{
  Tables1D tab(N - 1);
  WsTables<N> T;
  for (int d = 0; d < 3; ++d) {
    const real_t cd = (real_t)h[(d + 1) % 3] * h[(d + 2) % 3] / h[d];
    const real_t tau_hat = (real_t)tau_op * h[d];
    // 1-D SIPG operator of the cell's own unknowns: K - 1/2 sigma (l' e^T + e l'^T) + tau_hat e e^T at both ends
    std::vector<real_t> L(tab.K);
    for (int s = 0; s < 2; ++s) {
      const real_t sig = s ? 1 : -1; const int end = s ? N - 1 : 0;
      for (int i = 0; i < N; ++i) { L[i * N + end] -= sig * tab.fd[s][i] / 2; L[end * N + i] -= sig * tab.fd[s][i] / 2; }
      L[end * N + end] += tau_hat;
    }
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        real_t v = 0;
        for (int m = 0; m < N; ++m) v += tab.Minv[i * N + m] * L[m * N + j];
        T.G[d][i * N + j] = (double)(cd * v);
      }
    // coupling to the neighbour behind end s: + 1/2 sigma l'(s) v_nb  -  e_s (1/2 sigma g_nb + tau_hat v_nb)
    for (int s = 0; s < 2; ++s) {
      const real_t sig = s ? 1 : -1; const int end = s ? N - 1 : 0;
      for (int i = 0; i < N; ++i) {
        real_t md = 0;
        for (int m = 0; m < N; ++m) md += tab.Minv[i * N + m] * tab.fd[s][m];
        T.Pf[d][s][i] = (double)(cd * (sig / 2 * md - tau_hat * tab.Minv[i * N + end]));
        T.Qh[d][s][i] = (double)(-cd * sig / 2 * tab.Minv[i * N + end]);
      }
    }
  }
  for (int i = 0; i < N * N; ++i) T.M[i] = (double)tab.M[i];
  for (int s = 0; s < 2; ++s) for (int i = 0; i < N; ++i) T.fd[s][i] = (double)tab.fd[s][i];
  return T;
}

struct WsHostPlan
{
  int B = 0, HL = 0, n_batches = 0;
  int HA = 0, HB = 0;              // maxima over the batches of the x + y entries and of the z entries (staged variant)
  std::vector<i2> halo;            // [n_batches][HL]
  std::vector<int32_t> cnt, nloc;  // [n_batches], [n_batches][B * 6]
  std::vector<int64_t> nloc8;      // [n_batches][B], empty if a value does not fit a signed byte
};

// nb: [n_owned][6] local neighbour indices (ghost cells >= n_owned); every face has a neighbour on this path
inline WsHostPlan ws_build_plan(const int32_t * nb, int64_t n_owned, int B)
{
  WsHostPlan P;
  P.B = B;
  P.n_batches = (int)((n_owned + B - 1) / B);
  std::vector<std::vector<i2>> lists(P.n_batches);
  P.nloc.assign((size_t)P.n_batches * B * 6, 0);
  P.cnt.assign(P.n_batches, 0);
  for (int b = 0; b < P.n_batches; ++b) {
    const int64_t b0 = (int64_t)b * B, b1 = std::min<int64_t>(b0 + B, n_owned);
    int c[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) // entries sorted by direction: x faces, y faces, z faces
      for (int64_t cell = b0; cell < b1; ++cell)
        for (int f = 2 * d; f < 2 * d + 2; ++f) {
          const int32_t p = nb[cell * 6 + f];
          int32_t & nl = P.nloc[((size_t)b * B + (size_t)(cell - b0)) * 6 + f];
          if (p >= b0 && p < b1) nl = (int32_t)(p - b0);
          else { nl = -1 - (int32_t)lists[b].size(); lists[b].push_back(i2{(int)(((cell - b0) << 3) | f), p}); ++c[d]; }
        }
    P.cnt[b] = c[0] | (c[1] << 10) | (c[2] << 20);
    P.HL = std::max(P.HL, (int)lists[b].size());
    P.HA = std::max(P.HA, c[0] + c[1]); P.HB = std::max(P.HB, c[2]);
  }
  P.HL = std::max(P.HL, 1);
  P.HA = std::max(P.HA, 1); P.HB = std::max(P.HB, 1);
  if (B <= 127 && P.HL <= 128) {
    P.nloc8.assign((size_t)P.n_batches * B, 0);
    for (size_t i = 0; i < P.nloc8.size(); ++i) {
      uint64_t v = 0;
      for (int f = 0; f < 6; ++f) v |= (uint64_t)(uint8_t)(int8_t)P.nloc[i * 6 + f] << (8 * f);
      P.nloc8[i] = (int64_t)v;
    }
  }
  P.halo.assign((size_t)P.n_batches * P.HL, i2{0, 0});
  for (int b = 0; b < P.n_batches; ++b) std::copy(lists[b].begin(), lists[b].end(), P.halo.begin() + (size_t)b * P.HL);
  return P;
}

// ---------------------------------------------------------------------------------------------------------------
// device side (or its emulation): the body of one CTA
// ---------------------------------------------------------------------------------------------------------------

WS_FN int64_t ws_min(int64_t a, int64_t b) { return a < b ? a : b; }

WS_FN int ws_count_total(int c) { return (c & 1023) + ((c >> 10) & 1023) + ((c >> 20) & 1023); }

// what a producer lane holds of the batch after next: its packed entry counts and two entries of its halo list
struct WsPrefetch { i2 h[2]; int c; };

WS_FN void ws_prefetch(const WsArgs & A, int bt, int lane, WsPrefetch & pre)
{
  pre.c = A.cnt[bt];
  const int n = ws_count_total(pre.c);
  const i2 zero = {0, 0};
  pre.h[0] = lane < n ? A.halo[(size_t)bt * A.HL + lane] : zero;
  pre.h[1] = lane + 32 < n ? A.halo[(size_t)bt * A.HL + lane + 32] : zero;
}

// optional: the owned neighbour cells of the halo list just fetched (the batch after next) are requested into L2, one batch period before the
// producers load them; the entries are dealt to the producer warps in groups of eight
template<int N, int NP, class RT>
WS_FN void ws_l2_prefetch(RT & rt, const WsArgs & A, const WsPrefetch & pre, int pw, int lane)
{
  constexpr int N3 = N * N * N;
  if ((lane >> 3) % NP != pw % NP) return;
  const int n = ws_count_total(pre.c);
  WS_UNROLL
  for (int q = 0; q < 2; ++q)
    if (lane + 32 * q < n && pre.h[q].y < A.n_owned) rt.prefetch_l2(A.src + (size_t)pre.h[q].y * N3, N3 * (int)sizeof(double));
}

// one round: the lines (direction D) of the neighbour cells of entries e0 .. e0 + R - 1 -> registers -> traces.  All loads
// are issued before the first use; both end derivatives are computed and the one facing us is selected (no branches).
template<int N, int R, int D, bool GH>
WS_FN void ws_round(const WsTables<N> & T, const WsArgs & A, const i2 * hl, int e0, const double * own, const double * gho, int ab, bool act, double * TRV, double * TRG)
{
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int sd = (D == 0) ? 1 : (D == 1 ? N : N2); // stride along the line
  double x[R][N];
  int side[R];
  WS_UNROLL
  for (int q = 0; q < R; ++q) {
    const i2 h = hl[e0 + q];
    side[q] = h.x & 1;
    const double * line = ((!GH || h.y < A.n_owned) ? own : gho) + (size_t)h.y * N3;
    WS_UNROLL
    for (int i = 0; i < N; ++i) x[q][i] = line[i * sd];
  }
  WS_UNROLL
  for (int q = 0; q < R; ++q) {
    double g0 = T.fd[0][0] * x[q][0], g1 = T.fd[1][0] * x[q][0];
    WS_UNROLL
    for (int i = 1; i < N; ++i) { g0 = fma(T.fd[0][i], x[q][i], g0); g1 = fma(T.fd[1][i], x[q][i], g1); }
    // our lower face (side 0): the neighbour is entered through its upper end
    const double v = side[q] ? x[q][0] : x[q][N - 1];
    const double g = side[q] ? g0 : g1;
    if (act) { TRV[(e0 + q) * N2 + ab] = v; TRG[(e0 + q) * N2 + ab] = g; }
  }
}

// entries [e_begin, e_end) of direction D: full rounds of R neighbour cells, then single cells, dealt to the producer warps in turn
template<int N, int R, int D, bool GH, int NP>
WS_FN void ws_produce_dir(const WsTables<N> & T, const WsArgs & A, const i2 * hl, int e_begin, int e_end, int & round, int pw, int ab, bool act, double * TRV, double * TRG)
{
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int s1 = (D == 0) ? N : 1, s2 = (D == 2) ? N : N2; // strides across the face
  const double * const own = A.src + (ab % N) * s1 + (ab / N) * s2;
  const double * const gho = GH ? A.ghost - A.n_owned * N3 + (ab % N) * s1 + (ab / N) * s2 : own; // ghost cells are numbered from n_owned
  int e0 = e_begin;
  for (; e0 + R <= e_end; e0 += R, ++round)
    if (round % NP == pw) ws_round<N, R, D, GH>(T, A, hl, e0, own, gho, ab, act, TRV, TRG); // warp-uniform
  for (; e0 < e_end; ++e0, ++round)
    if (round % NP == pw) ws_round<N, 1, D, GH>(T, A, hl, e0, own, gho, ab, act, TRV, TRG);
}

// producer warp pw: index table + traces of the out-of-batch neighbours of batch bt -> the given halves of the buffers.
// pre holds the halo list of bt (fetched during the previous call) and leaves with that of bt_next (if >= 0); hl is the
// warp's private staging area, so a warp-level barrier orders its accesses.
template<int N, int R, bool GH, int NP, class RT>
WS_FN void ws_produce(RT & rt, const WsTables<N> & T, const WsArgs & A, int bt, int bt_next, int pw, int lane, WsPrefetch & pre, double * TRV, double * TRG, int * nlS, i2 * hl)
{
  constexpr int B = WsCfg<N>::B, N2 = N * N;
  constexpr int NL = (B * 6 + 32 * NP - 1) / (32 * NP);
  const int cx = pre.c & 1023, cy = (pre.c >> 10) & 1023, cz = (pre.c >> 20) & 1023;
  hl[lane] = pre.h[0]; hl[lane + 32] = pre.h[1];
  // index table of this batch: loads now, stores after the traces (their latency hides behind the rounds)
  int nl[NL];
  WS_UNROLL
  for (int j = 0; j < NL; ++j) { const int i = pw * 32 + lane + 32 * NP * j; nl[j] = i < B * 6 ? A.nloc[(size_t)bt * (B * 6) + i] : 0; }
  if (bt_next >= 0) { ws_prefetch(A, bt_next, lane, pre); if (A.l2pf) ws_l2_prefetch<N, NP>(rt, A, pre, pw, lane); }
  rt.sync_producer(pw);
  const bool act = lane < N2;
  const int ab = act ? lane : 0; // line within the face; the spare lanes shadow line 0 and store nothing
  int round = 0;
  ws_produce_dir<N, (R > 8 ? 8 : R), 0, GH, NP>(T, A, hl, 0, cx, round, pw, ab, act, TRV, TRG); // 16 x entries, 24 y and 24 z entries on aligned batches
  ws_produce_dir<N, R, 1, GH, NP>(T, A, hl, cx, cx + cy, round, pw, ab, act, TRV, TRG);
  ws_produce_dir<N, R, 2, GH, NP>(T, A, hl, cx + cy, cx + cy + cz, round, pw, ab, act, TRV, TRG);
  WS_UNROLL
  for (int j = 0; j < NL; ++j) { const int i = pw * 32 + lane + 32 * NP * j; if (i < B * 6) nlS[i] = nl[j]; }
  // hl is overwritten by the next call only behind the CTA-wide hand-over barrier
}

// ---- staged variant: producers ----
// per-warp state of the staging ring: the buffer the next round goes to
struct WsStage { int buf; };

// entries [e_begin, e_end) of the halo list hl -> trace slots slot0 + (e - e_begin).  Rounds of up to R entries are dealt to the producer
// warps in turn; a warp keeps two rounds in the ring: the asynchronous copies (cp.async, 16 bytes per lane and instruction, one commit
// group per round) of its next round are in flight while the current one is reduced - no registers are tied up by loads in flight.
template<int N, int R, bool GH, int NP, class RT>
WS_FN void ws_produce_staged(RT & rt, const WsTables<N> & T, const WsArgs & A, int pw, int lane, const i2 * hl, int e_begin, int e_end, int slot0, int & round,
                             double * TRV, double * TRG, double * stage, WsStage & st)
{
  constexpr int N2 = N * N, N3 = N2 * N;
  const bool act = lane < N2;
  const int ab = act ? lane : 0;
  const int nr = (e_end - e_begin + R - 1) / R;
  int j = ((pw - round) % NP + NP) % NP; // the j-th round of this call has the running index round + j
  round += nr;
  if (j >= nr) return;
  const double * const ghost0 = GH ? A.ghost - A.n_owned * N3 : A.src; // ghost cells are numbered from n_owned
  auto issue = [&](int jj, int b) {
    const int e0 = e_begin + jj * R;
    const int cnt = (e_end - e0 < R) ? e_end - e0 : R;
    WS_UNROLL
    for (int q = 0; q < R; ++q)
      if (q < cnt) { // warp-uniform
        const i2 h = hl[e0 + q];
        const double * p = ((!GH || h.y < A.n_owned) ? A.src : ghost0) + (size_t)h.y * N3;
        const int shift = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1); // cells start on odd multiples of 8 bytes every other time
        rt.stage_cell(stage + (size_t)(b * R + q) * WS_STAGE_SLOT, p - shift, lane); // 1008 bytes from the 16-byte aligned address below the cell
      }
    rt.stage_commit();
  };
  auto process = [&](int jj, int b) {
    const int e0 = e_begin + jj * R;
    const int cnt = (e_end - e0 < R) ? e_end - e0 : R;
    WS_UNROLL
    for (int q = 0; q < R; ++q)
      if (q < cnt) { // warp-uniform
        const i2 h = hl[e0 + q];
        const int f = h.x & 7, d = f >> 1, side = f & 1;
        const double * p = ((!GH || h.y < A.n_owned) ? A.src : ghost0) + (size_t)h.y * N3;
        const int shift = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
        const int sd = (d == 0) ? 1 : (d == 1 ? N : N2), s1 = (d == 0) ? N : 1, s2 = (d == 2) ? N : N2;
        const double * line = stage + (size_t)(b * R + q) * WS_STAGE_SLOT + shift + (ab % N) * s1 + (ab / N) * s2;
        double x[N];
        WS_UNROLL
        for (int i = 0; i < N; ++i) x[i] = line[i * sd];
        double g0 = T.fd[0][0] * x[0], g1 = T.fd[1][0] * x[0];
        WS_UNROLL
        for (int i = 1; i < N; ++i) { g0 = fma(T.fd[0][i], x[i], g0); g1 = fma(T.fd[1][i], x[i], g1); }
        // our lower face (side 0): the neighbour is entered through its upper end
        const double v = side ? x[0] : x[N - 1];
        const double g = side ? g0 : g1;
        const int slot = slot0 + (e0 - e_begin) + q;
        if (act) { TRV[slot * N2 + ab] = v; TRG[slot * N2 + ab] = g; }
      }
  };
  int b = st.buf;
  issue(j, b);
  for (; j < nr; j += NP) {
    const bool more = j + NP < nr;
    if (more) { issue(j + NP, b ^ 1); rt.stage_wait_prev(); } else rt.stage_wait_all();
    rt.sync_producer(pw); // the copies of every lane have landed
    process(j, b);
    rt.sync_producer(pw); // the staged cells of this round are read: the buffer may be refilled
    b ^= 1;
  }
  st.buf = b;
}

// x and y part of batch bt: halo list -> this warp's list buffer hb, index table -> nlS (z faces re-addressed to the slots behind HA),
// traces of the x and y entries -> slots [0, cx + cy).  pre holds the list of bt and leaves with that of bt_next; returns the packed counts.
template<int N, int R, bool GH, int NP, class RT>
WS_FN int ws_produce_staged_xy(RT & rt, const WsTables<N> & T, const WsArgs & A, int bt, int bt_next, int pw, int lane, WsPrefetch & pre, int & round, double * TRV, double * TRG,
                               int * nlS, i2 * hl, double * stage, WsStage & st)
{
  constexpr int B = WsCfg<N>::B;
  constexpr int NL = (B * 6 + 32 * NP - 1) / (32 * NP);
  const int c = pre.c;
  const int cx = c & 1023, cy = (c >> 10) & 1023;
  hl[lane] = pre.h[0]; hl[lane + 32] = pre.h[1];
  int nl[NL];
  WS_UNROLL
  for (int j = 0; j < NL; ++j) { const int i = pw * 32 + lane + 32 * NP * j; nl[j] = i < B * 6 ? A.nloc[(size_t)bt * (B * 6) + i] : 0; }
  if (bt_next >= 0) ws_prefetch(A, bt_next, lane, pre);
  rt.sync_producer(pw);
  ws_produce_staged<N, R, GH, NP>(rt, T, A, pw, lane, hl, 0, cx + cy, 0, round, TRV, TRG, stage, st);
  WS_UNROLL
  for (int j = 0; j < NL; ++j) {
    const int i = pw * 32 + lane + 32 * NP * j;
    if (i < B * 6) {
      int v = nl[j];
      if (i % 6 >= 4 && v < 0) v = -1 - (A.HA + (-1 - v) - (cx + cy)); // z faces: slot behind the x / y area
      nlS[i] = v;
    }
  }
  return c;
}

template<int N, int R, bool GH, int NP, class RT>
WS_FN void ws_cta(RT & rt, const WsTables<N> & T, const WsArgs & A)
{
  constexpr int B = WsCfg<N>::B, NC = WsCfg<N>::NC;
  constexpr int N2 = N * N, N3 = N2 * N;
  static_assert((N2 & 1) == 1, "odd n: contiguous cells in shared memory are conflict-free and bulk-copyable");
  static_assert(B * N <= NC, "one thread per cell plane");
  constexpr bool ST = WsStaged<N, R, NP>::value; // bulk-copied neighbour cells, single trace buffer (see WsStaged)
  double * const smem = rt.smem();
  const int trs = ST ? ((A.HT * N2 + 1) & ~1) : A.HL * N2; // doubles per trace array
  constexpr int OFF_T = B * N3;                // partial results, finally the result (bulk-store source)
  constexpr int OFF_GN = 2 * B * N3;           // [2][B][N2] own end derivatives of the current direction
  constexpr int OFF_TR = OFF_GN + 2 * B * N2;  // [2 buffers (staged: 1)][values, derivatives][HL][N2] traces of out-of-batch neighbours
  constexpr int STAGE = ST ? NP * 2 * R * WS_STAGE_SLOT : 0; // staged: [NP][2 rounds][R] cells of 1 KB
  double * const U = smem;                     // [B][N3] src values of the batch (bulk-copy destination)
  double * const Tt = smem + OFF_T;
  double * const GN = smem + OFF_GN;
  double * const stageS = smem + OFF_TR + 2 * trs; // (staged only)
  int * const nl2 = reinterpret_cast<int *>(smem + OFF_TR + (ST ? 2 * trs + STAGE : 4 * trs)); // [2][B * 6]
  i2 * const hlS = reinterpret_cast<i2 *>(nl2 + 2 * B * 6);           // [NP][HLMAX] (staged: [NP][2][HLMAX]) per-warp staging of the halo list
  void * const bar = hlS + NP * (ST ? 2 : 1) * WsCfg<N>::HLMAX;

  const int t = rt.tid();
  const bool producer = t >= NC;
  if (t == 0) rt.bar_init(bar);
  rt.sync_all();
  // item sequence of this CTA: item q(n) of its n-th iteration is cta + n * ncta, or - with a work counter - claimed one by one.
  // Claimed items travel through a ring of four in shared memory: thread 0 claims q(n + 3) during iteration n; the CTA-wide barrier
  // at the end of every iteration orders the write before the reads (iteration n + 1 needs q(n + 1), q(n + 2), q(n + 3)).
  int * const ring = reinterpret_cast<int *>(reinterpret_cast<char *>(bar) + 16);
  const bool dynamic = GH && A.counter != nullptr; // unpartitioned launches (GH = false) keep the static sequence at compile time
  if (dynamic) {
    if (t == 0) { ring[0] = rt.claim(A.counter); ring[1] = rt.claim(A.counter); ring[2] = rt.claim(A.counter); }
    rt.sync_all();
  }
  const int cta0 = rt.cta(), step = rt.ncta();
  auto q = [&](int n) { return dynamic ? ring[n & 3] : cta0 + n * step; };
  const int first = q(0);
  if (first >= A.n_items) return;
  const int lc = t / N, s = t % N;    // plane layout
  const int lz = t % B, sz = t / B;   // z-line layout: consecutive lanes = consecutive cells (stride n^3, odd -> conflict-free)

  // ---- prologue: traces + index table of the first batch, its bulk copy ----
  const int pw = producer ? (t - NC) / 32 : 0, lane = t % 32;
  WsPrefetch pre;
  pre.c = 0; pre.h[0] = i2{0, 0}; pre.h[1] = i2{0, 0};
  bool ghosts_acquired = false;
  // staged variant: state of this producer warp
  WsStage stg; stg.buf = 0;
  int st_round = 0, st_cnt[2] = {0, 0};
  {
    const int bt = A.batches ? A.batches[first] : first;
    if (producer && ST) {
      const int it1 = q(1);
      if (GH && A.flags && first >= A.first_ghost_item) { ws_acquire_ghosts(rt, A); ghosts_acquired = true; }
      ws_prefetch(A, bt, lane, pre);
      const int btn = it1 < A.n_items ? (A.batches ? A.batches[it1] : it1) : -1;
      double * const stw = stageS + (size_t)pw * 2 * R * WS_STAGE_SLOT;
      i2 * const hlw = hlS + (size_t)(pw * 2 + 0) * WsCfg<N>::HLMAX;
      if (GH && (!A.flags || first >= A.first_ghost_item))
        st_cnt[0] = ws_produce_staged_xy<N, R, GH, NP>(rt, T, A, bt, btn, pw, lane, pre, st_round, smem + OFF_TR, smem + OFF_TR + trs, nl2, hlw, stw, stg);
      else
        st_cnt[0] = ws_produce_staged_xy<N, R, false, NP>(rt, T, A, bt, btn, pw, lane, pre, st_round, smem + OFF_TR, smem + OFF_TR + trs, nl2, hlw, stw, stg);
    } else if (producer) {
      const int it1 = q(1);
      if (GH && A.flags && first >= A.first_ghost_item) { ws_acquire_ghosts(rt, A); ghosts_acquired = true; }
      ws_prefetch(A, bt, lane, pre);
      // with the in-kernel hand-over only the items from first_ghost_item on can have neighbours in the ghost buffer: the others run
      // the very producer code of an unpartitioned mesh (no pointer select per load)
      if (GH && (!A.flags || first >= A.first_ghost_item))
        ws_produce<N, R, GH, NP>(rt, T, A, bt, it1 < A.n_items ? (A.batches ? A.batches[it1] : it1) : -1, pw, lane, pre, smem + OFF_TR, smem + OFF_TR + trs, nl2,
                                 hlS + pw * WsCfg<N>::HLMAX);
      else
        ws_produce<N, R, false, NP>(rt, T, A, bt, it1 < A.n_items ? (A.batches ? A.batches[it1] : it1) : -1, pw, lane, pre, smem + OFF_TR, smem + OFF_TR + trs, nl2,
                                    hlS + pw * WsCfg<N>::HLMAX);
    } else if (t == 0) {
      const int64_t c0 = (int64_t)bt * B;
      const uint32_t by = (uint32_t)((int)ws_min(B, A.n_owned - c0) * N3 * sizeof(double));
      if (by % 16 == 0) rt.load_issue(bar, U, A.src + c0 * N3, by);
    }
  }
  rt.sync_all();

  // The two roles run their own loops over the same batch sequence and meet at the CTA-wide barrier once per batch.
  if (producer && ST) {
    rt.role_producer();
    double * const stw = stageS + (size_t)pw * 2 * R * WS_STAGE_SLOT;
    double * const TRV = smem + OFF_TR, * const TRG = smem + OFF_TR + trs;
    int buf = 0;
    for (int n = 0; q(n) < A.n_items; ++n, buf ^= 1) {
      const int itc = q(n), itn = q(n + 1), itnn = q(n + 2);
      // z traces of the current batch (its list is in this warp's list buffer `buf`), then "B ready"
      {
        const int c = st_cnt[buf];
        const int cxy = (c & 1023) + ((c >> 10) & 1023), cz = (c >> 20) & 1023;
        const i2 * const hlw = hlS + (size_t)(pw * 2 + buf) * WsCfg<N>::HLMAX;
        if (GH && (!A.flags || itc >= A.first_ghost_item)) ws_produce_staged<N, R, GH, NP>(rt, T, A, pw, lane, hlw, cxy, cxy + cz, A.HA, st_round, TRV, TRG, stw, stg);
        else ws_produce_staged<N, R, false, NP>(rt, T, A, pw, lane, hlw, cxy, cxy + cz, A.HA, st_round, TRV, TRG, stw, stg);
      }
      rt.arrive_b();
      rt.wait_a(); // the compute warps have left the y phase of the current batch: the x / y slots are free
      if (itn < A.n_items) {
        const int bn = A.batches ? A.batches[itn] : itn;
        const int bnn = itnn < A.n_items ? (A.batches ? A.batches[itnn] : itnn) : -1;
        if (GH && A.flags && !ghosts_acquired && itn >= A.first_ghost_item) { ws_acquire_ghosts(rt, A); ghosts_acquired = true; }
        i2 * const hlw = hlS + (size_t)(pw * 2 + (buf ^ 1)) * WsCfg<N>::HLMAX;
        if (GH && (!A.flags || itn >= A.first_ghost_item))
          st_cnt[buf ^ 1] = ws_produce_staged_xy<N, R, GH, NP>(rt, T, A, bn, bnn, pw, lane, pre, st_round, TRV, TRG, nl2 + (buf ^ 1) * B * 6, hlw, stw, stg);
        else
          st_cnt[buf ^ 1] = ws_produce_staged_xy<N, R, false, NP>(rt, T, A, bn, bnn, pw, lane, pre, st_round, TRV, TRG, nl2 + (buf ^ 1) * B * 6, hlw, stw, stg);
      }
      rt.sync_all(); // hand-over: x / y traces and index table of the next batch are complete
    }
    return;
  }
  if (producer) {
    rt.role_producer(); // register re-allocation between the roles where the run-time interface implements it
    int buf = 0;
    for (int n = 0; q(n) < A.n_items; ++n, buf ^= 1) {
      const int itn = q(n + 1), itnn = q(n + 2);
      if (itn < A.n_items) {
        const int bn = A.batches ? A.batches[itn] : itn;
        if (GH && A.flags && !ghosts_acquired && itn >= A.first_ghost_item) { ws_acquire_ghosts(rt, A); ghosts_acquired = true; }
        if (GH && (!A.flags || itn >= A.first_ghost_item))
          ws_produce<N, R, GH, NP>(rt, T, A, bn, itnn < A.n_items ? (A.batches ? A.batches[itnn] : itnn) : -1, pw, lane, pre, smem + OFF_TR + (buf ^ 1) * 2 * trs,
                             smem + OFF_TR + (buf ^ 1) * 2 * trs + trs, nl2 + (buf ^ 1) * B * 6, hlS + pw * WsCfg<N>::HLMAX);
        else
          ws_produce<N, R, false, NP>(rt, T, A, bn, itnn < A.n_items ? (A.batches ? A.batches[itnn] : itnn) : -1, pw, lane, pre, smem + OFF_TR + (buf ^ 1) * 2 * trs,
                             smem + OFF_TR + (buf ^ 1) * 2 * trs + trs, nl2 + (buf ^ 1) * B * 6, hlS + pw * WsCfg<N>::HLMAX);
      }
      rt.sync_all(); // hand-over: traces / index table of the next batch are complete, those of this batch are free
    }
    return;
  }
  rt.role_compute();
  int buf = 0;
  for (int n = 0; q(n) < A.n_items; ++n, buf ^= 1) {
    const int it = q(n), itn = q(n + 1);
    const bool has_next = itn < A.n_items;
    if (dynamic && t == 0) ring[(n + 3) & 3] = rt.claim(A.counter); // slot of q(n - 1): read for the last time before the previous barrier
    {
      const int batch = A.batches ? A.batches[it] : it;
      const int64_t b0 = (int64_t)batch * B;
      const int nvalid = (int)ws_min(B, A.n_owned - b0);
      const uint32_t bytes = (uint32_t)(nvalid * N3 * sizeof(double));
      const bool use_tma = (bytes % 16 == 0);
      const bool valid = lc < nvalid;               // also false for t >= B * N
      const bool validz = (sz < N) && (lz < nvalid);
      const int * nlS = nl2 + buf * B * 6;
      const int off_trv = OFF_TR + (ST ? 0 : buf * 2 * trs), off_trg = off_trv + trs;

      int nlp[4] = {0, 0, 0, 0};
      if (valid) {
        WS_UNROLL
        for (int f = 0; f < 4; ++f) nlp[f] = nlS[lc * 6 + f];
      }
      if (use_tma) rt.load_wait(bar);
      else {
        for (int i = t; i < nvalid * N3; i += NC) U[i] = A.src[b0 * N3 + i]; // ragged last batch
        rt.sync_compute();
      }

      double u[N][N], acc[N][N];
      if (valid) {
        WS_UNROLL
        for (int j = 0; j < N; ++j)
          WS_UNROLL
          for (int i = 0; i < N; ++i) { u[j][i] = U[lc * N3 + s * N2 + i + N * j]; acc[j][i] = 0.0; }
      }
      // ---- x and y sweeps on the register plane z = s ----
      WS_UNROLL
      for (int d = 0; d < 2; ++d) {
        if (valid) {
          double g0[N], g1[N];
          WS_UNROLL
          for (int l = 0; l < N; ++l) { const double x = (d == 0) ? u[l][0] : u[0][l]; g0[l] = T.fd[0][0] * x; g1[l] = T.fd[1][0] * x; }
          WS_UNROLL
          for (int m = 1; m < N; ++m)
            WS_UNROLL
            for (int l = 0; l < N; ++l) {
              const double x = (d == 0) ? u[l][m] : u[m][l];
              g0[l] = fma(T.fd[0][m], x, g0[l]); g1[l] = fma(T.fd[1][m], x, g1[l]);
            }
          WS_UNROLL
          for (int l = 0; l < N; ++l) { GN[(0 * B + lc) * N2 + s * N + l] = g0[l]; GN[(1 * B + lc) * N2 + s * N + l] = g1[l]; }
        }
        rt.sync_compute();
        if (valid) {
          WS_UNROLL
          for (int side = 0; side < 2; ++side) {
            const int nl = nlp[2 * d + side];
            const bool inb = nl >= 0;
            const int e = -1 - nl;
            const int endn = side ? 0 : N - 1; // the neighbour's end node facing us
            const int voff = inb ? nl * N3 + s * N2 + (d == 0 ? endn : N * endn) : off_trv + e * N2 + N * s;
            const int vstr = (inb && d == 0) ? N : 1;
            const int goff = inb ? OFF_GN + ((side ^ 1) * B + nl) * N2 + s * N : off_trg + e * N2 + N * s;
            double vn[N], gn[N];
            WS_UNROLL
            for (int l = 0; l < N; ++l) { vn[l] = smem[voff + l * vstr]; gn[l] = smem[goff + l]; }
            WS_UNROLL
            for (int m = 0; m < N; ++m)
              WS_UNROLL
              for (int l = 0; l < N; ++l) {
                if (d == 0) acc[l][m] = fma(T.Pf[d][side][m], vn[l], acc[l][m]); else acc[m][l] = fma(T.Pf[d][side][m], vn[l], acc[m][l]);
              }
            WS_UNROLL
            for (int m = 0; m < N; ++m)
              WS_UNROLL
              for (int l = 0; l < N; ++l) {
                if (d == 0) acc[l][m] = fma(T.Qh[d][side][m], gn[l], acc[l][m]); else acc[m][l] = fma(T.Qh[d][side][m], gn[l], acc[m][l]);
              }
          }
          WS_UNROLL
          for (int c = 0; c < N; ++c)
            WS_UNROLL
            for (int l = 0; l < N; ++l)
              WS_UNROLL
              for (int r = 0; r < N; ++r) {
                if (d == 0) acc[l][r] = fma(T.G[d][r * N + c], u[l][c], acc[l][r]);
                else acc[r][l] = fma(T.G[d][r * N + c], u[c][l], acc[r][l]);
              }
        }
        if (d == 1 && t == 0) rt.store_wait_read(); // the previous batch's bulk store has read Tt
        rt.sync_compute();                          // GN of this direction is consumed
      }
      if (ST) rt.arrive_a(); // this thread has read its last x / y trace of the batch
      if (valid) {
        WS_UNROLL
        for (int j = 0; j < N; ++j)
          WS_UNROLL
          for (int i = 0; i < N; ++i) Tt[lc * N3 + s * N2 + i + N * j] = acc[j][i];
      }
      // ---- z sweep: thread (cell lz, slice sz) owns the lines (i, j = sz) ----
      int nlz[2] = {0, 0};
      if (validz) {
        nlz[0] = nlS[lz * 6 + 4]; nlz[1] = nlS[lz * 6 + 5];
        WS_UNROLL
        for (int i = 0; i < N; ++i)
          WS_UNROLL
          for (int k = 0; k < N; ++k) u[i][k] = U[lz * N3 + k * N2 + i + N * sz];
        double g0[N], g1[N];
        WS_UNROLL
        for (int i = 0; i < N; ++i) { g0[i] = T.fd[0][0] * u[i][0]; g1[i] = T.fd[1][0] * u[i][0]; }
        WS_UNROLL
        for (int k = 1; k < N; ++k)
          WS_UNROLL
          for (int i = 0; i < N; ++i) { g0[i] = fma(T.fd[0][k], u[i][k], g0[i]); g1[i] = fma(T.fd[1][k], u[i][k], g1[i]); }
        WS_UNROLL
        for (int i = 0; i < N; ++i) { GN[(0 * B + lz) * N2 + sz * N + i] = g0[i]; GN[(1 * B + lz) * N2 + sz * N + i] = g1[i]; }
      }
      rt.sync_compute(); // Tt planes and z end derivatives visible
      if (ST) rt.wait_b();  // the z traces of this batch are complete
      if (validz) {
        WS_UNROLL
        for (int i = 0; i < N; ++i)
          WS_UNROLL
          for (int k = 0; k < N; ++k) acc[i][k] = Tt[lz * N3 + k * N2 + i + N * sz];
        WS_UNROLL
        for (int side = 0; side < 2; ++side) {
          const int nl = nlz[side];
          const bool inb = nl >= 0;
          const int e = -1 - nl;
          const int endn = side ? 0 : N - 1;
          const int voff = inb ? nl * N3 + endn * N2 + N * sz : off_trv + e * N2 + N * sz;
          const int goff = inb ? OFF_GN + ((side ^ 1) * B + nl) * N2 + sz * N : off_trg + e * N2 + N * sz;
          double vn[N], gn[N];
          WS_UNROLL
          for (int i = 0; i < N; ++i) { vn[i] = smem[voff + i]; gn[i] = smem[goff + i]; }
          WS_UNROLL
          for (int k = 0; k < N; ++k)
            WS_UNROLL
            for (int i = 0; i < N; ++i) acc[i][k] = fma(T.Pf[2][side][k], vn[i], acc[i][k]);
          WS_UNROLL
          for (int k = 0; k < N; ++k)
            WS_UNROLL
            for (int i = 0; i < N; ++i) acc[i][k] = fma(T.Qh[2][side][k], gn[i], acc[i][k]);
        }
        WS_UNROLL
        for (int c = 0; c < N; ++c)
          WS_UNROLL
          for (int i = 0; i < N; ++i)
            WS_UNROLL
            for (int r = 0; r < N; ++r) acc[i][r] = fma(T.G[2][r * N + c], u[i][c], acc[i][r]);
        // mass matrix along z: u <- M acc
        WS_UNROLL
        for (int i = 0; i < N; ++i)
          WS_UNROLL
          for (int r = 0; r < N; ++r) u[i][r] = T.M[r * N] * acc[i][0];
        WS_UNROLL
        for (int c = 1; c < N; ++c)
          WS_UNROLL
          for (int i = 0; i < N; ++i)
            WS_UNROLL
            for (int r = 0; r < N; ++r) u[i][r] = fma(T.M[r * N + c], acc[i][c], u[i][r]);
        WS_UNROLL
        for (int i = 0; i < N; ++i)
          WS_UNROLL
          for (int r = 0; r < N; ++r) Tt[lz * N3 + r * N2 + i + N * sz] = u[i][r];
      }
      rt.sync_compute(); // U is dead from here on
      if (has_next && t == 0) {
        const int64_t c0 = (int64_t)(A.batches ? A.batches[itn] : itn) * B;
        const uint32_t by = (uint32_t)((int)ws_min(B, A.n_owned - c0) * N3 * sizeof(double));
        if (by % 16 == 0) { rt.fence_async(); rt.load_issue(bar, U, A.src + c0 * N3, by); }
      }
      // ---- mass matrices along x and y on the register plane, in place in Tt ----
      if (valid) {
        WS_UNROLL
        for (int j = 0; j < N; ++j)
          WS_UNROLL
          for (int i = 0; i < N; ++i) u[j][i] = Tt[lc * N3 + s * N2 + i + N * j];
        WS_UNROLL
        for (int j = 0; j < N; ++j)
          WS_UNROLL
          for (int r = 0; r < N; ++r) acc[j][r] = T.M[r * N] * u[j][0];
        WS_UNROLL
        for (int c = 1; c < N; ++c)
          WS_UNROLL
          for (int j = 0; j < N; ++j)
            WS_UNROLL
            for (int r = 0; r < N; ++r) acc[j][r] = fma(T.M[r * N + c], u[j][c], acc[j][r]);
        WS_UNROLL
        for (int r = 0; r < N; ++r)
          WS_UNROLL
          for (int i = 0; i < N; ++i) u[r][i] = T.M[r * N] * acc[0][i];
        WS_UNROLL
        for (int c = 1; c < N; ++c)
          WS_UNROLL
          for (int r = 0; r < N; ++r)
            WS_UNROLL
            for (int i = 0; i < N; ++i) u[r][i] = fma(T.M[r * N + c], acc[c][i], u[r][i]);
        WS_UNROLL
        for (int r = 0; r < N; ++r)
          WS_UNROLL
          for (int i = 0; i < N; ++i) Tt[lc * N3 + s * N2 + i + N * r] = u[r][i];
      }
      if (use_tma) {
        rt.fence_async();
        rt.sync_compute(); // result complete
        if (t == 0) rt.store_issue(A.dst + b0 * N3, Tt, bytes, A.add != 0); // read of Tt awaited before the next batch writes it
      } else {
        rt.sync_compute();
        for (int i = t; i < nvalid * N3; i += NC) { if (A.add) A.dst[b0 * N3 + i] += Tt[i]; else A.dst[b0 * N3 + i] = Tt[i]; }
      }
    }
    rt.sync_all(); // hand-over: traces / index table of the next batch are complete, those of this batch are free
  }
  if (t == 0) rt.store_wait_all(); // all bulk stores complete before the CTA exits
}

} // namespace ws
} // namespace exadg_b200
