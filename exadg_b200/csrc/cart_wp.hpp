// Warp-private Cartesian vmult (n = 5, k = 4): the affine fast path restructured around what ncu showed for the warp-specialised
// kernel of cart_ws.hpp (profiles/r02_notes.md): both of its variants saturate the LSU / L1 data pipe (70-73 % of its peak, FP64
// pipe 42 %), a quarter of their shared-memory wavefronts are bank-conflict replays, and 1.3-1.7 of ~9 warp-cycles per issue wait
// at the seven named barriers a batch needs.
//
// Same operator and the same Kronecker formulation (cell_loop + face_loop of I/operators/operator_base.cpp:1349-1397, fluxes
// I/poisson/spatial_discretization/laplace_operator.h:180-197):
//     y = (M x M x M) sum_d c_d (Minv L_d) u
// What changes:
//  * a compute warp OWNS six cells of the 24-cell batch (30 of its 32 lanes: one xy-plane of a cell per lane in the plane layout,
//    one xz-plane per lane in the z layout).  The two transpositions plane <-> z layout therefore stay inside the warp
//    (__syncwarp instead of CTA-wide named barriers) and go through a skewed layout
//        addr(c, i, j, k) = j + 5 c + 33 k + 162 i          (c = cell of the warp, 810 doubles per warp)
//    whose lane-dependent part is the lane id itself in both layouts: no bank conflicts in either direction;
//  * everything other warps need from a cell is published once per batch, right after the batch has landed: both end derivatives
//    of every line in all three directions (GN) and the end values in z (VNz).  One split barrier (A) orders these writes before
//    the face terms; the traces of out-of-batch neighbours (TR, written by the producer warps) are handed over by the same barrier;
//  * barrier C (after the x/y face terms) releases the batch buffer U, so that the bulk copy of the NEXT batch is in flight during
//    the z sweep and the mass matrices; barrier B (after the z face terms) releases GN / VNz / TR;
//  * a warp stores its six result cells with its own bulk copy as soon as they are complete.
//  All barriers are mbarriers with one arrival per warp (elected lane after __syncwarp).
//
// Like cart_ws.hpp this header is written against a small run-time interface RT and compiled twice: by nvcc with the PTX
// implementation (vmult_cartesian_wp.cu) and by g++ on OS threads (tests/cpp/wp_emulate.cpp), which checks indexing, the barrier
// protocol (ThreadSanitizer) and the results against the CPU oracle without a GPU.
#pragma once
#include "cart_ws.hpp"

namespace exadg_b200
{
namespace wp
{
using ws::i2;
using ws::WsArgs;
using ws::WsTables;

template<int N, int NP_ = 2>
struct WpCfg
{
  static constexpr int B = 24;     // cells per batch (the batch plan of cart_ws.hpp is reused as is)
  static constexpr int CW = 6;     // cells per compute warp: 6 x 5 planes = 30 lanes
  static constexpr int NCW = B / CW; // compute warps
  static constexpr int NP = NP_;   // producer warps
  static constexpr int NT = 32 * (NCW + NP);
  static constexpr int HLMAX = 64; // out-of-batch faces per batch the producers can stage (two per lane)
  static constexpr int SK = 33, SI = 162, WT = 5 * SI; // skewed layout: strides of k and i, doubles per warp
};

// offsets into the dynamic shared memory, in doubles
template<int N>
struct WpSmem
{
  static constexpr int B = WpCfg<N>::B, N2 = N * N, N3 = N2 * N;
  static constexpr int U = 0;                                   // [B][N3] src values of the batch (bulk-copy destination)
  static constexpr int T = U + B * N3;                          // [NCW][WT] transposition area / result (bulk-store source)
  static constexpr int GN = T + WpCfg<N>::NCW * WpCfg<N>::WT;   // [3][2][B][N2] end derivatives of the batch's own lines
  static constexpr int VZ = GN + 6 * B * N2;                    // [2][B][N2] end values in z
  static constexpr int TRV = VZ + 2 * B * N2;                   // [HLMAX][N2] end values of out-of-batch neighbours
  static constexpr int TRG = TRV + WpCfg<N>::HLMAX * N2;        // [HLMAX][N2] end derivatives of out-of-batch neighbours
  static constexpr int END = TRG + WpCfg<N>::HLMAX * N2;
};

template<int N, int NP = 2>
inline size_t wp_smem_bytes()
{
  return (size_t)WpSmem<N>::END * sizeof(double) + (size_t)NP * WpCfg<N>::HLMAX * sizeof(i2) + (size_t)WpCfg<N>::B * sizeof(int64_t) /* neighbour table */
         + 4 * 8 /* mbarriers */ + 16;
}

// ---------------------------------------------------------------------------------------------------------------
// producer: traces of the out-of-batch neighbours.  Identical arithmetic to ws::ws_round; the difference is the single-buffered
// trace area: the loads of the first round are issued, THEN the warp waits for barrier B of the previous batch (its consumers
// are done with the area), then it reduces and stores.
// ---------------------------------------------------------------------------------------------------------------
template<int N, int R, int D, bool GH>
WS_FN void wp_round_load(const WsArgs & A, const i2 * hl, int e0, const double * own, const double * gho, double (&x)[R][N], int (&side)[R])
{
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int sd = (D == 0) ? 1 : (D == 1 ? N : N2); // stride along the line
  WS_UNROLL
  for (int q = 0; q < R; ++q) {
    const i2 h = hl[e0 + q];
    side[q] = h.x & 1;
    const double * line = ((!GH || h.y < A.n_owned) ? own : gho) + (size_t)h.y * N3;
    WS_UNROLL
    for (int i = 0; i < N; ++i) x[q][i] = line[i * sd];
  }
}

template<int N, int R>
WS_FN void wp_round_store(const WsTables<N> & T, int e0, const double (&x)[R][N], const int (&side)[R], int ab, bool act, double * TRV, double * TRG)
{
  constexpr int N2 = N * N;
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

// entries [e_begin, e_end) of direction D, dealt to the producer warps in rounds of R (then single) neighbour cells.
// `gate` is called once per batch by every producer warp before its first store (even if the warp has no round at all).
template<int N, int R, int D, bool GH, int NP, class Gate>
WS_FN void wp_produce_dir(const WsTables<N> & T, const WsArgs & A, const i2 * hl, int e_begin, int e_end, int & round, int pw, int ab, bool act, double * TRV, double * TRG,
                          Gate & gate)
{
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int s1 = (D == 0) ? N : 1, s2 = (D == 2) ? N : N2; // strides across the face
  const double * const own = A.src + (ab % N) * s1 + (ab / N) * s2;
  const double * const gho = GH ? A.ghost - A.n_owned * N3 + (ab % N) * s1 + (ab / N) * s2 : own; // ghost cells are numbered from n_owned
  int e0 = e_begin;
  for (; e0 + R <= e_end; e0 += R, ++round)
    if (round % NP == pw) { // warp-uniform
      double x[R][N]; int side[R];
      wp_round_load<N, R, D, GH>(A, hl, e0, own, gho, x, side);
      gate();
      wp_round_store<N, R>(T, e0, x, side, ab, act, TRV, TRG);
    }
  for (; e0 < e_end; ++e0, ++round)
    if (round % NP == pw) {
      double x[1][N]; int side[1];
      wp_round_load<N, 1, D, GH>(A, hl, e0, own, gho, x, side);
      gate();
      wp_round_store<N, 1>(T, e0, x, side, ab, act, TRV, TRG);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the CTA
// ---------------------------------------------------------------------------------------------------------------
template<int N, int R, bool GH, int NP, class RT>
WS_FN void wp_cta(RT & rt, const WsTables<N> & T, const WsArgs & A)
{
  using Cfg = WpCfg<N, NP>;
  using S = WpSmem<N>;
  constexpr int B = Cfg::B, CW = Cfg::CW, NCW = Cfg::NCW;
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int SK = Cfg::SK, SI = Cfg::SI, WT = Cfg::WT;
  static_assert(N == 5, "layout constants are those of n = 5");
  double * const smem = rt.smem();
  double * const U = smem + S::U;
  double * const GN = smem + S::GN;
  double * const VZ = smem + S::VZ;
  double * const TRV = smem + S::TRV;
  double * const TRG = smem + S::TRG;
  i2 * const hlS = reinterpret_cast<i2 *>(smem + S::END);                               // [NP][HLMAX]
  int64_t * const NL = reinterpret_cast<int64_t *>(hlS + NP * Cfg::HLMAX);              // [B] packed neighbour indices of the batch (handed over by barrier A)
  void * const barU = reinterpret_cast<char *>(NL + B);                                 // batch landed (bulk copy, tx count)
  void * const barA = reinterpret_cast<char *>(barU) + 8;                               // GN / VNz / TR of the batch complete
  void * const barB = reinterpret_cast<char *>(barU) + 16;                              // GN / VNz / TR of the batch consumed
  void * const barC = reinterpret_cast<char *>(barU) + 24;                              // U of the batch consumed

  const int t = rt.tid();
  const int warp = t / 32, lane = t % 32;
  const bool producer = warp >= NCW;
  if (t == 0) { rt.mbar_init(barU, 1); rt.mbar_init(barA, NCW + NP); rt.mbar_init(barB, NCW); rt.mbar_init(barC, NCW); }
  rt.sync_all();
  const int first = rt.cta(), step = rt.ncta();
  if (first >= A.n_items) return;

  auto batch_of = [&](int it) { return A.batches ? A.batches[it] : it; };
  auto batch_bytes = [&](int bt) { return (uint32_t)((int)ws::ws_min(B, A.n_owned - (int64_t)bt * B) * N3 * sizeof(double)); };

  if (producer) {
    // ------------------------------------------------------------------------------------------------------------
    const int pw = warp - NCW;
    i2 * const hl = hlS + pw * Cfg::HLMAX; // private staging of the halo list: a warp-level barrier orders its accesses
    const bool act = lane < N2;
    const int ab = act ? lane : 0;         // line within the face; the spare lanes shadow line 0 and store nothing
    rt.role_producer();
    ws::WsPrefetch pre;
    ws::ws_prefetch(A, batch_of(first), lane, pre);
    if (pw == 0 && lane == 0) rt.load_issue(barU, U, A.src + (int64_t)batch_of(first) * B * N3, batch_bytes(batch_of(first)));
    uint32_t n = 0; // batches handled so far
    bool ghosts_acquired = false;
    for (int it = first; it < A.n_items; it += step, ++n) {
      const int itn = it + step;
      if (GH && A.flags && !ghosts_acquired && it >= A.first_ghost_item) { ws::ws_acquire_ghosts(rt, A); ghosts_acquired = true; }
      // traces of batch `it`; the trace area was last read for the batch before (barrier B of iteration n - 1)
      const int cx = pre.c & 1023, cy = (pre.c >> 10) & 1023, cz = (pre.c >> 20) & 1023;
      hl[lane] = pre.h[0]; hl[lane + 32] = pre.h[1];
      if (itn < A.n_items) ws::ws_prefetch(A, batch_of(itn), lane, pre);
      rt.sync_warp();
      bool gated = false;
      auto gate = [&]() { if (!gated) { if (n > 0) rt.mbar_wait(barB, (n - 1) & 1); gated = true; } };
      int round = 0;
      if (GH && (!A.flags || it >= A.first_ghost_item)) { // only these items can have neighbours in the ghost buffer
        wp_produce_dir<N, (R > 8 ? 8 : R), 0, GH, NP>(T, A, hl, 0, cx, round, pw, ab, act, TRV, TRG, gate);
        wp_produce_dir<N, R, 1, GH, NP>(T, A, hl, cx, cx + cy, round, pw, ab, act, TRV, TRG, gate);
        wp_produce_dir<N, R, 2, GH, NP>(T, A, hl, cx + cy, cx + cy + cz, round, pw, ab, act, TRV, TRG, gate);
      } else {
        wp_produce_dir<N, (R > 8 ? 8 : R), 0, false, NP>(T, A, hl, 0, cx, round, pw, ab, act, TRV, TRG, gate);
        wp_produce_dir<N, R, 1, false, NP>(T, A, hl, cx, cx + cy, round, pw, ab, act, TRV, TRG, gate);
        wp_produce_dir<N, R, 2, false, NP>(T, A, hl, cx + cy, cx + cy + cz, round, pw, ab, act, TRV, TRG, gate);
      }
      gate(); // a warp without any round still orders itself behind the consumers (its arrival below completes barrier A)
      if (pw == 0 && lane < B) NL[lane] = A.nloc8[(size_t)batch_of(it) * B + lane]; // read by the compute warps between barriers A and B
      rt.sync_warp(); // all lanes' trace stores (and reads of hl) precede the elected arrival
      if (lane == 0) rt.mbar_arrive(barA);
      // the next batch may land as soon as the compute warps are done with U (barrier C of this iteration)
      if (pw == 0 && itn < A.n_items) {
        if (lane == 0) {
          rt.mbar_wait(barC, n & 1);
          rt.load_issue(barU, U, A.src + (int64_t)batch_of(itn) * B * N3, batch_bytes(batch_of(itn)));
        }
        rt.sync_warp();
      }
    }
    return;
  }

  // --------------------------------------------------------------------------------------------------------------
  // compute warp `warp`: cells 6 warp .. 6 warp + 5 of the batch; lane = 5 c + s
  rt.role_compute();
  const int c = lane / N, s = lane % N;           // cell of the warp, plane (plane layout: z = s; z layout: y = s)
  const bool lane_ok = lane < CW * N;
  const int lc = warp * CW + (lane_ok ? c : 0);   // cell of the batch
  double * const Tw = smem + S::T + warp * WT;    // the warp's transposition area
  const int sk = s + N * c;                       // lane-dependent part of the skewed address in the z layout (j = s)
  uint32_t n = 0;
  for (int it = first; it < A.n_items; it += step, ++n) {
    const int batch = batch_of(it);
    const int64_t b0 = (int64_t)batch * B;
    const int nvalid = (int)ws::ws_min(B, A.n_owned - b0);
    const bool valid = lane_ok && lc < nvalid;
    rt.mbar_wait(barU, n & 1);                       // the batch has landed
    if (n > 0) rt.mbar_wait(barB, (n - 1) & 1);      // GN / VNz of the previous batch are consumed

    double u[N][N], acc[N][N];
    // ---- T1: plane z = s.  End derivatives in x and y -> GN; cell part of the x and y sweeps ----
    if (valid) {
      WS_UNROLL
      for (int j = 0; j < N; ++j)
        WS_UNROLL
        for (int i = 0; i < N; ++i) u[j][i] = U[lc * N3 + s * N2 + i + N * j];
      WS_UNROLL
      for (int d = 0; d < 2; ++d) {
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
        for (int l = 0; l < N; ++l) { GN[((d * 2 + 0) * B + lc) * N2 + s * N + l] = g0[l]; GN[((d * 2 + 1) * B + lc) * N2 + s * N + l] = g1[l]; }
      }
      WS_UNROLL
      for (int l = 0; l < N; ++l)
        WS_UNROLL
        for (int r = 0; r < N; ++r) acc[l][r] = T.G[0][r * N] * u[l][0];
      WS_UNROLL
      for (int cc = 1; cc < N; ++cc)
        WS_UNROLL
        for (int l = 0; l < N; ++l)
          WS_UNROLL
          for (int r = 0; r < N; ++r) acc[l][r] = fma(T.G[0][r * N + cc], u[l][cc], acc[l][r]);
      WS_UNROLL
      for (int cc = 0; cc < N; ++cc)
        WS_UNROLL
        for (int l = 0; l < N; ++l)
          WS_UNROLL
          for (int r = 0; r < N; ++r) acc[r][l] = fma(T.G[1][r * N + cc], u[cc][l], acc[r][l]);
    }
    // ---- T2: xz-plane y = s.  uz[i][k]; end derivatives and end values in z -> GN, VNz (uz stays in registers until the z sweep) ----
    double uz[N][N];
    if (valid) {
      WS_UNROLL
      for (int i = 0; i < N; ++i)
        WS_UNROLL
        for (int k = 0; k < N; ++k) uz[i][k] = U[lc * N3 + k * N2 + i + N * s];
      double g0[N], g1[N];
      WS_UNROLL
      for (int i = 0; i < N; ++i) { g0[i] = T.fd[0][0] * uz[i][0]; g1[i] = T.fd[1][0] * uz[i][0]; }
      WS_UNROLL
      for (int k = 1; k < N; ++k)
        WS_UNROLL
        for (int i = 0; i < N; ++i) { g0[i] = fma(T.fd[0][k], uz[i][k], g0[i]); g1[i] = fma(T.fd[1][k], uz[i][k], g1[i]); }
      WS_UNROLL
      for (int i = 0; i < N; ++i) {
        GN[((2 * 2 + 0) * B + lc) * N2 + s * N + i] = g0[i]; GN[((2 * 2 + 1) * B + lc) * N2 + s * N + i] = g1[i];
        VZ[(0 * B + lc) * N2 + s * N + i] = uz[i][0]; VZ[(1 * B + lc) * N2 + s * N + i] = uz[i][N - 1];
      }
    }
    rt.sync_warp();
    if (lane == 0) rt.mbar_arrive(barA);
    rt.mbar_wait(barA, n & 1); // traces of every cell of the batch and of the out-of-batch neighbours are complete
    const int64_t nlc8 = NL[lc]; // the six neighbour indices of the lane's cell, one signed byte each

    // ---- P1: face terms in x and y on the register plane ----
    if (valid) {
      WS_UNROLL
      for (int d = 0; d < 2; ++d)
        WS_UNROLL
        for (int side = 0; side < 2; ++side) {
          const int nb = (int)(int8_t)(nlc8 >> (8 * (2 * d + side)));
          const bool inb = nb >= 0;
          const int e = -1 - nb;
          const int endn = side ? 0 : N - 1; // the neighbour's end node facing us
          const int voff = inb ? S::U + nb * N3 + s * N2 + (d == 0 ? endn : N * endn) : S::TRV + e * N2 + N * s;
          const int vstr = (inb && d == 0) ? N : 1;
          const int goff = inb ? S::GN + ((d * 2 + (side ^ 1)) * B + nb) * N2 + s * N : S::TRG + e * N2 + N * s;
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
    }
    rt.sync_warp();
    if (lane == 0) { rt.mbar_arrive(barC); rt.store_wait_read(); } // U is free; the previous result of this warp has left Tw
    rt.sync_warp();
    // ---- plane -> z layout through the skewed area: (i, j, k = s) ----
    if (valid) {
      WS_UNROLL
      for (int j = 0; j < N; ++j)
        WS_UNROLL
        for (int i = 0; i < N; ++i) Tw[j + N * c + SK * s + SI * i] = acc[j][i];
    }
    rt.sync_warp();
    // ---- P2: z sweep on the xz-plane y = s: face terms, cell part, mass matrix along z ----
    if (valid) {
      WS_UNROLL
      for (int i = 0; i < N; ++i)
        WS_UNROLL
        for (int k = 0; k < N; ++k) acc[i][k] = Tw[sk + SK * k + SI * i];
      WS_UNROLL
      for (int side = 0; side < 2; ++side) {
        const int nb = (int)(int8_t)(nlc8 >> (8 * (4 + side)));
        const bool inb = nb >= 0;
        const int e = -1 - nb;
        const int voff = inb ? S::VZ + ((side ^ 1) * B + nb) * N2 + s * N : S::TRV + e * N2 + N * s;
        const int goff = inb ? S::GN + ((2 * 2 + (side ^ 1)) * B + nb) * N2 + s * N : S::TRG + e * N2 + N * s;
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
    }
    rt.sync_warp();
    if (lane == 0) rt.mbar_arrive(barB); // GN / VNz / TR of this batch are consumed by this warp
    if (valid) {
      WS_UNROLL
      for (int cc = 0; cc < N; ++cc)
        WS_UNROLL
        for (int i = 0; i < N; ++i)
          WS_UNROLL
          for (int r = 0; r < N; ++r) acc[i][r] = fma(T.G[2][r * N + cc], uz[i][cc], acc[i][r]);
      // mass matrix along z: u <- M acc
      WS_UNROLL
      for (int i = 0; i < N; ++i)
        WS_UNROLL
        for (int r = 0; r < N; ++r) u[i][r] = T.M[r * N] * acc[i][0];
      WS_UNROLL
      for (int cc = 1; cc < N; ++cc)
        WS_UNROLL
        for (int i = 0; i < N; ++i)
          WS_UNROLL
          for (int r = 0; r < N; ++r) u[i][r] = fma(T.M[r * N + cc], acc[i][cc], u[i][r]);
      WS_UNROLL
      for (int i = 0; i < N; ++i)
        WS_UNROLL
        for (int r = 0; r < N; ++r) Tw[sk + SK * r + SI * i] = u[i][r];
    }
    rt.sync_warp();
    // ---- P3: mass matrices along x and y on the register plane z = s; result in the natural layout, stored by the warp ----
    if (valid) {
      WS_UNROLL
      for (int j = 0; j < N; ++j)
        WS_UNROLL
        for (int i = 0; i < N; ++i) u[j][i] = Tw[j + N * c + SK * s + SI * i];
      WS_UNROLL
      for (int j = 0; j < N; ++j)
        WS_UNROLL
        for (int r = 0; r < N; ++r) acc[j][r] = T.M[r * N] * u[j][0];
      WS_UNROLL
      for (int cc = 1; cc < N; ++cc)
        WS_UNROLL
        for (int j = 0; j < N; ++j)
          WS_UNROLL
          for (int r = 0; r < N; ++r) acc[j][r] = fma(T.M[r * N + cc], u[j][cc], acc[j][r]);
      WS_UNROLL
      for (int r = 0; r < N; ++r)
        WS_UNROLL
        for (int i = 0; i < N; ++i) u[r][i] = T.M[r * N] * acc[0][i];
      WS_UNROLL
      for (int cc = 1; cc < N; ++cc)
        WS_UNROLL
        for (int r = 0; r < N; ++r)
          WS_UNROLL
          for (int i = 0; i < N; ++i) u[r][i] = fma(T.M[r * N + cc], acc[cc][i], u[r][i]);
    }
    rt.sync_warp(); // every lane has read its skewed plane before the natural layout overwrites the area
    if (valid) {
      WS_UNROLL
      for (int r = 0; r < N; ++r)
        WS_UNROLL
        for (int i = 0; i < N; ++i) Tw[c * N3 + s * N2 + i + N * r] = u[r][i];
    }
    rt.fence_async();
    rt.sync_warp();
    if (lane == 0) {
      const int ncw = nvalid - warp * CW < CW ? nvalid - warp * CW : CW; // cells of this warp in a ragged last batch
      if (ncw > 0) rt.store_issue(A.dst + (b0 + warp * CW) * N3, Tw, (uint32_t)(ncw * N3 * sizeof(double)), A.add != 0);
    }
  }
  if (lane == 0) rt.store_wait_all(); // the warp's bulk stores are complete before the CTA exits
}

} // namespace wp
} // namespace exadg_b200
