// General-geometry SIPG Laplace vmult (curved / non-periodic meshes), FP64, degrees 1..7.
//
// One kernel fuses what the reference does in three MatrixFree loops
// (I/operators/operator_base.cpp:1349-1434): the cell integral, the own side of all six face
// integrals (interior faces and homogeneous Dirichlet/Neumann boundary faces) and the write of
// dst.  Evaluating every interior face from both sides (the reference's cell-based view,
// operator_base.cpp:1618-1704) removes the scatter-add into the neighbour and with it any atomics
// or colouring: each DoF of dst is written exactly once, by one thread, in a fixed order.
//
// Thread layout: n^2 threads per cell (n = k+1), CPB cells per CTA; every 1-D contraction is done by
// the thread that owns the line, which keeps the line in registers (n loads, n^2 FMAs, n stores);
// the 1-D matrices are kernel parameters, i.e. constant-bank operands of the DFMAs.
// Work is done in the collocation basis psi (Lagrange polynomials on the Gauss points):
//   u_q = S u;  cell: R = sum_e Dq_e^T [G (Dq u_q)];  faces: rank-1 updates of R along normal lines;
//   y = S^T R   (S^T maps test coefficients back to the nodal basis since span{psi} = span{l}).
// Quadrature-point physics follows laplace_operator.cpp:129-265 and laplace_operator.h:180-197.
#include <algorithm>
#include <cstdlib>
#include <stdexcept>

#include "operator.cuh"

namespace exadg_b200
{
namespace
{
template<int N>
struct GenTables
{
  double S[N * N], Dq[N * N], sv[2][N], sd[2][N], fd[2][N];
};

struct GenArgs
{
  const int32_t * nb; const int32_t * face_id; const uint8_t * face_info;
  const double * cellG; const double * faceG; const double * tau_f;
  const double * src; const double * ghost; double * dst;
  const int32_t * cells; int64_t n_items; int64_t n_owned; int add;
  // Helmholtz / viscous operator: mass * (v, u) + lap * a_SIPG(u, v) on each of ncomp components; an item is a
  // (cell, component) pair and doubles as the block index of the vector (cell-major, then component)
  int ncomp; double mass, lap; const double * cellJxW;
};

template<int N, bool TRANSPOSE>
__device__ __forceinline__ void sweep(const double * __restrict__ M, const double * in, double * out, int base, int stride)
{
  double u[N];
#pragma unroll
  for (int i = 0; i < N; ++i) u[i] = in[base + i * stride];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c) acc = fma(TRANSPOSE ? M[c * N + r] : M[r * N + c], u[c], acc);
    out[base + r * stride] = acc;
  }
}

// MODE 0: dst (+)= A src.  MODE 1: diagonal (unit vectors column by column, neighbour function zero:
// do_face_int_integral, laplace_operator.cpp:165-191; operator_base.cpp:1552-1616).
// resident CTAs per SM the register allocation is tuned for (measured per degree on B200, curved periodic box)
// N^2 <= 32: the lines of a cell live in ONE warp (32 / N^2 cells per warp; the spare lanes exit), so every barrier between the sweeps
// is a __syncwarp of the participating lanes and the warps of a CTA run decoupled from each other; larger N: N^2 threads per cell and
// block barriers
#ifndef EXADG_B200_CELL_FUSED_ALL
#define EXADG_B200_CELL_FUSED_ALL 0
#endif
constexpr bool CELL_FUSED_ALL = EXADG_B200_CELL_FUSED_ALL != 0;
template<int N> struct GenCfg
{
  static constexpr int MIN_BLOCKS = (N == 5 || N == 7) ? 2 : 4;
  static constexpr bool WARP_CELLS = (N * N <= 32);
  // face phase with line ownership (every 2-D operation on a face field line by line) or with a row per face point: measured per degree on
  // the curved box (scripts/r02_shot28.sh): k=2 13.3 -> 13.9, k=4 18.4 -> 21.5 GDoF/s, but k=3 24.9 -> 18.7, k=5 18.6 -> 12.9, k=6 17.2 -> 11.2
  static constexpr bool FACE_LINES = (N == 3 || N == 5);
  // cell part with the z sweeps in registers and chained sweeps (see the kernel), per degree after measurement on the curved box
  // (scripts/r02_shot28.sh, GDoF/s without -> with): k=2 13.9 -> 14.7, k=3 24.9 -> 26.9, k=5 18.6 -> 19.8, k=6 17.2 -> 17.3, k=7 22.2 -> 22.6,
  // but k=4 21.5 -> 20.9 (the long fused column step costs more latency than its shared-memory accesses saved)
  static constexpr bool CELL_FUSED = CELL_FUSED_ALL || (N != 5);
  static constexpr int CPW = WARP_CELLS ? 32 / (N * N) : 0;                      // cells per warp
  static constexpr int WARPS = (N == 5) ? 8 : 4;                                  // warps per CTA (warp-cell layout)
  static constexpr int CPB = WARP_CELLS ? CPW * WARPS : ((N * N >= 64) ? 2 : 4);  // cells per CTA
  static constexpr int THREADS = WARP_CELLS ? 32 * WARPS : N * N * CPB;
};
// warp-cell layout: only the lanes that own a line take part (the spare lanes of a warp exit at once)
template<int N>
__device__ __forceinline__ void gen_sync()
{
  constexpr int LANES = GenCfg<N>::CPW * N * N;
  if (GenCfg<N>::WARP_CELLS) __syncwarp(LANES >= 32 ? 0xffffffffu : ((1u << (LANES & 31)) - 1u)); else __syncthreads();
}

template<int N, int CPB, int MODE>
__global__ void __launch_bounds__(GenCfg<N>::THREADS, GenCfg<N>::MIN_BLOCKS) vmult_general_kernel(const __grid_constant__ GenTables<N> T, const GenArgs A)
{
  constexpr int NP = N | 1;          // odd x-extent: conflict-free 64-bit shared accesses in every direction
  constexpr int SZ = NP * N * N;
  constexpr int N2 = N * N, N3 = N * N * N;
  constexpr int FS = (GenCfg<N>::FACE_LINES ? 18 : 10) * N2; // face scratch per cell
  extern __shared__ double smem[];
  const int t = threadIdx.x;
  if (GenCfg<N>::WARP_CELLS && (t & 31) >= GenCfg<N>::CPW * N2) return; // spare lanes of the warp-cell layout
  const int lc = GenCfg<N>::WARP_CELLS ? (t >> 5) * GenCfg<N>::CPW + (t & 31) / N2 : t / N2;
  const int r = GenCfg<N>::WARP_CELLS ? (t & 31) % N2 : t % N2, a = r % N, b = r / N;
  double * Uq = smem + (size_t)lc * (4 * SZ + FS);
  double * F0 = Uq + SZ, * F1 = F0 + SZ, * F2 = F1 + SZ;
  double * W = F2 + SZ; // [10][N2]
  const int64_t item = (int64_t)blockIdx.x * CPB + lc;
  const bool valid = item < A.n_items;
  const int64_t block = valid ? (A.cells ? (int64_t)A.cells[item] : item) : (A.cells ? (int64_t)A.cells[0] : 0); // (cell, component)
  const int64_t cell = block / A.ncomp; const int comp = (int)(block % A.ncomp);

  const int lbase[3] = {NP * (a + N * b), a + NP * N * b, a + NP * b};
  const int lstr[3] = {1, NP, NP * N};
  const int nstr[3] = {1, N, N2}; // nodal (global) strides

  for (int col = 0; col < (MODE == 1 ? N3 : 1); ++col) {
    double * R = F0;
    if constexpr (GenCfg<N>::CELL_FUSED) {
      // Fused cell part: every sweep along z is done by the thread that owns the column (x, y) = (a, b) in registers - the nodal values come
      // straight from global memory, the z derivative, the metric terms and the tested z flux never see shared memory - and consecutive
      // sweeps along the same line are chained (S_y with Dq_y).  35 % fewer shared-memory accesses than the phase-by-phase form below.
      double uz[N], r2[N];
      { // P0 + S_z
        double u[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const int nodal = a + N * (b + N * k);
          u[k] = (MODE == 1) ? ((nodal == col) ? 1.0 : 0.0) : A.src[block * N3 + nodal];
        }
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < N; ++k) acc = fma(T.S[q * N + k], u[k], acc);
          Uq[a + NP * (b + N * q)] = acc;
        }
      }
      gen_sync<N>();
      sweep<N, false>(T.S, Uq, Uq, lbase[0], lstr[0]); gen_sync<N>();
      { // S_y chained with Dq_y: the y line (x = a, z = b)
        double u[N], v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) u[i] = Uq[lbase[1] + i * lstr[1]];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) acc = fma(T.S[q * N + i], u[i], acc);
          v[q] = acc;
          Uq[lbase[1] + q * lstr[1]] = acc;
        }
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) acc = fma(T.Dq[q * N + i], v[i], acc);
          F1[lbase[1] + q * lstr[1]] = acc;
        }
      }
      gen_sync<N>();
      sweep<N, false>(T.Dq, Uq, F0, lbase[0], lstr[0]);
      gen_sync<N>();
      { // column (a, b): Dq_z, flux = G grad (get_gradient + submit_gradient, laplace_operator.cpp:135), Dq_z^T, mass term
        double d2[N], f2[N];
#pragma unroll
        for (int k = 0; k < N; ++k) uz[k] = Uq[a + NP * (b + N * k)];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < N; ++k) acc = fma(T.Dq[q * N + k], uz[k], acc);
          d2[q] = acc;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const int q = a + N * (b + N * k), s = a + NP * (b + N * k);
          const double * g = A.cellG + (size_t)cell * 6 * N3 + q;
          const double gxx = g[0], gyy = g[N3], gzz = g[2 * N3], gxy = g[3 * N3], gxz = g[4 * N3], gyz = g[5 * N3];
          const double d0 = F0[s], d1 = F1[s];
          F0[s] = A.lap * (gxx * d0 + gxy * d1 + gxz * d2[k]);
          F1[s] = A.lap * (gxy * d0 + gyy * d1 + gyz * d2[k]);
          f2[k] = A.lap * (gxz * d0 + gyz * d1 + gzz * d2[k]);
        }
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < N; ++k) acc = fma(T.Dq[k * N + q], f2[k], acc);
          r2[q] = acc;
        }
        if (A.cellJxW) { // MassKernel::get_volume_flux (mass_kernel.h:76-82): submit_value(scaling_factor * u)
#pragma unroll
          for (int k = 0; k < N; ++k) r2[k] = fma(A.mass * A.cellJxW[(size_t)cell * N3 + a + N * (b + N * k)], uz[k], r2[k]);
        }
      }
      gen_sync<N>();
      sweep<N, true>(T.Dq, F0, F0, lbase[0], lstr[0]);
      sweep<N, true>(T.Dq, F1, F1, lbase[1], lstr[1]);
      gen_sync<N>();
#pragma unroll
      for (int k = 0; k < N; ++k) { const int s = a + NP * (b + N * k); R[s] += F1[s] + r2[k]; }
      gen_sync<N>();
    } else {
    // ---- P0: nodal values into shared memory ----
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const int nodal = a + N * (b + N * k);
      double v;
      if (MODE == 1) v = (nodal == col) ? 1.0 : 0.0;
      else v = A.src[block * N3 + nodal];
      Uq[a + NP * (b + N * k)] = v;
    }
    gen_sync<N>();
    // ---- P1: to the collocation basis ----
    sweep<N, false>(T.S, Uq, Uq, lbase[0], lstr[0]); gen_sync<N>();
    sweep<N, false>(T.S, Uq, Uq, lbase[1], lstr[1]); gen_sync<N>();
    sweep<N, false>(T.S, Uq, Uq, lbase[2], lstr[2]); gen_sync<N>();
    // ---- P2: reference gradient ----
    sweep<N, false>(T.Dq, Uq, F0, lbase[0], lstr[0]);
    sweep<N, false>(T.Dq, Uq, F1, lbase[1], lstr[1]);
    sweep<N, false>(T.Dq, Uq, F2, lbase[2], lstr[2]);
    gen_sync<N>();
    // ---- P3: flux = G grad (get_gradient + submit_gradient, laplace_operator.cpp:135) ----
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const int q = a + N * (b + N * k), s = a + NP * (b + N * k);
      const double * g = A.cellG + (size_t)cell * 6 * N3 + q;
      const double gxx = g[0], gyy = g[N3], gzz = g[2 * N3], gxy = g[3 * N3], gxz = g[4 * N3], gyz = g[5 * N3];
      const double d0 = F0[s], d1 = F1[s], d2 = F2[s];
      F0[s] = A.lap * (gxx * d0 + gxy * d1 + gxz * d2);
      F1[s] = A.lap * (gxy * d0 + gyy * d1 + gyz * d2);
      F2[s] = A.lap * (gxz * d0 + gyz * d1 + gzz * d2);
    }
    gen_sync<N>();
    // ---- P4: test with grad psi ----
    sweep<N, true>(T.Dq, F0, F0, lbase[0], lstr[0]);
    sweep<N, true>(T.Dq, F1, F1, lbase[1], lstr[1]);
    sweep<N, true>(T.Dq, F2, F2, lbase[2], lstr[2]);
    gen_sync<N>();
#pragma unroll
    for (int k = 0; k < N; ++k) { const int s = a + NP * (b + N * k); R[s] += F1[s] + F2[s]; }
    if (A.cellJxW) { // MassKernel::get_volume_flux (mass_kernel.h:76-82): submit_value(scaling_factor * u)
#pragma unroll
      for (int k = 0; k < N; ++k) { const int s = a + NP * (b + N * k); R[s] = fma(A.mass * A.cellJxW[(size_t)cell * N3 + a + N * (b + N * k)], Uq[s], R[s]); }
    }
    gen_sync<N>();

    }
    // ---- P5: faces, one direction (two faces) at a time ----
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double * Wvm = W, * Wv2 = W + 2 * N2, * Wd2 = W + 4 * N2, * Wx = W + 6 * N2, * Wy = W + 8 * N2;           // [2 sides][N2] each
      double * Wm1 = W + 10 * N2, * Wm2 = W + 12 * N2, * Wp1 = W + 14 * N2, * Wp2 = W + 16 * N2;
      double vm[2], dm[2];
      int32_t nbc[2]; int32_t F[2]; int info[2];
      {
        double u[N];
#pragma unroll
        for (int i = 0; i < N; ++i) u[i] = Uq[lbase[d] + i * lstr[d]];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          double v = 0.0, g = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) { v = fma(T.sv[s][i], u[i], v); g = fma(T.sd[s][i], u[i], g); }
          vm[s] = v; dm[s] = g;
          Wvm[s * N2 + r] = v;
        }
      }
      const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int f = 2 * d + s;
        nbc[s] = A.nb[cell * 6 + f]; F[s] = A.face_id[cell * 6 + f]; info[s] = A.face_info[cell * 6 + f];
        double v2 = 0.0, d2 = 0.0;
        if (MODE == 0 && nbc[s] >= 0) {
          const double * un = (nbc[s] < A.n_owned) ? A.src + ((size_t)nbc[s] * A.ncomp + comp) * N3 : A.ghost + ((size_t)(nbc[s] - A.n_owned) * A.ncomp + comp) * N3;
          const int sp = info[s] & 1; // side of the neighbour's face (standard orientation: same direction d)
          const int off = a * nstr[t1] + b * nstr[t2];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const double x = un[off + i * nstr[d]];
            d2 = fma(sp ? T.fd[1][i] : T.fd[0][i], x, d2);
            if (i == 0 && !sp) v2 = x;      // l_j(0) = delta_{j,0}
            if (i == N - 1 && sp) v2 = x;   // l_j(1) = delta_{j,n-1}
          }
        }
        Wv2[s * N2 + r] = v2; Wd2[s * N2 + r] = d2;
      }
      double zc[2], cgd[2];
      if constexpr (GenCfg<N>::FACE_LINES) {
      gen_sync<N>();
      // The 2-D operations on the face fields (n x n values, two faces per direction) are done LINE BY LINE: an item is one line of one
      // field, its owner loads the n values once and produces all n outputs (the first version gave every face point its own row:
      // n loads per output, and five such loops made the face phase half of the kernel's shared-memory traffic).
      // nodal -> collocation along t1: Wx = S Wv2, Wy = S Wd2   (4 n items: field f in {v2 s0, v2 s1, d2 s0, d2 s1}, line = fixed t2 index)
      for (int it = r; it < 4 * N; it += N2) {
        const int f = it / N, l = it % N;
        const double * in = ((f < 2) ? Wv2 : Wd2) + (f & 1) * N2 + N * l;
        double * out = ((f < 2) ? Wx : Wy) + (f & 1) * N2 + N * l;
        double u[N];
#pragma unroll
        for (int p = 0; p < N; ++p) u[p] = in[p];
#pragma unroll
        for (int o = 0; o < N; ++o) {
          double acc = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) acc = fma(T.S[o * N + p], u[p], acc);
          out[o] = acc;
        }
      }
      gen_sync<N>();
      // ... and along t2: Wv2 <- S Wx (= v+ at the face points), Wd2 <- S Wy (= reference normal derivative of u+)
      for (int it = r; it < 4 * N; it += N2) {
        const int f = it / N, l = it % N;
        const double * in = ((f < 2) ? Wx : Wy) + (f & 1) * N2 + l;
        double * out = ((f < 2) ? Wv2 : Wd2) + (f & 1) * N2 + l;
        double u[N];
#pragma unroll
        for (int p = 0; p < N; ++p) u[p] = in[N * p];
#pragma unroll
        for (int o = 0; o < N; ++o) {
          double acc = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) acc = fma(T.S[o * N + p], u[p], acc);
          out[N * o] = acc;
        }
      }
      gen_sync<N>();
      // tangential derivatives of v- and v+ (8 n items: field in {v- s0, v- s1, v+ s0, v+ s1} x direction in {t1, t2})
      for (int it = r; it < 8 * N; it += N2) {
        const int g = it / N, l = it % N, fld = g & 3, dir = g >> 2;
        const double * in = ((fld < 2) ? Wvm : Wv2) + (fld & 1) * N2;
        double * out = ((fld < 2) ? (dir ? Wm2 : Wm1) : (dir ? Wp2 : Wp1)) + (fld & 1) * N2;
        const int base = dir ? l : N * l, st = dir ? N : 1;
        double u[N];
#pragma unroll
        for (int p = 0; p < N; ++p) u[p] = in[base + p * st];
#pragma unroll
        for (int o = 0; o < N; ++o) {
          double acc = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) acc = fma(T.Dq[o * N + p], u[p], acc);
          out[base + o * st] = acc;
        }
      }
      gen_sync<N>();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const double m1 = Wm1[s * N2 + r], m2 = Wm2[s * N2 + r], p1 = Wp1[s * N2 + r], p2 = Wp2[s * N2 + r];
        const double vps = Wv2[s * N2 + r], dps = Wd2[s * N2 + r];
        const bool plus = info[s] & 8; const int bt = (info[s] >> 4) & 3;
        const double sgn = plus ? -1.0 : 1.0;
        const double * fg = A.faceG + (size_t)F[s] * 7 * N2 + r;
        const int om = plus ? 3 : 0, op = plus ? 0 : 3;
        double am[3], ap[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) { am[e] = sgn * fg[(om + e) * N2]; ap[e] = sgn * fg[(op + e) * N2]; }
        const double jxw = fg[6 * N2], tau = A.tau_f[F[s]];
        const double dnm = am[d] * dm[s] + am[t1] * m1 + am[t2] * m2;
        double dnp = ap[d] * dps + ap[t1] * p1 + ap[t2] * p2;
        double vpl = vps;
        if (bt == BT_DIRICHLET) { vpl = -vm[s]; dnp = dnm; }        // weak_boundary_conditions.h:44,118-121
        else if (bt == BT_NEUMANN) { vpl = vm[s]; dnp = -dnm; }     // :45,124-127
        else if (MODE == 1) { vpl = 0.0; dnp = 0.0; }               // exterior function zero
        const double jump = vm[s] - vpl;
        const double gf = A.lap * (-0.5 * jump);                    // laplace_operator.h:180-185 (viscous_operator.h:489-525 with nu)
        const double vf = A.lap * (0.5 * (dnm + dnp) - tau * jump); // laplace_operator.h:187-197
        zc[s] = -vf * jxw;                                          // submit_value(-value_flux)
        cgd[s] = am[d] * gf * jxw;                                  // submit_normal_derivative(gradient_flux)
        Wx[s * N2 + r] = am[t1] * gf * jxw;
        Wy[s * N2 + r] = am[t2] * gf * jxw;
      }
      gen_sync<N>();
      // tangential parts of submit_normal_derivative, tested with the derivatives of the collocation basis: Wm1 = Dq^T_t1 Wx, Wm2 = Dq^T_t2 Wy
      for (int it = r; it < 4 * N; it += N2) {
        const int f = it / N, l = it % N, dir = f >> 1;
        const double * in = (dir ? Wy : Wx) + (f & 1) * N2;
        double * out = (dir ? Wm2 : Wm1) + (f & 1) * N2;
        const int base = dir ? l : N * l, st = dir ? N : 1;
        double u[N];
#pragma unroll
        for (int p = 0; p < N; ++p) u[p] = in[base + p * st];
#pragma unroll
        for (int o = 0; o < N; ++o) {
          double acc = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) acc = fma(T.Dq[p * N + o], u[p], acc);
          out[base + o * st] = acc;
        }
      }
      gen_sync<N>();
#pragma unroll
      for (int s = 0; s < 2; ++s) zc[s] += Wm1[s * N2 + r] + Wm2[s * N2 + r];
      } else {
      gen_sync<N>();
      // nodal -> collocation on the face, first tangential direction
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        double x = 0.0, y = 0.0;
#pragma unroll
        for (int p = 0; p < N; ++p) { x = fma(T.S[a * N + p], Wv2[s * N2 + p + N * b], x); y = fma(T.S[a * N + p], Wd2[s * N2 + p + N * b], y); }
        Wx[s * N2 + r] = x; Wy[s * N2 + r] = y;
      }
      gen_sync<N>();
      double vp[2], dp[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        double x = 0.0, y = 0.0;
#pragma unroll
        for (int p = 0; p < N; ++p) { x = fma(T.S[b * N + p], Wx[s * N2 + a + N * p], x); y = fma(T.S[b * N + p], Wy[s * N2 + a + N * p], y); }
        vp[s] = x; dp[s] = y;
        Wv2[s * N2 + r] = x;
      }
      gen_sync<N>();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        double m1 = 0.0, m2 = 0.0, p1 = 0.0, p2 = 0.0;
#pragma unroll
        for (int p = 0; p < N; ++p) {
          m1 = fma(T.Dq[a * N + p], Wvm[s * N2 + p + N * b], m1); m2 = fma(T.Dq[b * N + p], Wvm[s * N2 + a + N * p], m2);
          p1 = fma(T.Dq[a * N + p], Wv2[s * N2 + p + N * b], p1); p2 = fma(T.Dq[b * N + p], Wv2[s * N2 + a + N * p], p2);
        }
        const bool plus = info[s] & 8; const int bt = (info[s] >> 4) & 3;
        const double sgn = plus ? -1.0 : 1.0;
        const double * fg = A.faceG + (size_t)F[s] * 7 * N2 + r;
        const int om = plus ? 3 : 0, op = plus ? 0 : 3;
        double am[3], ap[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) { am[e] = sgn * fg[(om + e) * N2]; ap[e] = sgn * fg[(op + e) * N2]; }
        const double jxw = fg[6 * N2], tau = A.tau_f[F[s]];
        const double dnm = am[d] * dm[s] + am[t1] * m1 + am[t2] * m2;
        double dnp = ap[d] * dp[s] + ap[t1] * p1 + ap[t2] * p2;
        double vpl = vp[s];
        if (bt == BT_DIRICHLET) { vpl = -vm[s]; dnp = dnm; }        // weak_boundary_conditions.h:44,118-121
        else if (bt == BT_NEUMANN) { vpl = vm[s]; dnp = -dnm; }     // :45,124-127
        else if (MODE == 1) { vpl = 0.0; dnp = 0.0; }               // exterior function zero
        const double jump = vm[s] - vpl;
        const double gf = A.lap * (-0.5 * jump);                    // laplace_operator.h:180-185 (viscous_operator.h:489-525 with nu)
        const double vf = A.lap * (0.5 * (dnm + dnp) - tau * jump); // laplace_operator.h:187-197
        zc[s] = -vf * jxw;                                          // submit_value(-value_flux)
        cgd[s] = am[d] * gf * jxw;                                  // submit_normal_derivative(gradient_flux)
        Wx[s * N2 + r] = am[t1] * gf * jxw;
        Wy[s * N2 + r] = am[t2] * gf * jxw;
      }
      gen_sync<N>();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        double z = zc[s];
#pragma unroll
        for (int p = 0; p < N; ++p) { z = fma(T.Dq[p * N + a], Wx[s * N2 + p + N * b], z); z = fma(T.Dq[p * N + b], Wy[s * N2 + a + N * p], z); }
        zc[s] = z;
      }
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double x = R[lbase[d] + i * lstr[d]];
        x = fma(T.sv[0][i], zc[0], x); x = fma(T.sd[0][i], cgd[0], x);
        x = fma(T.sv[1][i], zc[1], x); x = fma(T.sd[1][i], cgd[1], x);
        R[lbase[d] + i * lstr[d]] = x;
      }
      gen_sync<N>();
    }
    // ---- P6: back to the nodal basis, write ----
    sweep<N, true>(T.S, R, R, lbase[0], lstr[0]); gen_sync<N>();
    sweep<N, true>(T.S, R, R, lbase[1], lstr[1]); gen_sync<N>();
    if constexpr (GenCfg<N>::CELL_FUSED) {
      // S_z^T by the owner of the column (a, b) in registers, straight to global memory
      double rz[N];
#pragma unroll
      for (int k = 0; k < N; ++k) rz[k] = R[a + NP * (b + N * k)];
#pragma unroll
      for (int q = 0; q < N; ++q) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) v = fma(T.S[k * N + q], rz[k], v);
        const int nodal = a + N * (b + N * q);
        if (MODE == 1) {
          if (valid && nodal == col) { if (A.add) A.dst[block * N3 + col] += v; else A.dst[block * N3 + col] = v; }
        } else if (valid) {
          if (A.add) A.dst[block * N3 + nodal] += v; else A.dst[block * N3 + nodal] = v;
        }
      }
    } else {
      sweep<N, true>(T.S, R, R, lbase[2], lstr[2]); gen_sync<N>();
      if (MODE == 1) {
        const int k = col / N2;
        if (valid && r == col % N2) {
          const double v = R[a + NP * (b + N * k)];
          if (A.add) A.dst[block * N3 + col] += v; else A.dst[block * N3 + col] = v;
        }
      } else if (valid) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const int nodal = a + N * (b + N * k);
          const double v = R[a + NP * (b + N * k)];
          if (A.add) A.dst[block * N3 + nodal] += v; else A.dst[block * N3 + nodal] = v;
        }
      }
    }
    gen_sync<N>();
  }
}

template<int N>
GenTables<N> make_tables()
{
  Tables1D tab(N - 1);
  GenTables<N> T;
  for (int i = 0; i < N * N; ++i) { T.S[i] = (double)tab.S[i]; T.Dq[i] = (double)tab.Dq[i]; }
  for (int s = 0; s < 2; ++s) for (int i = 0; i < N; ++i) { T.sv[s][i] = (double)tab.sv[s][i]; T.sd[s][i] = (double)tab.sd[s][i]; T.fd[s][i] = (double)tab.fd[s][i]; }
  return T;
}

template<int N, int MODE>
void launch_n(const DeviceOperator & op, double * dst, const double * src, bool add, const int32_t * cells, int64_t n_items, cudaStream_t stream)
{
  constexpr int CPB = GenCfg<N>::CPB;
  constexpr int NP = N | 1;
  static const GenTables<N> T = make_tables<N>();
  const size_t smem = (size_t)CPB * (4 * NP * N * N + (GenCfg<N>::FACE_LINES ? 18 : 10) * N * N) * sizeof(double);
  if (first_use_on_device((const void *)vmult_general_kernel<N, CPB, MODE>))
    CUDA_CHECK(cudaFuncSetAttribute(vmult_general_kernel<N, CPB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GenArgs A;
  A.nb = op.nb; A.face_id = op.face_id; A.face_info = op.face_info; A.cellG = op.cellG; A.faceG = op.faceG; A.tau_f = op.tau_f;
  A.src = src; A.ghost = op.ghost; A.dst = dst; A.cells = cells; A.n_items = n_items; A.n_owned = op.n_owned; A.add = add ? 1 : 0;
  A.ncomp = op.n_components; A.mass = op.mass_coeff; A.lap = op.laplace_coeff; A.cellJxW = op.mass_coeff != 0.0 ? op.cellJxW : nullptr;
  if (n_items == 0) return;
  const unsigned grid = (unsigned)((n_items + CPB - 1) / CPB);
  vmult_general_kernel<N, CPB, MODE><<<grid, GenCfg<N>::THREADS, smem, stream>>>(T, A);
  CUDA_CHECK(cudaGetLastError());
}

template<int MODE>
void dispatch(const DeviceOperator & op, double * dst, const double * src, bool add, const int32_t * cells, int64_t n_items, cudaStream_t stream)
{
  switch (op.n) {
    case 2: launch_n<2, MODE>(op, dst, src, add, cells, n_items, stream); break;
    case 3: launch_n<3, MODE>(op, dst, src, add, cells, n_items, stream); break;
    case 4: launch_n<4, MODE>(op, dst, src, add, cells, n_items, stream); break;
    case 5: launch_n<5, MODE>(op, dst, src, add, cells, n_items, stream); break;
    case 6: launch_n<6, MODE>(op, dst, src, add, cells, n_items, stream); break;
    case 7: launch_n<7, MODE>(op, dst, src, add, cells, n_items, stream); break;
    case 8: launch_n<8, MODE>(op, dst, src, add, cells, n_items, stream); break;
    default: throw std::runtime_error("unsupported degree (1..7)");
  }
}
struct InvMassTable { int n; double Sinv[EXADG_MAX_N * EXADG_MAX_N]; };

// one CTA per (cell, component) block: t = S^-T r (three sweeps), t /= JxW, dst = S^-1 t (three sweeps); fixed summation order
__global__ void __launch_bounds__(128) inverse_mass_kernel(const InvMassTable T, const double * __restrict__ jxw, int ncomp, double * __restrict__ dst,
                                                           const double * __restrict__ src, int64_t n_blocks)
{
  __shared__ double A[512], B[512];
  const int n = T.n, n2 = n * n, n3 = n2 * n;
  for (int64_t block = blockIdx.x; block < n_blocks; block += gridDim.x) {
    const int64_t cell = block / ncomp;
    for (int i = threadIdx.x; i < n3; i += blockDim.x) A[i] = src[block * n3 + i];
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: matrix S^-T (entry [o][i] = Sinv[i][o]); pass 1: matrix S^-1
      for (int e = threadIdx.x; e < n3; e += blockDim.x) { // x
        const int o = e % n, jk = e / n;
        double v = 0.0;
        for (int i = 0; i < n; ++i) v = fma(pass == 0 ? T.Sinv[i * n + o] : T.Sinv[o * n + i], A[i + n * jk], v);
        B[e] = v;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < n3; e += blockDim.x) { // y
        const int i = e % n, o = (e / n) % n, k = e / n2;
        double v = 0.0;
        for (int j = 0; j < n; ++j) v = fma(pass == 0 ? T.Sinv[j * n + o] : T.Sinv[o * n + j], B[i + n * (j + n * k)], v);
        A[e] = v;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < n3; e += blockDim.x) { // z
        const int ij = e % n2, o = e / n2;
        double v = 0.0;
        for (int k = 0; k < n; ++k) v = fma(pass == 0 ? T.Sinv[k * n + o] : T.Sinv[o * n + k], A[ij + n2 * k], v);
        B[e] = pass == 0 ? v / jxw[cell * n3 + e] : v;
      }
      __syncthreads();
      if (pass == 0) { for (int e = threadIdx.x; e < n3; e += blockDim.x) A[e] = B[e]; __syncthreads(); }
    }
    for (int e = threadIdx.x; e < n3; e += blockDim.x) dst[block * n3 + e] = B[e];
    __syncthreads();
  }
}
} // namespace

void launch_inverse_mass(const DeviceOperator & op, double * dst, const double * src, cudaStream_t stream)
{
  if (!op.cellJxW) throw std::runtime_error("inverse mass: the operator was created without the mass data (exadg_b200_create_*_helmholtz)");
  Tables1D tab(op.degree);
  InvMassTable T; T.n = op.n;
  const std::vector<real_t> Si = invert(tab.S, op.n);
  for (int i = 0; i < op.n * op.n; ++i) T.Sinv[i] = (double)Si[i];
  const int64_t n_blocks = op.n_owned * op.n_components;
  if (n_blocks == 0) return;
  inverse_mass_kernel<<<(unsigned)std::min<int64_t>(n_blocks, 148 * 16), 128, 0, stream>>>(T, op.cellJxW, op.n_components, dst, src, n_blocks);
  CUDA_CHECK(cudaGetLastError());
}

namespace
{
} // namespace

void launch_vmult_general(const DeviceOperator & op, double * dst, const double * src, bool add, const int32_t * cells, int64_t n_cells, cudaStream_t stream)
{
  dispatch<0>(op, dst, src, add, cells, cells ? n_cells : op.n_owned * op.n_components, stream);
}

void launch_diagonal_general(const DeviceOperator & op, double * diag, bool add, cudaStream_t stream)
{
  dispatch<1>(op, diag, nullptr, add, nullptr, op.n_owned * op.n_components, stream);
}

} // namespace exadg_b200
