// Cartesian fast path of the SIPG Laplace vmult: uniform axis-aligned cells, all faces interior
// (the periodic box of applications/poisson/throughput, I/grid/periodic_box.h:35-88), FP64.
//
// Same operator as vmult_general.cu (cell_loop + face_loop of I/operators/operator_base.cpp:1349-1397
// with the fluxes of I/poisson/spatial_discretization/laplace_operator.h:180-197), restructured for the
// FP64 pipe, which - not HBM - bounds this case.  On a uniform box the operator is a sum of
// Kronecker products,
//     A = sum_d c_d (M x M x L_d),      c_d = h_e h_f / h_d,
// with the 1-D mass matrix M and the 1-D SIPG operator L_d (cell stiffness + both faces, coupling a
// line of a cell to the same line of its two neighbours in direction d only through the neighbour's
// end value v and end derivative g).  Factoring the mass matrices out,
//     y = (M x M x M) sum_d c_d (Minv L_d) u,
// costs 6 n^4 + 18 n^3 FMAs per cell (n = k+1) instead of ~12 n^4 + 100 n^3 for the quadrature-point
// evaluation, with identical results up to round-off (M and L_d are the exact Gauss(k+1) integrals the
// reference evaluates).  The Gauss-Lobatto end nodes make the value trace a plain nodal read.
//
// Mapping to the SM: a CTA owns a batch of B consecutive cells (a 4x4x2 brick in Morton order) staged in
// shared memory; a thread owns one xy-plane of one cell in registers (n^2 values: the x and y sweeps and
// the final M_x, M_y sweeps never leave the register file) and n z-lines for the z sweeps.  1-D matrices
// are kernel parameters = constant-bank operands of the DFMAs.  Neighbour traces inside the batch are
// exchanged through shared memory; traces of cells outside the batch (host-precomputed list) are
// computed once per batch from global memory/L2.  Every DoF of dst is written once, coalesced.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "cart_ws.hpp"
#include "operator.cuh"

namespace exadg_b200
{
namespace
{
template<int N>
struct CartTables
{
  double G[3][N * N]; // c_d Minv (K + own-side face terms, tau_hat_d folded in)
  double P[3][2][N];  // c_d * 1/2 sigma_s * Minv l'(s)          times neighbour end value
  double Q[3][2][N];  // -c_d * Minv e_s                           times (1/2 sigma_s g_nb + tau_hat_d v_nb)
  double M[N * N];
  double fd[2][N];    // l_j'(s)
  double tau_hat[3];
};

struct CartArgs
{
  const int32_t * nb;        // [owned][6]
  const int2 * halo;         // [n_batches][H] (lc<<3|f, neighbour cell)
  const int32_t * halo_cnt;  // [n_batches]
  const int32_t * batches;   // optional list of batch ids
  const double * src; const double * ghost; double * dst;
  int64_t n_owned; int n_items; int H; int add;
  double mass; // Helmholtz / viscous operator on a uniform box: scaling_factor_mass x cell volume, added before the mass sweeps (0: Laplace)
  int ncomp;   // components of the Helmholtz operator (1: scalar).  Vectors are cell-major with component blocks of n^3 values inside a cell; a CTA
               // applies the operator to ONE component of its batch of cells (blockIdx = item * ncomp + component): the batch plan is that of the scalar mesh
};

// ---- TMA (bulk async copy) + mbarrier helpers, sm_90+/sm_100a PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
// global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void * smem_dst, const void * gsrc, uint32_t bytes, uint64_t * bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (plain store or FP64 add-reduction for vmult_add)
__device__ __forceinline__ void tma_store_1d(void * gdst, const void * smem_src, uint32_t bytes, bool add)
{
  if (add) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// cells per batch = consecutive cells of the Morton curve: 4x4x4 bricks (1.5 out-of-batch faces per cell) while 2 CTAs of B n threads fit an SM
template<int N> struct CartCfg { static constexpr int B = (N >= 6) ? 16 : ((N >= 5) ? 32 : 64); };

// BB: cells per batch (default CartCfg; n = 4 also with 32-cell batches: measurement switch EXADG_B200_PLANE_B=32, three CTAs per SM instead of two but
// 2.0 instead of 1.5 out-of-batch faces per cell - measured slower, 91.8 against 99.8 GDoF/s, scripts/r02_shot50.sh)
// COMP: component blocks (Helmholtz operator with ncomp > 1); the scalar instantiation is the code it was before the blocks existed
template<int N, int BB = CartCfg<N>::B, bool COMP = false>
__global__ void __launch_bounds__(BB * N, (N == 5) ? 2 : ((N == 4 && BB == 32) ? 3 : 1)) vmult_cartesian_kernel(const __grid_constant__ CartTables<N> T, const CartArgs A)
{
  constexpr int B = BB, NT = B * N;
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int PS = N2 | 1, CS = N * PS; // odd plane stride: conflict-free plane- and line-wise access
  // n^2 a multiple of 16 (n = 4): the trace arrays [cell][n^2] would put every cell on the same banks (8-way conflicts: ncu of the 64^3 box had
  // 19.0 M conflict wavefronts out of 36.8 M).  Their in-face index is XOR-swizzled with the low bits of the cell that READS the entry (HV / HG)
  // or owns it (GN): the lanes of a warp - 8 cells x 4 planes - then cover all 16 eight-byte banks twice.
  constexpr bool SWZ = (N2 % 16 == 0);
  auto sw = [](int idx, int cell) { return SWZ ? (idx ^ (cell & 15)) : idx; };
  extern __shared__ __align__(128) double smem[];
  double * U = smem;                 // [B][CS]  src values, later the staging buffer of the result
  double * Tt = U + B * CS;          // [B][CS]  partial results
  double * GN = Tt + B * CS;         // [B][2][N2] own end derivatives of the current direction
  double * HV = GN + B * 2 * N2;     // [H][N2] end values of out-of-batch neighbours
  double * HG = HV + (size_t)A.H * N2; // [H][N2] end derivatives of out-of-batch neighbours
  int2 * hlS = reinterpret_cast<int2 *>(HG + (size_t)A.H * N2); // [H] halo list of this batch
  int * nbS = reinterpret_cast<int *>(hlS + A.H); // [B][6]
  int * slotS = nbS + B * 6;         // [B][6]
  uint64_t * bar = reinterpret_cast<uint64_t *>(slotS + B * 6);

  const int t = threadIdx.x, lc = t / N, s = t % N;
  const int ncomp = COMP ? A.ncomp : 1;
  const int bitem = COMP ? (int)blockIdx.x / ncomp : (int)blockIdx.x, comp = COMP ? (int)blockIdx.x % ncomp : 0;
  const int batch = A.batches ? A.batches[bitem] : bitem;
  const int64_t b0 = (int64_t)batch * B;
  // offset of value i of the batch (cell i / n^3 of the batch, component comp) in src / dst
  auto goff = [&](int i) { return COMP ? ((b0 + i / N3) * ncomp + comp) * N3 + i % N3 : b0 * N3 + i; };
  const int nvalid = (int)min((int64_t)B, A.n_owned - b0);
  const bool valid = lc < nvalid;
  // contiguous cell data (odd n: no padding) goes through one TMA bulk copy; 16-byte granularity
  const uint32_t bytes = (uint32_t)(nvalid * N3 * sizeof(double));
  const bool use_tma = (PS == N2) && (bytes % 16 == 0) && !COMP; // (component blocks of a batch are not contiguous)

  // ---- phase L: stage the batch, its neighbour table and the traces of out-of-batch neighbours ----
  if (use_tma) {
    if (t == 0) mbar_init(bar, 1);
    __syncthreads();
    if (t == 0) { mbar_expect_tx(bar, bytes); tma_load_1d(U, A.src + b0 * N3, bytes, bar); }
  }
  const int cnt = A.halo_cnt[batch];
  for (int i = t; i < B * 6; i += NT) { nbS[i] = (i / 6 < nvalid) ? A.nb[b0 * 6 + i] : -1; slotS[i] = -1; }
  for (int i = t; i < cnt; i += NT) hlS[i] = A.halo[(size_t)batch * A.H + i];
  if (!use_tma) {
    constexpr int UNR = 8; // independent loads in flight per thread
    for (int i0 = t; i0 < nvalid * N3; i0 += NT * UNR) {
      double v[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) { const int i = i0 + q * NT; v[q] = (i < nvalid * N3) ? A.src[goff(i)] : 0.0; }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int i = i0 + q * NT;
        if (i < nvalid * N3) { const int c = i / N3, rem = i % N3, k = rem / N2, e = rem % N2; U[c * CS + k * PS + e] = v[q]; }
      }
    }
  }
  __syncthreads();
  {
    constexpr int UNR = 5; // 5 items x n loads in flight per thread
    for (int it0 = t; it0 < cnt * N2; it0 += NT * UNR) {
      double x[UNR][N]; int sp[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int item = it0 + q * NT;
        sp[q] = 0;
        if (item < cnt * N2) {
          const int e = item / N2, ab = item % N2, a = ab % N, b = ab / N;
          const int2 h = hlS[e];
          const int f = h.x & 7, d = f >> 1;
          sp[q] = (f & 1) ^ 1; // neighbour is entered through its face (d, sp)
          const double * un = (h.y < A.n_owned) ? A.src + ((size_t)h.y * ncomp + comp) * N3 : A.ghost + ((size_t)(h.y - A.n_owned) * ncomp + comp) * N3;
          const int sd = (d == 0) ? 1 : (d == 1 ? N : N2);
          const int s1 = (d == 0) ? N : 1, s2 = (d == 2) ? N : N2;
          const double * line = un + a * s1 + b * s2;
#pragma unroll
          for (int i = 0; i < N; ++i) x[q][i] = line[i * sd];
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) x[q][i] = 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int item = it0 + q * NT;
        if (item < cnt * N2) {
          const int e = item / N2, ab = item % N2;
          double g = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) g = fma(sp[q] ? T.fd[1][i] : T.fd[0][i], x[q][i], g);
          const int2 h = hlS[e];
          const int abs_ = sw(ab, h.x >> 3); // swizzled with the reading cell
          HV[e * N2 + abs_] = sp[q] ? x[q][N - 1] : x[q][0];
          HG[e * N2 + abs_] = g;
          if (ab == 0) slotS[(h.x >> 3) * 6 + (h.x & 7)] = e;
        }
      }
    }
  }
  if (use_tma) mbar_wait(bar, 0);
  __syncthreads();

  double u[N][N], acc[N][N];
  if (valid) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = 0; i < N; ++i) { u[j][i] = U[lc * CS + s * PS + i + N * j]; acc[j][i] = A.mass * u[j][i]; }
  }

  // ---- x and y sweeps on the register plane z = s ----
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    if (valid) {
#pragma unroll
      for (int l = 0; l < N; ++l) { // line l: d=0 -> row j=l (runs over i); d=1 -> column i=l (runs over j)
        double g0 = 0.0, g1 = 0.0;
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const double x = (d == 0) ? u[l][m] : u[m][l];
          g0 = fma(T.fd[0][m], x, g0); g1 = fma(T.fd[1][m], x, g1);
        }
        GN[(0 * B + lc) * N2 + sw(s * N + l, lc)] = g0;
        GN[(1 * B + lc) * N2 + sw(s * N + l, lc)] = g1;
      }
    }
    __syncthreads();
    if (valid) {
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int f = 2 * d + side;
        const int nbl = nbS[lc * 6 + f] - (int)b0;
        const bool inb = (nbl >= 0 && nbl < nvalid);
        const int slot = slotS[lc * 6 + f];
        const double hs = side ? 0.5 : -0.5; // 1/2 sigma_s
#pragma unroll
        for (int l = 0; l < N; ++l) {
          double vn, gn;
          if (inb) {
            const int endn = side ? 0 : N - 1; // neighbour's end node facing us
            vn = (d == 0) ? U[nbl * CS + s * PS + endn + N * l] : U[nbl * CS + s * PS + l + N * endn];
            gn = GN[((side ^ 1) * B + nbl) * N2 + sw(s * N + l, nbl)];
          } else {
            vn = HV[slot * N2 + sw(l + N * s, lc)]; gn = HG[slot * N2 + sw(l + N * s, lc)];
          }
          const double tt = fma(hs, gn, T.tau_hat[d] * vn);
#pragma unroll
          for (int m = 0; m < N; ++m) {
            double & y = (d == 0) ? acc[l][m] : acc[m][l];
            y = fma(T.P[d][side][m], vn, y);
            y = fma(T.Q[d][side][m], tt, y);
          }
        }
      }
#pragma unroll
      for (int l = 0; l < N; ++l)
#pragma unroll
        for (int r = 0; r < N; ++r) {
          double y = (d == 0) ? acc[l][r] : acc[r][l];
#pragma unroll
          for (int c = 0; c < N; ++c) y = fma(T.G[d][r * N + c], (d == 0) ? u[l][c] : u[c][l], y);
          if (d == 0) acc[l][r] = y; else acc[r][l] = y;
        }
    }
    __syncthreads(); // GN is reused by the next direction
  }
  if (valid) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = 0; i < N; ++i) Tt[lc * CS + s * PS + i + N * j] = acc[j][i];
  }

  // ---- z sweep: thread (cell lz = t % B, slice sz = t / B) owns the n lines (i, j = sz).  The thread-to-cell map
  // differs from the plane sweeps on purpose: consecutive lanes touch consecutive cells (stride n^3, odd), which
  // makes every shared-memory access of this phase bank-conflict free; nothing is carried over in registers.
  constexpr bool REMAP_Z = (N != 4); // measured: n = 4 (cell stride 68 doubles) is faster with the plane map
  const int lz = REMAP_Z ? t % B : lc, sz = REMAP_Z ? t / B : s;
  const bool validz = (sz < N) && (lz < nvalid);
  // n = 4 keeps the plane map (lane = 4 lz + sz); with the slice along y its addresses 68 lz + 4 sz fall on four banks (8-way conflicts), with
  // the slice along x (the thread owns the lines (x = sz, y = i)) they are 68 lz + sz: the conflict-free pattern of the plane sweeps
  constexpr bool ZSWAP = (N == 4);
  auto zo = [sz](int i) { return ZSWAP ? sz + N * i : i + N * sz; }; // in-plane offset = in-face index of line i of this thread
  if (validz) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double x = U[lz * CS + k * PS + zo(i)];
        u[i][k] = x; // reuse the plane registers: u[i][k] = value of line i at height k
        g0 = fma(T.fd[0][k], x, g0); g1 = fma(T.fd[1][k], x, g1);
      }
      GN[(0 * B + lz) * N2 + sw(zo(i), lz)] = g0;
      GN[(1 * B + lz) * N2 + sw(zo(i), lz)] = g1;
    }
  }
  __syncthreads(); // Tt planes and z traces visible
  if (validz) {
    int nbl[2], slot[2]; bool inb[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      nbl[side] = nbS[lz * 6 + 4 + side] - (int)b0;
      inb[side] = (nbl[side] >= 0 && nbl[side] < nvalid);
      slot[side] = slotS[lz * 6 + 4 + side];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double w[N];
#pragma unroll
      for (int k = 0; k < N; ++k) w[k] = Tt[lz * CS + k * PS + zo(i)];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        double vn, gn;
        if (inb[side]) {
          const int endn = side ? 0 : N - 1;
          vn = U[nbl[side] * CS + endn * PS + zo(i)];
          gn = GN[((side ^ 1) * B + nbl[side]) * N2 + sw(zo(i), nbl[side])];
        } else {
          vn = HV[slot[side] * N2 + sw(zo(i), lz)]; gn = HG[slot[side] * N2 + sw(zo(i), lz)];
        }
        const double tt = fma(side ? 0.5 : -0.5, gn, T.tau_hat[2] * vn);
#pragma unroll
        for (int k = 0; k < N; ++k) { w[k] = fma(T.P[2][side][k], vn, w[k]); w[k] = fma(T.Q[2][side][k], tt, w[k]); }
      }
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) w[r] = fma(T.G[2][r * N + c], u[i][c], w[r]);
      // mass matrix along z, in place (this thread owns the whole line)
#pragma unroll
      for (int r = 0; r < N; ++r) {
        double y = 0.0;
#pragma unroll
        for (int c = 0; c < N; ++c) y = fma(T.M[r * N + c], w[c], y);
        Tt[lz * CS + r * PS + zo(i)] = y;
      }
    }
  }
  __syncthreads();
  // ---- mass matrices along x and y on the register plane, staged into U ----
  if (valid) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = 0; i < N; ++i) u[j][i] = Tt[lc * CS + s * PS + i + N * j];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int r = 0; r < N; ++r) {
        double y = 0.0;
#pragma unroll
        for (int c = 0; c < N; ++c) y = fma(T.M[r * N + c], u[j][c], y);
        acc[j][r] = y;
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int r = 0; r < N; ++r) {
        double y = 0.0;
#pragma unroll
        for (int c = 0; c < N; ++c) y = fma(T.M[r * N + c], acc[c][i], y);
        U[lc * CS + s * PS + i + N * r] = y;
      }
  }
  if (use_tma) {
    // result batch is contiguous in dst: one TMA bulk store (or FP64 add-reduction for vmult_add)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) tma_store_1d(A.dst + b0 * N3, U, bytes, A.add != 0);
    return;
  }
  __syncthreads();
  // ---- coalesced store ----
  for (int i = t; i < nvalid * N3; i += NT) {
    const int c = i / N3, rem = i % N3, k = rem / N2, e = rem % N2;
    const double v = U[c * CS + k * PS + e];
    if (A.add) A.dst[goff(i)] += v; else A.dst[goff(i)] = v;
  }
}


// =====================================================================================================
// Line variant (used for n = 6, 7, 8; k = 5, 6, 7).  A plane of a cell in registers costs 2 n^2 doubles: 192 / 242 / 255 registers
// (+ 680 bytes of spills) for n = 6 / 7 / 8, which left 3 - 9 warps per SM to the plane kernel above (round 1: 14 / 13 / 8 % of
// the HBM roofline).  Here a thread owns ONE line of a cell per sweep (2 n doubles in registers), a CTA of B n^2 threads a batch
// of B cells, every sweep goes through shared memory:
//     T = sum_d K_d u (line by line, face terms from the traces of the neighbour lines)  ->  T <- M_z M_y M_x T
// 14 shared-memory accesses per DoF against 6 n + 18 FMAs per DoF - at n >= 6 a better ratio than the plane kernel has at n = 5 -
// and 18 - 32 resident warps per SM.  Rows are padded to an odd stride so that consecutive lanes (consecutive lines) never share a
// bank in any of the three sweep directions.
// =====================================================================================================
// BB cells per CTA: 16 (two octets) or 8 (one octet: half the shared memory and threads per CTA, so two to four CTAs share an SM and
// the load / trace / sweep / store phases of different batches overlap, at the price of 3.0 instead of 2.5 out-of-batch faces per cell)
// MINB: resident CTAs the register budget is tuned for.  n = 6 with 16 cells: 576 threads x 92 registers left ONE CTA per SM (ncu: occupancy limited by
// registers, 18 warps, load / trace / sweep / store phases strictly serial); 56 registers admit two (shared memory: 2 x 98 KB).  8-cell batches:
// 4 / 3 / 2 CTAs for n = 6 / 7 / 8 (56 / 48 / 60 registers).
// Shared-memory layouts with fewer bank conflicts in the y / z sweeps (enumerated by scripts/line_layout_search.py: n = 8 rows of 8 rotated by y,
// planes of 68: no conflicts at all; n = 7 planes of 51, lanes along the second coordinate) were measured SLOWER (scripts/r02_shot48.sh: k = 6
// 84.4 -> 76.8, k = 7 89.3 -> 84.8 GDoF/s): the per-node address arithmetic costs more than the 1.3 - 1.5 x wavefronts of those sweeps.
template<int N, int BB> struct LineCfg { static constexpr int B = BB; static constexpr int NT = B * N * N; static constexpr int RS = N | 1; static constexpr int CS = RS * N * N;
                                         static constexpr int MINB = (BB == 16) ? (N == 6 ? 2 : 1) : (N == 6 ? 4 : (N == 7 ? 3 : 2)); };

template<int N, int BB, bool COMP = false>
__global__ void __launch_bounds__(LineCfg<N, BB>::NT, LineCfg<N, BB>::MINB) vmult_cartesian_line_kernel(const __grid_constant__ CartTables<N> T, const CartArgs A)
{
  constexpr int B = LineCfg<N, BB>::B, NT = LineCfg<N, BB>::NT, RS = LineCfg<N, BB>::RS, CS = LineCfg<N, BB>::CS;
  constexpr int N2 = N * N, N3 = N2 * N;
  extern __shared__ __align__(128) double smem[];
  double * U = smem;                   // [B][N][N][RS] src values (x fastest, rows padded to RS)
  double * Tt = U + B * CS;            // [B][N][N][RS] accumulated result
  double * GN = Tt + B * CS;           // [2][B][N2] end derivatives of the batch's own lines of the current direction
  double * HV = GN + 2 * B * N2;       // [H][N2] end values of out-of-batch neighbours
  double * HG = HV + (size_t)A.H * N2; // [H][N2] end derivatives of out-of-batch neighbours
  int2 * hlS = reinterpret_cast<int2 *>(HG + (size_t)A.H * N2); // [H]
  int * nbS = reinterpret_cast<int *>(hlS + A.H);               // [B][6]
  int * slotS = nbS + B * 6;                                     // [B][6]

  const int t = threadIdx.x, c = t / N2, ab = t % N2, a = ab % N, b = ab / N;
  const int ncomp = COMP ? A.ncomp : 1;
  const int bitem = COMP ? (int)blockIdx.x / ncomp : (int)blockIdx.x, comp = COMP ? (int)blockIdx.x % ncomp : 0;
  const int batch = A.batches ? A.batches[bitem] : bitem;
  const int64_t b0 = (int64_t)batch * B;
  // offset of value i of the batch (cell i / n^3 of the batch, component comp) in src / dst
  auto goff = [&](int i) { return COMP ? ((b0 + i / N3) * ncomp + comp) * N3 + i % N3 : b0 * N3 + i; };
  const int nvalid = (int)min((int64_t)B, A.n_owned - b0);
  const bool valid = c < nvalid;

  // ---- stage the batch (coalesced loads, 8 in flight per thread), its neighbour table and the halo list ----
  const int cnt = (A.add & 2) ? 0 : A.halo_cnt[batch]; // bit 1: timing experiment without the out-of-batch traces (results wrong)
  for (int i = t; i < B * 6; i += NT) { nbS[i] = (i / 6 < nvalid) ? A.nb[b0 * 6 + i] : -1; slotS[i] = -1; }
  for (int i = t; i < cnt; i += NT) hlS[i] = A.halo[(size_t)batch * A.H + i];
  {
    constexpr int UNR = 8;
    for (int i0 = t; i0 < nvalid * N3; i0 += NT * UNR) {
      double v[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) { const int i = i0 + q * NT; v[q] = (i < nvalid * N3) ? A.src[goff(i)] : 0.0; }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int i = i0 + q * NT;
        if (i < nvalid * N3) { const int cc = i / N3, rem = i % N3; U[cc * CS + (rem / N) * RS + rem % N] = v[q]; }
      }
    }
  }
  __syncthreads();
  // ---- traces of the out-of-batch neighbours: one line per item, UNR items x n loads in flight per thread (register budget) ----
  {
    constexpr int UNR = (N >= 8 || LineCfg<N, BB>::MINB > 1) ? 2 : (N == 7 ? 3 : 4);
    for (int it0 = t; it0 < cnt * N2; it0 += NT * UNR) {
      double x[UNR][N]; int sp[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int item = it0 + q * NT;
        sp[q] = 0;
        if (item < cnt * N2) {
          const int e = item / N2, lab = item % N2, la = lab % N, lb = lab / N;
          const int2 h = hlS[e];
          const int f = h.x & 7, d = f >> 1;
          sp[q] = (f & 1) ^ 1; // the neighbour is entered through its face (d, sp)
          const double * un = (h.y < A.n_owned) ? A.src + ((size_t)h.y * ncomp + comp) * N3 : A.ghost + ((size_t)(h.y - A.n_owned) * ncomp + comp) * N3;
          const int sd = (d == 0) ? 1 : (d == 1 ? N : N2);
          const int s1 = (d == 0) ? N : 1, s2 = (d == 2) ? N : N2;
          const double * line = un + la * s1 + lb * s2;
#pragma unroll
          for (int i = 0; i < N; ++i) x[q][i] = line[i * sd];
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) x[q][i] = 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int item = it0 + q * NT;
        if (item < cnt * N2) {
          const int e = item / N2, lab = item % N2;
          double g = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) g = fma(sp[q] ? T.fd[1][i] : T.fd[0][i], x[q][i], g);
          HV[e * N2 + lab] = sp[q] ? x[q][N - 1] : x[q][0];
          HG[e * N2 + lab] = g;
          if (lab == 0) { const int2 h = hlS[e]; slotS[(h.x >> 3) * 6 + (h.x & 7)] = e; }
        }
      }
    }
  }
  __syncthreads();

  // ---- three sweeps: thread (c, ab) owns line ab of cell c in direction d ----
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // offset of the line's first node and stride along it in the padded cell
    const int base = c * CS + ((d == 0) ? (a + N * b) * RS : (d == 1 ? a + b * N * RS : a + b * RS));
    const int sd = (d == 0) ? 1 : (d == 1 ? RS : N * RS);
    double x[N], y[N];
    if (valid) {
      double g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int m = 0; m < N; ++m) { x[m] = U[base + m * sd]; g0 = fma(T.fd[0][m], x[m], g0); g1 = fma(T.fd[1][m], x[m], g1); }
      GN[(0 * B + c) * N2 + ab] = g0;
      GN[(1 * B + c) * N2 + ab] = g1;
#pragma unroll
      for (int r = 0; r < N; ++r) {
        double v = (d == 0) ? A.mass * x[r] : 0.0; // mass term of the Helmholtz operator (in front of the mass sweeps like the rest)
#pragma unroll
        for (int m = 0; m < N; ++m) v = fma(T.G[d][r * N + m], x[m], v);
        y[r] = v;
      }
    }
    __syncthreads(); // end derivatives of this direction visible
    if (valid) {
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int f = 2 * d + side;
        const int nbl = nbS[c * 6 + f] - (int)b0;
        const bool inb = (nbl >= 0 && nbl < nvalid);
        double vn, gn;
        if (inb) {
          const int endn = side ? 0 : N - 1; // the neighbour's end node facing us
          vn = U[base - c * CS + nbl * CS + endn * sd];
          gn = GN[((side ^ 1) * B + nbl) * N2 + ab];
        } else {
          const int slot = slotS[c * 6 + f];
          vn = HV[slot * N2 + ab]; gn = HG[slot * N2 + ab];
        }
        const double tt = fma(side ? 0.5 : -0.5, gn, T.tau_hat[d] * vn);
#pragma unroll
        for (int m = 0; m < N; ++m) { y[m] = fma(T.P[d][side][m], vn, y[m]); y[m] = fma(T.Q[d][side][m], tt, y[m]); }
      }
      if (d == 0) {
#pragma unroll
        for (int m = 0; m < N; ++m) Tt[base + m * sd] = y[m];
      } else if (d == 1) {
#pragma unroll
        for (int m = 0; m < N; ++m) Tt[base + m * sd] += y[m];
      } else {
        // last direction: complete the sum and apply the mass matrix along z in place (this thread owns the whole line)
#pragma unroll
        for (int m = 0; m < N; ++m) x[m] = Tt[base + m * sd] + y[m];
#pragma unroll
        for (int r = 0; r < N; ++r) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) v = fma(T.M[r * N + m], x[m], v);
          Tt[base + r * sd] = v;
        }
      }
    }
    __syncthreads(); // GN is reused by the next direction; the lines of the next direction cross these
  }
  // ---- mass matrices along x, then y (in place, line by line) ----
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int base = c * CS + ((d == 0) ? (a + N * b) * RS : a + b * N * RS);
    const int sd = (d == 0) ? 1 : RS;
    if (valid) {
      double x[N];
#pragma unroll
      for (int m = 0; m < N; ++m) x[m] = Tt[base + m * sd];
#pragma unroll
      for (int r = 0; r < N; ++r) {
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < N; ++m) v = fma(T.M[r * N + m], x[m], v);
        Tt[base + r * sd] = v;
      }
    }
    __syncthreads();
  }
  // ---- coalesced store ----
  for (int i = t; i < nvalid * N3; i += NT) {
    const int cc = i / N3, rem = i % N3;
    const double v = Tt[cc * CS + (rem / N) * RS + rem % N];
    if (A.add & 1) A.dst[goff(i)] += v; else A.dst[goff(i)] = v;
  }
}

// =====================================================================================================
// Pipelined variant (used for n = 5): persistent CTAs of 4 warps on 24-cell batches.
//  * 24 cells x 5 planes = 120 plane threads = 4 warps: one warp per SM sub-partition, so the FP64 pipes of the
//    four sub-partitions carry equal work (5-warp CTAs load them 2:1:1:1 and stall at every barrier);
//  * 3 CTAs per SM (<= 75 KB shared memory, <= 168 registers) instead of 2;
//  * the lines of out-of-batch neighbours are fetched one sweep ahead into registers (x lines at the end of the
//    previous batch, y lines before the x sweep, z lines before the y sweep) and reduced to traces into one
//    shared buffer that is reused by the three directions: their latency hides behind the FP64 work;
//  * the bulk copy (TMA) of the next batch starts as soon as the z sweep has consumed the current one; the result
//    is finished in place in Tt and leaves through one TMA bulk store.
// =====================================================================================================
// n = 5: 24 cells (3 octets of the Morton curve) x 5 planes = 120 of 128 threads; n = 3: 40 cells x 3 planes = 120 of 128 threads.
// E = neighbour cells per thread group and direction fetched ahead; MINB = resident CTAs the register budget is tuned for.
template<int N> struct PipeCfg { static constexpr int B = (N == 3) ? 40 : 24; static constexpr int NT = 128; static constexpr int E = (N == 3) ? 3 : 5; static constexpr int MINB = (N == 3) ? 4 : 2; };

struct PipeArgs
{
  const int32_t * nb;        // [owned][6]
  const int2 * halo;         // [n_batches][HL] (lc<<3|f, neighbour cell), sorted by direction
  const int4 * halo_cnt;     // [n_batches] numbers of x-, y-, z-face entries
  const int32_t * batches;   // optional list of batch ids
  const double * src; const double * ghost; double * dst;
  int64_t n_owned; int n_items; int HL; int HD; int add;
};

// lines (direction D) of the neighbour cells of entries e0 + grp + q * n_grp (q < E) -> registers
template<int N, int D, int E>
__device__ __forceinline__ void pipe_load(double (&x)[E][N], const int2 * hl, int e0, int e1, int grp, int n_grp, int ab, const double * src, const double * ghost, int64_t n_owned)
{
  constexpr int N2 = N * N, N3 = N2 * N;
  constexpr int sd = (D == 0) ? 1 : (D == 1 ? N : N2);
  constexpr int s1 = (D == 0) ? N : 1, s2 = (D == 2) ? N : N2;
  const int off = (ab % N) * s1 + (ab / N) * s2;
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int e = e0 + grp + q * n_grp;
    if (e < e1) {
      const int2 h = hl[e];
      const double * line = ((h.y < n_owned) ? src + (size_t)h.y * N3 : ghost + (size_t)(h.y - n_owned) * N3) + off;
#pragma unroll
      for (int i = 0; i < N; ++i) x[q][i] = line[i * sd];
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) x[q][i] = 0.0; // defined on every path: the registers are dead until the next load
    }
  }
}
// registers -> end value / end derivative in the trace buffer (slot = index within the direction's entry list)
template<int N, int E, typename Tab>
__device__ __forceinline__ void pipe_reduce(const Tab & T, const double (&x)[E][N], const int2 * hl, int e0, int e1, int grp, int n_grp, int ab, double * HV, double * HG)
{
  constexpr int N2 = N * N;
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int e = e0 + grp + q * n_grp;
    if (e < e1) {
      const int2 h = hl[e];
      double g0 = T.fd[0][0] * x[q][0], g1 = T.fd[1][0] * x[q][0];
#pragma unroll
      for (int i = 1; i < N; ++i) { g0 = fma(T.fd[0][i], x[q][i], g0); g1 = fma(T.fd[1][i], x[q][i], g1); }
      const bool sp = !(h.x & 1); // the neighbour is entered through the side opposite to ours
      HV[(e - e0) * N2 + ab] = sp ? x[q][N - 1] : x[q][0];
      HG[(e - e0) * N2 + ab] = sp ? g1 : g0;
    }
  }
}
// entries beyond the pipelined depth (irregular batches only): fetched and reduced on the spot
template<int N, int D, typename Tab>
__device__ __forceinline__ void pipe_rest(const Tab & T, const int2 * hl, int e0, int e_from, int e1, int grp, int n_grp, int ab, const double * src, const double * ghost,
                                          int64_t n_owned, double * HV, double * HG)
{
  for (int eb = e_from; eb < e1; eb += n_grp) {
    double x[1][N];
    pipe_load<N, D, 1>(x, hl, eb, e1, grp, n_grp, ab, src, ghost, n_owned);
    const int e = eb + grp;
    if (e < e1) {
      constexpr int N2 = N * N;
      const int2 h = hl[e];
      double g0 = T.fd[0][0] * x[0][0], g1 = T.fd[1][0] * x[0][0];
#pragma unroll
      for (int i = 1; i < N; ++i) { g0 = fma(T.fd[0][i], x[0][i], g0); g1 = fma(T.fd[1][i], x[0][i], g1); }
      const bool sp = !(h.x & 1);
      HV[(e - e0) * N2 + ab] = sp ? x[0][N - 1] : x[0][0];
      HG[(e - e0) * N2 + ab] = sp ? g1 : g0;
    }
  }
}

#define PIPE_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory")

template<int N>
__global__ void __launch_bounds__(PipeCfg<N>::NT, PipeCfg<N>::MINB) vmult_cartesian_pipe_kernel(const __grid_constant__ CartTables<N> T, const PipeArgs A)
{
  constexpr int B = PipeCfg<N>::B, NT = PipeCfg<N>::NT, E = PipeCfg<N>::E;
  constexpr int N2 = N * N, N3 = N2 * N;
  static_assert((N2 & 1) == 1, "pipelined kernel: odd n (contiguous cells, TMA)");
  constexpr int NGRP = NT / N2;
  extern __shared__ __align__(128) double smem[];
  double * U = smem;                   // [B][N3] src values of the batch (TMA destination)
  double * Tt = U + B * N3;            // [B][N3] partial results, finally the result (TMA source)
  double * GN = Tt + B * N3;           // [B][2][N2] own end derivatives of the current direction
  double * HV = GN + B * 2 * N2;       // [HD][N2] end values of out-of-batch neighbours, current direction
  double * HG = HV + (size_t)A.HD * N2; // [HD][N2] end derivatives
  int2 * hl2 = reinterpret_cast<int2 *>(HG + (size_t)A.HD * N2); // [2][HL] halo lists (current / next batch)
  int * nb2 = reinterpret_cast<int *>(hl2 + 2 * A.HL);           // [2][B][6] neighbour tables
  int * slotS = nb2 + 2 * B * 6;                                 // [B][6] index of the face in its direction's entry list
  int4 * cntS = reinterpret_cast<int4 *>((reinterpret_cast<uintptr_t>(slotS + B * 6) + 15) & ~uintptr_t(15)); // [2]
  uint64_t * bar = reinterpret_cast<uint64_t *>(cntS + 2);

  const int t = threadIdx.x, lc = t / N, s = t % N;
  const int grp = t / N2, ab = t % N2;
  const bool hthread = grp < NGRP;
  uint32_t upar = 0;
  if (t == 0) mbar_init(bar, 1);
  auto batch_of = [&](int it) { return A.batches ? A.batches[it] : it; };
  double hx[E][N]; // lines of out-of-batch neighbours, one sweep ahead

  // ---- prologue: tables, bulk copy and x lines of the first batch ----
  int cur = 0;
  if ((int)blockIdx.x < A.n_items) {
    const int bt = batch_of(blockIdx.x);
    const int64_t c0 = (int64_t)bt * B;
    const int nv = (int)min((int64_t)B, A.n_owned - c0);
    const int4 hc0 = A.halo_cnt[bt];
    if (t == 0) cntS[0] = hc0;
    for (int i = t; i < B * 6; i += NT) nb2[i] = (i / 6 < nv) ? A.nb[c0 * 6 + i] : -1;
    for (int i = t; i < hc0.x + hc0.y + hc0.z; i += NT) hl2[i] = A.halo[(size_t)bt * A.HL + i];
  }
  PIPE_SYNC();
  if ((int)blockIdx.x < A.n_items) {
    const int bt = batch_of(blockIdx.x);
    const int64_t c0 = (int64_t)bt * B;
    const uint32_t by = (uint32_t)((int)min((int64_t)B, A.n_owned - c0) * N3 * sizeof(double));
    if (t == 0 && by % 16 == 0) { mbar_expect_tx(bar, by); tma_load_1d(U, A.src + c0 * N3, by, bar); }
    const int4 c = cntS[0];
    if (hthread) pipe_load<N, 0, E>(hx, hl2, 0, c.x, grp, NGRP, ab, A.src, A.ghost, A.n_owned);
  }

  for (int it = blockIdx.x; it < A.n_items; it += gridDim.x, cur ^= 1) {
    const int batch = batch_of(it);
    const int64_t b0 = (int64_t)batch * B;
    const int nvalid = (int)min((int64_t)B, A.n_owned - b0);
    const bool valid = lc < nvalid;
    const int2 * hl = hl2 + cur * A.HL;
    const int * nbS = nb2 + cur * B * 6;
    const uint32_t bytes = (uint32_t)(nvalid * N3 * sizeof(double));
    const bool use_tma = (bytes % 16 == 0);
    const int4 hc = cntS[cur];
    const int ex = hc.x, ey = hc.x + hc.y, ez = hc.x + hc.y + hc.z;

    // slot table: position of every out-of-batch face in its direction's list
    for (int e = t; e < ez; e += NT) { const int2 h = hl[e]; const int d = (h.x & 7) >> 1; slotS[(h.x >> 3) * 6 + (h.x & 7)] = e - (d == 0 ? 0 : (d == 1 ? ex : ey)); }
    if (!use_tma) for (int i = t; i < nvalid * N3; i += NT) U[i] = A.src[b0 * N3 + i]; // ragged last batch
    // tables of the next batch -> registers
    const int itn = it + gridDim.x;
    const bool has_next = itn < A.n_items;
    int pre_nb[(B * 6 + NT - 1) / NT]; int2 pre_hl = make_int2(0, 0); int4 pre_cnt = make_int4(0, 0, 0, 0);
    if (has_next) {
      const int bn = batch_of(itn);
      const int64_t c0 = (int64_t)bn * B;
      const int nv = (int)min((int64_t)B, A.n_owned - c0);
#pragma unroll
      for (int q = 0; q < (B * 6 + NT - 1) / NT; ++q) { const int i = t + q * NT; pre_nb[q] = (i < B * 6 && i / 6 < nv) ? A.nb[c0 * 6 + i] : -1; }
      pre_cnt = A.halo_cnt[bn];
      if (t < A.HL) pre_hl = A.halo[(size_t)bn * A.HL + t];
    }
    // x traces (lines were fetched during the previous batch), y lines on their way
    if (hthread) {
      pipe_reduce<N, E>(T, hx, hl, 0, ex, grp, NGRP, ab, HV, HG);
      if (ex > E * NGRP) pipe_rest<N, 0>(T, hl, 0, E * NGRP, ex, grp, NGRP, ab, A.src, A.ghost, A.n_owned, HV, HG);
      pipe_load<N, 1, E>(hx, hl, ex, ey, grp, NGRP, ab, A.src, A.ghost, A.n_owned);
    }
    if (use_tma) { mbar_wait(bar, upar); upar ^= 1; }
    PIPE_SYNC();

    double u[N][N], acc[N][N];
    if (valid) {
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) { u[j][i] = U[lc * N3 + s * N2 + i + N * j]; acc[j][i] = 0.0; }
    }
    // ---- x and y sweeps on the register plane z = s ----
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      if (valid) {
        double g0[N], g1[N];
#pragma unroll
        for (int l = 0; l < N; ++l) { const double x = (d == 0) ? u[l][0] : u[0][l]; g0[l] = T.fd[0][0] * x; g1[l] = T.fd[1][0] * x; }
#pragma unroll
        for (int m = 1; m < N; ++m)
#pragma unroll
          for (int l = 0; l < N; ++l) {
            const double x = (d == 0) ? u[l][m] : u[m][l];
            g0[l] = fma(T.fd[0][m], x, g0[l]); g1[l] = fma(T.fd[1][m], x, g1[l]);
          }
#pragma unroll
        for (int l = 0; l < N; ++l) { GN[(0 * B + lc) * N2 + s * N + l] = g0[l]; GN[(1 * B + lc) * N2 + s * N + l] = g1[l]; }
      }
      PIPE_SYNC();
      if (valid) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int f = 2 * d + side;
          const int nbl = nbS[lc * 6 + f] - (int)b0;
          const bool inb = (nbl >= 0 && nbl < nvalid);
          const int slot = slotS[lc * 6 + f];
          const double hs = side ? 0.5 : -0.5; // 1/2 sigma_s
          double vn[N], tt[N];
#pragma unroll
          for (int l = 0; l < N; ++l) {
            double gn;
            if (inb) {
              const int endn = side ? 0 : N - 1; // neighbour's end node facing us
              vn[l] = (d == 0) ? U[nbl * N3 + s * N2 + endn + N * l] : U[nbl * N3 + s * N2 + l + N * endn];
              gn = GN[((side ^ 1) * B + nbl) * N2 + s * N + l];
            } else {
              vn[l] = HV[slot * N2 + l + N * s]; gn = HG[slot * N2 + l + N * s];
            }
            tt[l] = fma(hs, gn, T.tau_hat[d] * vn[l]);
          }
#pragma unroll
          for (int m = 0; m < N; ++m)
#pragma unroll
            for (int l = 0; l < N; ++l) {
              if (d == 0) acc[l][m] = fma(T.P[d][side][m], vn[l], acc[l][m]); else acc[m][l] = fma(T.P[d][side][m], vn[l], acc[m][l]);
            }
#pragma unroll
          for (int m = 0; m < N; ++m)
#pragma unroll
            for (int l = 0; l < N; ++l) {
              if (d == 0) acc[l][m] = fma(T.Q[d][side][m], tt[l], acc[l][m]); else acc[m][l] = fma(T.Q[d][side][m], tt[l], acc[m][l]);
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
          for (int l = 0; l < N; ++l)
#pragma unroll
            for (int r = 0; r < N; ++r) {
              if (d == 0) acc[l][r] = fma(T.G[d][r * N + c], u[l][c], acc[l][r]);
              else acc[r][l] = fma(T.G[d][r * N + c], u[c][l], acc[r][l]);
            }
      }
      if (d == 1 && t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // previous batch's store has read Tt
      PIPE_SYNC(); // traces and GN of this direction are consumed
      // traces of the next direction from the lines fetched one sweep ago; lines of the direction after that
      if (hthread) {
        if (d == 0) {
          pipe_reduce<N, E>(T, hx, hl, ex, ey, grp, NGRP, ab, HV, HG);
          if (ey - ex > E * NGRP) pipe_rest<N, 1>(T, hl, ex, ex + E * NGRP, ey, grp, NGRP, ab, A.src, A.ghost, A.n_owned, HV, HG);
          pipe_load<N, 2, E>(hx, hl, ey, ez, grp, NGRP, ab, A.src, A.ghost, A.n_owned);
        } else {
          pipe_reduce<N, E>(T, hx, hl, ey, ez, grp, NGRP, ab, HV, HG);
          if (ez - ey > E * NGRP) pipe_rest<N, 2>(T, hl, ey, ey + E * NGRP, ez, grp, NGRP, ab, A.src, A.ghost, A.n_owned, HV, HG);
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) Tt[lc * N3 + s * N2 + i + N * j] = acc[j][i];
    }
    // ---- z sweep: thread (cell lz = t % B, slice sz = t / B) owns the lines (i, j = sz): consecutive lanes touch
    // consecutive cells (stride n^3, odd) -> bank-conflict free; nothing is carried over in registers ----
    const int lz = t % B, sz = t / B;
    const bool validz = (sz < N) && (lz < nvalid);
    if (validz) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < N; ++k) u[i][k] = U[lz * N3 + k * N2 + i + N * sz];
      double g0[N], g1[N];
#pragma unroll
      for (int i = 0; i < N; ++i) { g0[i] = T.fd[0][0] * u[i][0]; g1[i] = T.fd[1][0] * u[i][0]; }
#pragma unroll
      for (int k = 1; k < N; ++k)
#pragma unroll
        for (int i = 0; i < N; ++i) { g0[i] = fma(T.fd[0][k], u[i][k], g0[i]); g1[i] = fma(T.fd[1][k], u[i][k], g1[i]); }
#pragma unroll
      for (int i = 0; i < N; ++i) { GN[(0 * B + lz) * N2 + sz * N + i] = g0[i]; GN[(1 * B + lz) * N2 + sz * N + i] = g1[i]; }
    }
    PIPE_SYNC(); // Tt planes, z traces (GN and HV/HG) visible
    if (validz) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < N; ++k) acc[i][k] = Tt[lz * N3 + k * N2 + i + N * sz];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int nbl = nbS[lz * 6 + 4 + side] - (int)b0;
        const bool inb = (nbl >= 0 && nbl < nvalid);
        const int slot = slotS[lz * 6 + 4 + side];
        double vn[N], tt[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double gn;
          if (inb) {
            const int endn = side ? 0 : N - 1;
            vn[i] = U[nbl * N3 + endn * N2 + i + N * sz];
            gn = GN[((side ^ 1) * B + nbl) * N2 + sz * N + i];
          } else {
            vn[i] = HV[slot * N2 + i + N * sz]; gn = HG[slot * N2 + i + N * sz];
          }
          tt[i] = fma(side ? 0.5 : -0.5, gn, T.tau_hat[2] * vn[i]);
        }
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int i = 0; i < N; ++i) acc[i][k] = fma(T.P[2][side][k], vn[i], acc[i][k]);
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int i = 0; i < N; ++i) acc[i][k] = fma(T.Q[2][side][k], tt[i], acc[i][k]);
      }
#pragma unroll
      for (int c = 0; c < N; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int r = 0; r < N; ++r) acc[i][r] = fma(T.G[2][r * N + c], u[i][c], acc[i][r]);
      // mass matrix along z: u <- M acc
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int r = 0; r < N; ++r) u[i][r] = T.M[r * N] * acc[i][0];
#pragma unroll
      for (int c = 1; c < N; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int r = 0; r < N; ++r) u[i][r] = fma(T.M[r * N + c], acc[i][c], u[i][r]);
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int r = 0; r < N; ++r) Tt[lz * N3 + r * N2 + i + N * sz] = u[i][r];
    }
    PIPE_SYNC(); // U and the trace buffers are dead from here on
    // park the tables of the next batch and start its bulk copy
    if (has_next) {
      int * nbN = nb2 + (cur ^ 1) * B * 6;
#pragma unroll
      for (int q = 0; q < (B * 6 + NT - 1) / NT; ++q) { const int i = t + q * NT; if (i < B * 6) nbN[i] = pre_nb[q]; }
      if (t < A.HL) hl2[(cur ^ 1) * A.HL + t] = pre_hl;
      if (t == 0) {
        cntS[cur ^ 1] = pre_cnt;
        const int64_t c0 = (int64_t)batch_of(itn) * B;
        const uint32_t by = (uint32_t)((int)min((int64_t)B, A.n_owned - c0) * N3 * sizeof(double));
        if (by % 16 == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(bar, by); tma_load_1d(U, A.src + c0 * N3, by, bar);
        }
      }
    }
    // ---- mass matrices along x and y on the register plane, in place in Tt ----
    if (valid) {
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) u[j][i] = Tt[lc * N3 + s * N2 + i + N * j];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int r = 0; r < N; ++r) acc[j][r] = T.M[r * N] * u[j][0];
#pragma unroll
      for (int c = 1; c < N; ++c)
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
          for (int r = 0; r < N; ++r) acc[j][r] = fma(T.M[r * N + c], u[j][c], acc[j][r]);
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) u[r][i] = T.M[r * N] * acc[0][i];
#pragma unroll
      for (int c = 1; c < N; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
          for (int i = 0; i < N; ++i) u[r][i] = fma(T.M[r * N + c], acc[c][i], u[r][i]);
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) Tt[lc * N3 + s * N2 + i + N * r] = u[r][i];
    }
    if (use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    PIPE_SYNC(); // result complete, parked tables visible
    // x lines of the next batch -> registers (consumed at the top of the next iteration)
    {
      const int4 cn = has_next ? cntS[cur ^ 1] : make_int4(0, 0, 0, 0);
      if (hthread) pipe_load<N, 0, E>(hx, hl2 + (cur ^ 1) * A.HL, 0, cn.x, grp, NGRP, ab, A.src, A.ghost, A.n_owned);
    }
    if (use_tma) {
      if (t == 0) { // issue only; the wait for the source read sits right before Tt is written again (next batch / kernel end)
        if (A.add) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(A.dst + b0 * N3), "r"(smem_u32(Tt)), "r"(bytes) : "memory");
        else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A.dst + b0 * N3), "r"(smem_u32(Tt)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      for (int i = t; i < nvalid * N3; i += NT) { if (A.add) A.dst[b0 * N3 + i] += Tt[i]; else A.dst[b0 * N3 + i] = Tt[i]; }
    }
    // no barrier here: Tt is next written after the y sweep of the following batch, behind the store wait + barriers
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // all bulk stores complete before the CTA exits
}

struct CartPlan
{
  int n = 0, B = 0, H = 0, n_batches = 0;
  int2 * d_halo = nullptr; int32_t * d_cnt = nullptr;
  int32_t * d_interior = nullptr, * d_boundary = nullptr; int n_interior = 0, n_boundary = 0;
  int32_t * d_ordered = nullptr; // interior batches, then the batches with ghost neighbours (single-launch partitioned vmult)
  size_t smem = 0, smem_line = 0; // smem_line: line kernel (n >= 6), 0 if it does not apply
  std::vector<char> tables; // CartTables<n> of this operator (depends on h and tau)
  // pipelined kernel (n = 5)
  bool pipe = false; int HL = 0, HD = 0, n_sm = 148; int4 * d_cnt4 = nullptr; int pipe_ctas_per_sm = 1;
  // warp-specialised kernel (n = 5, same 24-cell batches): vmult_cartesian_ws.cu
  void * ws = nullptr;
};

// kernel of the fast path for n = 5: 0 pipelined 4-warp kernel, 1 (default) / 2 warp-specialised kernel with producer depth 8 / 12,
// 3 warp-specialised kernel with 4 producer warps and register re-allocation (EXADG_B200_CART_KERNEL=pipe / ws / ws12 / ws4p)
int g_cart_kernel = -1;

template<int N>
CartTables<N> make_cart_tables(const DeviceOperator & op)
{
  Tables1D tab(N - 1);
  CartTables<N> T;
  for (int d = 0; d < 3; ++d) {
    const int e = (d + 1) % 3, f = (d + 2) % 3;
    const real_t cd = (real_t)op.laplace_coeff * op.h[e] * op.h[f] / op.h[d]; // laplace_coeff: viscosity of the Helmholtz operator (1 otherwise)
    const real_t tau_hat = (real_t)op.tau_hat * op.h[d];
    T.tau_hat[d] = (double)tau_hat;
    // own-side 1-D operator K + sum_s [ -1/2 sigma (d e^T + e d^T) + tau_hat e e^T ]
    std::vector<real_t> L(tab.K);
    for (int s = 0; s < 2; ++s) {
      const real_t sig = s ? 1 : -1; const int end = s ? N - 1 : 0;
      for (int i = 0; i < N; ++i) { L[i * N + end] -= sig * tab.fd[s][i] / 2; L[end * N + i] -= sig * tab.fd[s][i] / 2; }
      L[end * N + end] += tau_hat;
    }
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) {
      real_t v = 0;
      for (int m = 0; m < N; ++m) v += tab.Minv[i * N + m] * L[m * N + j];
      T.G[d][i * N + j] = (double)(cd * v);
    }
    for (int s = 0; s < 2; ++s) {
      const real_t sig = s ? 1 : -1; const int end = s ? N - 1 : 0;
      for (int i = 0; i < N; ++i) {
        real_t md = 0;
        for (int m = 0; m < N; ++m) md += tab.Minv[i * N + m] * tab.fd[s][m];
        T.P[d][s][i] = (double)(cd * sig / 2 * md);
        T.Q[d][s][i] = (double)(-cd * tab.Minv[i * N + end]);
      }
    }
  }
  for (int i = 0; i < N * N; ++i) T.M[i] = (double)tab.M[i];
  for (int s = 0; s < 2; ++s) for (int i = 0; i < N; ++i) T.fd[s][i] = (double)tab.fd[s][i];
  return T;
}

// list / n_list: explicit batch list (chunked host-buffer vmult); otherwise `which` selects all / interior / boundary batches
template<int N>
void launch_n(const DeviceOperator & op, const CartPlan & plan, double * dst, const double * src, bool add, int which, cudaStream_t stream, const int32_t * list = nullptr,
              int n_list = 0)
{
  constexpr int B = CartCfg<N>::B;
  const CartTables<N> & T = *reinterpret_cast<const CartTables<N> *>(plan.tables.data());
  if (first_use_on_device((const void *)vmult_cartesian_kernel<N>))
    CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
  if (N == 4 && first_use_on_device((const void *)vmult_cartesian_kernel<4, 32>))
    CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_kernel<4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
  CartArgs A;
  A.nb = op.nb; A.halo = plan.d_halo; A.halo_cnt = plan.d_cnt; A.src = src; A.ghost = op.ghost; A.dst = dst;
  A.n_owned = op.n_owned; A.H = plan.H; A.add = add ? 1 : 0; A.mass = op.helmholtz ? op.mass_coeff * op.h[0] * op.h[1] * op.h[2] : 0.0;
  A.ncomp = op.helmholtz ? op.n_components : 1;
  A.batches = which == 0 ? nullptr : (which == 1 ? plan.d_interior : plan.d_boundary);
  A.n_items = which == 0 ? plan.n_batches : (which == 1 ? plan.n_interior : plan.n_boundary);
  if (list) { A.batches = list; A.n_items = n_list; }
  if (A.n_items == 0) return;
  if (A.ncomp > 1) { // Helmholtz operator: one component of a batch per CTA
    if (first_use_on_device((const void *)vmult_cartesian_kernel<N, CartCfg<N>::B, true>))
      CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_kernel<N, CartCfg<N>::B, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    vmult_cartesian_kernel<N, CartCfg<N>::B, true><<<A.n_items * A.ncomp, B * N, plan.smem, stream>>>(T, A);
  }
  else if (N == 4 && plan.B == 32) vmult_cartesian_kernel<4, 32><<<A.n_items, 32 * 4, plan.smem, stream>>>(*reinterpret_cast<const CartTables<4> *>(plan.tables.data()), A);
  else vmult_cartesian_kernel<N><<<A.n_items, B * N, plan.smem, stream>>>(T, A);
  CUDA_CHECK(cudaGetLastError());
}
template<int N, int BB>
void launch_line_b(const DeviceOperator & op, const CartPlan & plan, double * dst, const double * src, bool add, int which, cudaStream_t stream, const int32_t * list,
                   int n_list)
{
  const CartTables<N> & T = *reinterpret_cast<const CartTables<N> *>(plan.tables.data());
  if (first_use_on_device((const void *)vmult_cartesian_line_kernel<N, BB>))
    CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_line_kernel<N, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
  CartArgs A;
  A.nb = op.nb; A.halo = plan.d_halo; A.halo_cnt = plan.d_cnt; A.src = src; A.ghost = op.ghost; A.dst = dst;
  A.n_owned = op.n_owned; A.H = plan.H; A.add = add ? 1 : 0; A.mass = op.helmholtz ? op.mass_coeff * op.h[0] * op.h[1] * op.h[2] : 0.0;
  A.ncomp = op.helmholtz ? op.n_components : 1;
  A.batches = which == 0 ? nullptr : (which == 1 ? plan.d_interior : plan.d_boundary);
  A.n_items = which == 0 ? plan.n_batches : (which == 1 ? plan.n_interior : plan.n_boundary);
  if (list) { A.batches = list; A.n_items = n_list; }
  if (A.n_items == 0) return;
  static const bool skip_halo = getenv("EXADG_B200_LINE_SKIP_HALO") != nullptr; // timing experiment only (results wrong)
  if (skip_halo) A.add |= 2;
  if (A.ncomp > 1) {
    if (first_use_on_device((const void *)vmult_cartesian_line_kernel<N, BB, true>))
      CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_line_kernel<N, BB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    vmult_cartesian_line_kernel<N, BB, true><<<A.n_items * A.ncomp, LineCfg<N, BB>::NT, plan.smem_line, stream>>>(T, A);
  }
  else vmult_cartesian_line_kernel<N, BB><<<A.n_items, LineCfg<N, BB>::NT, plan.smem_line, stream>>>(T, A);
  CUDA_CHECK(cudaGetLastError());
}
template<int N>
void launch_line(const DeviceOperator & op, const CartPlan & plan, double * dst, const double * src, bool add, int which, cudaStream_t stream, const int32_t * list = nullptr,
                 int n_list = 0)
{
  if (plan.B == 8) launch_line_b<N, 8>(op, plan, dst, src, add, which, stream, list, n_list);
  else launch_line_b<N, 16>(op, plan, dst, src, add, which, stream, list, n_list);
}
template<int N>
void launch_pipe(const DeviceOperator & op, const CartPlan & plan, double * dst, const double * src, bool add, int which, cudaStream_t stream, const int32_t * list = nullptr,
                 int n_list = 0)
{
  const CartTables<N> & T = *reinterpret_cast<const CartTables<N> *>(plan.tables.data());
  if (first_use_on_device((const void *)vmult_cartesian_pipe_kernel<N>))
    CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_pipe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
  const int ctas_per_sm = plan.pipe_ctas_per_sm; // occupancy of this plan's shared-memory size (plan_create)
  PipeArgs A;
  A.nb = op.nb; A.halo = plan.d_halo; A.halo_cnt = plan.d_cnt4; A.src = src; A.ghost = op.ghost; A.dst = dst;
  A.n_owned = op.n_owned; A.HL = plan.HL; A.HD = plan.HD; A.add = add ? 1 : 0;
  A.batches = which == 0 ? nullptr : (which == 1 ? plan.d_interior : plan.d_boundary);
  A.n_items = which == 0 ? plan.n_batches : (which == 1 ? plan.n_interior : plan.n_boundary);
  if (list) { A.batches = list; A.n_items = n_list; }
  if (A.n_items == 0) return;
  const int grid = std::min(A.n_items, plan.n_sm * ctas_per_sm);
  vmult_cartesian_pipe_kernel<N><<<grid, PipeCfg<N>::NT, plan.smem, stream>>>(T, A);
  CUDA_CHECK(cudaGetLastError());
}
} // namespace

bool cartesian_supported(int n) { return n >= 2 && n <= 8; }

int cartesian_kernel_variant(int set)
{
  if (g_cart_kernel < 0) { const char * e = getenv("EXADG_B200_CART_KERNEL"); g_cart_kernel = (e && std::strcmp(e, "pipe") == 0) ? 0 : ((e && std::strcmp(e, "ws12") == 0) ? 2 : ((e && std::strcmp(e, "ws4p") == 0) ? 3 : ((e && std::strcmp(e, "wp") == 0) ? 4 : ((e && std::strcmp(e, "ws") == 0) ? 1 : ((e && std::strcmp(e, "wst") == 0) ? 6 : 3))))); }
  const int previous = g_cart_kernel;
  if (set >= 0) g_cart_kernel = set;
  return previous;
}

template<int N>
void store_tables(CartPlan & P, const DeviceOperator & op)
{
  const CartTables<N> T = make_cart_tables<N>(op);
  P.tables.resize(sizeof(T));
  std::memcpy(P.tables.data(), &T, sizeof(T));
}

static size_t plan_create(DeviceOperator & op, const HostMesh & mesh, bool allow_pipe);
// measured (scripts/r02_shot46.sh, register budgets of LineCfg::MINB): 8-cell batches 91.5 / 84.4 / 89.3 GDoF/s at k = 5 / 6 / 7 against 84.6 / 62.6 / 71.8
// with 16-cell batches (4 / 3 / 2 resident CTAs instead of 2 / 1 / 1)
static int line_batch_default(int n) { (void)n; return 8; }

// builds the batch plan (halo lists, interior/boundary batches, tables); returns the dynamic shared
// memory per CTA, or 0 if the batch does not fit (caller falls back to the general kernel)
size_t cartesian_plan_create(DeviceOperator & op, const HostMesh & mesh) { return plan_create(op, mesh, true); }

static size_t plan_create(DeviceOperator & op, const HostMesh & mesh, bool allow_pipe)
{
  CartPlan * Pp = new CartPlan;
  CartPlan & P = *Pp;
  P.n = op.n;
  const int N = op.n;
  P.B = (N >= 6) ? 16 : ((N >= 5) ? 32 : 64);
  if (N == 4) { const char * e = getenv("EXADG_B200_PLANE_B"); if (e && std::atoi(e) == 32) P.B = 32; }
  if (N >= 6 && !getenv("EXADG_B200_NO_LINE")) { // line kernel: 8 or 16 cells per CTA (EXADG_B200_LINE_B; default per degree from the measurements)
    const char * e = getenv("EXADG_B200_LINE_B");
    const int want = e ? std::atoi(e) : line_batch_default(N);
    P.B = (want == 8) ? 8 : 16;
  }
  P.pipe = allow_pipe && (N == 5) && !op.helmholtz && !getenv("EXADG_B200_NO_PIPE"); // (the mass term lives in the plane and line kernels only) // n = 3 measured slower than the 64-cell kernel (0.99 vs 0.86 ms) // pipelined 4-warp kernel (EXADG_B200_NO_PIPE=1: the 5-warp kernel)
  if (P.pipe) P.B = (N == 3) ? PipeCfg<3>::B : PipeCfg<5>::B;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&P.n_sm, cudaDevAttrMultiProcessorCount, dev); if (P.n_sm < 1) P.n_sm = 148; }
  P.n_batches = (int)((mesh.n_owned + P.B - 1) / P.B);
  std::vector<std::vector<int2>> lists(P.n_batches);
  std::vector<int32_t> interior, boundary;
  for (int b = 0; b < P.n_batches; ++b) {
    const int64_t b0 = (int64_t)b * P.B, b1 = std::min<int64_t>(b0 + P.B, mesh.n_owned);
    bool ghost = false;
    for (int64_t c = b0; c < b1; ++c) for (int f = 0; f < 6; ++f) {
      const int32_t p = mesh.nb[c * 6 + f];
      if (p >= b0 && p < b1) continue;
      lists[b].push_back(make_int2((int)(((c - b0) << 3) | f), p));
      ghost |= (p >= mesh.n_owned);
    }
    P.H = std::max<int>(P.H, (int)lists[b].size());
    (ghost ? boundary : interior).push_back(b);
  }
  P.H = std::max(P.H, 1);
  if (P.pipe) {
    // entries sorted by direction; the trace buffer holds one direction at a time
    std::vector<int4> cnt4(P.n_batches);
    for (int b = 0; b < P.n_batches; ++b) {
      std::stable_sort(lists[b].begin(), lists[b].end(), [](const int2 & x, const int2 & y) { return ((x.x & 7) >> 1) < ((y.x & 7) >> 1); });
      int c[3] = {0, 0, 0};
      for (auto & e : lists[b]) c[(e.x & 7) >> 1]++;
      cnt4[b] = make_int4(c[0], c[1], c[2], 0);
      P.HD = std::max(P.HD, std::max(c[0], std::max(c[1], c[2])));
    }
    P.HL = P.H; P.HD = std::max(P.HD, 1);
    const int N2p = N * N, N3p = N2p * N;
    P.smem = ((size_t)2 * P.B * N3p + (size_t)P.B * 2 * N2p + (size_t)2 * P.HD * N2p) * sizeof(double) + (size_t)2 * P.HL * sizeof(int2)
             + (size_t)P.B * 18 * sizeof(int) + 64 + 16;
    if (P.smem > 227 * 1024 - 1024 || P.HL > PipeCfg<5>::NT) { delete Pp; return plan_create(op, mesh, false); } // irregular batches: the 5-warp kernel has no such limits
    {
      int occ = 0;
      if (N == 3) {
        CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_pipe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vmult_cartesian_pipe_kernel<3>, PipeCfg<3>::NT, P.smem));
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(vmult_cartesian_pipe_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vmult_cartesian_pipe_kernel<5>, PipeCfg<5>::NT, P.smem));
      }
      P.pipe_ctas_per_sm = std::max(occ, 1);
    }
    CUDA_CHECK(cudaMalloc(&P.d_cnt4, cnt4.size() * sizeof(int4)));
    CUDA_CHECK(cudaMemcpy(P.d_cnt4, cnt4.data(), cnt4.size() * sizeof(int4), cudaMemcpyHostToDevice));
  }
  const int N2 = N * N, PS = N2 | 1, CS = N * PS;
  if (!P.pipe) P.smem = ((size_t)2 * P.B * CS + (size_t)P.B * 2 * N2 + (size_t)2 * P.H * N2) * sizeof(double) + (size_t)P.H * sizeof(int2) + (size_t)P.B * 12 * sizeof(int) + 16;
  if (N >= 6 && (P.B == 16 || P.B == 8) && !getenv("EXADG_B200_NO_LINE")) {
    const int RS = N | 1;
    const size_t sl = ((size_t)2 * P.B * RS * N2 + (size_t)2 * P.B * N2 + (size_t)2 * P.H * N2) * sizeof(double) + (size_t)P.H * sizeof(int2) + (size_t)P.B * 12 * sizeof(int) + 16;
    if (sl <= 227 * 1024 - 1024) P.smem_line = sl;
  }
  if (N >= 6 && P.B == 8 && !P.smem_line) { delete Pp; return 0; } // the plane kernel has 16-cell batches only
  if (P.smem > 227 * 1024 - 1024) { delete Pp; return 0; } // does not fit: caller falls back to the general kernel
  std::vector<int2> flat((size_t)P.n_batches * P.H, make_int2(0, 0));
  std::vector<int32_t> cnt(P.n_batches);
  for (int b = 0; b < P.n_batches; ++b) { cnt[b] = (int32_t)lists[b].size(); std::copy(lists[b].begin(), lists[b].end(), flat.begin() + (size_t)b * P.H); }
  CUDA_CHECK(cudaMalloc(&P.d_halo, flat.size() * sizeof(int2)));
  CUDA_CHECK(cudaMemcpy(P.d_halo, flat.data(), flat.size() * sizeof(int2), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&P.d_cnt, cnt.size() * sizeof(int32_t)));
  CUDA_CHECK(cudaMemcpy(P.d_cnt, cnt.data(), cnt.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  P.n_interior = (int)interior.size(); P.n_boundary = (int)boundary.size();
  if (mesh.world > 1) {
    if (P.n_interior) { CUDA_CHECK(cudaMalloc(&P.d_interior, interior.size() * 4)); CUDA_CHECK(cudaMemcpy(P.d_interior, interior.data(), interior.size() * 4, cudaMemcpyHostToDevice)); }
    if (P.n_boundary) { CUDA_CHECK(cudaMalloc(&P.d_boundary, boundary.size() * 4)); CUDA_CHECK(cudaMemcpy(P.d_boundary, boundary.data(), boundary.size() * 4, cudaMemcpyHostToDevice)); }
    std::vector<int32_t> ordered(interior);
    ordered.insert(ordered.end(), boundary.begin(), boundary.end());
    if (!ordered.empty()) { CUDA_CHECK(cudaMalloc(&P.d_ordered, ordered.size() * 4)); CUDA_CHECK(cudaMemcpy(P.d_ordered, ordered.data(), ordered.size() * 4, cudaMemcpyHostToDevice)); }
  }
  switch (N) {
    case 2: store_tables<2>(P, op); break;
    case 3: store_tables<3>(P, op); break;
    case 4: store_tables<4>(P, op); break;
    case 5: store_tables<5>(P, op); break;
    case 6: store_tables<6>(P, op); break;
    case 7: store_tables<7>(P, op); break;
    case 8: store_tables<8>(P, op); break;
    default: delete Pp; return 0;
  }
  if (P.pipe && P.B == ws::WsCfg<5>::B) {
    // optional second kernel for the same batches; stays null if the batches are too irregular for it or anything goes wrong
    try { P.ws = ws_plan_create(op, mesh); } catch (const std::exception &) { P.ws = nullptr; cudaGetLastError(); }
  }
  op.cart_plan = Pp;
  return P.smem;
}

void cartesian_plan_destroy(DeviceOperator & op)
{
  CartPlan * P = static_cast<CartPlan *>(op.cart_plan);
  if (!P) return;
  cudaFree(P->d_halo); cudaFree(P->d_cnt); cudaFree(P->d_cnt4); cudaFree(P->d_interior); cudaFree(P->d_boundary); cudaFree(P->d_ordered);
  ws_plan_destroy(P->ws);
  delete P;
  op.cart_plan = nullptr;
}

int cartesian_batch_size(const DeviceOperator & op)
{
  const CartPlan * plan = static_cast<const CartPlan *>(op.cart_plan);
  return plan ? plan->B : 0;
}
int cartesian_n_batches(const DeviceOperator & op)
{
  const CartPlan * plan = static_cast<const CartPlan *>(op.cart_plan);
  return plan ? plan->n_batches : 0;
}

static void launch_cart(const DeviceOperator & op, double * dst, const double * src, bool add, int which, const int32_t * list, int n_list, cudaStream_t stream);

static int ws_depth_of(int variant) { return variant == 2 ? 12 : (variant == 3 ? 4 : (variant == 4 ? 100 : (variant == 5 ? 101 : (variant == 6 ? 3 : 8)))); }

bool launch_vmult_cartesian_fused(const DeviceOperator & op, double * dst, const double * src, bool add, const GhostSync & gs, cudaStream_t stream)
{
  const CartPlan * plan = static_cast<const CartPlan *>(op.cart_plan);
  const int variant = op.cart_variant >= 0 ? op.cart_variant : cartesian_kernel_variant(-1);
  static const bool off = getenv("EXADG_B200_NO_FUSED_HALO") != nullptr; // measurement switch: the multi-launch path
  if (off || !plan || op.n != 5 || !plan->ws || variant < 1 || !plan->d_ordered || gs.n_peers > 16) return false;
  ws_launch(op, plan->ws, dst, src, add, plan->d_ordered, plan->n_interior + plan->n_boundary, plan->n_sm, ws_depth_of(variant), true, stream, &gs, plan->n_interior);
  return true;
}

// which: 0 all batches, 1 batches that touch no ghost cell, 2 batches that do
void launch_vmult_cartesian_part(const DeviceOperator & op, double * dst, const double * src, bool add, int which, cudaStream_t stream)
{
  launch_cart(op, dst, src, add, which, nullptr, 0, stream);
}

// explicit list of batches that read no ghost cell (steps of the pipelined host-buffer vmult; on a partitioned operator the batches with
// ghost neighbours are not in these lists - they run behind the ghost import)
void launch_vmult_cartesian_list(const DeviceOperator & op, double * dst, const double * src, bool add, const int32_t * list, int n_list, cudaStream_t stream)
{
  if (n_list > 0) launch_cart(op, dst, src, add, 1, list, n_list, stream);
}

static void launch_cart(const DeviceOperator & op, double * dst, const double * src, bool add, int which, const int32_t * list, int n_list, cudaStream_t stream)
{
  const CartPlan * plan = static_cast<const CartPlan *>(op.cart_plan);
  if (!plan) throw std::runtime_error("Cartesian plan missing");
  switch (op.n) {
    case 2: launch_n<2>(op, *plan, dst, src, add, which, stream, list, n_list); break;
    case 3: if (plan->pipe) launch_pipe<3>(op, *plan, dst, src, add, which, stream, list, n_list); else launch_n<3>(op, *plan, dst, src, add, which, stream, list, n_list); break;
    case 4: launch_n<4>(op, *plan, dst, src, add, which, stream, list, n_list); break;
    case 5: {
      // warp-specialised kernel for full launches of an unpartitioned mesh and for the interior launch of a partition (no ghost
      // reads by construction: the very kernel instantiation that is verified on one GPU); the batches that touch ghost cells
      // stay on the pipelined kernel until the ghost path of the warp-specialised kernel has run on >= 2 GPUs
      const int variant = op.cart_variant >= 0 ? op.cart_variant : cartesian_kernel_variant(-1);
      static const bool pipe_ghost = getenv("EXADG_B200_PIPE_GHOST") != nullptr; // measurement switch: batches with ghost neighbours on the pipelined kernel (round-1 behaviour)
      const bool with_ghosts = which == 2 || (which == 0 && op.n_ghost > 0);
      const bool ws_ok = plan->ws && variant >= 1 && (!with_ghosts || !pipe_ghost);
      if (ws_ok)
        ws_launch(op, plan->ws, dst, src, add, list ? list : (which == 0 ? nullptr : (which == 1 ? plan->d_interior : plan->d_boundary)),
                  list ? n_list : (which == 0 ? plan->n_batches : (which == 1 ? plan->n_interior : plan->n_boundary)), plan->n_sm,
                  ws_depth_of(variant), with_ghosts, stream);
      else if (plan->pipe) launch_pipe<5>(op, *plan, dst, src, add, which, stream, list, n_list);
      else launch_n<5>(op, *plan, dst, src, add, which, stream, list, n_list);
      break;
    }
    case 6: if (plan->smem_line) launch_line<6>(op, *plan, dst, src, add, which, stream, list, n_list); else launch_n<6>(op, *plan, dst, src, add, which, stream, list, n_list); break;
    case 7: if (plan->smem_line) launch_line<7>(op, *plan, dst, src, add, which, stream, list, n_list); else launch_n<7>(op, *plan, dst, src, add, which, stream, list, n_list); break;
    case 8: if (plan->smem_line) launch_line<8>(op, *plan, dst, src, add, which, stream, list, n_list); else launch_n<8>(op, *plan, dst, src, add, which, stream, list, n_list); break;
    default: throw std::runtime_error("Cartesian fast path supports degrees 1..7");
  }
}

} // namespace exadg_b200
