// 1-D reference-element tables for FE_DGQ(k) with Gauss(k+1) quadrature (host side).
//
// Reference semantics: I/operators/finite_element.h:58-61 (FE_DGQ(k): tensor Lagrange basis on the
// k+1 Gauss-Lobatto points, lexicographic) and I/operators/quadrature.h:43-46 (QGauss(k+1)).
// The arithmetic lives in deal.II (not vendored in the reference); the tables are rebuilt here
// from the definitions, in long double.
#pragma once
#include <cmath>
#include <vector>

namespace exadg_b200
{
typedef long double real_t;

inline void legendre_pd(int n, real_t x, real_t & p, real_t & dp)
{
  if (n == 0) { p = 1; dp = 0; return; }
  real_t pm = 1, pc = x;
  for (int j = 2; j <= n; ++j) { real_t pn = ((2 * j - 1) * x * pc - (j - 1) * pm) / j; pm = pc; pc = pn; }
  p = pc;
  dp = n * (x * pc - pm) / (x * x - 1);
}

// Gauss-Legendre on [0,1]
inline void gauss_points(int n, std::vector<real_t> & x, std::vector<real_t> & w)
{
  x.assign(n, 0); w.assign(n, 0);
  const real_t pi = acosl(-1.0L);
  for (int i = 0; i < (n + 1) / 2; ++i) {
    real_t z = cosl(pi * (i + 0.75L) / (n + 0.5L)), p, dp;
    for (int it = 0; it < 100; ++it) { legendre_pd(n, z, p, dp); real_t dz = p / dp; z -= dz; if (fabsl(dz) < 1e-19L) break; }
    legendre_pd(n, z, p, dp);
    real_t wi = 2 / ((1 - z * z) * dp * dp);
    x[n - 1 - i] = (1 + z) / 2; x[i] = (1 - z) / 2;
    w[n - 1 - i] = wi / 2; w[i] = wi / 2;
  }
  if (n % 2) x[n / 2] = 0.5L;
}

// Gauss-Lobatto points on [0,1] (n >= 2)
inline void lobatto_points(int n, std::vector<real_t> & x)
{
  x.assign(n, 0);
  x[n - 1] = 1;
  const int N = n - 1;
  const real_t pi = acosl(-1.0L);
  for (int i = 1; i <= (n - 1) / 2; ++i) {
    real_t z = -cosl(pi * i / N), p, dp;
    for (int it = 0; it < 200; ++it) {
      legendre_pd(N, z, p, dp);
      real_t ddp = (2 * z * dp - N * (N + 1) * p) / (1 - z * z);
      real_t dz = dp / ddp; z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    x[i] = (1 + z) / 2; x[n - 1 - i] = (1 - z) / 2;
  }
  if (n % 2) x[n / 2] = 0.5L;
}

// Lagrange polynomials on `nodes`, value and derivative at x
inline void lagrange_at(const std::vector<real_t> & nodes, real_t x, std::vector<real_t> & v, std::vector<real_t> & d)
{
  const int n = (int)nodes.size();
  v.assign(n, 0); d.assign(n, 0);
  for (int j = 0; j < n; ++j) {
    real_t pv = 1;
    for (int i = 0; i < n; ++i) if (i != j) pv *= (x - nodes[i]) / (nodes[j] - nodes[i]);
    real_t pd = 0;
    for (int m = 0; m < n; ++m) if (m != j) {
      real_t t = 1 / (nodes[j] - nodes[m]);
      for (int i = 0; i < n; ++i) if (i != j && i != m) t *= (x - nodes[i]) / (nodes[j] - nodes[i]);
      pd += t;
    }
    v[j] = pv; d[j] = pd;
  }
}

// dense inverse (Gauss-Jordan with partial pivoting), row-major n x n
inline std::vector<real_t> invert(std::vector<real_t> A, int n)
{
  std::vector<real_t> I(n * n, 0);
  for (int i = 0; i < n; ++i) I[i * n + i] = 1;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabsl(A[r * n + c]) > fabsl(A[piv * n + c])) piv = r;
    for (int j = 0; j < n; ++j) { std::swap(A[c * n + j], A[piv * n + j]); std::swap(I[c * n + j], I[piv * n + j]); }
    real_t s = 1 / A[c * n + c];
    for (int j = 0; j < n; ++j) { A[c * n + j] *= s; I[c * n + j] *= s; }
    for (int r = 0; r < n; ++r) if (r != c) {
      real_t f = A[r * n + c];
      if (f == 0) continue;
      for (int j = 0; j < n; ++j) { A[r * n + j] -= f * A[c * n + j]; I[r * n + j] -= f * I[c * n + j]; }
    }
  }
  return I;
}

struct Tables1D
{
  int n = 0;                    // k+1 basis functions = quadrature points
  std::vector<real_t> xn, xq, w; // Gauss-Lobatto nodes, Gauss points, weights (on [0,1])
  std::vector<real_t> S, D;     // S[q*n+j] = l_j(x_q), D[q*n+j] = l_j'(x_q)       (nodal basis at Gauss points)
  std::vector<real_t> Dq;       // Dq[q*n+p] = psi_p'(x_q)   (collocation derivative, psi = Lagrange on Gauss points)
  std::vector<real_t> sv[2], sd[2]; // psi_p(s), psi_p'(s) at s = 0, 1
  std::vector<real_t> fd[2];    // l_j'(s) at s = 0, 1 (nodal); l_j(s) = delta with the end node
  std::vector<real_t> M, K, Minv; // reference mass S^T W S, stiffness D^T W D, inverse mass

  explicit Tables1D(int degree)
  {
    n = degree + 1;
    if (degree == 0) xn.assign(1, 0.5L); else lobatto_points(n, xn);
    gauss_points(n, xq, w);
    S.assign(n * n, 0); D.assign(n * n, 0); Dq.assign(n * n, 0);
    std::vector<real_t> v, d;
    for (int q = 0; q < n; ++q) {
      lagrange_at(xn, xq[q], v, d);
      for (int j = 0; j < n; ++j) { S[q * n + j] = v[j]; D[q * n + j] = d[j]; }
      lagrange_at(xq, xq[q], v, d);
      for (int j = 0; j < n; ++j) Dq[q * n + j] = d[j];
    }
    for (int s = 0; s < 2; ++s) {
      lagrange_at(xq, (real_t)s, sv[s], sd[s]);
      lagrange_at(xn, (real_t)s, v, fd[s]);
    }
    M.assign(n * n, 0); K.assign(n * n, 0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
      real_t m = 0, k = 0;
      for (int q = 0; q < n; ++q) { m += S[q * n + i] * w[q] * S[q * n + j]; k += D[q * n + i] * w[q] * D[q * n + j]; }
      M[i * n + j] = m; K[i * n + j] = k;
    }
    Minv = invert(M, n);
  }
};

} // namespace exadg_b200
