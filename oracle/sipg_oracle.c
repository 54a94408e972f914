/*
 * sipg_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle / reported CPU baseline).
 *
 * A dependency-free, plain-C restatement of the algorithm behind ExaDG's
 * Poisson::LaplaceOperator::vmult (3-D SIPG Laplace, FE_DGQ(k), Gauss(k+1)) and of
 * the callers around it (diagonal, dealii::SolverCG, dealii::PreconditionChebyshev,
 * rhs / L2 error of applications/poisson/sine).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (exadg_b200/) never links, imports or calls it.  The timed CPU baseline is the
 * vectorised variant in sipg_fast.inc (orc_vmult_fast), checked against the functions here.
 *
 * Parity pin: vmult itself is pinned by no reference fixture ("parity unpinned" at the
 * operator level); the oracle as a whole (cell + interior-face + Dirichlet + Neumann
 * terms, penalty, MappingQ(3) geometry, rhs, L2 error) IS pinned to 6 digits by the
 * golden L2 errors of applications/poisson/sine/tests/{cartesian,curvilinear}.output
 * (tests/test_oracle_golden.py).
 *
 * Reference lines restated (paths relative to /root/reference, I/ = include/exadg/):
 *   I/operators/operator_base.cpp:264-310   apply(): loop(cell_loop, face_loop, boundary_face_loop_hom_operator), zero dst
 *   I/operators/operator_base.cpp:312-354   apply_add()
 *   I/operators/operator_base.cpp:1349-1370 cell_loop
 *   I/operators/operator_base.cpp:1372-1397 face_loop
 *   I/operators/operator_base.cpp:1399-1434 boundary_face_loop_hom_operator
 *   I/operators/operator_base.cpp:1618-1704 cell_based_loop_diagonal (cell-wise view of the faces)
 *   I/poisson/spatial_discretization/laplace_operator.cpp:129-265 do_cell/face/boundary integrals
 *   I/poisson/spatial_discretization/laplace_operator.h:128-197   tau per face, fluxes
 *   I/poisson/spatial_discretization/weak_boundary_conditions.h:34-234 mirror values
 *   I/operators/interior_penalty_parameter.h:43-128 tau_K and (k+1)^2 factor
 *   I/solvers_and_preconditioners/utilities/invert_diagonal.h:35-46
 *   I/grid/periodic_box.h:35-88, I/grid/deformed_cube_manifold.h:47-60,127-157
 *   applications/poisson/sine/application.h:32-99 (solution, Neumann data, rhs)
 *   I/postprocessor/error_calculation.cpp:36-115 (relative L2 error, Gauss(k+3))
 * deal.II conventions (deal.II is not vendored in the reference; restated from its
 * documented behaviour): FE_DGQ(k) = tensor Lagrange basis on k+1 Gauss-Lobatto points,
 * lexicographic numbering (x fastest); DoFs numbered cell by cell in active-cell order
 * (coarse cell lexicographic, children in z-order); MappingQ(m) support points = tensor
 * Gauss-Lobatto points pushed through the manifold; SolverCG / ReductionControl;
 * PreconditionChebyshev.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 16 /* max 1-D points (degree+1, quadrature, mapping) */

/* ------------------------------------------------------------------------- */
/* 1-D tables                                                                */
/* ------------------------------------------------------------------------- */

static void legendre(int n, long double x, long double *p, long double *dp)
{
  /* P_n(x) and P_n'(x) on [-1,1] */
  long double p0 = 1.0L, p1 = x;
  if (n == 0) { *p = 1.0L; *dp = 0.0L; return; }
  for (int j = 2; j <= n; ++j) {
    long double pj = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
    p0 = p1; p1 = pj;
  }
  *p = p1;
  *dp = n * (x * p1 - p0) / (x * x - 1.0L);
}

/* Gauss-Legendre points/weights on [0,1] */
static void gauss_legendre(int n, double *x, double *w)
{
  for (int i = 0; i < n; ++i) {
    long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
    for (int it = 0; it < 100; ++it) {
      long double p, dp; legendre(n, z, &p, &dp);
      long double dz = p / dp; z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    long double p, dp; legendre(n, z, &p, &dp);
    long double wi = 2.0L / ((1.0L - z * z) * dp * dp);
    x[n - 1 - i] = (double)(0.5L * (z + 1.0L));
    w[n - 1 - i] = (double)(0.5L * wi);
  }
}

/* Gauss-Lobatto points on [0,1], n >= 2 */
static void gauss_lobatto(int n, double *x)
{
  x[0] = 0.0; x[n - 1] = 1.0;
  int N = n - 1; /* interior points are roots of P_N' */
  for (int i = 1; i < n - 1; ++i) {
    long double z = -cosl(3.14159265358979323846264338327950288L * i / N);
    for (int it = 0; it < 200; ++it) {
      /* f = P_N'(z), f' from Legendre ODE: (1-z^2) P'' = 2 z P' - N(N+1) P */
      long double p, dp; legendre(N, z, &p, &dp);
      long double ddp = (2.0L * z * dp - N * (N + 1) * p) / (1.0L - z * z);
      long double dz = dp / ddp; z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    x[i] = (double)(0.5L * (z + 1.0L));
  }
  /* symmetrise */
  for (int i = 0; i < n / 2; ++i) {
    double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
    x[i] = a; x[n - 1 - i] = 1.0 - a;
  }
  if (n % 2) x[n / 2] = 0.5;
}

/* Lagrange basis l_j on nodes xn[0..n) at point x: value and derivative */
static void lagrange(int n, const double *xn, double x, double *val, double *der)
{
  for (int j = 0; j < n; ++j) {
    long double v = 1.0L, d = 0.0L;
    for (int i = 0; i < n; ++i) if (i != j) v *= ((long double)x - xn[i]) / ((long double)xn[j] - xn[i]);
    for (int m = 0; m < n; ++m) if (m != j) {
      long double t = 1.0L / ((long double)xn[j] - xn[m]);
      for (int i = 0; i < n; ++i) if (i != j && i != m) t *= ((long double)x - xn[i]) / ((long double)xn[j] - xn[i]);
      d += t;
    }
    val[j] = (double)v; der[j] = (double)d;
  }
}

typedef struct {
  int n;  /* k+1 basis functions per direction */
  int nq; /* quadrature points per direction    */
  double xn[MAXN], xq[MAXN], w[MAXN];
  double S[MAXN * MAXN];  /* S[q*n+j] = l_j(x_q)  */
  double D[MAXN * MAXN];  /* D[q*n+j] = l_j'(x_q) */
  double fv[2][MAXN];     /* l_j(0), l_j(1)   */
  double fd[2][MAXN];     /* l_j'(0), l_j'(1) */
} Basis;

static void basis_init(Basis *b, int degree, int nq)
{
  b->n = degree + 1; b->nq = nq;
  if (degree == 0) b->xn[0] = 0.5; else gauss_lobatto(b->n, b->xn);
  gauss_legendre(nq, b->xq, b->w);
  for (int q = 0; q < nq; ++q) lagrange(b->n, b->xn, b->xq[q], &b->S[q * b->n], &b->D[q * b->n]);
  lagrange(b->n, b->xn, 0.0, b->fv[0], b->fd[0]);
  lagrange(b->n, b->xn, 1.0, b->fv[1], b->fd[1]);
}

/* ------------------------------------------------------------------------- */
/* Mesh                                                                      */
/* ------------------------------------------------------------------------- */

enum { BT_INTERIOR = 0, BT_DIRICHLET = 1, BT_NEUMANN = 2 };

typedef struct {
  int m;            /* mapping degree */
  long n_cells;
  double *xmap;     /* [n_cells][(m+1)^3][3], lexicographic, Gauss-Lobatto support points */
  long *nb;         /* [n_cells][6] neighbour cell (periodic neighbours included) or -1 */
  unsigned char *nbface; /* [n_cells][6] face number of the neighbour */
  unsigned char *bt;     /* [n_cells][6] BT_* */
} Mesh;

/* I/grid/deformed_cube_manifold.h:47-60 push_forward */
static void push_forward(const double X[3], double left, double right, double deformation, int frequency, double x[3])
{
  double sinval = deformation;
  for (int d = 0; d < 3; ++d) sinval *= sin(frequency * M_PI * (X[d] - left) / (right - left));
  for (int d = 0; d < 3; ++d) x[d] = X[d] + sinval;
}

/* position of cell `c` (active-cell order: coarse cell lexicographic x fastest, then
 * hierarchical children, child = x + 2y + 4z) -> integer coordinates on the n^3 grid */
static void cell_to_ijk(long c, int n_sub, int refine, int ijk[3])
{
  long per_coarse = 1L << (3 * refine);
  long coarse = c / per_coarse, mort = c % per_coarse;
  int cx = (int)(coarse % n_sub), cy = (int)((coarse / n_sub) % n_sub), cz = (int)(coarse / ((long)n_sub * n_sub));
  int x = 0, y = 0, z = 0;
  for (int l = 0; l < refine; ++l) {
    int child = (int)((mort >> (3 * (refine - 1 - l))) & 7);
    x = (x << 1) | (child & 1); y = (y << 1) | ((child >> 1) & 1); z = (z << 1) | ((child >> 2) & 1);
  }
  ijk[0] = (cx << refine) + x; ijk[1] = (cy << refine) + y; ijk[2] = (cz << refine) + z;
}

static long ijk_to_cell(const int ijk[3], int n_sub, int refine)
{
  int mask = (1 << refine) - 1;
  int cx = ijk[0] >> refine, cy = ijk[1] >> refine, cz = ijk[2] >> refine;
  int x = ijk[0] & mask, y = ijk[1] & mask, z = ijk[2] & mask;
  long mort = 0;
  for (int l = refine - 1; l >= 0; --l) {
    int child = ((x >> l) & 1) | (((y >> l) & 1) << 1) | (((z >> l) & 1) << 2);
    mort = (mort << 3) | child;
  }
  long coarse = cx + (long)n_sub * (cy + (long)n_sub * cz);
  return coarse * (1L << (3 * refine)) + mort;
}

/* subdivided_hyper_cube(n_sub,-1,1) + refine_global(refine) (I/grid/periodic_box.h:47,86),
 * bc[f] per domain face f (0/1 = x low/high, 2/3 = y, 4/5 = z): 0 periodic, 1 Dirichlet, 2 Neumann */
static Mesh *mesh_hypercube(int n_sub, int refine, int m, double deformation, int frequency, const int bc[6])
{
  const double left = -1.0, right = 1.0;
  int n = n_sub << refine;
  Mesh *M = (Mesh *)calloc(1, sizeof(Mesh));
  M->m = m; M->n_cells = (long)n * n * n;
  int np = m + 1, np3 = np * np * np;
  M->xmap = (double *)malloc(sizeof(double) * M->n_cells * np3 * 3);
  M->nb = (long *)malloc(sizeof(long) * M->n_cells * 6);
  M->nbface = (unsigned char *)malloc(M->n_cells * 6);
  M->bt = (unsigned char *)malloc(M->n_cells * 6);
  double gl[MAXN];
  if (m >= 1) gauss_lobatto(np, gl);
  double h = (right - left) / n;
  for (long c = 0; c < M->n_cells; ++c) {
    int ijk[3]; cell_to_ijk(c, n_sub, refine, ijk);
    for (int a2 = 0; a2 < np; ++a2) for (int a1 = 0; a1 < np; ++a1) for (int a0 = 0; a0 < np; ++a0) {
      double X[3] = { left + h * (ijk[0] + gl[a0]), left + h * (ijk[1] + gl[a1]), left + h * (ijk[2] + gl[a2]) };
      /* snap shared faces/vertices exactly (avoid h*(i+1) vs h*i+h drift) */
      if (a0 == np - 1) X[0] = left + h * (ijk[0] + 1);
      if (a1 == np - 1) X[1] = left + h * (ijk[1] + 1);
      if (a2 == np - 1) X[2] = left + h * (ijk[2] + 1);
      double *x = &M->xmap[(c * np3 + a0 + np * (a1 + np * a2)) * 3];
      if (deformation != 0.0) push_forward(X, left, right, deformation, frequency, x);
      else { x[0] = X[0]; x[1] = X[1]; x[2] = X[2]; }
    }
    for (int f = 0; f < 6; ++f) {
      int d = f / 2, s = f % 2;
      int nijk[3] = { ijk[0], ijk[1], ijk[2] };
      nijk[d] += s ? 1 : -1;
      M->nbface[c * 6 + f] = (unsigned char)(f ^ 1);
      if (nijk[d] < 0 || nijk[d] >= n) {
        if (bc[f] == 0) { /* periodic: counts as interior (interior_penalty_parameter.h:88-89) */
          nijk[d] = (nijk[d] + n) % n;
          M->nb[c * 6 + f] = ijk_to_cell(nijk, n_sub, refine);
          M->bt[c * 6 + f] = BT_INTERIOR;
        } else {
          M->nb[c * 6 + f] = -1;
          M->bt[c * 6 + f] = (unsigned char)bc[f];
        }
      } else {
        M->nb[c * 6 + f] = ijk_to_cell(nijk, n_sub, refine);
        M->bt[c * 6 + f] = BT_INTERIOR;
      }
    }
  }
  return M;
}

static void mesh_free(Mesh *M)
{
  if (!M) return;
  free(M->xmap); free(M->nb); free(M->nbface); free(M->bt); free(M);
}

/* ------------------------------------------------------------------------- */
/* Operator object: geometry at quadrature points, penalty                   */
/* ------------------------------------------------------------------------- */

typedef struct {
  int k;
  Basis b;
  Mesh *mesh;
  double ip_factor;
  long n_cells, n_dofs;
  /* geometry, uncompressed (oracle favours clarity): */
  double *Jinv_c; /* [cell][nq^3][9]   J^{-1}[i][j] = d xi_i / d x_j */
  double *JxW_c;  /* [cell][nq^3] */
  double *Jinv_f; /* [cell][6][nq^2][9] own-side J^{-1} at face points */
  double *nrm_f;  /* [cell][6][nq^2][3] outward unit normal */
  double *JxW_f;  /* [cell][6][nq^2] */
  double *xq_f;   /* [cell][6][nq^2][3] physical face quadrature points (boundary data) */
  double *xq_c;   /* [cell][nq^3][3] physical cell quadrature points */
  double *tauK;   /* [cell] surface/volume */
  /* unique interior faces (face-centric loop like MatrixFree::loop) */
  long n_faces; long *face_m, *face_p; unsigned char *face_fm, *face_fp;
  long n_bfaces; long *bface_c; unsigned char *bface_f;
  /* vectorised baseline (sipg_fast.inc): applicability and cell sizes of the uniform box */
  int fast_checked, fast_ok; double fast_h[3];
  int lean; /* orc_create_periodic_box_lean: no stored geometry */
} Op;

static double det3(const double J[9])
{
  return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
}
static void inv3(const double J[9], double det, double Ji[9])
{
  double id = 1.0 / det;
  Ji[0] = (J[4] * J[8] - J[5] * J[7]) * id; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * id; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * id;
  Ji[3] = (J[5] * J[6] - J[3] * J[8]) * id; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * id; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * id;
  Ji[6] = (J[3] * J[7] - J[4] * J[6]) * id; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * id; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * id;
}

/* Jacobian J[i][j] = d x_i / d xi_j and position of the MappingQ(m) interpolant at xi */
static void mapping_eval(const Mesh *M, long c, const double *gl, const double xi[3], double J[9], double x[3])
{
  int np = M->m + 1;
  double v[3][MAXN], d[3][MAXN];
  for (int e = 0; e < 3; ++e) lagrange(np, gl, xi[e], v[e], d[e]);
  for (int i = 0; i < 9; ++i) J[i] = 0.0;
  x[0] = x[1] = x[2] = 0.0;
  const double *X = &M->xmap[c * np * np * np * 3];
  for (int a2 = 0; a2 < np; ++a2) for (int a1 = 0; a1 < np; ++a1) for (int a0 = 0; a0 < np; ++a0) {
    const double *p = &X[(a0 + np * (a1 + np * a2)) * 3];
    double N = v[0][a0] * v[1][a1] * v[2][a2];
    double g[3] = { d[0][a0] * v[1][a1] * v[2][a2], v[0][a0] * d[1][a1] * v[2][a2], v[0][a0] * v[1][a1] * d[2][a2] };
    for (int i = 0; i < 3; ++i) { x[i] += p[i] * N; for (int j = 0; j < 3; ++j) J[i * 3 + j] += p[i] * g[j]; }
  }
}

static void op_setup_geometry(Op *op)
{
  const Mesh *M = op->mesh; const Basis *b = &op->b;
  int nq = b->nq, nq2 = nq * nq, nq3 = nq2 * nq;
  long nc = op->n_cells;
  op->Jinv_c = (double *)malloc(sizeof(double) * nc * nq3 * 9);
  op->JxW_c = (double *)malloc(sizeof(double) * nc * nq3);
  op->xq_c = (double *)malloc(sizeof(double) * nc * nq3 * 3);
  op->Jinv_f = (double *)malloc(sizeof(double) * nc * 6 * nq2 * 9);
  op->nrm_f = (double *)malloc(sizeof(double) * nc * 6 * nq2 * 3);
  op->JxW_f = (double *)malloc(sizeof(double) * nc * 6 * nq2);
  op->xq_f = (double *)malloc(sizeof(double) * nc * 6 * nq2 * 3);
  op->tauK = (double *)malloc(sizeof(double) * nc);
  double gl[MAXN]; gauss_lobatto(M->m + 1, gl);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < nc; ++c) {
    double volume = 0.0, surface = 0.0;
    for (int q2 = 0; q2 < nq; ++q2) for (int q1 = 0; q1 < nq; ++q1) for (int q0 = 0; q0 < nq; ++q0) {
      int q = q0 + nq * (q1 + nq * q2);
      double xi[3] = { b->xq[q0], b->xq[q1], b->xq[q2] }, J[9], x[3];
      mapping_eval(M, c, gl, xi, J, x);
      double det = det3(J);
      inv3(J, det, &op->Jinv_c[(c * nq3 + q) * 9]);
      op->JxW_c[c * nq3 + q] = det * b->w[q0] * b->w[q1] * b->w[q2];
      for (int i = 0; i < 3; ++i) op->xq_c[(c * nq3 + q) * 3 + i] = x[i];
      volume += op->JxW_c[c * nq3 + q];
    }
    for (int f = 0; f < 6; ++f) {
      int d = f / 2, s = f % 2, t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
      /* interior_penalty_parameter.h:88-89: 1 on true boundary faces, 1/2 otherwise (periodic = interior) */
      double factor = (M->bt[c * 6 + f] != BT_INTERIOR) ? 1.0 : 0.5;
      for (int qb = 0; qb < nq; ++qb) for (int qa = 0; qa < nq; ++qa) {
        int q = qa + nq * qb;
        double xi[3]; xi[d] = (double)s; xi[t1] = b->xq[qa]; xi[t2] = b->xq[qb];
        double J[9], x[3], Ji[9];
        mapping_eval(M, c, gl, xi, J, x);
        double det = det3(J); inv3(J, det, Ji);
        /* n ~ J^{-T} n_ref, n_ref = +-e_d ; area element = |det J| |J^{-T} n_ref| */
        double nv[3] = { Ji[d * 3 + 0], Ji[d * 3 + 1], Ji[d * 3 + 2] };
        double len = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        double sgn = s ? 1.0 : -1.0;
        long o = (c * 6 + f) * nq2 + q;
        for (int i = 0; i < 9; ++i) op->Jinv_f[o * 9 + i] = Ji[i];
        for (int i = 0; i < 3; ++i) { op->nrm_f[o * 3 + i] = sgn * nv[i] / len; op->xq_f[o * 3 + i] = x[i]; }
        op->JxW_f[o] = fabs(det) * len * b->w[qa] * b->w[qb];
        surface += op->JxW_f[o] * factor;
      }
    }
    op->tauK[c] = surface / volume; /* interior_penalty_parameter.h:96 */
  }
}

static void op_build_faces(Op *op)
{
  const Mesh *M = op->mesh; long nc = op->n_cells;
  long nf = 0, nbf = 0;
  for (long c = 0; c < nc; ++c) for (int f = 0; f < 6; ++f) {
    long p = M->nb[c * 6 + f];
    if (p < 0) ++nbf; else if (c < p || (c == p && f > M->nbface[c * 6 + f])) ++nf;
  }
  /* a face between c and p>c may be listed from both (c,f) only once: c<p rule; when two cells
     share two faces (n=2 periodic) both faces appear with distinct f, fine. */
  op->face_m = (long *)malloc(sizeof(long) * (nf + 1)); op->face_p = (long *)malloc(sizeof(long) * (nf + 1));
  op->face_fm = (unsigned char *)malloc(nf + 1); op->face_fp = (unsigned char *)malloc(nf + 1);
  op->bface_c = (long *)malloc(sizeof(long) * (nbf + 1)); op->bface_f = (unsigned char *)malloc(nbf + 1);
  nf = 0; nbf = 0;
  for (long c = 0; c < nc; ++c) for (int f = 0; f < 6; ++f) {
    long p = M->nb[c * 6 + f];
    if (p < 0) { op->bface_c[nbf] = c; op->bface_f[nbf] = (unsigned char)f; ++nbf; }
    else if (c < p || (c == p && f > M->nbface[c * 6 + f])) {
      op->face_m[nf] = c; op->face_fm[nf] = (unsigned char)f; op->face_p[nf] = p; op->face_fp[nf] = M->nbface[c * 6 + f]; ++nf;
    }
  }
  op->n_faces = nf; op->n_bfaces = nbf;
}

/* ------------------------------------------------------------------------- */
/* Sum-factorised evaluation / integration (what FEEvaluation does)          */
/* ------------------------------------------------------------------------- */

/* out[.., r, ..] (+)= sum_c A[r*cols+c] in[.., c, ..] along direction dir; dims = extents of `in` */
static void apply1d(const double *A, int rows, int cols, int dir, const int dims[3], const double *in, double *out, int transpose, int add)
{
  /* transpose: use A^T, i.e. out[c] += sum_r A[r*cols+c] in[r]; then in-extent along dir is rows, out-extent cols */
  int nin = transpose ? rows : cols, nout = transpose ? cols : rows;
  int s = 1; for (int e = 0; e < dir; ++e) s *= dims[e];
  int outer = 1; for (int e = dir + 1; e < 3; ++e) outer *= dims[e];
  for (int o = 0; o < outer; ++o) for (int i = 0; i < s; ++i) {
    const double *pin = in + (long)o * nin * s + i;
    double *pout = out + (long)o * nout * s + i;
    for (int r = 0; r < nout; ++r) {
      double acc = add ? pout[(long)r * s] : 0.0;
      for (int c = 0; c < nin; ++c) acc += (transpose ? A[c * cols + r] : A[r * cols + c]) * pin[(long)c * s];
      pout[(long)r * s] = acc;
    }
  }
}

#define MAXP (MAXN * MAXN * MAXN)

/* reference gradient of u (nodal, n^3) at the nq^3 quadrature points: g[e][q] = d u / d xi_e */
static void eval_cell_grad(const Basis *b, const double *u, double *g0, double *g1, double *g2)
{
  int n = b->n, nq = b->nq;
  double a[MAXP], bb[MAXP], t[MAXP];
  int d0[3] = { n, n, n }, d1[3] = { nq, n, n }, d2[3] = { nq, nq, n };
  apply1d(b->S, nq, n, 0, d0, u, a, 0, 0);  /* Sx u */
  apply1d(b->D, nq, n, 0, d0, u, bb, 0, 0); /* Dx u */
  apply1d(b->S, nq, n, 1, d1, bb, t, 0, 0); apply1d(b->S, nq, n, 2, d2, t, g0, 0, 0);
  apply1d(b->D, nq, n, 1, d1, a, t, 0, 0);  apply1d(b->S, nq, n, 2, d2, t, g1, 0, 0);
  apply1d(b->S, nq, n, 1, d1, a, t, 0, 0);  apply1d(b->D, nq, n, 2, d2, t, g2, 0, 0);
}

/* values of u at quadrature points */
static void eval_cell_val(const Basis *b, const double *u, double *v)
{
  int n = b->n, nq = b->nq; double a[MAXP], t[MAXP];
  int d0[3] = { n, n, n }, d1[3] = { nq, n, n }, d2[3] = { nq, nq, n };
  apply1d(b->S, nq, n, 0, d0, u, a, 0, 0); apply1d(b->S, nq, n, 1, d1, a, t, 0, 0); apply1d(b->S, nq, n, 2, d2, t, v, 0, 0);
}

/* y_i += sum_q [ d phi_i/d xi_e (q) f_e(q) ] (+ phi_i(q) fv(q) if fv) */
static void integrate_cell(const Basis *b, const double *f0, const double *f1, const double *f2, const double *fv, double *y)
{
  int n = b->n, nq = b->nq; double t[MAXP], t2[MAXP];
  int q3[3] = { nq, nq, nq }, q2[3] = { nq, nq, n }, q1[3] = { nq, n, n };
  if (f0) { apply1d(b->S, nq, n, 2, q3, f0, t, 1, 0); apply1d(b->S, nq, n, 1, q2, t, t2, 1, 0); apply1d(b->D, nq, n, 0, q1, t2, y, 1, 1); }
  if (f1) { apply1d(b->S, nq, n, 2, q3, f1, t, 1, 0); apply1d(b->D, nq, n, 1, q2, t, t2, 1, 0); apply1d(b->S, nq, n, 0, q1, t2, y, 1, 1); }
  if (f2) { apply1d(b->D, nq, n, 2, q3, f2, t, 1, 0); apply1d(b->S, nq, n, 1, q2, t, t2, 1, 0); apply1d(b->S, nq, n, 0, q1, t2, y, 1, 1); }
  if (fv) { apply1d(b->S, nq, n, 2, q3, fv, t, 1, 0); apply1d(b->S, nq, n, 1, q2, t, t2, 1, 0); apply1d(b->S, nq, n, 0, q1, t2, y, 1, 1); }
}

static void face_dirs(int f, int *d, int *s, int *t1, int *t2)
{
  *d = f / 2; *s = f % 2; *t1 = (*d == 0) ? 1 : 0; *t2 = (*d == 2) ? 1 : 2;
}

/* trace on face f: value and reference gradient (3 comps, in cell reference directions) at nq^2 points */
static void eval_face(const Basis *b, const double *u, int f, double *val, double *g /* [3][nq2] */)
{
  int n = b->n, nq = b->nq, nq2 = nq * nq, d, s, t1, t2; face_dirs(f, &d, &s, &t1, &t2);
  int str[3] = { 1, n, n * n };
  double v2[MAXN * MAXN], dn2[MAXN * MAXN], tmp[MAXN * MAXN];
  for (int jb = 0; jb < n; ++jb) for (int ja = 0; ja < n; ++ja) {
    double v = 0.0, dn = 0.0;
    for (int i = 0; i < n; ++i) { double uu = u[i * str[d] + ja * str[t1] + jb * str[t2]]; v += b->fv[s][i] * uu; dn += b->fd[s][i] * uu; }
    v2[ja + n * jb] = v; dn2[ja + n * jb] = dn;
  }
  int e0[3] = { n, n, 1 }, e1[3] = { nq, n, 1 };
  apply1d(b->S, nq, n, 0, e0, v2, tmp, 0, 0);  apply1d(b->S, nq, n, 1, e1, tmp, val, 0, 0);
  apply1d(b->S, nq, n, 0, e0, dn2, tmp, 0, 0); apply1d(b->S, nq, n, 1, e1, tmp, g + d * nq2, 0, 0);
  apply1d(b->D, nq, n, 0, e0, v2, tmp, 0, 0);  apply1d(b->S, nq, n, 1, e1, tmp, g + t1 * nq2, 0, 0);
  apply1d(b->S, nq, n, 0, e0, v2, tmp, 0, 0);  apply1d(b->D, nq, n, 1, e1, tmp, g + t2 * nq2, 0, 0);
}

/* y_i += sum_q [ phi_i cv(q) + d phi_i / d xi_e cg[e](q) ] on face f */
static void integrate_face(const Basis *b, int f, const double *cv, const double *cg /* [3][nq2] */, double *y)
{
  int n = b->n, nq = b->nq, nq2 = nq * nq, d, s, t1, t2; face_dirs(f, &d, &s, &t1, &t2);
  int str[3] = { 1, n, n * n };
  double v2[MAXN * MAXN], dn2[MAXN * MAXN], tmp[MAXN * MAXN];
  int e0[3] = { nq, nq, 1 }, e1[3] = { nq, n, 1 };
  /* nodal-in-plane coefficients multiplying l_i(s) (v2) and l_i'(s) (dn2) */
  apply1d(b->S, nq, n, 1, e0, cv, tmp, 1, 0);           apply1d(b->S, nq, n, 0, e1, tmp, v2, 1, 0);
  apply1d(b->S, nq, n, 1, e0, cg + t1 * nq2, tmp, 1, 0); apply1d(b->D, nq, n, 0, e1, tmp, v2, 1, 1);
  apply1d(b->D, nq, n, 1, e0, cg + t2 * nq2, tmp, 1, 0); apply1d(b->S, nq, n, 0, e1, tmp, v2, 1, 1);
  apply1d(b->S, nq, n, 1, e0, cg + d * nq2, tmp, 1, 0);  apply1d(b->S, nq, n, 0, e1, tmp, dn2, 1, 0);
  for (int jb = 0; jb < n; ++jb) for (int ja = 0; ja < n; ++ja)
    for (int i = 0; i < n; ++i)
      y[i * str[d] + ja * str[t1] + jb * str[t2]] += b->fv[s][i] * v2[ja + n * jb] + b->fd[s][i] * dn2[ja + n * jb];
}

/* ------------------------------------------------------------------------- */
/* SIPG Laplace pieces (laplace_operator.cpp)                                */
/* ------------------------------------------------------------------------- */

/* cell_loop body: gather_evaluate(gradients) -> do_cell_integral -> integrate_scatter  (operator_base.cpp:1349-1370) */
static void cell_integral(const Op *op, long c, const double *u, double *y)
{
  const Basis *b = &op->b; int nq3 = b->nq * b->nq * b->nq;
  double g[3][MAXP];
  eval_cell_grad(b, u, g[0], g[1], g[2]);
  for (int q = 0; q < nq3; ++q) {
    const double *Ji = &op->Jinv_c[(c * nq3 + q) * 9];
    double w = op->JxW_c[c * nq3 + q];
    /* get_gradient: grad_x = J^{-T} grad_xi ; submit_gradient: flux_xi = J^{-1} grad_x * JxW (laplace_operator.cpp:135) */
    double gx[3];
    for (int j = 0; j < 3; ++j) gx[j] = Ji[0 * 3 + j] * g[0][q] + Ji[1 * 3 + j] * g[1][q] + Ji[2 * 3 + j] * g[2][q];
    for (int i = 0; i < 3; ++i) g[i][q] = (Ji[i * 3 + 0] * gx[0] + Ji[i * 3 + 1] * gx[1] + Ji[i * 3 + 2] * gx[2]) * w;
  }
  integrate_cell(b, g[0], g[1], g[2], NULL, y);
}

/* value and normal derivative (w.r.t. normal nrm) of cell c's function on its face f */
static void face_value_and_normal_derivative(const Op *op, long c, int f, const double *u, const double *nrm /* [nq2][3] */, double *val, double *dn)
{
  const Basis *b = &op->b; int nq2 = b->nq * b->nq;
  double g[3 * MAXN * MAXN];
  eval_face(b, u, f, val, g);
  for (int q = 0; q < nq2; ++q) {
    const double *Ji = &op->Jinv_f[((c * 6 + f) * nq2 + q) * 9];
    double s = 0.0;
    for (int j = 0; j < 3; ++j) {
      double gx = Ji[0 * 3 + j] * g[0 * nq2 + q] + Ji[1 * 3 + j] * g[1 * nq2 + q] + Ji[2 * 3 + j] * g[2 * nq2 + q];
      s += gx * nrm[q * 3 + j];
    }
    dn[q] = s;
  }
}

/* submit_normal_derivative(gf) + submit_value(sv) on (c,f) w.r.t. normal nrm, then integrate (adds into y) */
static void face_submit_integrate(const Op *op, long c, int f, const double *nrm, const double *gf, const double *sv, double *y)
{
  const Basis *b = &op->b; int nq2 = b->nq * b->nq;
  double cv[MAXN * MAXN], cg[3 * MAXN * MAXN];
  for (int q = 0; q < nq2; ++q) {
    long o = (c * 6 + f) * nq2 + q;
    const double *Ji = &op->Jinv_f[o * 9];
    double w = op->JxW_f[o];
    cv[q] = sv[q] * w;
    for (int e = 0; e < 3; ++e)
      cg[e * nq2 + q] = (Ji[e * 3 + 0] * nrm[q * 3 + 0] + Ji[e * 3 + 1] * nrm[q * 3 + 1] + Ji[e * 3 + 2] * nrm[q * 3 + 2]) * gf[q] * w;
  }
  integrate_face(b, f, cv, cg, y);
}

static double penalty_factor(const Op *op) { return op->ip_factor * (op->k + 1.0) * (op->k + 1.0); } /* interior_penalty_parameter.h:124 */

/* face_loop body for one interior face: do_face_integral (laplace_operator.cpp:139-163) */
static void interior_face_integral(const Op *op, long fi, const double *src, double *dst)
{
  const Basis *b = &op->b; int n3 = b->n * b->n * b->n, nq2 = b->nq * b->nq;
  long cm = op->face_m[fi], cp = op->face_p[fi]; int fm = op->face_fm[fi], fp = op->face_fp[fi];
  const double *nrm = &op->nrm_f[(cm * 6 + fm) * nq2 * 3]; /* n = n^- for both sides */
  double vm[MAXN * MAXN], vp[MAXN * MAXN], dm[MAXN * MAXN], dp[MAXN * MAXN], gf[MAXN * MAXN], vf[MAXN * MAXN], mvf[MAXN * MAXN];
  face_value_and_normal_derivative(op, cm, fm, src + cm * n3, nrm, vm, dm);
  face_value_and_normal_derivative(op, cp, fp, src + cp * n3, nrm, vp, dp);
  double tau = fmax(op->tauK[cm], op->tauK[cp]) * penalty_factor(op); /* laplace_operator.h:128-140 */
  for (int q = 0; q < nq2; ++q) {
    gf[q] = -0.5 * (vm[q] - vp[q]);                              /* laplace_operator.h:180-185 */
    vf[q] = 0.5 * (dm[q] + dp[q]) - tau * (vm[q] - vp[q]);       /* laplace_operator.h:187-197 */
    mvf[q] = -vf[q];
  }
  face_submit_integrate(op, cm, fm, nrm, gf, mvf, dst + cm * n3); /* m: normal_derivative(gf), value(-vf) */
  face_submit_integrate(op, cp, fp, nrm, gf, vf, dst + cp * n3);  /* p: normal_derivative(gf), value(+vf) */
}

/* homogeneous boundary integral (laplace_operator.cpp:221-265 with weak_boundary_conditions.h tables);
   mode 0: full homogeneous operator; used also by the diagonal */
static void boundary_face_integral(const Op *op, long c, int f, const double *u, double *y)
{
  const Basis *b = &op->b; int nq2 = b->nq * b->nq;
  const double *nrm = &op->nrm_f[(c * 6 + f) * nq2 * 3];
  int bt = op->mesh->bt[c * 6 + f];
  double vm[MAXN * MAXN], dm[MAXN * MAXN], gf[MAXN * MAXN], mvf[MAXN * MAXN];
  face_value_and_normal_derivative(op, c, f, u, nrm, vm, dm);
  double tau = op->tauK[c] * penalty_factor(op); /* laplace_operator.h:142-151 */
  for (int q = 0; q < nq2; ++q) {
    double vp, dp;
    if (bt == BT_DIRICHLET) { vp = -vm[q]; dp = dm[q]; }   /* weak_boundary_conditions.h:44,118-121 */
    else { vp = vm[q]; dp = -dm[q]; }                      /* Neumann: :45,124-127 and :207-222 */
    gf[q] = -0.5 * (vm[q] - vp);
    mvf[q] = -(0.5 * (dm[q] + dp) - tau * (vm[q] - vp));
  }
  face_submit_integrate(op, c, f, nrm, gf, mvf, y);
}

/* own-side contribution of interior face (c,f) as seen from cell c (outward normal of c), neighbour
   function given (cell-based view: operator_base.cpp:858-880,1665-1696 with integrator_m = current cell).
   If u_nb == NULL the exterior function is zero (do_face_int_integral, laplace_operator.cpp:165-191). */
static void face_integral_cellwise(const Op *op, long c, int f, const double *u, const double *u_nb, double *y)
{
  const Basis *b = &op->b; int nq2 = b->nq * b->nq;
  const double *nrm = &op->nrm_f[(c * 6 + f) * nq2 * 3];
  long p = op->mesh->nb[c * 6 + f]; int fp = op->mesh->nbface[c * 6 + f];
  double vm[MAXN * MAXN], vp[MAXN * MAXN], dm[MAXN * MAXN], dp[MAXN * MAXN], gf[MAXN * MAXN], mvf[MAXN * MAXN];
  face_value_and_normal_derivative(op, c, f, u, nrm, vm, dm);
  if (u_nb) face_value_and_normal_derivative(op, p, fp, u_nb, nrm, vp, dp);
  else for (int q = 0; q < nq2; ++q) { vp[q] = 0.0; dp[q] = 0.0; }
  double tau = fmax(op->tauK[c], op->tauK[p]) * penalty_factor(op);
  for (int q = 0; q < nq2; ++q) {
    gf[q] = -0.5 * (vm[q] - vp[q]);
    mvf[q] = -(0.5 * (dm[q] + dp[q]) - tau * (vm[q] - vp[q]));
  }
  face_submit_integrate(op, c, f, nrm, gf, mvf, y);
}

/* ------------------------------------------------------------------------- */
/* Public API                                                                */
/* ------------------------------------------------------------------------- */

static Op *op_from_mesh(Mesh *M, int degree, double ip_factor)
{
  Op *op = (Op *)calloc(1, sizeof(Op));
  op->k = degree; op->mesh = M; op->ip_factor = ip_factor;
  basis_init(&op->b, degree, degree + 1); /* QGauss(k+1): I/operators/quadrature.h:45 */
  op->n_cells = M->n_cells; op->n_dofs = M->n_cells * op->b.n * op->b.n * op->b.n;
  op_setup_geometry(op);
  op_build_faces(op);
  return op;
}

/* Lean operator for the vectorised CPU baseline at benchmark size (bench.py, 96^3 cells): periodic Cartesian box only.  Holds the
 * mesh connectivity and tau_K = sum_d 1/h_d (interior_penalty_parameter.h:68-98 on a uniform box with weights 1/2 on all six
 * faces) - none of the per-quadrature-point geometry, which would take 36 KB per cell.  Only orc_vmult_fast, orc_n_dofs and
 * orc_n_cells may be called on it (everything else needs the stored geometry and is refused by op->lean). */
void *orc_create_periodic_box_lean(int degree, int n_sub, int refine, double ip_factor)
{
  if (degree < 1 || degree + 1 > MAXN - 3) return NULL;
  const int bc[6] = {0, 0, 0, 0, 0, 0};
  Mesh *M = mesh_hypercube(n_sub, refine, 1, 0.0, 2, bc);
  Op *op = (Op *)calloc(1, sizeof(Op));
  op->k = degree; op->mesh = M; op->ip_factor = ip_factor; op->lean = 1;
  basis_init(&op->b, degree, degree + 1);
  op->n_cells = M->n_cells; op->n_dofs = M->n_cells * op->b.n * op->b.n * op->b.n;
  /* cell 0: vertices 0 and 7 of the trilinear mapping span the box */
  const double *X = M->xmap;
  double tk = 0.0;
  for (int e = 0; e < 3; ++e) { op->fast_h[e] = X[7 * 3 + e] - X[e]; tk += 1.0 / op->fast_h[e]; }
  op->tauK = (double *)malloc(sizeof(double) * op->n_cells);
  for (long c = 0; c < op->n_cells; ++c) op->tauK[c] = tk;
  op->fast_checked = 1; op->fast_ok = 1;
  return op;
}

void *orc_create_hypercube(int degree, int n_sub, int refine, int mapping_degree, double deformation, int frequency, const int *bc, double ip_factor)
{
  if (degree < 1 || degree + 1 > MAXN - 3 || mapping_degree < 1 || mapping_degree + 1 > MAXN) return NULL;
  Mesh *M = mesh_hypercube(n_sub, refine, mapping_degree, deformation, frequency, bc);
  return op_from_mesh(M, degree, ip_factor);
}

/* generic mesh: arrays are copied */
void *orc_create(int degree, int mapping_degree, long n_cells, const double *xmap, const long *nb, const unsigned char *nbface, const unsigned char *bt, double ip_factor)
{
  Mesh *M = (Mesh *)calloc(1, sizeof(Mesh));
  int np3 = (mapping_degree + 1) * (mapping_degree + 1) * (mapping_degree + 1);
  M->m = mapping_degree; M->n_cells = n_cells;
  M->xmap = (double *)malloc(sizeof(double) * n_cells * np3 * 3); memcpy(M->xmap, xmap, sizeof(double) * n_cells * np3 * 3);
  M->nb = (long *)malloc(sizeof(long) * n_cells * 6); memcpy(M->nb, nb, sizeof(long) * n_cells * 6);
  M->nbface = (unsigned char *)malloc(n_cells * 6); memcpy(M->nbface, nbface, n_cells * 6);
  M->bt = (unsigned char *)malloc(n_cells * 6); memcpy(M->bt, bt, n_cells * 6);
  return op_from_mesh(M, degree, ip_factor);
}

void orc_destroy(void *h)
{
  Op *op = (Op *)h; if (!op) return;
  mesh_free(op->mesh);
  free(op->Jinv_c); free(op->JxW_c); free(op->xq_c); free(op->Jinv_f); free(op->nrm_f); free(op->JxW_f); free(op->xq_f); free(op->tauK);
  free(op->face_m); free(op->face_p); free(op->face_fm); free(op->face_fp); free(op->bface_c); free(op->bface_f);
  free(op);
}

long orc_n_dofs(void *h) { return ((Op *)h)->n_dofs; }
long orc_n_cells(void *h) { return ((Op *)h)->n_cells; }
long orc_n_interior_faces(void *h) { return ((Op *)h)->n_faces; }
long orc_n_boundary_faces(void *h) { return ((Op *)h)->n_bfaces; }
/* MassKernel (I/operators/mass_kernel.h:32-93): dst = (phi_i, u) with QGauss(k+1): gather_evaluate(values) -> submit_value(u JxW)
 * -> integrate_scatter; one component.  The Helmholtz / viscous operator of the incompressible Navier-Stokes module in Laplace
 * formulation with constant viscosity is scaling_factor_mass * M + nu * A_SIPG, component by component (SURVEY 8 f-3;
 * momentum_operator.cpp:376-426, viscous_operator.h:365-386, 489-560). */
void orc_mass_vmult(void *h, double *dst, const double *src)
{
  Op *op = (Op *)h; const Basis *b = &op->b; int n3 = b->n * b->n * b->n, nq3 = b->nq * b->nq * b->nq;
#pragma omp parallel for schedule(static)
  for (long c = 0; c < op->n_cells; ++c) {
    double v[MAXP];
    eval_cell_val(b, src + c * n3, v);
    for (int q = 0; q < nq3; ++q) v[q] *= op->JxW_c[c * nq3 + q];
    for (int i = 0; i < n3; ++i) dst[c * n3 + i] = 0.0;
    integrate_cell(b, NULL, NULL, NULL, v, dst + c * n3);
  }
}
void orc_get_cell_jxw(void *h, double *jxw) { Op *op = (Op *)h; int nq3 = op->b.nq * op->b.nq * op->b.nq; memcpy(jxw, op->JxW_c, sizeof(double) * op->n_cells * nq3); }
void orc_get_tau(void *h, double *tau) { Op *op = (Op *)h; memcpy(tau, op->tauK, sizeof(double) * op->n_cells); }
void orc_get_mesh(void *h, double *xmap, long *nb, unsigned char *nbface, unsigned char *bt)
{
  Op *op = (Op *)h; Mesh *M = op->mesh; int np3 = (M->m + 1) * (M->m + 1) * (M->m + 1);
  if (xmap) memcpy(xmap, M->xmap, sizeof(double) * M->n_cells * np3 * 3);
  if (nb) memcpy(nb, M->nb, sizeof(long) * M->n_cells * 6);
  if (nbface) memcpy(nbface, M->nbface, M->n_cells * 6);
  if (bt) memcpy(bt, M->bt, M->n_cells * 6);
}

/* apply_add: the three loops of MatrixFree::loop, face-centric (operator_base.cpp:312-354). Serial. */
void orc_vmult_add(void *h, double *dst, const double *src)
{
  Op *op = (Op *)h; int n3 = op->b.n * op->b.n * op->b.n;
  if (op->lean) { fprintf(stderr, "oracle: operation needs the stored geometry (lean operator)\n"); abort(); }
  for (long c = 0; c < op->n_cells; ++c) cell_integral(op, c, src + c * n3, dst + c * n3);
  for (long f = 0; f < op->n_faces; ++f) interior_face_integral(op, f, src, dst);
  for (long f = 0; f < op->n_bfaces; ++f) {
    long c = op->bface_c[f];
    boundary_face_integral(op, c, op->bface_f[f], src + c * n3, dst + c * n3);
  }
}

/* vmult -> apply: zero dst, then loops (operator_base.cpp:156-168, 264-310) */
void orc_vmult(void *h, double *dst, const double *src)
{
  Op *op = (Op *)h;
  memset(dst, 0, sizeof(double) * op->n_dofs);
  orc_vmult_add(h, dst, src);
}

/* Same operator evaluated cell by cell (each cell adds its own side of its 6 faces); OpenMP over cells.
   Used as the threaded CPU baseline and cross-checked against orc_vmult in the tests. */
void orc_vmult_cellwise(void *h, double *dst, const double *src, int n_threads)
{
  Op *op = (Op *)h; int n3 = op->b.n * op->b.n * op->b.n;
  if (op->lean) { fprintf(stderr, "oracle: operation needs the stored geometry (lean operator)\n"); abort(); }
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
  for (long c = 0; c < op->n_cells; ++c) {
    double y[MAXP];
    for (int i = 0; i < n3; ++i) y[i] = 0.0;
    cell_integral(op, c, src + c * n3, y);
    for (int f = 0; f < 6; ++f) {
      long p = op->mesh->nb[c * 6 + f];
      if (p >= 0) face_integral_cellwise(op, c, f, src + c * n3, src + p * n3, y);
      else boundary_face_integral(op, c, f, src + c * n3, y);
    }
    for (int i = 0; i < n3; ++i) dst[c * n3 + i] = y[i];
  }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* calculate_diagonal (operator_base.cpp:608-646): column-by-column, cell term + own side of each
   face with exterior function zero (laplace_operator.cpp:165-219), boundary faces homogeneous. */
void orc_calculate_diagonal(void *h, double *diag)
{
  Op *op = (Op *)h; int n3 = op->b.n * op->b.n * op->b.n;
  if (op->lean) { fprintf(stderr, "oracle: operation needs the stored geometry (lean operator)\n"); abort(); }
#pragma omp parallel for schedule(static)
  for (long c = 0; c < op->n_cells; ++c) {
    double e[MAXP], y[MAXP];
    for (int j = 0; j < n3; ++j) {
      for (int i = 0; i < n3; ++i) { e[i] = 0.0; y[i] = 0.0; }
      e[j] = 1.0;
      cell_integral(op, c, e, y);
      for (int f = 0; f < 6; ++f) {
        if (op->mesh->nb[c * 6 + f] >= 0) face_integral_cellwise(op, c, f, e, NULL, y);
        else boundary_face_integral(op, c, f, e, y);
      }
      diag[c * n3 + j] = y[j];
    }
  }
}

/* invert_diagonal.h:35-46 */
void orc_invert_diagonal(double *d, long n)
{
  for (long i = 0; i < n; ++i) d[i] = (fabs(d[i]) > 1.0e-10) ? 1.0 / d[i] : 1.0;
}

void orc_calculate_inverse_diagonal(void *h, double *diag)
{
  orc_calculate_diagonal(h, diag);
  orc_invert_diagonal(diag, ((Op *)h)->n_dofs);
}

/* ------------------------------------------------------------------------- */
/* Vector helpers (fixed, left-to-right summation order: the oracle is serial */
/* here on purpose so residual histories are reproducible)                    */
/* ------------------------------------------------------------------------- */
static double vdot(const double *a, const double *b, long n)
{
  /* pairwise-blocked summation to keep round-off growth small for 1e8-long vectors */
  double total = 0.0;
  for (long i0 = 0; i0 < n; i0 += 4096) {
    long i1 = i0 + 4096 < n ? i0 + 4096 : n; double s = 0.0;
    for (long i = i0; i < i1; ++i) s += a[i] * b[i];
    total += s;
  }
  return total;
}

typedef void (*apply_fn)(void *ctx, double *dst, const double *src);

/* ------------------------------------------------------------------------- */
/* dealii::PreconditionChebyshev restated             */
/* ------------------------------------------------------------------------- */
typedef struct {
  Op *op; double *inv_diag; int degree; double smoothing_range; int eig_cg_n_iterations;
  double lambda_max_est, lambda_min_est; /* raw CG/Lanczos estimates */
  double theta, delta;                   /* centre and half-width of the smoothing interval */
  int use_cellwise;
} Cheb;

static void op_apply(Op *op, double *dst, const double *src, int cellwise)
{
  if (cellwise) orc_vmult_cellwise(op, dst, src, 0); else orc_vmult(op, dst, src);
}

/* eigenvalues of the symmetric tridiagonal matrix (diag a[0..n), offdiag b[0..n-1)) by bisection: largest & smallest */
static void tridiag_extreme_eigs(const double *a, const double *b, int n, double *emin, double *emax)
{
  double lo = a[0], hi = a[0];
  for (int i = 0; i < n; ++i) {
    double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < n - 1 ? fabs(b[i]) : 0.0);
    if (a[i] - r < lo) lo = a[i] - r;
    if (a[i] + r > hi) hi = a[i] + r;
  }
  for (int which = 0; which < 2; ++which) {
    int target = which == 0 ? 1 : n; /* number of eigenvalues < x must reach `target` */
    double l = lo, u = hi;
    for (int it = 0; it < 200; ++it) {
      double x = 0.5 * (l + u);
      int cnt = 0; double q = 1.0;
      for (int i = 0; i < n; ++i) {
        double bb = i > 0 ? b[i - 1] * b[i - 1] : 0.0;
        q = a[i] - x - (i > 0 ? bb / q : 0.0);
        if (q == 0.0) q = 1e-300;
        if (q < 0.0) ++cnt;
      }
      if (cnt >= target) u = x; else l = x;
    }
    if (which == 0) *emin = 0.5 * (l + u); else *emax = 0.5 * (l + u);
  }
}

/* dealii::SolverCG with ReductionControl; optional Jacobi / Chebyshev preconditioner; optional Lanczos
   coefficient capture for eigenvalue estimates.  Returns last_step(); res_hist[it] = ||g_it||_2. */
static int cg_solve(Op *op, double *x, const double *b, int precond /*0 identity, 1 jacobi, 2 chebyshev*/, const double *inv_diag, Cheb *cheb,
                    double abs_tol, double rel_tol, int max_it, int cellwise, double *res_hist, int hist_len,
                    double *lanczos_a, double *lanczos_b, int *lanczos_n, int *converged);

static void cheb_vmult(Cheb *ch, double *dst, const double *src);

static void precond_apply(int precond, const double *inv_diag, Cheb *cheb, double *h, const double *g, long n)
{
  if (precond == 1) for (long i = 0; i < n; ++i) h[i] = inv_diag[i] * g[i]; /* jacobi_preconditioner.h:50-62 */
  else if (precond == 2) cheb_vmult(cheb, h, g);
}

static int cg_solve(Op *op, double *x, const double *b, int precond, const double *inv_diag, Cheb *cheb,
                    double abs_tol, double rel_tol, int max_it, int cellwise, double *res_hist, int hist_len,
                    double *lanczos_a, double *lanczos_b, int *lanczos_n, int *converged)
{
  long n = op->n_dofs;
  double *g = (double *)malloc(sizeof(double) * n), *d = (double *)malloc(sizeof(double) * n), *hh = (double *)malloc(sizeof(double) * n);
  int it = 0, x_zero = 1;
  for (long i = 0; i < n; ++i) if (x[i] != 0.0) { x_zero = 0; break; }
  if (!x_zero) { op_apply(op, g, x, cellwise); for (long i = 0; i < n; ++i) g[i] -= b[i]; }
  else for (long i = 0; i < n; ++i) g[i] = -b[i];
  double res = sqrt(vdot(g, g, n)), res0 = res, reduced_tol = rel_tol * res0;
  if (res_hist && hist_len > 0) res_hist[0] = res;
  int state = 0; /* 0 iterate, 1 success, 2 failure */
  /* ReductionControl::check then SolverControl::check */
  if (res < reduced_tol || res <= abs_tol) state = 1; else if (0 >= max_it || isnan(res)) state = 2;
  double gh = 0.0, alpha_prev = 0.0, beta_prev = 0.0; int nl = 0;
  if (state == 0) {
    if (precond) { precond_apply(precond, inv_diag, cheb, hh, g, n); for (long i = 0; i < n; ++i) d[i] = -hh[i]; gh = vdot(g, hh, n); }
    else { for (long i = 0; i < n; ++i) d[i] = -g[i]; gh = res * res; }
  }
  while (state == 0) {
    ++it;
    op_apply(op, hh, d, cellwise);
    double alpha = vdot(d, hh, n);
    alpha = gh / alpha;
    for (long i = 0; i < n; ++i) x[i] += alpha * d[i];
    for (long i = 0; i < n; ++i) g[i] += alpha * hh[i];
    res = sqrt(vdot(g, g, n));
    if (res_hist && it < hist_len) res_hist[it] = res;
    /* Lanczos tridiagonal from CG coefficients (dealii SolverCG eigenvalue slot) */
    if (lanczos_a) {
      lanczos_a[nl] = 1.0 / alpha + (nl > 0 ? beta_prev / alpha_prev : 0.0);
    }
    if (res < reduced_tol || res <= abs_tol) state = 1; else if (it >= max_it || isnan(res)) state = 2;
    if (state != 0) { if (lanczos_a) ++nl; break; }
    double beta = gh;
    if (precond) { precond_apply(precond, inv_diag, cheb, hh, g, n); gh = vdot(g, hh, n); beta = gh / beta; for (long i = 0; i < n; ++i) d[i] = beta * d[i] - hh[i]; }
    else { gh = res * res; beta = gh / beta; for (long i = 0; i < n; ++i) d[i] = beta * d[i] - g[i]; }
    if (lanczos_a) { lanczos_b[nl] = sqrt(beta) / alpha; ++nl; }
    alpha_prev = alpha; beta_prev = beta;
  }
  if (lanczos_n) *lanczos_n = nl;
  if (converged) *converged = (state == 1);
  free(g); free(d); free(hh);
  return it;
}

int orc_cg(void *h, double *x, const double *b, int use_jacobi, double abs_tol, double rel_tol, int max_it, int cellwise, double *res_hist, int hist_len, int *converged)
{
  Op *op = (Op *)h; double *inv_diag = NULL;
  if (use_jacobi) { inv_diag = (double *)malloc(sizeof(double) * op->n_dofs); orc_calculate_inverse_diagonal(h, inv_diag); }
  int it = cg_solve(op, x, b, use_jacobi ? 1 : 0, inv_diag, NULL, abs_tol, rel_tol, max_it, cellwise, res_hist, hist_len, NULL, NULL, NULL, converged);
  free(inv_diag);
  return it;
}

/* PreconditionChebyshev::estimate_eigenvalues + parameters as set in chebyshev_smoother.h:149-172 */
void *orc_cheb_create(void *h, int degree, double smoothing_range, int eig_cg_n_iterations, int cellwise)
{
  Op *op = (Op *)h; long n = op->n_dofs;
  Cheb *ch = (Cheb *)calloc(1, sizeof(Cheb));
  ch->op = op; ch->degree = degree; ch->smoothing_range = smoothing_range; ch->eig_cg_n_iterations = eig_cg_n_iterations; ch->use_cellwise = cellwise;
  ch->inv_diag = (double *)malloc(sizeof(double) * n);
  orc_calculate_inverse_diagonal(h, ch->inv_diag);
  /* start vector: (global index mod 11), mean removed */
  double *rhs = (double *)malloc(sizeof(double) * n), *sol = (double *)calloc(n, sizeof(double));
  double mean = 0.0;
  for (long i = 0; i < n; ++i) { rhs[i] = (double)(i % 11); mean += rhs[i]; }
  mean /= (double)n;
  for (long i = 0; i < n; ++i) rhs[i] -= mean;
  double la[256], lb[256]; int nl = 0, conv = 0;
  int its = eig_cg_n_iterations < 250 ? eig_cg_n_iterations : 250;
  /* ReductionControl(eig_cg_n_iterations, sqrt(eps), eig_cg_residual = 1e-2) */
  cg_solve(op, sol, rhs, 1, ch->inv_diag, NULL, 1.4901161193847656e-08, 1e-2, its, cellwise, NULL, 0, la, lb, &nl, &conv);
  if (nl > 0) tridiag_extreme_eigs(la, lb, nl, &ch->lambda_min_est, &ch->lambda_max_est);
  else { ch->lambda_min_est = 1.0; ch->lambda_max_est = 1.0; }
  double max_ev = 1.2 * ch->lambda_max_est;
  double alpha = smoothing_range > 1.0 ? max_ev / smoothing_range : fmin(0.9 * max_ev, ch->lambda_min_est);
  ch->delta = 0.5 * (max_ev - alpha); ch->theta = 0.5 * (max_ev + alpha);
  free(rhs); free(sol);
  return ch;
}

void orc_cheb_destroy(void *c) { Cheb *ch = (Cheb *)c; if (!ch) return; free(ch->inv_diag); free(ch); }
void orc_cheb_get(void *c, double *out4) { Cheb *ch = (Cheb *)c; out4[0] = ch->lambda_min_est; out4[1] = ch->lambda_max_est; out4[2] = ch->theta; out4[3] = ch->delta; }
void orc_cheb_set(void *c, double theta, double delta) { Cheb *ch = (Cheb *)c; ch->theta = theta; ch->delta = delta; }

/* shared recurrence; zero_start: vmult (x0 = 0), else step (x0 = dst) */
static void cheb_run(Cheb *ch, double *x, const double *b, int zero_start)
{
  Op *op = ch->op; long n = op->n_dofs;
  double *xold = (double *)malloc(sizeof(double) * n), *r = (double *)malloc(sizeof(double) * n);
  double theta = ch->theta, delta = ch->delta;
  /* first update */
  if (zero_start) {
    for (long i = 0; i < n; ++i) { xold[i] = 0.0; x[i] = ch->inv_diag[i] * b[i] / theta; }
  } else {
    op_apply(op, r, x, ch->use_cellwise);
    for (long i = 0; i < n; ++i) { double xi = x[i]; xold[i] = xi; x[i] = xi + ch->inv_diag[i] * (b[i] - r[i]) / theta; }
  }
  if (ch->degree < 2 || fabs(delta) < 1e-40) { free(xold); free(r); return; }
  double rhok = delta / theta, sigma = theta / delta;
  for (int k = 0; k < ch->degree - 1; ++k) {
    op_apply(op, r, x, ch->use_cellwise);
    double rhokp = 1.0 / (2.0 * sigma - rhok);
    double factor1 = rhokp * rhok, factor2 = 2.0 * rhokp / delta;
    rhok = rhokp;
    for (long i = 0; i < n; ++i) {
      double xi = x[i];
      x[i] = xi + factor1 * (xi - xold[i]) + factor2 * ch->inv_diag[i] * (b[i] - r[i]);
      xold[i] = xi;
    }
  }
  free(xold); free(r);
}
static void cheb_vmult(Cheb *ch, double *dst, const double *src) { cheb_run(ch, dst, src, 1); }
void orc_cheb_vmult(void *c, double *dst, const double *src) { cheb_run((Cheb *)c, dst, src, 1); }
void orc_cheb_step(void *c, double *dst, const double *src) { cheb_run((Cheb *)c, dst, src, 0); }

int orc_cg_chebyshev(void *h, void *c, double *x, const double *b, double abs_tol, double rel_tol, int max_it, int cellwise, double *res_hist, int hist_len, int *converged)
{
  return cg_solve((Op *)h, x, b, 2, NULL, (Cheb *)c, abs_tol, rel_tol, max_it, cellwise, res_hist, hist_len, NULL, NULL, NULL, converged);
}

/* ------------------------------------------------------------------------- */
/* applications/poisson/sine: rhs and L2 error (golden-output pin)           */
/* ------------------------------------------------------------------------- */
static const double SINE_FREQ = 3.0 * M_PI; /* application.h:32 */
static double sine_solution(const double *p) { return sin(SINE_FREQ * p[0]) * sin(SINE_FREQ * p[1]) * sin(SINE_FREQ * p[2]); }
static double sine_neumann(const double *p) { return SINE_FREQ * cos(SINE_FREQ * p[0]) * sin(SINE_FREQ * p[1]) * sin(SINE_FREQ * p[2]); }
static double sine_rhs(const double *p) { return SINE_FREQ * SINE_FREQ * 3.0 * sine_solution(p); }

/* Poisson::Operator::rhs (operator.cpp:414-423): -(inhomogeneous boundary integrals) + (f, v) */
void orc_rhs_sine(void *h, double *rhs)
{
  Op *op = (Op *)h; const Basis *b = &op->b; int n3 = b->n * b->n * b->n, nq2 = b->nq * b->nq, nq3 = nq2 * b->nq;
  memset(rhs, 0, sizeof(double) * op->n_dofs);
  for (long c = 0; c < op->n_cells; ++c) {
    double fv[MAXP];
    for (int q = 0; q < nq3; ++q) fv[q] = sine_rhs(&op->xq_c[(c * nq3 + q) * 3]) * op->JxW_c[c * nq3 + q];
    integrate_cell(b, NULL, NULL, NULL, fv, rhs + c * n3);
  }
  for (long fi = 0; fi < op->n_bfaces; ++fi) {
    long c = op->bface_c[fi]; int f = op->bface_f[fi]; int bt = op->mesh->bt[c * 6 + f];
    const double *nrm = &op->nrm_f[(c * 6 + f) * nq2 * 3];
    double tau = op->tauK[c] * penalty_factor(op);
    double gf[MAXN * MAXN], sv[MAXN * MAXN], tmp[MAXP];
    for (int q = 0; q < nq2; ++q) {
      const double *xq = &op->xq_f[((c * 6 + f) * nq2 + q) * 3];
      /* inhomogeneous operator: u^- = 0, grad u^- n = 0 (weak_boundary_conditions.h tables) */
      double vm = 0.0, dm = 0.0, vp, dp;
      if (bt == BT_DIRICHLET) { vp = -vm + 2.0 * sine_solution(xq); dp = dm; }
      else { vp = vm; dp = -dm + 2.0 * sine_neumann(xq); }
      gf[q] = -0.5 * (vm - vp);
      sv[q] = -(0.5 * (dm + dp) - tau * (vm - vp));
    }
    for (int i = 0; i < n3; ++i) tmp[i] = 0.0;
    face_submit_integrate(op, c, f, nrm, gf, sv, tmp);
    for (int i = 0; i < n3; ++i) rhs[c * n3 + i] -= tmp[i]; /* shifted to the right-hand side (operator_base.cpp:533-535) */
  }
}

/* relative L2 error with Gauss(k+3) (error_calculation.cpp:36-115) */
double orc_l2_error_sine(void *h, const double *u)
{
  Op *op = (Op *)h; const Mesh *M = op->mesh;
  Basis be; basis_init(&be, op->k, op->k + 3);
  int n3 = be.n * be.n * be.n, nq = be.nq;
  double gl[MAXN]; gauss_lobatto(M->m + 1, gl);
  double err2 = 0.0, nrm2 = 0.0;
  for (long c = 0; c < op->n_cells; ++c) {
    double v[MAXP];
    eval_cell_val(&be, u + c * n3, v);
    double e_c = 0.0, n_c = 0.0;
    for (int q2 = 0; q2 < nq; ++q2) for (int q1 = 0; q1 < nq; ++q1) for (int q0 = 0; q0 < nq; ++q0) {
      double xi[3] = { be.xq[q0], be.xq[q1], be.xq[q2] }, J[9], x[3];
      mapping_eval(M, c, gl, xi, J, x);
      double w = det3(J) * be.w[q0] * be.w[q1] * be.w[q2];
      double ex = sine_solution(x), df = v[q0 + nq * (q1 + nq * q2)] - ex;
      e_c += df * df * w; n_c += ex * ex * w;
    }
    err2 += e_c; nrm2 += n_c;
  }
  return sqrt(err2) / sqrt(nrm2);
}

/* physical coordinates of all nodal (Gauss-Lobatto) DoF positions: xyz[dof][3] */
void orc_dof_coordinates(void *h, double *xyz)
{
  Op *op = (Op *)h; const Mesh *M = op->mesh; const Basis *b = &op->b; int n = b->n, n3 = n * n * n;
  double gl[MAXN]; gauss_lobatto(M->m + 1, gl);
  for (long c = 0; c < op->n_cells; ++c)
    for (int i2 = 0; i2 < n; ++i2) for (int i1 = 0; i1 < n; ++i1) for (int i0 = 0; i0 < n; ++i0) {
      double xi[3] = { b->xn[i0], b->xn[i1], b->xn[i2] }, J[9];
      mapping_eval(M, c, gl, xi, J, &xyz[(c * n3 + i0 + n * (i1 + n * i2)) * 3]);
    }
}

/* 1-D tables for tests: xn[n], xq[n], w[n], S[n*n], D[n*n], fv[2n], fd[2n] */
void orc_get_basis(int degree, double *xn, double *xq, double *w, double *S, double *D, double *fv, double *fd)
{
  Basis b; basis_init(&b, degree, degree + 1); int n = b.n;
  memcpy(xn, b.xn, sizeof(double) * n); memcpy(xq, b.xq, sizeof(double) * n); memcpy(w, b.w, sizeof(double) * n);
  memcpy(S, b.S, sizeof(double) * n * n); memcpy(D, b.D, sizeof(double) * n * n);
  for (int s = 0; s < 2; ++s) { memcpy(fv + s * n, b.fv[s], sizeof(double) * n); memcpy(fd + s * n, b.fd[s], sizeof(double) * n); }
}

/* vectorised CPU baseline for uniform boxes (bench.py's cpu_baseline / --impl reference legs) */
#include "sipg_fast.inc"
