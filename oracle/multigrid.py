"""CPU restatement of the reference's multigrid preconditioner on DG levels -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Follows
  MultigridPreconditionerBase::initialize_levels   I/solvers_and_preconditioners/multigrid/multigrid_preconditioner_base.cpp:97-323
  MultigridAlgorithm::vmult / v_cycle              I/solvers_and_preconditioners/multigrid/multigrid_algorithm.h:88-109, 173-243
  ChebyshevSmoother (point Jacobi)                 I/solvers_and_preconditioners/multigrid/smoothers/chebyshev_smoother.h:79-172
  MGCoarseKrylov (CG + point Jacobi)               I/solvers_and_preconditioners/multigrid/coarse_grid_solvers.h:62-232
  MGTransfer (dealii::MGTwoLevelTransfer)          I/solvers_and_preconditioners/multigrid/transfer.cpp:28-69
for the multigrid types that stay in the DG space (hMG, pMG, hpMG, phMG).  Level operators, smoothers and the coarse CG are the
oracle's (oracle/sipg_oracle.c); the transfers are numpy tensor products of 1-D embedding matrices.  Arithmetic is FP64 on every
level (the reference instantiates the levels in float, multigrid_preconditioner_base.h:60): parity with the reference's iteration
counts is therefore unpinned; the pin is the converged solution (golden L2 errors of applications/poisson/sine/tests).
"""
import numpy as np

from .oracle import OracleChebyshev, OracleOperator, basis_tables


def initialize_levels(mg_type, p_sequence, degree, n_h_levels):
    """[(h_level, degree)] coarse -> fine (is_dg = true, no c-transfer)."""
    if mg_type == "hMG":
        p_levels = [degree]
    else:
        p_levels, p = [], degree
        while True:
            p_levels.append(p)
            q = {"GoToOne": 1, "DecreaseByOne": max(p - 1, 1), "Bisect": max(p // 2, 1)}[p_sequence]
            if q == p_levels[-1]:
                break
            p = q
        p_levels.reverse()
    hs = list(range(n_h_levels))
    if mg_type == "hMG":
        return [(h, p_levels[0]) for h in hs]
    if mg_type == "pMG":
        return [(hs[-1], p) for p in p_levels]
    if mg_type == "phMG":
        return [(h, p_levels[0]) for h in hs[:-1]] + [(hs[-1], p) for p in p_levels]
    if mg_type == "hpMG":
        return [(hs[0], p) for p in p_levels[:-1]] + [(h, p_levels[-1]) for h in hs]
    raise ValueError("This multigrid type is not implemented!")


def lagrange_matrix(nodes, x):
    """L[i, j] = l_j(x_i) for the Lagrange basis on `nodes`."""
    L = np.ones((len(x), len(nodes)))
    for j in range(len(nodes)):
        for m in range(len(nodes)):
            if m != j:
                L[:, j] *= (x - nodes[m]) / (nodes[j] - nodes[m])
    return L


class Transfer:
    """prolongation = embedding, restriction = transpose; vectors are cell-major, lexicographic (x fastest) inside a cell."""

    def __init__(self, k_fine, k_coarse, h):
        xf, xc = basis_tables(k_fine)["xn"], basis_tables(k_coarse)["xn"]
        self.nf, self.nc, self.h = k_fine + 1, k_coarse + 1, h
        if h:
            assert k_fine == k_coarse
            self.I = [lagrange_matrix(xc, 0.5 * xf), lagrange_matrix(xc, 0.5 + 0.5 * xf)]
        else:
            self.I = [lagrange_matrix(xc, xf)]

    def prolongate_add(self, fine, coarse):
        nf, nc = self.nf, self.nc
        uc = coarse.reshape(-1, nc, nc, nc)  # [cell, z, y, x]
        if self.h:
            uf = fine.reshape(-1, 8, nf, nf, nf)
            for child in range(8):
                Ix, Iy, Iz = self.I[child & 1], self.I[(child >> 1) & 1], self.I[(child >> 2) & 1]
                uf[:, child] += np.einsum("ai,bj,ck,nkji->ncba", Ix, Iy, Iz, uc, optimize=True)
        else:
            Ix = self.I[0]
            fine.reshape(-1, nf, nf, nf)[...] += np.einsum("ai,bj,ck,nkji->ncba", Ix, Ix, Ix, uc, optimize=True)
        return fine

    def restrict_add(self, coarse, fine):
        nf, nc = self.nf, self.nc
        uc = coarse.reshape(-1, nc, nc, nc)
        if self.h:
            uf = fine.reshape(-1, 8, nf, nf, nf)
            for child in range(8):
                Ix, Iy, Iz = self.I[child & 1], self.I[(child >> 1) & 1], self.I[(child >> 2) & 1]
                uc += np.einsum("ai,bj,ck,ncba->nkji", Ix, Iy, Iz, uf[:, child], optimize=True)
        else:
            Ix = self.I[0]
            uc += np.einsum("ai,bj,ck,ncba->nkji", Ix, Ix, Ix, fine.reshape(-1, nf, nf, nf), optimize=True)
        return coarse


class OracleMultigrid:
    def __init__(self, levels, n_sub=1, mapping_degree=1, deformation=0.0, frequency=2, bc=(0,) * 6, ip_factor=1.0,
                 smoother_iterations=5, smoothing_range=20.0, iterations_eigenvalue_estimation=20,
                 coarse_abs_tol=1e-12, coarse_rel_tol=1e-3, coarse_max_iter=10000):
        self.levels = list(levels)
        self.ops = [OracleOperator(k, n_sub, h, mapping_degree, deformation, frequency, bc, ip_factor) for (h, k) in self.levels]
        self.smoothers = [None] + [OracleChebyshev(op, smoother_iterations, smoothing_range, iterations_eigenvalue_estimation) for op in self.ops[1:]]
        self.transfers = [None]
        for (hc, kc), (hf, kf) in zip(self.levels[:-1], self.levels[1:]):
            assert (hc != hf) != (kc != kf), "Between two consecutive multigrid levels, only one type of transfer is allowed."
            self.transfers.append(Transfer(kf, kc, hc != hf))
        self.singular = all(b != 1 for b in bc)
        self.coarse = (coarse_abs_tol, coarse_rel_tol, coarse_max_iter)
        self.solution = [np.zeros(op.n_dofs) for op in self.ops]
        self.coarse_iterations = 0

    def v_cycle(self, level, defect):
        op = self.ops[level]
        if level == 0:
            r = defect[0].copy()
            if self.singular:
                r -= r.mean()
            a, rel, mx = self.coarse
            x, it, _, conv = op.cg(r, x0=self.solution[0], jacobi=True, abs_tol=a, rel_tol=rel, max_it=mx)
            assert conv, "coarse solver did not converge"
            self.coarse_iterations += it
            self.solution[0] = x
            return
        sm = self.smoothers[level]
        self.solution[level] = sm.vmult(defect[level])
        t = defect[level] - op.vmult_cellwise(self.solution[level])
        self.transfers[level].restrict_add(defect[level - 1], t)
        self.v_cycle(level - 1, defect)
        self.transfers[level].prolongate_add(self.solution[level], self.solution[level - 1])
        self.solution[level] = sm.step(self.solution[level], defect[level])

    def vmult(self, src):
        defect = [np.zeros(op.n_dofs) for op in self.ops]
        defect[-1][:] = src
        self.v_cycle(len(self.ops) - 1, defect)
        return self.solution[-1].copy()

    def pcg(self, A, b, abs_tol=1e-20, rel_tol=1e-10, max_it=10000):
        """dealii::SolverCG with this preconditioner (same recurrences as cg_solve of oracle/sipg_oracle.c)."""
        x = np.zeros_like(b)
        g = A.vmult_cellwise(x) - b
        res0 = res = np.sqrt(g @ g)
        hist = [res]
        if res < rel_tol * res0 or res <= abs_tol:
            return x, 0, np.array(hist)
        h = self.vmult(g)
        d = -h
        gh = g @ h
        it = 0
        while True:
            it += 1
            Ad = A.vmult_cellwise(d)
            alpha = gh / (d @ Ad)
            x += alpha * d
            g += alpha * Ad
            res = np.sqrt(g @ g)
            hist.append(res)
            if res < rel_tol * res0 or res <= abs_tol or it >= max_it:
                break
            h = self.vmult(g)
            gh_new = g @ h
            beta = gh_new / gh
            d = beta * d - h
            gh = gh_new
        return x, it, np.array(hist)
