"""Small vmult / diagonal / CG / Chebyshev runs for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(SURVEY section 5): every kernel family once, on meshes small enough for the sanitizer.

    compute-sanitizer --tool racecheck python tests/sanitizer_cases.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import exadg_b200  # noqa: E402

ROUND = os.environ.get("SANITIZER_ROUND", "")
for (degree, n_sub, refine, deformation, bc) in [] if ROUND == "2" else [(4, 3, 0, 0.0, (0,) * 6), (4, 1, 2, 0.0, (0,) * 6), (4, 3, 1, 0.0, (0,) * 6), (2, 5, 0, 0.0, (0,) * 6), (3, 1, 1, 0.0, (0,) * 6),
                                                 (5, 1, 1, 0.0, (0,) * 6), (3, 2, 0, 0.1, (0,) * 6), (2, 2, 0, 0.15, (1, 2, 1, 1, 1, 1))]:
    # k=4 on the affine path has three kernels: 5-warp (32-cell batches), pipelined and warp-specialised (the latter needs <= 64
    # out-of-batch faces per batch: 4^3 and 6^3 cells qualify, 3^3 does not and exercises the fall-back)
    for pipe in (("5warp", "pipe", "ws") if (degree == 4 and deformation == 0.0) else ("default",)):
        if pipe == "5warp":
            os.environ["EXADG_B200_NO_PIPE"] = "1"
        else:
            os.environ.pop("EXADG_B200_NO_PIPE", None)
        exadg_b200.cartesian_kernel(0 if pipe == "pipe" else 1)
        op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, 1, deformation, 2, bc)
        x = torch.rand(op.local_size(), dtype=torch.float64, device="cuda")
        y = op.initialize_dof_vector()
        op.vmult(y, x)
        op.vmult_add(y, x)
        d = op.initialize_dof_vector()
        op.calculate_inverse_diagonal(d)
        if bc != (0,) * 6:
            s = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), exadg_b200.SolverData(5, 0.0, 0.0))
            try:
                s.solve(op.initialize_dof_vector(), y)
            except exadg_b200.ExaDGError:
                pass
            ch = exadg_b200.ChebyshevSmoother(op, 3, 20.0, 4)
            ch.vmult(d, y)
        torch.cuda.synchronize()
        print("ok k=%d cells=%d deformation=%g pipe=%s cartesian=%d" % (degree, op.n_cells_owned, deformation, pipe, op.is_cartesian_path), flush=True)
# ---- round 2: the kernels added since (SANITIZER_ROUND=1 runs only the block above, =2 only this one) ----


import numpy as np  # noqa: E402

if ROUND != "1":
    os.environ.pop("EXADG_B200_NO_PIPE", None)
    # every k=4 kernel variant on 4^3 and 6^3 cells (3 / 9 batches): 3 = default (four producer warps, setmaxnreg), 4 = warp-private,
    # 6 = staged (cp.async ring, single trace buffer, named barriers 3 / 4)
    # (SANITIZER_SKIP_WP=1 leaves the warp-private kernel out: its hand-over between the warps is an mbarrier protocol, which racecheck does not
    # model - it reports the producer's trace stores against the consumer's loads; the protocol is checked by ThreadSanitizer on the emulation)
    for variant in ((3, 6) if os.environ.get("SANITIZER_SKIP_WP") else (3, 4, 6)):
        for (n_sub, refine) in ((1, 2), (3, 1)):
            op = exadg_b200.LaplaceOperator.hypercube(4, n_sub, refine)
            op.set_kernel_variant(variant)
            x = torch.rand(op.local_size(), dtype=torch.float64, device="cuda")
            y = op.initialize_dof_vector()
            op.vmult(y, x)
            op.vmult_add(y, x)
            torch.cuda.synchronize()
            print("ok k=4 variant=%d cells=%d" % (variant, op.n_cells_owned), flush=True)
    # line kernel (k = 5, 6, 7), plane kernel with 64-cell batches (k = 3)
    for (degree, n_sub, refine) in ((5, 1, 2), (6, 3, 0), (7, 1, 1), (3, 1, 2)):
        op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine)
        x = torch.rand(op.local_size(), dtype=torch.float64, device="cuda")
        y = op.initialize_dof_vector()
        op.vmult(y, x)
        torch.cuda.synchronize()
        print("ok line/plane kernel k=%d cells=%d cartesian=%d" % (degree, op.n_cells_owned, op.is_cartesian_path), flush=True)
    # general kernel with warp-owned cells (k = 1..4) and block barriers (k = 5), curved mesh with boundaries; Helmholtz, inverse mass
    WALLS = (1, 1, 2, 1, 0, 0)
    for degree in (1, 2, 3, 4, 5):
        op = exadg_b200.LaplaceOperator.hypercube(degree, 2, 0, 2, 0.1, 2, WALLS)
        x = torch.rand(op.local_size(), dtype=torch.float64, device="cuda")
        y = op.initialize_dof_vector()
        op.vmult(y, x)
        op.calculate_diagonal(y)
        hel = exadg_b200.LaplaceOperator.hypercube_helmholtz(degree, 3, 10.0, 0.1, 2, 0, 2, 0.1, 2, WALLS)
        xv = torch.rand(hel.local_size(), dtype=torch.float64, device="cuda")
        yv = hel.initialize_dof_vector()
        hel.vmult(yv, xv)
        hel.inverse_mass_vmult(yv, xv)
        torch.cuda.synchronize()
        print("ok general kernel / Helmholtz / inverse mass k=%d" % degree, flush=True)
    # hybrid launch of a bounded uniform box (k = 2, 16^3 cells), rhs / evaluate / L2 error kernels
    op = exadg_b200.LaplaceOperator.hypercube(2, 1, 4, 1, 0.0, 2, (1, 2, 1, 1, 1, 1))
    x = torch.rand(op.local_size(), dtype=torch.float64, device="cuda")
    y = op.initialize_dof_vector()
    op.vmult(y, x)
    xyz, bt = op.boundary_quadrature_points()
    op.set_boundary_values(np.ones(xyz.shape[:2]))
    op.rhs(y)
    op.evaluate(y, x)
    op.integrate_source_add(y, np.ones(op.cell_quadrature_points(3).shape[:2]))
    op.l2_error(x, np.ones(op.cell_quadrature_points(5).shape[:2]))
    torch.cuda.synchronize()
    print("ok hybrid path=%d, rhs / evaluate / error" % op.is_cartesian_path, flush=True)
    # multigrid: p- and h-transfer kernels, Chebyshev smoothers, coarse CG (k = 2, 4^3 cells, curved, phMG)
    args = dict(degree=2, n_subdivisions=1, n_refinements=2, mapping_degree=1, deformation=0.1, boundary=WALLS)
    mg = exadg_b200.MultigridPreconditioner.hypercube(args, "phMG", "Bisect", smoother_iterations=2, iterations_eigenvalue_estimation=4)
    b = torch.rand(mg.op.local_size(), dtype=torch.float64, device="cuda")
    z = mg.op.initialize_dof_vector()
    mg.vmult(z, b)
    torch.cuda.synchronize()
    print("ok multigrid levels=%s" % (mg.levels,), flush=True)
print("SANITIZER_CASES_DONE")
