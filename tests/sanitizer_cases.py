"""Small vmult / diagonal / CG / Chebyshev runs for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(SURVEY section 5): every kernel family once, on meshes small enough for the sanitizer.

    compute-sanitizer --tool racecheck python tests/sanitizer_cases.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import exadg_b200  # noqa: E402

for (degree, n_sub, refine, deformation, bc) in [(4, 3, 0, 0.0, (0,) * 6), (4, 1, 2, 0.0, (0,) * 6), (4, 3, 1, 0.0, (0,) * 6), (2, 5, 0, 0.0, (0,) * 6), (3, 1, 1, 0.0, (0,) * 6),
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
print("SANITIZER_CASES_DONE")
