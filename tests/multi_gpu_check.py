"""Multi-GPU parity check (run under torchrun, one rank per GPU; tests/test_gpu_multi.py spawns it from pytest when >= 2 GPUs are
visible): the partitioned vmult must match the CPU ORACLE (cases up to 16^3 cells) and the single-partition run of the same library,
CG counts must equal the single-partition solve, for both halo transports: NCCL send/recv and NVLink peer-memory stores (CUDA IPC).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py

Kernel selection for k = 4 on the affine path: by default the interior batches run on the warp-specialised kernel and the batches with
ghost neighbours on the pipelined one; EXADG_B200_WS_GHOST=1 routes the latter through the warp-specialised kernel too (its ghost path,
so far verified on the CPU emulation only), EXADG_B200_CART_KERNEL=pipe uses the pipelined kernel everywhere.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import exadg_b200  # noqa: E402
from exadg_b200.laplace_operator import nccl_unique_id  # noqa: E402
from oracle.oracle import OracleOperator  # noqa: E402  (the checker; never on the product path)

P6 = (0,) * 6
# (4, 1, 3), (4, 1, 4), (4, 3, 3): octet-aligned partitions (the warp-specialised k=4 kernel applies on every rank);
# (4, 5, 0), (5, 3, 0), (2, 5, 0): partitions that end inside a cell batch (ghost indices directly follow a ragged last batch)
CASES = [(4, 1, 3, 0.0, P6), (4, 1, 4, 0.0, P6), (4, 3, 3, 0.0, P6), (4, 5, 0, 0.0, P6), (5, 3, 0, 0.0, P6), (2, 5, 0, 0.0, P6), (4, 3, 2, 0.0, P6), (3, 1, 3, 0.1, P6),
         (2, 2, 2, 0.15, (1, 2, 1, 1, 1, 1)), (5, 1, 2, 0.0, P6),
         # k=3 plane kernel with the swizzled trace arrays and k=6 / k=7 line kernels with 8-cell batches on a partition (ghost reads in the trace phase)
         (3, 1, 3, 0.0, P6), (6, 1, 2, 0.0, P6), (7, 1, 2, 0.0, P6)]


def fresh_nccl_id(rank):
    # one unique id per communicator (an id cannot be reused for a second ncclCommInitRank)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().numpy().tobytes())


def check_case(case, transport, rank, world):
    degree, n_sub, refine, deformation, bc = case
    op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, 1, deformation, 2, bc, 1.0, rank=rank, world=world)
    op.init_nccl(fresh_nccl_id(rank))  # all-reduces of the dot products; ghost import too unless p2p is enabled
    if transport == "p2p":
        op.enable_p2p(dist)            # ghost import by NVLink peer-memory stores
    n3 = (degree + 1) ** 3
    n_global = op.n()
    g = torch.Generator().manual_seed(123)
    x_global = torch.rand(n_global, dtype=torch.float64, generator=g) * 2 - 1
    lo = (n_global // n3) * rank // world * n3
    hi = lo + op.local_size()
    ok = True
    # reference: the whole problem on this GPU alone
    ref = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, 1, deformation, 2, bc, 1.0)
    y_ref = ref.initialize_dof_vector()
    ref.vmult(y_ref, x_global.cuda())
    # the oracle on the same global vector (every rank computes it: small cases only)
    y_orc = None
    if (n_sub << refine) <= 16:
        y_orc = torch.from_numpy(OracleOperator(degree, n_sub, refine, 1, deformation, 2, bc).vmult(x_global.numpy())).cuda()
        assert ((y_ref - y_orc).norm() / y_orc.norm()).item() < 1e-12
    dst = op.initialize_dof_vector()
    worst = 0.0
    for rep in range(3):  # repeated: exercises the double-buffered ghost ranges / epochs
        src = (x_global[lo:hi] * (rep + 1)).cuda()
        op.vmult(dst, src)
        worst = max(worst, ((dst - (rep + 1) * y_ref[lo:hi]).norm() / ((rep + 1) * y_ref.norm())).item())
        if y_orc is not None:
            worst = max(worst, ((dst - (rep + 1) * y_orc[lo:hi]).norm() / ((rep + 1) * y_orc.norm())).item())
    # vmult_add on top of the last result
    op.vmult_add(dst, src)
    worst = max(worst, ((dst - 6 * y_ref[lo:hi]).norm() / (6 * y_ref.norm())).item())
    # host-buffer vmult, direct variant on the partition (piece-wise upload, interior units behind the uploads with dst stored straight into
    # the pinned host buffer, units with ghost neighbours behind the ghost import): bit for bit the device vmult, repeatedly (epochs)
    op.vmult(dst, src)
    h_src = src.cpu().pin_memory()
    h_dst = torch.empty_like(h_src).pin_memory()
    op.set_host_pipeline_mode("direct")
    pipe_diff = 0.0
    for rep in range(3):
        h_dst.fill_(float("nan"))
        op.vmult_host_pipelined(h_dst, h_src)
        d = (h_dst.cuda() - dst).abs().max().item()
        pipe_diff = max(pipe_diff, d if d == d else float("inf"))
    op.set_host_pipeline_mode("auto")
    op.vmult_host(h_dst, h_src)  # the sequential entry point after it
    pipe_diff = max(pipe_diff, (h_dst.cuda() - dst).abs().max().item())
    pflag = torch.tensor([pipe_diff], device="cuda")
    dist.all_reduce(pflag, op=dist.ReduceOp.MAX)
    if pflag.item() != 0.0:
        ok = False
        if rank == 0:
            print("host-buffer vmult (direct variant) differs from the device vmult: %g" % pflag.item(), flush=True)
    its = None
    if bc != P6:  # CG with Jacobi: iteration counts equal to the single-partition solve
        b = y_ref.clone()
        s1 = exadg_b200.KrylovSolverCG(ref, exadg_b200.JacobiPreconditioner(ref), exadg_b200.SolverData(2000, 1e-20, 1e-8))
        x1 = ref.initialize_dof_vector()
        n1 = s1.solve(x1, b)
        s2 = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), exadg_b200.SolverData(2000, 1e-20, 1e-8))
        x2 = op.initialize_dof_vector()
        n2 = s2.solve(x2, b[lo:hi].clone())
        its = (n1, n2)
        ok &= (n1 == n2)
        ok &= ((x2 - x1[lo:hi]).norm() / x1.norm()).item() < 1e-6
    flag = torch.tensor([worst], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%s k=%d cells=%d^3 deformation=%g bc=%s cartesian_path=%d ghosts(rank0)=%d max rel err %.3e cg its %s"
              % (transport, degree, n_sub << refine, deformation, bc, op.is_cartesian_path, op.n_cells_ghost, flag.item(), its), flush=True)
    ok &= flag.item() < 1e-12
    del op, ref
    return ok


def check_helmholtz(transport, rank, world):
    """Config 4 on a partition: the viscous / Helmholtz operator (three velocity components; whole cell blocks travel in the ghost import) must
    match the single-partition operator - on the curved walled box (general kernel on both sides) and on the uniform periodic box, where the
    single-partition operator runs the affine fast kernels and the partition the general kernel - and CG must take the same number of
    iterations."""
    ok = True
    for (k, n_sub, refine, deformation, bc) in [(3, 1, 2, 0.1, (1, 1, 1, 1, 2, 2)), (2, 3, 1, 0.0, P6), (5, 1, 2, 0.0, P6)]:
        args = (k, 3, 40.0, 0.05, n_sub, refine, 1, deformation, 2, bc)
        ref = exadg_b200.LaplaceOperator.hypercube_helmholtz(*args)
        op = exadg_b200.LaplaceOperator.hypercube_helmholtz(*args, rank=rank, world=world)
        op.init_nccl(fresh_nccl_id(rank))
        if transport == "p2p":
            op.enable_p2p(dist)
        blk = 3 * (k + 1) ** 3
        g = torch.Generator().manual_seed(11)
        x_global = torch.rand(ref.local_size(), dtype=torch.float64, generator=g) * 2 - 1
        lo = (ref.local_size() // blk) * rank // world * blk
        hi = lo + op.local_size()
        y_ref = ref.initialize_dof_vector()
        ref.vmult(y_ref, x_global.cuda())
        dst = op.initialize_dof_vector()
        worst = 0.0
        for rep in range(3):
            op.vmult(dst, (x_global[lo:hi] * (rep + 1)).cuda())
            worst = max(worst, ((dst - (rep + 1) * y_ref[lo:hi]).norm() / ((rep + 1) * y_ref.norm())).item())
        d_ref, d = ref.initialize_dof_vector(), op.initialize_dof_vector()
        ref.calculate_diagonal(d_ref); op.calculate_diagonal(d)
        worst = max(worst, ((d - d_ref[lo:hi]).norm() / d_ref.norm()).item())
        data = exadg_b200.SolverData(500, 1e-20, 1e-9)
        x1, x2 = ref.initialize_dof_vector(), op.initialize_dof_vector()
        n1 = exadg_b200.KrylovSolverCG(ref, exadg_b200.JacobiPreconditioner(ref), data).solve(x1, y_ref)
        n2 = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), data).solve(x2, y_ref[lo:hi].clone())
        worst_t = torch.tensor([worst, ((x2 - x1[lo:hi]).norm() / x1.norm()).item()], device="cuda")
        dist.all_reduce(worst_t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("%s Helmholtz k=%d cells=%d^3 deformation=%g: vmult / diagonal rel err %.2e, CG its %d / %d, solution diff %.2e"
                  % (transport, k, n_sub << refine, deformation, worst_t[0].item(), n1, n2, worst_t[1].item()), flush=True)
        ok &= worst_t[0].item() < 1e-12 and n1 == n2 and worst_t[1].item() < 1e-7
        del op, ref
    return ok


def check_multigrid(transport, rank, world):
    """Config 3 on a partition: CG preconditioned by the p-multigrid V-cycle (DG levels k = 4, 2, 1 on the same cells - the p-transfer is
    cell-local on any partition; Chebyshev smoothers, CG + point Jacobi on the coarsest level, all with global dot products) must
    take the iteration count of the single-partition solve and give its solution."""
    from exadg_b200.laplace_operator import MultigridPreconditioner, multigrid_levels
    bc = (1, 2, 1, 1, 1, 1)
    kw = dict(degree=4, n_subdivisions=1, n_refinements=2, mapping_degree=3, deformation=0.15, frequency=2, boundary=bc, ip_factor=1.0)
    levels = multigrid_levels("pMG", "Bisect", 4, 3)
    data = exadg_b200.SolverData(200, 1e-20, 1e-10)
    ref_mg = MultigridPreconditioner.hypercube(kw, "pMG")
    ref = ref_mg.op
    g = torch.Generator().manual_seed(7)
    x_global = torch.rand(ref.n(), dtype=torch.float64, generator=g) * 2 - 1
    b = ref.initialize_dof_vector()
    ref.vmult(b, x_global.cuda())
    x1 = ref.initialize_dof_vector()
    n1 = exadg_b200.KrylovSolverCG(ref, ref_mg, data).solve(x1, b)
    ops = []
    for (h, k) in levels:
        op = exadg_b200.LaplaceOperator.hypercube(**dict(kw, degree=k, n_refinements=h, rank=rank, world=world))
        op.init_nccl(fresh_nccl_id(rank))
        if transport == "p2p":
            op.enable_p2p(dist)
        ops.append(op)
    mg = MultigridPreconditioner(ops)
    fine = ops[-1]
    n3 = 125
    lo = (fine.n() // n3) * rank // world * n3
    hi = lo + fine.local_size()
    x2 = fine.initialize_dof_vector()
    n2 = exadg_b200.KrylovSolverCG(fine, mg, data).solve(x2, b[lo:hi].clone())
    err = torch.tensor([((x2 - x1[lo:hi]).norm() / x1.norm()).item()], device="cuda")
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%s CG + pMG %s: iterations single partition %d, %d ranks %d; solution difference %.2e" % (transport, levels, n1, world, n2, err.item()), flush=True)
    return n1 == n2 and err.item() < 1e-8


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for case in CASES:
        for transport in ("nccl", "p2p"):
            ok &= check_case(case, transport, rank, world)
    for transport in ("nccl", "p2p"):
        ok &= check_multigrid(transport, rank, world)
        ok &= check_helmholtz(transport, rank, world)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
