#!/usr/bin/env python
"""bench.py -- DoFs/s of the 3-D SIPG Laplace matrix-free vmult (FP64) on B200.

Contract (see the task description): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A step = one vmult over the whole DoF vector.  Workload at N=1: BASELINE.json configs[1] at the degree the
metric is quoted on (k=4): periodic Cartesian box, 96^3 cells, 110,592,000 DoFs (885 MB per vector, i.e.
inputs larger than the 126 MB L2).  N>1 is weak scaling on the reference's admissible grids n_1d in {1,3,5} x 2^l
(hypercube_resolution_parameters.h:114-165) closest to a constant per-GPU size: 128^3, 160^3, 192^3 cells for N=2,4,8
(131 / 128 / 111 M DoFs per GPU; N=8 is BASELINE.json configs[4]), cells partitioned p4est-style, ghost import by NVLink
peer-memory stores inside the operator launch.
`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP over all host cores) on a
bounded sample of the same workload; the real deal.II/ExaDG binary cannot be built here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DoFs/s of 3D SIPG Laplace matrix-free vmult (fp64, k=4)"
B_ALG = 16.0  # algorithmic bytes per DoF of the affine vmult: read src once + write dst once
GRIDS = {1: (3, 5), 2: (1, 7), 4: (5, 5), 8: (3, 6)}  # n_gpus -> (n_subdivisions, n_refinements): 96, 128, 160, 192 cells per direction


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload_key)
        except Exception:
            return None
    return None


def tune_k4_kernel(n_sub, refine):
    """Times the k=4 kernels of the affine fast path side by side in a child process (build/ws_tournament, C ABI only) and returns
    (chosen variant, {variant: ms}): the fastest variant whose vmult and vmult_add agree with the pipelined kernel (which the GPU
    tests pin to the oracle) to 1e-14, the default unless another one is at least 2 % faster. None if the tool is not built or fails."""
    exe = os.path.join(ROOT, "build", "ws_tournament")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, "-1", str(n_sub), str(refine), "30"], capture_output=True, text=True, timeout=90)
    except Exception:
        return None
    if r.returncode != 0 or "TOURNAMENT DONE" not in r.stdout:
        return None
    ms, worst = {}, {}
    for line in r.stdout.splitlines():
        w = line.split()
        try:
            if line.startswith("variant ") and "ms/vmult" in line:
                ms[int(w[1].rstrip(":"))] = float(w[2])
            elif "rel l2 (variant" in line:
                v = int(line.split("(variant")[1].split()[0].rstrip(",)"))
                worst[v] = max(worst.get(v, 0.0), float(w[-1]))
        except (ValueError, IndexError):
            return None
    default = 1
    if default not in ms or worst.get(default, 1.0) > 1e-14:
        return None
    chosen = default
    for v, t in ms.items():
        if v != 0 and worst.get(v, 1.0) <= 1e-14 and t < 0.98 * ms[chosen]:
            chosen = v
    return chosen, ms


def probe_pipelined_e2e(degree, n_sub, refine, deformation):
    """Child process: both variants of vmult_host_pipelined ("staged": chunk plan with a copy-engine download per chunk; "direct": piece-wise
    upload, the kernels store dst straight into the pinned host buffer) on the same workload must reproduce the device vmult bit for bit
    (twice: plan and events are reused).  A fault or a hang in the child cannot take the benchmark down; the parent only times the variants
    the child has validated and keeps the sequential entry point otherwise.  Returns the list of validated variants."""
    code = (
        "import sys, torch\n"
        "sys.path.insert(0, %r)\n"
        "import exadg_b200\n"
        "torch.cuda.set_device(0)\n"
        "op = exadg_b200.LaplaceOperator.hypercube(%d, %d, %d, 1, %r, 2, (0,) * 6, 1.0)\n"
        "src = torch.rand(op.local_size(), dtype=torch.float64, device='cuda') * 2 - 1\n"
        "dst = op.initialize_dof_vector(); op.vmult(dst, src)\n"
        "h_src = torch.empty(op.local_size(), dtype=torch.float64).pin_memory(); h_src.copy_(src.cpu())\n"
        "h_dst = torch.empty(op.local_size(), dtype=torch.float64).pin_memory()\n"
        "op.vmult_host(h_dst, h_src)\n"
        "assert (h_dst.cuda() - dst).abs().max().item() == 0.0\n"
        "for mode in ('staged', 'direct'):\n"
        "    op.set_host_pipeline_mode(mode)\n"
        "    for rep in range(2):\n"
        "        h_dst.fill_(float('nan')); op.vmult_host_pipelined(h_dst, h_src)\n"
        "        assert (h_dst.cuda() - dst).abs().max().item() == 0.0\n"
        "    print('PIPELINED_OK_' + mode, flush=True)\n" % (ROOT, degree, n_sub, refine, deformation))
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=200)
        out = r.stdout
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
    except Exception:
        return []
    return [m for m in ("staged", "direct") if "PIPELINED_OK_" + m in out]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


CPU_PORT = "vectorised port of the reference algorithm: 8 cells per SIMD batch (AVX-512 or 2 x AVX2, chosen at load time), sum factorisation with compile-time degree, OpenMP over cell batches"


def _cpu_operator(degree, cells_1d):
    """The oracle's lean periodic-box operator (connectivity + penalty only) at the bench's own mesh size."""
    from oracle.oracle import OracleOperator
    n_sub, refine = cells_1d, 0
    while n_sub % 2 == 0 and n_sub > 1:
        n_sub //= 2
        refine += 1
    return OracleOperator(degree, n_sub, refine, 1, 0.0, lean=True)


def cpu_baseline(degree, seconds=12.0, cells_1d=96):
    """The oracle (port of the reference algorithm, OpenMP over cells) on the host cores, same mesh as the GPU arm at N=1; the sample
    is bounded in time (about `seconds` of vmults after the warm-up), not in size."""
    import numpy as np
    from oracle.oracle import synthetic_vector
    op = _cpu_operator(degree, cells_1d)
    x = synthetic_vector(op.n_dofs)
    y = np.zeros_like(x)
    threads = len(os.sched_getaffinity(0))  # all host cores (torchrun exports OMP_NUM_THREADS=1; the count is passed explicitly)
    for _ in range(3):  # warm-up: OpenMP team, page faults of dst, caches
        op.vmult_fast(x, threads, dst=y)
    best, reps, t_start = float("inf"), 0, time.time()
    while time.time() - t_start < seconds or reps < 5:
        t0 = time.perf_counter()
        op.vmult_fast(x, threads, dst=y)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    return {"value": op.n_dofs / best, "unit": "DoFs/s", "cores": threads, "kind": "port",
            "sample": "k=%d periodic Cartesian box, %d^3 cells (%d DoFs, the GPU arm's own mesh), min over %d vmults in %.0f s after 3 warm-up vmults, oracle/sipg_fast.inc orc_vmult_fast (%s), gcc -O3 -fopenmp"
                      % (degree, cells_1d, op.n_dofs, reps, time.time() - t_start, CPU_PORT if op.fast_path_used else "scalar fallback")}, op.n_dofs / best, best


def run_reference(args):
    """Reference arm: the CPU implementation of the path on the box's host cores (oracle port) on the GPU arm's own config
    (N=1 workload: 96^3 cells at k=4).  W warm-up steps (OpenMP team spin-up, page faults), then blocks of exactly K steps: the
    reported block is the fastest of up to three (a noisy neighbour on the host costs a block, not the figure)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    degree = args.degree
    from oracle.oracle import synthetic_vector
    import numpy as np
    cells = args.cells if args.cells else (GRIDS[1][0] << GRIDS[1][1])
    op = _cpu_operator(degree, cells)
    x = synthetic_vector(op.n_dofs)
    y = np.zeros_like(x)
    threads = len(os.sched_getaffinity(0))  # all host cores (torchrun exports OMP_NUM_THREADS=1; the count is passed explicitly)
    for _ in range(max(args.warmup, 3)):
        op.vmult_fast(x, threads, dst=y)
    blocks = []
    t_all = time.perf_counter()
    while len(blocks) < 3 and (not blocks or time.perf_counter() - t_all + blocks[0] < 90.0):
        t0 = time.perf_counter()
        for _ in range(args.steps):
            op.vmult_fast(x, threads, dst=y)
        blocks.append(time.perf_counter() - t0)
    dt = min(blocks)
    value = op.n_dofs * args.steps / dt
    sample = ("each step = one vmult on the GPU arm's N=1 workload: k=%d periodic Cartesian box, %d^3 cells (%d DoFs); fastest of %d block(s) of %d steps "
              "(block times %s s); %s" % (degree, cells, op.n_dofs, len(blocks), args.steps, ", ".join("%.3f" % b for b in blocks),
                                          CPU_PORT if op.fast_path_used else "scalar fallback"))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "DoFs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "SIPG Laplace vmult, FE_DGQ(%d), Gauss(%d), periodic Cartesian box, %d^3 cells, %d DoFs (CPU port of the reference algorithm, not deal.II)"
                                  % (degree, degree + 1, cells, op.n_dofs)},
           "cpu_baseline": {"value": value, "unit": "DoFs/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def run_callers(args, op, src, dst, world, rank, barrier, dist):
    """Secondary lines (not the headline): the callers of vmult on the same workload:
    cg        - dealii::SolverCG iterations without preconditioner: vmult + fused vector updates, 16 + 56 B/DoF
    chebyshev - one application of the Chebyshev(5)/point-Jacobi smoother: 4 vmults + fused updates"""
    import torch
    import exadg_b200
    n_global = op.n()
    if args.mode == "cg":
        b = op.initialize_dof_vector()
        op.vmult(b, src)  # consistent right-hand side (the periodic operator is singular)
        its = max(args.steps, 5)
        solver = exadg_b200.KrylovSolverCG(op, None, exadg_b200.SolverData(its, 0.0, 0.0))
        for timed in (False, True):
            x = op.initialize_dof_vector()
            barrier()
            t0 = time.perf_counter()
            try:
                solver.solve(x, b)
            except exadg_b200.ExaDGError:
                pass  # max_iter reached on purpose: exactly `its` iterations are timed
            barrier()
            dt = time.perf_counter() - t0
        per_it = dt / its
        b_alg = 16.0 + 56.0
        label = "CG iteration (dealii::SolverCG restated, no preconditioner): vmult + dot + fused x/g update + d update"
    else:
        ch = exadg_b200.ChebyshevSmoother(op, 5, 20.0, 20)
        reps = max(args.steps // 5, 3)
        ch.vmult(dst, src)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            ch.vmult(dst, src)
        barrier()
        per_it = (time.perf_counter() - t0) / reps
        b_alg = 4 * 16.0 + 3 * 8.0 + 4 * 7 * 8.0
        label = "Chebyshev(5)/point-Jacobi smoother application: 4 vmults + 5 fused update passes"
    t = torch.tensor([per_it], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_it = t.item()
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = b_alg * (n_global / world) / per_it / 1e9
        print(json.dumps({"metric": "DoFs/s per " + ("CG iteration" if args.mode == "cg" else "smoother application"), "value": n_global / per_it,
                          "unit": "DoFs/s", "n_gpus": world, "ms_per_step": per_it * 1e3, "dtype": "f64", "data": "synthetic", "secondary": True,
                          "config": {"workload": label, "degree": args.degree, "dofs": n_global, "timing": "host clock around the library call (includes one host read of the residual per CG iteration)"},
                          "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "algorithmic_bytes_per_dof": b_alg,
                                       "peak_source": peak_src}}), flush=True)


def callers_block(op, src, n_global, degree):
    """Secondary figures inside the one JSON line (N=1): the callers of vmult on the same workload, device-timed with CUDA events on
    the operator's stream (= torch's current stream), and a solve-level end-to-end figure (host rhs in, host solution out)."""
    import torch
    import exadg_b200
    peak, _ = measured_peak()
    out = {}

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    b = op.initialize_dof_vector()
    op.vmult(b, src)  # consistent right-hand side (the periodic operator is singular)
    its = 20
    solver = exadg_b200.KrylovSolverCG(op, None, exadg_b200.SolverData(its, 0.0, 0.0))

    def cg_run(x):
        try:
            solver.solve(x, b)
        except exadg_b200.ExaDGError:
            pass  # max_iter reached on purpose: exactly `its` iterations

    x = op.initialize_dof_vector()
    cg_run(x)  # warm-up (work vectors)
    x.zero_()
    t = timed(lambda: cg_run(x))
    b_alg = 16.0 + 16.0 + 48.0 + 24.0  # vmult + d.Ad + (x, g update with |g|^2) + d update
    out["cg_iteration"] = {"ms": t / its * 1e3, "dofs_per_s": n_global * its / t, "iterations": its, "algorithmic_bytes_per_dof": b_alg,
                           "hbm_frac": b_alg * n_global * its / t / 1e9 / peak,
                           "what": "dealii::SolverCG restated (no preconditioner): vmult, d.Ad, fused x/g update with |g|^2, d update; one host read of the residual per iteration (ReductionControl::check)"}
    ch = exadg_b200.ChebyshevSmoother(op, 5, 20.0, 20)
    y = op.initialize_dof_vector()
    ch.vmult(y, b)
    reps = 5
    t = timed(lambda: [ch.vmult(y, b) for _ in range(reps)])
    b_alg = 4 * 16.0 + 3 * 8.0 + 4 * 7 * 8.0
    out["chebyshev_application"] = {"ms": t / reps * 1e3, "dofs_per_s": n_global * reps / t, "algorithmic_bytes_per_dof": b_alg, "hbm_frac": b_alg * n_global * reps / t / 1e9 / peak,
                                    "what": "ChebyshevSmoother(5 iterations, point Jacobi): 4 vmults + 5 fused update passes"}
    # solve-level end to end: pinned host rhs -> H2D -> 20 CG iterations -> D2H of the solution
    h_b = torch.empty(b.numel(), dtype=torch.float64).pin_memory()
    h_b.copy_(b.cpu())
    h_x = torch.empty(b.numel(), dtype=torch.float64).pin_memory()

    def solve_e2e():
        bb = h_b.cuda(non_blocking=True)
        xx = torch.zeros_like(bb)
        try:
            exadg_b200.KrylovSolverCG(op, None, exadg_b200.SolverData(its, 0.0, 0.0)).solve(xx, bb)
        except exadg_b200.ExaDGError:
            pass
        h_x.copy_(xx, non_blocking=True)
        torch.cuda.synchronize()

    solve_e2e()
    t0 = time.perf_counter()
    solve_e2e()
    t = time.perf_counter() - t0
    out["solve_e2e"] = {"ms": t * 1e3, "dofs_times_iterations_per_s": n_global * its / t, "iterations": its, "h2d_bytes": n_global * 8, "d2h_bytes": n_global * 8,
                        "what": "host rhs (pinned) -> device, 20 CG iterations with all vectors resident, solution -> host: the transfers a Krylov solve actually pays, amortised over its iterations"}
    del ch, solver, x, y, b
    return out


def applications_block():
    """Secondary figures for the solver-level configurations of BASELINE.json (N=1, after the timed region; freed before returning):
    config 3 - full Poisson solve of applications/poisson/sine (curved mesh, k=4, CG + multigrid with Chebyshev/point-Jacobi smoothers),
    config 4 - the two operators of the dual-splitting Navier-Stokes solver (pressure Poisson k=4, viscous Helmholtz k=5, 3 components)."""
    import numpy as np
    import torch
    import exadg_b200
    out = {}
    w = 3.0 * np.pi

    def solution(x):
        return np.sin(w * x[..., 0]) * np.sin(w * x[..., 1]) * np.sin(w * x[..., 2])

    def timed(fn, reps=1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    # ---- config 3: applications/poisson/sine (application.h:134-178): deformation 0.15, MappingQ(3), Dirichlet + Neumann, rel 1e-10 ----
    args = dict(degree=4, n_subdivisions=2, n_refinements=4, mapping_degree=3, deformation=0.15, boundary=(1, 2, 1, 1, 1, 1))
    t0 = time.perf_counter()
    op = exadg_b200.LaplaceOperator.hypercube(**args)
    op.use_torch_stream()
    mg = exadg_b200.MultigridPreconditioner.hypercube(args, "phMG", "Bisect", fine_operator=op)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    xyz, bt = op.boundary_quadrature_points()
    neumann = w * np.cos(w * xyz[..., 0]) * np.sin(w * xyz[..., 1]) * np.sin(w * xyz[..., 2])
    op.set_boundary_values(np.where((bt == exadg_b200.DIRICHLET)[:, None], solution(xyz), neumann))
    b = op.initialize_dof_vector()
    op.rhs(b)
    op.integrate_source_add(b, 3.0 * w * w * solution(op.cell_quadrature_points(5)))
    x = op.initialize_dof_vector()
    solver = exadg_b200.KrylovSolverCG(op, mg, exadg_b200.SolverData(1000, 1e-20, 1e-10))
    solver.solve(x, b)  # warm-up (work vectors, coarse solution)
    x.zero_()
    t = timed(lambda: solver.solve(x, b))
    err = op.l2_error(x, solution(op.cell_quadrature_points(7)))
    jac = exadg_b200.KrylovSolverCG(op, exadg_b200.JacobiPreconditioner(op), exadg_b200.SolverData(10000, 1e-20, 1e-10))
    xj = op.initialize_dof_vector()
    tj = timed(lambda: jac.solve(xj, b))
    out["poisson_solve"] = {"dofs": op.n(), "levels": mg.levels, "cg_iterations": solver.n, "solve_ms": t * 1e3, "dofs_per_s": op.n() / t, "setup_s": t_setup,
                            "relative_l2_error": err, "coarse_cg_iterations_total": mg.info()["coarse_iterations"],
                            "jacobi_cg_iterations": jac.n, "jacobi_solve_ms": tj * 1e3,
                            "what": "applications/poisson/sine: k=4, 32^3 cells, sine-deformed (0.15) MappingQ(3) mesh, Dirichlet + Neumann, CG rel 1e-10 with phMG "
                                    "(DG levels k=4,2,1 then h-coarsening, Chebyshev(5)/point-Jacobi smoothers, CG + point Jacobi to 1e-3 on the coarsest level); FP64 levels"}
    del solver, jac, mg, op, x, xj, b
    torch.cuda.empty_cache()

    # ---- config 4: Taylor-Green vortex operators (periodic box, k_u = 5, k_p = 4): pressure Poisson vmult, viscous Helmholtz vmult ----
    cells = dict(n_subdivisions=3, n_refinements=4)  # 48^3 cells
    pres = exadg_b200.LaplaceOperator.hypercube(degree=4, **cells)
    pres.use_torch_stream()
    g = torch.Generator(device="cuda").manual_seed(7)
    u = torch.rand(pres.local_size(), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = pres.initialize_dof_vector()
    pres.vmult_async(y, u)
    tp = timed(lambda: pres.vmult_async(y, u), 20)
    visc = exadg_b200.LaplaceOperator.hypercube_helmholtz(5, 3, 200.0, 1.0 / 1600.0, **cells)
    visc.use_torch_stream()
    uv = torch.rand(visc.local_size(), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    yv = visc.initialize_dof_vector()
    visc.vmult_async(yv, uv)
    tv = timed(lambda: visc.vmult_async(yv, uv), 5)
    out["ins_operators"] = {"cells": 48 ** 3, "pressure_poisson": {"degree": 4, "dofs": pres.n(), "vmult_ms": tp * 1e3, "dofs_per_s": pres.n() / tp, "singular": bool(pres.operator_is_singular())},
                            "viscous_helmholtz": {"degree": 5, "components": 3, "dofs": visc.n(), "vmult_ms": tv * 1e3, "dofs_per_s": visc.n() / tv,
                                                  "kernel": "affine line kernel, one component of a cell batch per CTA, mass term in front of the mass sweeps (uniform periodic box; general kernel elsewhere)"},
                            "what": "operators of the dual-splitting scheme on the Taylor-Green box (periodic, 48^3 cells): pressure Poisson = the SIPG Laplace fast path at k=4; "
                                    "viscous step = gamma0/dt M + nu A_SIPG on 3 velocity components at k=5 (affine fast path, one component of a cell batch per CTA)"}
    del pres, visc, u, y, uv, yv
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch
    import exadg_b200
    from exadg_b200.laplace_operator import nccl_unique_id

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    degree = args.degree
    if args.cells:
        n_sub, refine = args.cells, 0
        while n_sub % 2 == 0 and n_sub > 1:
            n_sub //= 2
            refine += 1
    else:
        n_sub, refine = GRIDS.get(world, (3, 5))
    deformation = 0.1 if args.mesh == "curvilinear" else 0.0
    kernel_selection = None
    if world == 1 and degree == 4 and deformation == 0.0 and args.tune and "EXADG_B200_CART_KERNEL" not in os.environ:
        # the k=4 fast path has several validated kernels; pick the fastest on this box before anything is allocated here
        tuned = tune_k4_kernel(n_sub, refine)
        if tuned is not None:
            exadg_b200.cartesian_kernel(tuned[0])
            kernel_selection = {"chosen_variant": tuned[0], "ms_per_vmult_by_variant": tuned[1],
                                "variants": "0 pipelined, 1 warp-specialised 2 producer warps, 2 deeper producers, 3 four producer warps + setmaxnreg (library default), 4/5 warp-private"}
    op = exadg_b200.LaplaceOperator.hypercube(degree, n_sub, refine, 1, deformation, 2, (0,) * 6, 1.0, rank=rank, world=world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        op.init_nccl(bytes(idt.cpu().numpy().tobytes()))
        halo_used = "nccl"
        if args.halo == "p2p":
            try:
                op.enable_p2p(dist)  # ghost import by NVLink peer-memory stores fused into the pack kernel
                halo_used = "p2p"
            except Exception as e:  # no peer access on this box: the NCCL transport is equivalent
                sys.stderr.write("peer-memory halo unavailable (%s); using NCCL send/recv\n" % e)
        # every rank must use the same transport
        flag = torch.tensor([1 if halo_used == "p2p" else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if args.halo == "p2p" and flag.item() == 0 and halo_used == "p2p":
            raise RuntimeError("peer-memory halo enabled on some ranks only")
        args.halo = halo_used
    op.use_torch_stream()
    n_local, n_global = op.local_size(), op.n()
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    src = torch.rand(n_local, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    dst = op.initialize_dof_vector()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode != "vmult":
        run_callers(args, op, src, dst, world, rank, barrier, dist)
        return
    for _ in range(max(args.warmup, 3)):
        op.vmult_async(dst, src)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = op.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        op.vmult_async(dst, src)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = op.kernel_launches() - launches0
    # reference protocol as a cross-check (throughput_parameters.h:53-84): min over repeats of the mean over inner reps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    clocks = sampler.stop() if rank == 0 else None

    # cheap size-independent checks of the dst that was just timed (every N, asserted): A 1 = 0 on the periodic box and
    # <A u, v> = <u, A v> with global dot products
    def gsum(t):
        t = t.reshape(1).clone()
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    def gmax(t):
        t = t.reshape(1).clone()
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    op.synchronize()
    scale = gmax(dst.abs().max())
    w1, w2 = op.initialize_dof_vector(), op.initialize_dof_vector()
    op.vmult(w1, torch.ones_like(src))
    a_one = gmax(w1.abs().max()) / scale
    v = torch.rand(n_local, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    op.vmult(w2, v)
    auv, uav = gsum(torch.dot(dst, v)), gsum(torch.dot(src, w2))
    sym = abs(auv - uav) / max(abs(auv), 1e-300)
    invariants = {"A1_over_Au_max": a_one, "symmetry_rel": sym, "ranks": world}
    assert a_one < 1e-9 and sym < 1e-9, invariants
    del w1, w2, v

    # end to end through the host-buffer entry point: pinned host src -> H2D -> vmult -> D2H pinned dst
    h_src = torch.empty(n_local, dtype=torch.float64).pin_memory()
    h_src.copy_(src.cpu())
    h_dst = torch.empty(n_local, dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    op.vmult_host(h_dst, h_src)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        op.vmult_host(h_dst, h_src)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = t.item()
    if world == 1:
        assert (h_dst.cuda() - dst).abs().max().item() == 0.0
    e2e_api = "exadg_b200_vmult_host (pinned host src/dst, copies inside the timed region)"
    e2e_plain_s, e2e_pipe_s = e2e_s, {}
    # the same call with upload, operator and download overlapped inside it (unpartitioned operators), in its two variants; a variant is
    # used for the e2e figure only if a child process has first reproduced the device result bit for bit with it, if it does so again
    # here, and if it is faster
    e2e_apis = {"staged": "exadg_b200_vmult_host_pipelined, staged variant (pinned host src/dst; upload, vmult and copy-engine download overlap chunk by chunk inside the call)",
                "direct": "exadg_b200_vmult_host_pipelined, direct variant (pinned host src/dst; src uploaded in 24 MB pieces, behind every piece one launch applies the "
                          "cell batches whose neighbours have arrived, the kernel's bulk stores write dst straight into the pinned host buffer over PCIe)"}
    if world == 1 and args.e2e_api != "plain":
        for mode in probe_pipelined_e2e(degree, n_sub, refine, deformation):
            try:
                op.set_host_pipeline_mode(mode)
                h_dst.zero_()
                op.vmult_host_pipelined(h_dst, h_src)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    op.vmult_host_pipelined(h_dst, h_src)
                pipe_s = time.perf_counter() - t0
                if (h_dst.cuda() - dst).abs().max().item() == 0.0:  # counted only if it reproduces the device result bit for bit here as well
                    e2e_pipe_s[mode] = pipe_s
                    if pipe_s < e2e_s:
                        e2e_s = pipe_s
                        e2e_api = e2e_apis[mode]
            except Exception:  # the figures measured so far stand
                pass
        op.set_host_pipeline_mode("auto")
    elif world > 1 and args.e2e_api != "plain" and op.is_cartesian_path:
        # partitioned operator: the direct variant (interior batches behind the piece-wise uploads, batches with ghost neighbours behind the ghost
        # import, dst stored into the pinned host buffer by the kernels).  All ranks enter every call; it is used only if every rank reproduces
        # its device vmult bit for bit with it and if it is faster
        op.set_host_pipeline_mode("direct")
        h_dst.fill_(float("nan"))
        op.vmult_host_pipelined(h_dst, h_src)
        bad = torch.tensor([0.0 if (h_dst.cuda() - dst).abs().max().item() == 0.0 else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            op.vmult_host_pipelined(h_dst, h_src)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if bad.item() == 0.0:
            e2e_pipe_s["direct"] = t.item()
            if t.item() < e2e_s:
                e2e_s = t.item()
                e2e_api = e2e_apis["direct"] + "; partitioned: the batches with ghost neighbours follow the NVLink ghost import behind the last upload"
        op.set_host_pipeline_mode("auto")

    if rank == 0:
        value = n_global * args.steps / (ms * 1e-3)
        peak, peak_src = measured_peak()
        b_alg = B_ALG if deformation == 0.0 else 16.0 + 48.0 + 168.0 / (degree + 1)
        # the vmult kernel is the only kernel of a step at N=1 (halo pack/exchange are separate launches at N>1),
        # so its average launch duration is ms / steps, measured with CUDA events on the launching stream
        achieved = b_alg * (n_global / world) * args.steps / (ms * 1e-3) / 1e9
        key = "k%d_%s" % (degree, args.mesh)
        kernel = "general (stored metrics)"
        if op.is_cartesian_path:
            kernel = "cartesian"
            if degree == 4:
                variant = op.get_kernel_variant()
                names = {0: "pipelined vmult_cartesian_pipe_kernel<5>",
                         1: "warp-specialised vmult_cartesian_ws_kernel<5,8,2 producer warps>", 2: "warp-specialised vmult_cartesian_ws_kernel<5,12,2 producer warps>",
                         3: "warp-specialised vmult_cartesian_ws_kernel<5,4,4 producer warps, setmaxnreg> (library default)",
                         4: "warp-private vmult_cartesian_wp_kernel<5,8,2 producer warps>", 5: "warp-private vmult_cartesian_wp_kernel<5,12,2 producer warps>"}
                kernel = names.get(variant, str(variant))
                if variant >= 1:
                    key += "_ws" if variant <= 3 else "_wp"
                    if world > 1 and args.halo == "p2p":
                        kernel += "; ONE launch per vmult: 64 CTAs export this rank's cells to the peers' ghost buffers over NVLink first, batches without ghost neighbours next (items claimed from a work counter), the producers acquire the peers' flags before the first batch that reads ghost cells"
        out = {"metric": METRIC if degree == 4 else METRIC.replace("k=4", "k=%d" % degree), "value": value, "unit": "DoFs/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "SIPG Laplace vmult, FE_DGQ(%d), Gauss(%d), periodic %s box, %d^3 cells, %d DoFs, src uniform(-1,1)"
                                      % (degree, degree + 1, "Cartesian" if deformation == 0.0 else "sine-deformed (trilinear)", n_sub << refine, n_global),
                          "l2_policy": "inputs larger than L2 (%.0f MB per vector per GPU)" % (n_local * 8 / 1e6),
                          "kernel_path": "cartesian" if op.is_cartesian_path else "general", "kernel": kernel, "kernel_selection": kernel_selection, "invariants": invariants, "partition": "p4est-style contiguous Morton ranges, %d rank(s)" % world,
                          "halo": "none" if world == 1 else ("NVLink peer-memory stores (CUDA IPC), overlapped with interior cells" if args.halo == "p2p" else "NCCL send/recv, overlapped with interior cells")},
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(key),
                            "peak_source": peak_src, "algorithmic_bytes_per_dof": b_alg, "dofs_per_launch": n_global // world},
               "e2e": {"value": n_global * e2e_steps / e2e_s, "unit": "DoFs/s", "h2d_bytes_per_step": n_global * 8, "d2h_bytes_per_step": n_global * 8,
                       "api": e2e_api, "sequential_dofs_per_s": n_global * e2e_steps / e2e_plain_s,
                       "pipelined_dofs_per_s": {m: n_global * e2e_steps / t for m, t in e2e_pipe_s.items()} or None},
               "gpu_launches": launches, "clocks": clocks}
        if world == 1 and not args.no_callers and deformation == 0.0:
            try:
                out["callers"] = callers_block(op, src, n_global, degree)
            except Exception as e:  # secondary figures must not take the headline down
                out["callers"] = {"error": str(e)[:200]}
            try:
                del src, dst
                torch.cuda.empty_cache()
                out["applications"] = applications_block()
            except Exception as e:
                out["applications"] = {"error": str(e)[:300]}
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(degree, seconds=args.cpu_seconds, cells_1d=n_sub << refine)[0]
        if not args.no_fp64_peak:
            # the affine kernel is bounded by the FP64 pipe before HBM: report the measured DFMA / DMMA rates and the fraction of the
            # DFMA rate the kernel's algorithmic flops reach (6 n^4 + 18 n^3 FMA per cell)
            dfma, dmma = exadg_b200.fp64_peak()
            n1 = degree + 1
            flop_per_dof = 2.0 * (6 * n1 ** 4 + 18 * n1 ** 3) / n1 ** 3 if op.is_cartesian_path else None
            out["fp64"] = {"dfma_tflops_measured": dfma, "dmma_tflops_measured": dmma, "algorithmic_flop_per_dof": flop_per_dof,
                           "achieved_tflops": (flop_per_dof * value / world / 1e12) if flop_per_dof else None,
                           "frac_of_dfma": (flop_per_dof * value / world / 1e12 / dfma) if flop_per_dof else None}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--mesh", default="cartesian", choices=["cartesian", "curvilinear"])
    ap.add_argument("--cells", type=int, default=0, help="cells per direction (default: the workload of the contract)")
    ap.add_argument("--mode", default="vmult", choices=["vmult", "cg", "chebyshev"], help="vmult = the headline metric; cg / chebyshev = secondary lines for the callers")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="ghost import transport for N>1")
    ap.add_argument("--e2e-api", default="auto", choices=["auto", "plain"], help="auto: also try the pipelined host-buffer entry point for the e2e figure (validated in a child process first)")
    ap.add_argument("--tune", action="store_true", help="k=4: time the validated kernel variants in a child process first and use the fastest (default: the library's default kernel)")
    ap.add_argument("--no-tune", action="store_true", help="(default behaviour, kept for older command lines)")
    ap.add_argument("--no-callers", action="store_true", help="skip the secondary measurements of the callers (CG iteration, Chebyshev application, solve-level e2e)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-fp64-peak", action="store_true", help="skip the measured DFMA / DMMA rates")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
