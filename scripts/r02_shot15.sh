#!/bin/bash
# 1 GPU: hybrid fast path of bounded uniform boxes: parity, timing against the general-only run (k=4, 64^3 cells, sine boundary conditions)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_vmult.py tests/test_gpu_multigrid.py tests/test_gpu_rhs_error.py tests/test_gpu_solvers.py -q -x > gpurun_out/r02_s15_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s15_pytest.log )
tail -n 8 gpurun_out/r02_s15_pytest.log
python - <<'PY'
import torch, exadg_b200, time
for k, refine in ((4, 6), (2, 6), (3, 6), (5, 5)):
    for fg in (False, True):
        op = exadg_b200.LaplaceOperator.hypercube(k, 1, refine, 1, 0.0, 2, (1, 2, 1, 1, 1, 1), 1.0, force_general=fg)
        op.use_torch_stream()
        src = torch.rand(op.local_size(), dtype=torch.float64, device="cuda") * 2 - 1
        dst = op.initialize_dof_vector()
        for _ in range(3): op.vmult_async(dst, src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): op.vmult_async(dst, src)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("k=%d cells=%d^3 path=%d ms %.3f GDoF/s %.1f" % (k, 1 << refine, op.is_cartesian_path, ms, op.n() / ms / 1e6), flush=True)
        del op, src, dst
PY
