#!/bin/bash
# Helmholtz / viscous operator on the affine fast kernels (uniform periodic box): parity, then the INS operators block of the bench
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_helmholtz.py tests/test_gpu_vmult.py -q -x > gpurun_out/r02_s52_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s52_pytest.log )
tail -n 12 gpurun_out/r02_s52_pytest.log
cat > /tmp/helm.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import exadg_b200
torch.cuda.set_device(0)
for fast in (1, 0):
    if not fast: os.environ["EXADG_B200_NO_HELMHOLTZ_FAST"] = "1"
    for (k, cells) in ((5, (3, 4)), (3, (1, 6)), (4, (3, 4)), (2, (1, 6))):
        op = exadg_b200.LaplaceOperator.hypercube_helmholtz(k, 3, 100.0, 1e-3, cells[0], cells[1])
        n = op.local_size()
        src = torch.rand(n, dtype=torch.float64, device="cuda") * 2 - 1
        dst = op.initialize_dof_vector()
        for _ in range(3): op.vmult(dst, src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        op.use_torch_stream() if hasattr(op, "use_torch_stream") else None
        e0.record()
        for _ in range(10): op.vmult(dst, src)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("fast=%d k=%d cells=%d^3 dofs=%d: %.3f ms %.1f GDoF/s  |dst|=%.12e" % (fast, k, cells[0] << cells[1], n, ms, n / ms / 1e6, dst.norm().item()), flush=True)
        del op
PY
timeout 300 python /tmp/helm.py 2>&1 | tail -n 9 | tee gpurun_out/r02_s52_helm.log
