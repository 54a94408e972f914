#!/bin/bash
# Helmholtz fast path with one component of a cell batch per CTA: full GPU suite, then timings
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s55_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s55_pytest.log )
tail -n 3 gpurun_out/r02_s55_pytest.log
cat > /tmp/helm.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import exadg_b200
torch.cuda.set_device(0)
for (k, cells) in ((5, (3, 4)), (3, (1, 6)), (4, (3, 4)), (2, (1, 6))):
    op = exadg_b200.LaplaceOperator.hypercube_helmholtz(k, 3, 100.0, 1e-3, cells[0], cells[1])
    n = op.local_size()
    src = torch.rand(n, dtype=torch.float64, device="cuda") * 2 - 1
    dst = op.initialize_dof_vector()
    for _ in range(3): op.vmult(dst, src)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20): op.vmult_async(dst, src)
    op.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
    print("k=%d cells=%d^3 dofs=%d: %.3f ms %.1f GDoF/s" % (k, cells[0] << cells[1], n, ms, n / ms / 1e6), flush=True)
    del op
PY
timeout 100 python /tmp/helm.py 2>&1 | tail -n 4 | tee gpurun_out/r02_s55_helm.log
