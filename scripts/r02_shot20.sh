#!/bin/bash
# 1 GPU: k=3 with 64-cell batches (parity + bench); line kernel without the out-of-batch trace phase (timing experiment, results wrong)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_vmult.py tests/test_gpu_solvers.py -q -x > gpurun_out/r02_s20_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s20_pytest.log )
tail -n 4 gpurun_out/r02_s20_pytest.log
timeout 200 python bench.py --degree 3 --cells 128 --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s20_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k=3 B=64 ms %.3f GDoF/s %.1f frac %.3f' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac']))"
for k in 5 6 7; do
  case $k in 5) c=80;; 6) c=64;; 7) c=64;; esac
  EXADG_B200_LINE_SKIP_HALO=1 timeout 200 python - <<PY
import torch, exadg_b200
k, c = $k, $c
n_sub, refine = c, 0
while n_sub % 2 == 0: n_sub //= 2; refine += 1
op = exadg_b200.LaplaceOperator.hypercube(k, n_sub, refine); op.use_torch_stream()
src = torch.rand(op.local_size(), dtype=torch.float64, device="cuda"); dst = op.initialize_dof_vector()
for _ in range(3): op.vmult_async(dst, src)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): op.vmult_async(dst, src)
e1.record(); torch.cuda.synchronize()
print("k=%d WITHOUT halo traces: ms %.3f" % (k, e0.elapsed_time(e1) / 20))
PY
done
tail -3 gpurun_out/r02_s20_err.log
