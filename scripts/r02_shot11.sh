#!/bin/bash
# 1 GPU: multigrid / Helmholtz tests, general-kernel parity, L2 prefetch distance of the general kernel on the curved k=4 box
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_helmholtz.py tests/test_gpu_vmult.py -q > gpurun_out/r02_s11_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s11_pytest.log )
tail -n 8 gpurun_out/r02_s11_pytest.log
rm -f gpurun_out/r02_s11_curved.jsonl
for d in 0 1 2 4; do
  EXADG_B200_GENERAL_PREFETCH=$d timeout 300 python bench.py --degree 4 --cells 64 --mesh curvilinear --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s11_err.log | python -c "
import sys, json
l = sys.stdin.read().strip().splitlines()[-1]; d = json.loads(l); d['prefetch_distance'] = $d
print(json.dumps(d))" >> gpurun_out/r02_s11_curved.jsonl
done
for k in 2 3 5; do
  EXADG_B200_GENERAL_PREFETCH=2 timeout 300 python bench.py --degree $k --cells 64 --mesh curvilinear --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s11_err.log >> gpurun_out/r02_s11_curved.jsonl
  EXADG_B200_GENERAL_PREFETCH=0 timeout 300 python bench.py --degree $k --cells 64 --mesh curvilinear --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s11_err.log >> gpurun_out/r02_s11_curved.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_s11_curved.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:60], d.get('prefetch_distance'), 'ms %.3f' % d['ms_per_step'], 'GDoF/s %.1f' % (d['value'] / 1e9), 'frac %.3f' % d['roofline']['frac'])
PY
tail -5 gpurun_out/r02_s11_err.log
