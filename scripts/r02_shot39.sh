#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_vmult.py tests/test_gpu_helmholtz.py tests/test_gpu_multigrid.py -q -x > gpurun_out/r02_s39_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s39_pytest.log )
tail -n 3 gpurun_out/r02_s39_pytest.log
for k in 4 2; do
  case $k in 2) c=96;; 4) c=64;; esac
  timeout 300 python bench.py --degree $k --cells $c --mesh curvilinear --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s39_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k=$k ms %.3f GDoF/s %.1f frac %.3f' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac']))"
done
