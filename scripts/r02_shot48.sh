#!/bin/bash
# line kernel with the enumerated shared-memory layouts (n = 7, 8): parity, then the sweep lines k = 5, 6, 7
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_vmult.py -q -x > gpurun_out/r02_s48_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s48_pytest.log )
tail -n 4 gpurun_out/r02_s48_pytest.log
for k in 5 6 7; do
  case $k in 5) c=80;; 6) c=64;; 7) c=64;; esac
  timeout 300 python bench.py --degree $k --cells $c --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s48_err.log | tee -a gpurun_out/r02_s48_sweep.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k=$k ms %.3f GDoF/s %.1f frac %.3f' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac']))"
done
