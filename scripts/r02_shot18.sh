#!/bin/bash
# 1 GPU: line kernel with 8-cell batches (k = 5, 6, 7): parity + timing against 16-cell batches
mkdir -p gpurun_out
( EXADG_B200_LINE_B=8 timeout 900 python -m pytest tests/test_gpu_vmult.py -q -x -k "not hybrid" > gpurun_out/r02_s18_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s18_pytest.log )
tail -n 4 gpurun_out/r02_s18_pytest.log
rm -f gpurun_out/r02_s18_line.jsonl
for b in 8 16; do
for k in 5 6 7; do
  case $k in 5) c=80;; 6) c=64;; 7) c=64;; esac
  EXADG_B200_LINE_B=$b timeout 200 python bench.py --degree $k --cells $c --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s18_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$b k=$k ms %.3f GDoF/s %.1f frac %.3f' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac']))"
done
done
tail -3 gpurun_out/r02_s18_err.log
