#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2; do
 for k in 4 3; do
  case $k in 3) c=80;; 4) c=64;; esac
  EXADG_B200_GENERAL_PREFETCH=$d timeout 300 python bench.py --degree $k --cells $c --mesh curvilinear --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s36_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prefetch $d k=$k ms %.3f GDoF/s %.1f' % (d['ms_per_step'], d['value'] / 1e9))"
 done
done
