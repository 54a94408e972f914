#!/bin/bash
# line kernel, 8-cell batches with 2-4 CTAs per SM (register budgets) against 16-cell batches: k = 5, 6, 7
mkdir -p gpurun_out
for lb in 16 8; do
for k in 5 6 7; do
  case $k in 5) c=80;; 6) c=64;; 7) c=64;; esac
  EXADG_B200_LINE_B=$lb timeout 300 python bench.py --degree $k --cells $c --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s46_err.log | tee -a gpurun_out/r02_s46_sweep_b$lb.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LINE_B=$lb k=$k ms %.3f GDoF/s %.1f frac %.3f inv %s' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac'], d['config']['invariants']))"
done
done
