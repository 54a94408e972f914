#!/bin/bash
# n = 4 plane kernel with 32-cell batches (3 CTAs per SM) against 64-cell batches
mkdir -p gpurun_out
for pb in 64 32; do
  EXADG_B200_PLANE_B=$pb timeout 300 python bench.py --degree 3 --cells 128 --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s50_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PLANE_B=$pb k=3 ms %.3f GDoF/s %.1f frac %.3f inv %s' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac'], d['config']['invariants']))"
done
EXADG_B200_PLANE_B=32 timeout 300 python -m pytest tests/test_gpu_vmult.py -q -x -k "3" 2>&1 | tail -n 2
