#!/bin/bash
# scalar kernels after the COMP templating: k = 3 and 5 (same numbers as the sweep expected)
for k in 3 5; do
  case $k in 3) c=128;; 5) c=80;; esac
  timeout 60 python bench.py --degree $k --cells $c --steps 10 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k=$k ms %.3f GDoF/s %.1f' % (d['ms_per_step'], d['value'] / 1e9))"
done
