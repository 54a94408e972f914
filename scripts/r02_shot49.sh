#!/bin/bash
# full GPU suite on the final tree, then the Cartesian degree sweep k = 2..7
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s49_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s49_pytest.log ) 2> gpurun_out/r02_s49_time.log
tail -n 5 gpurun_out/r02_s49_pytest.log; tail -n 3 gpurun_out/r02_s49_time.log
rm -f gpurun_out/r02_s49_sweep_cart.jsonl
for k in 2 3 4 5 6 7; do
  case $k in 2) c=160;; 3) c=128;; 4) c=96;; 5) c=80;; 6) c=64;; 7) c=64;; esac
  timeout 300 python bench.py --degree $k --cells $c --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s49_err.log | tee -a gpurun_out/r02_s49_sweep_cart.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k=$k ms %.3f GDoF/s %.1f frac %.3f' % (d['ms_per_step'], d['value'] / 1e9, d['roofline']['frac']))"
done
