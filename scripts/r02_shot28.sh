#!/bin/bash
# 1 GPU: general kernel with line ownership in the face phase: parity (whole GPU suite), curved sweep
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s28_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s28_pytest.log )
tail -n 8 gpurun_out/r02_s28_pytest.log
rm -f gpurun_out/r02_s28_curved.jsonl
for k in 2 3 4 5 6 7; do
  case $k in 2) c=96;; 3) c=80;; 4) c=64;; 5) c=64;; 6) c=48;; 7) c=48;; esac
  timeout 300 python bench.py --degree $k --cells $c --mesh curvilinear --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain 2>> gpurun_out/r02_s28_err.log >> gpurun_out/r02_s28_curved.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_s28_curved.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:80], 'ms %.3f' % d['ms_per_step'], 'GDoF/s %.1f' % (d['value'] / 1e9), 'frac %.3f' % d['roofline']['frac'])
PY
tail -5 gpurun_out/r02_s28_err.log
