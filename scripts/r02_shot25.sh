#!/bin/bash
# 1 GPU: pipelined host-buffer vmult as a CUDA graph: parity, chunk sizes
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_host_pipeline.py -q -x > gpurun_out/r02_s25_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s25_pytest.log )
tail -n 4 gpurun_out/r02_s25_pytest.log
for c in 768 1536 3072 6144 12288; do
  EXADG_B200_HP_CELLS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-callers --no-fp64-peak 2>> gpurun_out/r02_s25_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']; print('chunk cells $c: e2e %.3f GDoF/s (sequential %.3f, pipelined %s)' % (e['value'] / 1e9, e['sequential_dofs_per_s'] / 1e9, e['pipelined_dofs_per_s']))"
done
tail -3 gpurun_out/r02_s25_err.log
