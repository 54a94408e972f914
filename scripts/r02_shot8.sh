#!/bin/bash
# 1 GPU: parity of all degrees (line kernel for k = 5, 6, 7), degree sweep on the Cartesian box, bench with callers block
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_vmult.py tests/test_gpu_solvers.py -q > gpurun_out/r02_s8_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s8_pytest.log )
tail -n 5 gpurun_out/r02_s8_pytest.log
rm -f gpurun_out/r02_sweep_cart.jsonl
for k in 2 3 4 5 6 7; do
  case $k in 2) c=160;; 3) c=128;; 4) c=96;; 5) c=80;; 6) c=64;; 7) c=64;; esac
  timeout 200 python bench.py --degree $k --cells $c --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain >> gpurun_out/r02_sweep_cart.jsonl 2>> gpurun_out/r02_sweep_err.log
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep_cart.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:40], 'ms %.3f'%d['ms_per_step'], 'GDoF/s %.1f'%(d['value']/1e9), 'frac %.3f'%d['roofline']['frac'])
PY
for k in 5 6 7; do
  case $k in 5) c=80;; 6) c=64;; 7) c=64;; esac
  EXADG_B200_NO_LINE=1 timeout 200 python bench.py --degree $k --cells $c --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('plane kernel k=$k', 'GDoF/s %.1f'%(d['value']/1e9))"
done
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_s8_bench_n1.json 2> gpurun_out/r02_s8_bench_n1.err
python -c "import json;d=json.loads(open('gpurun_out/r02_s8_bench_n1.json').read().strip().splitlines()[-1]);print('n1',d['value']/1e9,d['ms_per_step'],json.dumps(d.get('callers'))[:1500])"
