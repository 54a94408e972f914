#!/bin/bash
# degree sweep of BASELINE.json configs[1] (+ curved mesh), one JSON line per run -> gpurun_out/sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for spec in "2 160" "3 128" "4 96" "5 80" "6 64" "7 64"; do
  set -- $spec
  timeout 600 python bench.py --degree $1 --cells $2 --steps 10 --warmup 3 --no-cpu 2>gpurun_out/sweep_err.log | tail -1 >> gpurun_out/sweep.jsonl
done
for spec in "2 160" "3 128" "4 96" "5 80" "6 64" "7 64"; do
  set -- $spec
  timeout 900 python bench.py --degree $1 --cells $2 --mesh curvilinear --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/sweep_err.log | tail -1 >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['config']['workload'][:70], '| %.1f GDoF/s | %.3f ms | frac %.3f | path %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['config']['kernel_path']))
PY
