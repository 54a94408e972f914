#!/bin/bash
# 1 GPU: default bench line (incl. callers + applications blocks), ncu capture of the general kernel on the curved k=4 box
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_s13_bench_n1.json 2> gpurun_out/r02_s13_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_s13_bench_n1.json').read().strip().splitlines()[-1])
print('n1', d['value']/1e9, d['ms_per_step'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['value']/1e9)
print(json.dumps(d.get('applications'))[:3000])
print(json.dumps(d.get('cpu_baseline'))[:600])
"
tail -3 gpurun_out/r02_s13_bench_n1.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:vmult_general_kernel -s 3 -c 1 -f -o gpurun_out/r02_general_k4_curved_64 python bench.py --degree 4 --cells 64 --mesh curvilinear --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s13_ncu.log 2>&1
tail -3 gpurun_out/r02_s13_ncu.log
